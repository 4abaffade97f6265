"""Data parallelism for the alternated step: one process per GPU, sample-sharded batches, full model replicas.

The reference is single-process; this is the one strategy that applies (section 8e).  Exchange
steps per iteration, all over `torch.distributed` (NCCL on NVLink/NVSwitch on the GPU box, gloo in the CPU tests):
  * mean of the flat netC gradient after the C-step backward   (train_generator.py:211 -> :212)
  * mean of the BatchNorm running-stat buffer after the C-step (local-BN policy: batch statistics are rank-local, the
    buffers the eval-mode G-step reads at :227-228 are kept identical across ranks)
  * mean of the flat netG gradient after the G-step backward   (:254 -> :255)
Each rank draws its own poison count / blur sigmas (seed + rank), as W independent reference processes would.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment (1 process -> (0, 1, 0))."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(device=None, backend=None):
    """Initialises the default process group when WORLD_SIZE > 1.  Returns (rank, world, local_rank)."""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if (device is not None and torch.device(device).type == "cuda") else "gloo"
        kw = {"device_id": torch.device(device)} if backend == "nccl" and device is not None else {}
        dist.init_process_group(backend, **kw)
    return rank, world, local


def seed_rank(seed: int, rank: int):
    """Per-rank RNG streams for the host-side draws of make_plan (numpy global + torch CPU generator)."""
    np.random.seed(seed + rank)
    torch.manual_seed(seed + rank)


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous row range of this rank (the last ranks get the remainder-free floor share; n_rows % world must be 0
    for the averaged gradient to equal the global-batch mean)."""
    if n_rows % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (n_rows, world))
    per = n_rows // world
    return rank * per, (rank + 1) * per


def allreduce_mean_(t: torch.Tensor, group=None):
    """In-place mean over ranks.  NCCL: one AVG all-reduce (graph-capturable); gloo has no AVG: SUM then scale."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    if t.is_cuda:
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.mul_(1.0 / dist.get_world_size(group))
    return t


class GradSync:
    """The two hooks `engine.AlternatedStep` calls: grad_hook(net_name, flat_grad) and buf_hook(flat_running_stats).
    Also accepts dicts / lists of tensors (the CPU oracle's per-parameter gradients in tests/test_dp_cpu.py)."""

    def __init__(self, group=None):
        self.group = group
        self.bytes = 0  # payload all-reduced so far (reporting)

    def _each(self, obj):
        if torch.is_tensor(obj):
            yield obj
        elif isinstance(obj, dict):
            for k in sorted(obj):
                if torch.is_tensor(obj[k]) and obj[k].is_floating_point():
                    yield obj[k]
        else:
            for t in obj:
                yield t

    def grad_hook(self, name, grads):
        for t in self._each(grads):
            allreduce_mean_(t, self.group)
            self.bytes += t.numel() * t.element_size()

    def buf_hook(self, bufs):
        for t in self._each(bufs):
            allreduce_mean_(t, self.group)
            self.bytes += t.numel() * t.element_size()
