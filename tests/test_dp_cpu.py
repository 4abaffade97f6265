"""Data-parallel semantics on the CPU (gloo, world size 2): combat_b200.parallel's exchange points reproduce
"the reference step per shard from identical weights, gradients averaged, BatchNorm buffers averaged" (SURVEY.md 8e,
local-BN policy).  The arithmetic is the CPU oracle's; what is under test is the host-side DP logic the CUDA engine
uses unchanged (GradSync hooks, sharding, per-rank RNG streams)."""
import copy
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import combat_oracle as O

WORLD = 2
B_GLOBAL = 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch():
    g = torch.Generator().manual_seed(7)
    x = torch.rand(B_GLOBAL, 3, 32, 32, generator=g) * 2 - 1
    y = torch.randint(0, 10, (B_GLOBAL,), generator=g)
    y[0] = y[1] = y[2] = 0   # target-class rows in both shards
    y[8] = y[9] = y[10] = 0
    return x, y


def _worker(rank, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from combat_b200 import parallel
    r, w, _ = parallel.init(device="cpu")
    assert (r, w) == (rank, WORLD)
    sync = parallel.GradSync()
    state = O.init_step_state(3)
    x, y = _batch()
    lo, hi = parallel.shard_rows(B_GLOBAL, rank, WORLD)
    parallel.seed_rank(11, rank)
    out = O.alternated_step(state, x[lo:hi], y[lo:hi], O.default_opt(), with_metrics=False, grad_hook=sync.grad_hook,
                            buf_hook=sync.buf_hook)
    torch.save({"netC_p": state["netC_p"], "netC_b": state["netC_b"], "netG_p": state["netG_p"], "gradsC": out["gradsC"],
                "gradsG": out["gradsG"], "total_x": out["total_x"], "num_bd": out["num_bd"], "bytes": sync.bytes},
               os.path.join(outdir, "rank%d.pt" % rank))
    torch.distributed.destroy_process_group()


@pytest.fixture(scope="module")
def dp_run(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("dp"))
    mp.spawn(_worker, args=(_free_port(), d), nprocs=WORLD, join=True)
    return [torch.load(os.path.join(d, "rank%d.pt" % r)) for r in range(WORLD)]


def test_replicas_stay_identical(dp_run):
    a, b = dp_run
    for key in ("netC_p", "netG_p", "netC_b"):
        for n in a[key]:
            assert torch.equal(a[key][n], b[key][n]), (key, n)
    # payload: both flat gradients + the floating-point BatchNorm buffers, once each
    n_c = sum(v.numel() for v in a["netC_p"].values())
    n_g = sum(v.numel() for v in a["netG_p"].values())
    n_b = sum(v.numel() for v in a["netC_b"].values() if v.is_floating_point())
    assert a["bytes"] == 4 * (n_c + n_g + n_b)


def test_update_is_sgd_on_the_mean_of_shard_gradients(dp_run):
    init = O.init_step_state(3)
    for net, gkey, lr in (("netC_p", "gradsC", 1e-2), ("netG_p", "gradsG", 1e-2)):
        p = {k: v.clone() for k, v in init[net].items()}
        mean = {k: (dp_run[0][gkey][k] + dp_run[1][gkey][k]) / 2 for k in p}
        O.sgd_nesterov_step(p, mean, {}, lr)
        for k in p:
            assert torch.allclose(p[k], dp_run[0][net][k], rtol=0, atol=1e-7), (net, k)


def test_shard_gradients_equal_the_single_process_reference_step(dp_run):
    """Before the exchange point each rank IS the reference step on its shard (same weights, seed + rank)."""
    from combat_b200 import parallel
    x, y = _batch()
    for rank in range(WORLD):
        state = O.init_step_state(3)
        lo, hi = parallel.shard_rows(B_GLOBAL, rank, WORLD)
        parallel.seed_rank(11, rank)
        out = O.alternated_step(state, x[lo:hi], y[lo:hi], O.default_opt(), with_metrics=False)
        assert out["num_bd"] == dp_run[rank]["num_bd"]
        for k, g in out["gradsC"].items():  # not bit-equal: the workers run with another host thread count
            ref = dp_run[rank]["gradsC"][k]
            assert float((g - ref).norm()) <= 1e-3 * float(ref.norm()) + 1e-7, k


def test_batchnorm_buffers_are_the_mean_of_local_statistics(dp_run):
    init = O.init_step_state(3)
    local = []
    for rank in range(WORLD):
        b = copy.deepcopy(init["netC_b"])
        with torch.no_grad():
            O.preact_resnet18_forward(init["netC_p"], b, dp_run[rank]["total_x"], True)
        local.append(b)
    for n, v in dp_run[0]["netC_b"].items():
        if v.is_floating_point():
            assert torch.allclose(v, (local[0][n] + local[1][n]) / 2, rtol=1e-5, atol=1e-7), n


def test_shard_rows_and_seeds():
    from combat_b200 import parallel
    assert [parallel.shard_rows(4096, r, 8) for r in (0, 7)] == [(0, 512), (3584, 4096)]
    with pytest.raises(ValueError):
        parallel.shard_rows(10, 0, 4)
    parallel.seed_rank(5, 1)
    a = (np.random.rand(), float(torch.rand(1)))
    parallel.seed_rank(5, 1)
    assert a == (np.random.rand(), float(torch.rand(1)))
    parallel.seed_rank(5, 2)
    assert a != (np.random.rand(), float(torch.rand(1)))
