"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu`):
every rank runs the CUDA engine on its shard with NCCL exchange AND the CPU oracle on the same shard with gloo
exchange through the same combat_b200.parallel.GradSync hooks, then compares losses, selections and updates."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

WORLD = 2
B_LOCAL = 32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, outdir, use_graph):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from combat_b200 import parallel
    from combat_b200.engine import AlternatedStep, make_plan
    from oracle import combat_oracle as O
    torch.cuda.set_device(rank)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", rank))
    sync = parallel.GradSync()
    g = torch.Generator().manual_seed(100 + rank)
    batches = [(torch.rand(B_LOCAL, 3, 32, 32, generator=g) * 2 - 1, torch.randint(0, 10, (B_LOCAL,), generator=g)) for _ in range(2)]
    # oracle with gloo exchange
    state = O.init_step_state(3)
    init = {k: {n: v.clone() for n, v in state[k].items()} for k in ("netC_p", "netG_p")}
    parallel.seed_rank(11, rank)
    refs = [O.alternated_step(state, x, y, O.default_opt(), grad_hook=sync.grad_hook, buf_hook=sync.buf_hook) for x, y in batches]
    # engine with NCCL exchange
    st0 = O.init_step_state(3)
    eng = AlternatedStep(device="cuda:%d" % rank, dtype=torch.float32, grad_hook=sync.grad_hook, buf_hook=sync.buf_hook)
    eng.load_state(netC={**st0["netC_p"], **st0["netC_b"]}, clean={**st0["clean_p"], **st0["clean_b"]}, netG=st0["netG_p"],
                   netF={**st0["netF_p"], **st0["netF_b"]})
    parallel.seed_rank(11, rank)
    res = []
    for (x, y), r in zip(batches, refs):
        plan = make_plan(y.numpy(), eng.opt)
        assert plan.num_bd == r["num_bd"]
        out = AlternatedStep.unpack(eng.step(x.cuda(), y.numpy(), plan, use_graph=use_graph))
        res.append({k: (out[k], r[k]) for k in ("loss_c", "loss_ce", "loss_l2", "clean_model_loss")})
    sdC, sdG = eng.netC.state_dict(), eng.netG.state_dict()

    def delta_err(sd, key):
        num = den = 0.0
        for n, v0 in init[key].items():
            if key == "netG_p" and n.endswith("bias") and n not in ("conv0_0.bias", "upconv0_0.bias"):
                continue
            d_ref = (state[key][n] - v0).double()
            d_dev = (sd[n].cpu() - v0).double()
            num += float(((d_dev - d_ref) ** 2).sum())
            den += float((d_ref ** 2).sum())
        return (num / den) ** 0.5

    flatC = eng.netC.store.flat.clone()
    other = [torch.empty_like(flatC) for _ in range(WORLD)]
    dist.all_gather(other, flatC)
    torch.save({"losses": res, "errC": delta_err(sdC, "netC_p"), "errG": delta_err(sdG, "netG_p"),
                "replicas_equal": bool(torch.equal(other[0], other[1])),
                "bn_err": float((sdC["layer1.0.bn1.running_mean"].cpu() - state["netC_b"]["layer1.0.bn1.running_mean"]).abs().max())},
               os.path.join(outdir, "rank%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("use_graph", [False, True])
def test_two_gpu_step_matches_the_data_parallel_oracle(tmp_path, use_graph):
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs %d GPUs" % WORLD)
    mp.spawn(_worker, args=(_free_port(), str(tmp_path), use_graph), nprocs=WORLD, join=True)
    for rank in range(WORLD):
        r = torch.load(os.path.join(str(tmp_path), "rank%d.pt" % rank))
        assert r["replicas_equal"]
        for it, losses in enumerate(r["losses"]):
            for k, (dev, ref) in losses.items():
                assert abs(dev - ref) < (2e-5 if it == 0 else 1e-3) * max(1.0, abs(ref)), (rank, it, k, dev, ref)
        # fp32 mode, two iterations: same bar as the single-GPU test (tests/test_step_gpu.py)
        assert r["errC"] < 2e-2 and r["errG"] < 2e-2, (r["errC"], r["errG"])
        assert r["bn_err"] < 1e-4
