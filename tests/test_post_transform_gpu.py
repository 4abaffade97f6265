"""GPU parity of PostTensorTransform (csrc/augment.cu) against the oracle's restatement of the kornia pipeline
(utils/dataloader.py:45-60): pixels and gradients for given parameters, the nn.Module drop-in, and the whole alternated
step under --post_transform_option use (the reference's default) against the oracle step that draws the same decisions."""
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import combat_oracle as O  # noqa: E402


def _opt(option="use", dataset="cifar10"):
    return SimpleNamespace(post_transform_option=option, random_crop=5, random_rotation=10, dataset=dataset)


def _seed(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


@pytest.mark.parametrize("shape", [(7, 3, 32, 32), (5, 3, 64, 64), (3, 1, 20, 28), (2, 3, 224, 224)])
def test_kernel_vs_restatement_fwd_bwd(shape):
    from combat_b200 import ops
    from combat_b200.utils.dataloader import draw_params
    rows = shape[0]
    seen = set()
    for seed in range(8):
        opt = _opt("use" if seed % 3 else "use_modified", "cifar10" if seed % 2 else "celeba")
        _seed(100 + seed)
        prm = O.draw_post_transform(rows, opt)
        _seed(100 + seed)
        P = draw_params(rows, opt)
        seen.add((prm["crop"], prm["rot"], bool(prm["flip"].any())))
        x = (torch.rand(shape) * 2 - 1).requires_grad_(True)
        ref = O.apply_post_transform(x, prm)
        g = torch.randn(shape)
        ref.backward(g)
        Pd = torch.from_numpy(P).cuda()
        got = ops.post_transform_fwd(x.detach().cuda(), Pd)
        exact = not prm["rot"]
        # rotation: the restatement goes through kornia's normalised coordinates ((W-1)/2 scaling, two 3x3 inverses, affine_grid)
        # in float32, the kernel rotates in pixel space: the sample position differs by ~1e-7 * W pixels, times the image
        # slope (values in [-1, 1]) -- measured 2.1e-5 at 64x64
        tol = 0.0 if exact else 1e-6 * max(shape[2], shape[3]) + 1e-5
        assert (got.cpu() - ref.detach()).abs().max() <= tol, (seed, prm["crop"], prm["rot"])
        dx = ops.post_transform_bwd(g.cuda(), Pd)
        assert (dx.cpu() - x.grad).abs().max() <= (1e-6 if exact else 6 * tol), seed
        # accumulate=True adds a second adjoint into the same buffer (T4 + T5 of the G-step)
        ops.post_transform_bwd(g.cuda(), Pd, out=dx, accumulate=True)
        assert (dx.cpu() - 2 * x.grad).abs().max() <= max(1e-4, 12 * tol)
    assert len(seen) >= 3


def test_module_drop_in_autograd():
    from combat_b200.utils.dataloader import PostTensorTransform
    opt = _opt("use", "cifar10")
    tf = PostTensorTransform(opt)
    x = (torch.rand(6, 3, 32, 32) * 2 - 1)
    _seed(7)
    ref_in = x.clone().requires_grad_(True)
    ref = O.post_transform(ref_in, opt)
    (ref * ref).sum().backward()
    _seed(7)
    xin = x.cuda().requires_grad_(True)
    y = tf(xin)
    (y * y).sum().backward()
    assert (y.detach().cpu() - ref.detach()).abs().max() < 2e-5
    assert (xin.grad.cpu() - ref_in.grad).abs().max() < 1e-4
    assert PostTensorTransform(_opt("no_use"))(xin) is xin
    with pytest.raises(RuntimeError):
        tf(x)  # CPU tensor: no fallback


@pytest.mark.parametrize("use_graph", [False, True])
def test_step_with_default_transform_option_vs_oracle(use_graph):
    """One (eager) / three (graph-replayed) alternated iterations with the reference's DEFAULT --post_transform_option use,
    float32 path: the engine and the oracle consume `random`, numpy and the torch generator in the same order, so every
    crop / rotation / flip decision is identical and the step agrees to the float32 bars of test_step_gpu."""
    from test_step_gpu import make_engine, rel, seeded_state
    from combat_b200.engine import AlternatedStep, default_opt, make_plan
    B = 32
    state = seeded_state(33)
    opt_o = O.default_opt(post_transform_option="use")
    eng = make_engine(state, torch.float32, opt=default_opt(post_transform_option="use"))
    n_it = 3 if use_graph else 1
    batches = [(torch.rand(B, 3, 32, 32) * 2 - 1, torch.randint(0, 10, (B,))) for _ in range(n_it)]
    _seed(11)
    refs = [O.alternated_step(state, x, y, opt_o) for x, y in batches]
    assert all(len(r["tf"]) == 5 for r in refs)
    _seed(11)
    for it, ((x, y), r) in enumerate(zip(batches, refs)):
        plan = make_plan(y.numpy(), eng.opt)
        assert plan.num_bd == r["num_bd"] and plan.sigma_g == r["sigma_g"] and plan.sigma_c == r["sigma_c"]
        # storage order T1 | T3 | T4 | T2 | T5 vs call order T1, T2, T3, T4, T5
        for slot, call in ((0, 0), (3, 1), (1, 2), (2, 3), (4, 4)):
            prm = r["tf"][call]
            assert np.array_equal(plan.tf[slot][:, 0], (prm["xs"] - prm["pad"]).numpy().astype(np.float32))
            assert np.array_equal(plan.tf[slot][:, 5] != 0, prm["flip"].numpy())
            assert bool(plan.tf[slot][0, 4]) == prm["rot"]
        out = eng.step(x.cuda(), y.numpy(), plan, use_graph=use_graph, keep_debug=not use_graph)
        s = AlternatedStep.unpack(out)
        tol = 2e-5 if it == 0 else 2e-3
        for k in ("loss_c", "loss_ce", "loss_l2", "clean_model_loss", "loss_grad_l2"):
            assert abs(s[k] - r[k]) < tol * max(1.0, abs(r[k])), (it, k, s[k], r[k])
        if not use_graph:
            d = out["debug"]
            for k in ("x_bd", "logits_c", "pred_bd", "clean_model_preds", "clean_preds", "pred_clean"):
                assert rel(d[k], r[k]) < 1e-4, (k, rel(d[k], r[k]))
            for k in ("n_clean_correct", "n_bd_correct", "n_clean_model_correct", "n_clean_model_bd_ba", "n_clean_model_bd_asr"):
                assert s[k] == r[k], k
    # the generator update went through the adjoint of T4 / T5
    sdG = eng.netG.state_dict()
    num = sum(float(((sdG[n].cpu() - state["netG_p"][n]) ** 2).sum()) for n in state["netG_p"])
    den = sum(float((state["netG_p"][n] ** 2).sum()) for n in state["netG_p"])
    assert (num / den) ** 0.5 < 1e-3
