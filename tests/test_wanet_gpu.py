"""GPU parity of the WaNet variant (train_generator_wanet.py / train_victim_wanet.py): the fused warp kernel and its backward
(csrc/warp.cu) against torch's F.interpolate / F.grid_sample / autograd on the CPU (what the reference calls), nets.GridGenerator
against the oracle restatement, the engine's step / evaluation against oracle.alternated_step_wanet / eval_batch (pinned to the
unmodified reference variant by tests/golden/step_wanet_b32x2.npz), and the public train() / eval() / main() surface against
that fixture."""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import combat_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel2(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _seed(s):
    torch.manual_seed(s); np.random.seed(s); random.seed(s)


@pytest.mark.parametrize("S,H,r,N", [(2, 32, 0.15, 37), (4, 16, 0.9, 5), (3, 64, 2.5, 3), (1, 32, 0.5, 2)])
def test_warp_kernels_vs_torch(S, H, r, N):
    """forward (image, noise grid, both partial sums, batch assembly with a permutation) and backward (gradient w.r.t. the flow,
    clamp mask and zero padding exercised by the large grid_rescale cases) at float32 rounding level."""
    from combat_b200 import ops
    _seed(S * 100 + H)
    x = torch.rand(N, 3, H, H) * 2 - 1
    flow = torch.tanh(torch.randn(N, 2, S, S) * 1.5).requires_grad_(True)
    g1, g2 = torch.randn(N, 3, H, H), torch.randn(N, 3, H, H)
    opt = O.default_opt(input_height=H, input_width=H, grid_rescale=r, s=S)
    ref, ng = O.wanet_warp(x, flow, opt)
    l2_scale = 0.37
    ((ref * (g1 + g2)).sum() + 0.5 * l2_scale * (ng ** 2).sum()).backward()
    ident = torch.linspace(-1, 1, steps=H).cuda()
    xd, fd = x.cuda(), flow.detach().cuda().contiguous()
    ngd = torch.empty(N, H, H, 2, device="cuda")
    sq, gl = torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    out = ops.wanet_warp_fwd(xd, fd, ident, None, N, r, S, noise_grid=ngd, sq_partial=sq, gl_partial=gl)
    # float32 on both sides; with grid_rescale 2.5 most coordinates sit ON the clamp, where one ulp moves a tap boundary
    assert rel(out, ref) < (2e-5 if r < 1 else 1e-4) and rel(ngd, ng) < 2e-6
    assert abs(float(sq.sum()) - float((ng.detach() ** 2).sum())) < 1e-5 * float((ng.detach() ** 2).sum())
    gref = float(O.wanet_grad_l2(ng.detach()))
    assert abs(float(gl.sum()) / N - gref) < 1e-5 * gref
    dz = ops.wanet_warp_bwd(xd, fd, ident, g1.cuda(), g2.cuda(), r, l2_scale, S)
    assert rel(dz.view(N, 2, S, S), flow.grad) < 2e-4, rel(dz.view(N, 2, S, S), flow.grad)
    dz1 = ops.wanet_warp_bwd(xd, fd, ident, (g1 + g2).cuda(), None, r, l2_scale, S)
    assert rel(dz1, dz) < 1e-5
    # C-step batch assembly: rows i < num_bd are warped images of perm[i], the rest bit-exact copies (num_bd on the device)
    perm = torch.randperm(N).int().cuda()
    nbd = max(1, N // 3)
    tot = ops.wanet_warp_fwd(xd, fd, ident, perm, 0, r, S, num_bd_dev=torch.tensor([nbd], dtype=torch.int32, device="cuda"))
    assert torch.equal(tot[nbd:], xd[perm.long()[nbd:]])
    assert torch.equal(tot[:nbd], out[perm.long()[:nbd]])


@pytest.mark.parametrize("N,H", [(6, 32), (1, 32), (3, 64)])
def test_grid_generator_vs_oracle(N, H):
    """nets.GridGenerator forward / backward (float32 path) against the oracle's restatement with autograd, per tensor; N = 1 is
    the `.squeeze()` batch-of-one case of networks/models.py:381."""
    from combat_b200 import nets
    _seed(5)
    p = O.init_grid_generator_state(torch.default_generator, S=2)
    net = nets.GridGenerator(3, 64, 2, device="cuda", dtype=torch.float32)
    net.load_state_dict({k: v.cuda() for k, v in p.items()})
    x = torch.rand(N, 3, H, H) * 2 - 1
    for t in p.values():
        t.requires_grad_(True)
    ref = O.grid_generator_forward(p, x, 2)
    dout = torch.randn_like(ref)
    (ref * dout).sum().backward()
    out, ctx = net.forward(x.cuda(), None, save=True)
    assert tuple(out.shape) == (N, 2, 2, 2) and rel(out, ref) < 1e-4
    net.zero_grad()
    net.backward(ctx, dout.cuda())
    # A property of the reference's architecture: the average pool sits right behind a non-affine InstanceNorm (models.py:380-381),
    # whose output has zero mean per (sample, channel) -- so the pooled feature is 0 up to rounding, the flow does not depend on
    # the image, and every encoder gradient (and fc1.weight's) is rounding noise in the reference too (~1e-8 here).  Only
    # fc1.bias and fc2 carry a signal; all tensors are compared on the scale of the largest gradient.
    assert float(ctx["pooled"].abs().max()) < 1e-5
    scale = max(float(t.grad.norm()) for t in p.values())
    for k, t in p.items():
        err = float((net.store.g(k).cpu().double() - t.grad.double()).norm())
        assert err < 2e-3 * scale, (k, err, scale)
    for k in ("fc1.bias", "fc2.weight", "fc2.bias"):
        assert rel2(net.store.g(k), p[k].grad) < 1e-4, (k, rel2(net.store.g(k), p[k].grad))


def _state(seed):
    _seed(seed)
    gen = torch.default_generator
    netC_p, netC_b = O.init_preact_resnet18_state(gen)
    clean_p, clean_b = O.init_preact_resnet18_state(gen)
    netG_p = O.init_grid_generator_state(gen, S=2)
    netF_p, netF_b = O.init_frequency_model_state(gen)
    return dict(netC_p=netC_p, netC_b=netC_b, clean_p=clean_p, clean_b=clean_b, netG_p=netG_p, netF_p=netF_p, netF_b=netF_b,
                momC={}, momG={})


@pytest.mark.parametrize("tf", ["no_use", "use"])
@pytest.mark.parametrize("name,dtype,use_graph", [("fp32", torch.float32, False), ("fp32", torch.float32, True),
                                                  ("bf16", torch.bfloat16, False)])
def test_wanet_iterations_vs_oracle(name, dtype, use_graph, tf):
    from test_step_gpu import make_engine, net_delta
    from combat_b200.engine import AlternatedStep, make_plan
    B = 32
    state = _state(29)
    opt = O.default_opt(variant="wanet", post_transform_option=tf)
    eng = make_engine(state, dtype, opt=opt)
    assert eng.wanet and eng.S == 2
    before = {k: {n: v.clone() for n, v in state[k].items()} for k in ("netC_p", "netG_p")}
    n_it = 3 if use_graph else 2
    g = torch.Generator().manual_seed(2)
    batches = []
    for _ in range(n_it):
        y = torch.randint(0, 10, (B,), generator=g)
        y[:6] = 0   # enough target-class rows that every iteration poisons some (the reference raises for num_bd == 0)
        batches.append((torch.rand(B, 3, 32, 32, generator=g) * 2 - 1, y))
    _seed(8)
    refs, snaps = [], []
    for x, y in batches:
        refs.append(O.alternated_step_wanet(state, x, y, opt))
        snaps.append({k: {n: v.clone() for n, v in state[k].items()} for k in ("netC_p", "netG_p")})
    _seed(8)
    fp32 = name == "fp32"
    for it, ((x, y), r) in enumerate(zip(batches, refs)):
        plan = make_plan(y.numpy(), eng.opt)
        assert plan.num_bd == r["num_bd"] > 0 and plan.sigma_c is None and plan.sigma_g is None
        assert np.array_equal(plan.total_targets, r["total_y"].numpy())
        if tf == "use":
            assert len(r["tf"]) == 5
            for slot, call in ((0, 0), (3, 1), (1, 2), (2, 3), (4, 4)):
                prm = r["tf"][call]
                assert np.array_equal(plan.tf[slot][:, 0], (prm["xs"] - prm["pad"]).numpy().astype(np.float32))
                assert np.array_equal(plan.tf[slot][:, 5] != 0, prm["flip"].numpy())
        out = eng.step(x.cuda(), y.numpy(), plan, use_graph=use_graph, keep_debug=not use_graph)
        s = AlternatedStep.unpack(out)
        ltol = (2e-5 if it == 0 else 2e-3) if fp32 else 1e-2
        for k in ("loss_c", "loss_ce", "loss_l2", "clean_model_loss", "loss_grad_l2"):
            assert abs(s[k] - r[k]) < ltol * max(1.0, abs(r[k])), (it, k, s[k], r[k])
        if not use_graph:
            d = out["debug"]
            if tf == "no_use":
                assert torch.equal(d["total_x"][plan.num_bd:].cpu(), r["total_x"][plan.num_bd:])   # gathered rows: bit-exact
            for k, dk in (("flow", "noise_raw"), ("x_bd", "x_bd"), ("logits_c", "logits_c"), ("pred_bd", "pred_bd"),
                          ("clean_model_preds", "clean_model_preds"), ("clean_preds", "clean_preds"), ("pred_clean", "pred_clean")):
                if fp32:
                    assert rel(d[dk], r[k]) < (1e-4 if it == 0 else 3e-3), (it, k, rel(d[dk], r[k]))
                else:
                    assert rel2(d[dk], r[k]) < 5e-2, (it, k, rel2(d[dk], r[k]))
            if fp32 and it == 0:
                for k in ("n_clean_correct", "n_bd_correct", "n_clean_model_correct", "n_clean_model_bd_ba", "n_clean_model_bd_asr"):
                    assert s[k] == r[k], k
        eC, cC = net_delta(before["netC_p"], snaps[it]["netC_p"], eng.netC.state_dict())
        eG, cG = net_delta(before["netG_p"], snaps[it]["netG_p"], eng.netG.state_dict(), skip_dead=False)
        print("wanet update after iteration %d: netC L2 err %.3e cos %.5f | netG L2 err %.3e cos %.5f" % (it + 1, eC, cC, eG, cG))
        if fp32:
            bar = 2e-2 if it == 0 else 8e-2
            assert eC < bar and eG < bar, (it, eC, eG)
        else:
            assert cC > 0.9 and cG > 0.9, (it, cC, cG)


@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "graph"])
def test_wanet_eval_step_vs_oracle(use_graph):
    from combat_b200.engine import AlternatedStep
    st = _state(31)
    opt = O.default_opt(variant="wanet")
    eng = AlternatedStep(opt=opt, device="cuda", dtype=torch.float32)
    j = lambda p, b: {**p, **b}
    eng.load_state(netC=j(st["netC_p"], st["netC_b"]), clean=j(st["clean_p"], st["clean_b"]), netG=st["netG_p"],
                   netF=j(st["netF_p"], st["netF_b"]))
    g = torch.Generator().manual_seed(4)
    for it in range(3):
        x = torch.rand(40, 3, 32, 32, generator=g) * 2 - 1
        y = torch.randint(0, 10, (40,), generator=g)
        state0 = torch.get_rng_state()
        r = O.eval_batch(st, x, y, opt)
        out = eng.eval_step(x.cuda(), y.numpy(), use_graph=use_graph)
        assert torch.equal(torch.get_rng_state(), state0)    # nothing drawn on either side
        c = out["counts"].cpu().numpy()
        got = dict(clean_correct=c[0], bd_correct=c[2], F_correct=c[4], cm_correct=c[6], cm_bd_ba=c[8], cm_bd_asr=c[9])
        for k, v in got.items():
            assert int(v) == r[k], (it, k, int(v), r[k])
        if out["debug"] is not None:
            nt = r["ntrg"].cuda()
            assert rel(out["debug"]["x_bd"][nt], r["x_bd"]) < 2e-5 and rel(out["debug"]["preds_bd"][nt], r["preds_bd"]) < 2e-4


def test_wanet_train_and_eval_reproduce_the_reference_fixture(golden, tmp_path):
    """train_generator_wanet.{get_model, train, eval} through the public API against two iterations of the UNMODIFIED reference
    variant (tests/golden/step_wanet_b32x2.npz)."""
    from test_api_gpu import _opt, _Writer
    from combat_b200 import train_generator_wanet as tw
    g = golden("step_wanet_b32x2.npz")
    seed, B, nb = int(g["seed"]), int(g["B"]), int(g["n_batches"])
    opt = _opt(["--dtype", "fp32", "--no_graph", "--log_every", "1"])
    assert opt.s == int(g["s"]) and opt.grid_rescale == float(g["grid_rescale"])
    _seed(seed)
    netC, optC, schC, netG, optG, schG, netF, clean = tw.get_model(opt)
    batches = [(torch.rand(B, 3, 32, 32) * 2 - 1, torch.randint(0, 10, (B,))) for _ in range(nb)]
    for i in range(nb):
        assert np.array_equal(batches[i][1].numpy(), g["y_%d" % i])
    # same seed, same construction order: the reference's initial flow on the first batch
    with torch.no_grad():
        flow0 = netG(batches[0][0].cuda())
    assert rel(flow0, torch.from_numpy(g["flow_0"])) < 1e-4
    a = torch.linspace(-1, 1, steps=32)
    gx, gy = torch.meshgrid(a, a, indexing="ij")
    ident = torch.stack((gy, gx), 2)[None, ...].cuda()
    sd0 = {n: {k: v.detach().clone().cpu() for k, v in m.state_dict().items()} for n, m in (("netC_", netC), ("netG_", netG))}
    w = _Writer()
    tw.train(netC, optC, schC, netG, optG, schG, netF, clean, batches, ident, w, 1, opt)
    torch.cuda.synchronize()
    sc = w.scalars[0][1]
    vals = g["loss_values"]           # per iteration: ce(C), ce(bd), mse(noise grid), mse, mse (logged pair), ce(clean)
    l2 = vals[2] + vals[8]
    assert abs(sc["L2 Loss"] * B * nb - l2) < 2e-3 * l2
    gl = vals[3] + vals[4] + vals[9] + vals[10]
    assert abs(sc["Grad L2 Loss"] * B * nb - gl) < 2e-3 * gl
    bad = []
    for pre, mod in (("netC_", netC), ("netG_", netG)):
        sd = mod.state_dict()
        for n, v0 in sd0[pre].items():
            if not torch.is_floating_point(v0) or (pre + "dnorm_" + n) not in g.files:
                continue
            if pre == "netG_" and n.endswith("bias") and n.startswith("conv") and n != "conv0_0.bias":
                continue
            d = float((sd[n].detach().cpu() - v0).double().norm())
            ref = g[pre + "dnorm_" + n][0]
            if abs(d - ref) > 3e-2 * ref + 1e-12:
                bad.append((pre + n, d, float(ref)))
    assert not bad, bad
    opt.ckpt_path = str(tmp_path / "wn.pth.tar")
    bests = tw.eval(netC, optC, schC, netG, optG, schG, netF, clean, batches, ident, -1.0, 0.0, 0.0, 0.0, 0.0, 0.0, w, 1, opt)
    assert len(bests) == 6
    ck = torch.load(opt.ckpt_path, map_location="cpu", weights_only=False)
    assert set(ck["netG"]) == set(netG.state_dict()) and "fc2.weight" in ck["netG"]
    with pytest.raises(NotImplementedError):
        tw.train(netC, optC, schC, netG, optG, schG, netF, clean, batches, ident * 0.5, w, 1, opt)


def test_wanet_victim_trainer_public_api(tmp_path, capsys):
    """train_victim_wanet: eval_batch against the oracle's warp on the gathered non-target rows, then main() on synthetic data."""
    from combat_b200 import config
    from combat_b200 import train_victim_wanet as tvw
    opt = config.get_arguments().parse_args(["--device", "cuda", "--dtype", "fp32"])
    opt.input_height = opt.input_width = 32
    opt.input_channel = 3
    _seed(4)
    netC, optC, schC, netG = tvw.get_model(opt)
    sdC = {k: v.detach().cpu().clone() for k, v in netC.state_dict().items()}
    sdG = {k: v.detach().cpu().clone() for k, v in netG.state_dict().items()}
    netC_p, netC_b = O.split_state(sdC)
    o = O.default_opt(variant="wanet")
    g = torch.Generator().manual_seed(9)
    x = torch.rand(24, 3, 32, 32, generator=g) * 2 - 1
    y = torch.randint(0, 10, (24,), generator=g)
    counts, nb, d = tvw.eval_batch(netC, netG, x, y, tvw._variant(opt))
    ntrg = (y != 0).nonzero()[:, 0]
    with torch.no_grad():
        x_bd = O.wanet_warp(x[ntrg], O.grid_generator_forward(sdG, x[ntrg], 2), o)[0]
        preds_bd = O.preact_resnet18_forward(netC_p, netC_b, x_bd, False)
        preds_clean = O.preact_resnet18_forward(netC_p, netC_b, x, False)
    assert nb == len(ntrg) and rel(d["x_bd"][ntrg.cuda()], x_bd) < 2e-5 and rel(d["preds_bd"][ntrg.cuda()], preds_bd) < 2e-4
    c = counts.cpu().numpy()
    assert int(c[0]) == int((preds_clean.argmax(1) == y).sum()) and int(c[2]) == int((preds_bd.argmax(1) == 0).sum())
    args = ["--synthetic_data", "--debug", "--bs", "32", "--n_iters", "1", "--log_every", "4", "--saving_prefix", "vw",
            "--checkpoints", str(tmp_path), "--load_checkpoint", "none"]
    _seed(0)
    best = tvw.main(args)
    out = capsys.readouterr().out
    assert "CE Loss" in out and "Bd Acc" in out and len(best) == 2
    ck = torch.load(str(tmp_path / "vw_clean" / "cifar10" / "cifar10_vw_clean.pth.tar"), map_location="cpu", weights_only=False)
    assert ck["grid_rescale"] == opt.grid_rescale
