"""GPU parity of the evaluation loop (train_generator.py:321-465): the engine's fixed-shape
eval_step against the oracle's gathered sub-batch, and the public eval() against the fixture recorded from the unmodified
reference (tests/golden/eval_b64x2.npz)."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import combat_oracle as O  # noqa: E402


def rel(a, b):
    a = a.detach().float().cpu().double()
    b = b.detach().float().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "graph"])
def test_eval_step_vs_oracle(dtype, use_graph):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from combat_b200.engine import AlternatedStep
    st = O.init_step_state(6)
    eng = AlternatedStep(device="cuda", dtype=dtype)
    j = lambda p, b: {**p, **b}
    eng.load_state(netC=j(st["netC_p"], st["netC_b"]), clean=j(st["clean_p"], st["clean_b"]), netG=st["netG_p"],
                   netF=j(st["netF_p"], st["netF_b"]))
    g = torch.Generator().manual_seed(3)
    fp32 = dtype == torch.float32
    for it in range(3):
        x = torch.rand(48, 3, 32, 32, generator=g) * 2 - 1
        y = torch.randint(0, 10, (48,), generator=g)
        torch.manual_seed(40 + it)
        r = O.eval_batch(st, x, y, O.default_opt())
        torch.manual_seed(40 + it)
        out = eng.eval_step(x.cuda(), y.numpy(), use_graph=use_graph)
        c = out["counts"].cpu().numpy()
        assert out["sigma"] == r["sigma"] and out["n_bd"] == r["n_bd"]            # RNG draw and row selection: bit-exact
        got = dict(clean_correct=c[0], bd_correct=c[2], F_correct=c[4], cm_correct=c[6], cm_bd_ba=c[8], cm_bd_asr=c[9])
        slack = 0 if fp32 else 3   # bf16 logits of a random-init net flip a few near-tied argmaxes
        for k, v in got.items():
            assert abs(int(v) - r[k]) <= slack, (it, k, int(v), r[k])
        if out["debug"] is not None:
            d, nt = out["debug"], r["ntrg"].cuda()
            tol = 2e-4 if fp32 else 6e-2
            assert rel(d["preds_clean"], r["preds_clean"]) < tol and rel(d["cm_clean"], r["cm_clean"]) < tol
            assert rel(d["x_bd"][nt], r["x_bd"]) < (1e-5 if fp32 else 3e-2)
            assert rel(d["preds_bd"][nt], r["preds_bd"]) < tol and rel(d["cm_bd"][nt], r["cm_bd"]) < tol


@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "graph"])
def test_inputaware_eval_step_vs_oracle(use_graph):
    """train_generator_inputaware.py:376-413: the base evaluation + the cross-trigger accuracy (trigger of the second loader's
    rows on this batch, its own sigma draw, non-target rows against their TRUE labels)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from combat_b200.engine import AlternatedStep
    st = O.init_step_state(8)
    opt = O.default_opt(variant="inputaware")
    eng = AlternatedStep(opt=opt, device="cuda", dtype=torch.float32)
    j = lambda p, b: {**p, **b}
    eng.load_state(netC=j(st["netC_p"], st["netC_b"]), clean=j(st["clean_p"], st["clean_b"]), netG=st["netG_p"],
                   netF=j(st["netF_p"], st["netF_b"]))
    g = torch.Generator().manual_seed(4)
    for it in range(3):
        x = torch.rand(40, 3, 32, 32, generator=g) * 2 - 1
        x2 = torch.rand(40, 3, 32, 32, generator=g) * 2 - 1
        y = torch.randint(0, 10, (40,), generator=g)
        torch.manual_seed(50 + it)
        r = O.eval_batch(st, x, y, opt, x2=x2)
        torch.manual_seed(50 + it)
        out = eng.eval_step(x.cuda(), y.numpy(), use_graph=use_graph, x2=x2.cuda())
        c = out["counts"].cpu().numpy()
        assert out["sigma"] == r["sigma"] and out["sigma2"] == r["sigma2"] and out["n_bd"] == r["n_bd"]
        got = dict(clean_correct=c[0], bd_correct=c[2], cross_correct=c[12], F_correct=c[4], cm_correct=c[6], cm_bd_ba=c[8], cm_bd_asr=c[9])
        for k, v in got.items():
            assert int(v) == r[k], (it, k, int(v), r[k])
        if out["debug"] is not None:
            d = out["debug"]
            assert rel(d["x_bd2"], r["x_bd2"]) < 1e-5 and rel(d["preds_cross"], r["preds_cross"]) < 2e-4
    with pytest.raises(ValueError):
        eng.eval_step(x.cuda(), y.numpy())   # the second batch is not optional in this variant


def test_eval_api_reproduces_the_reference_fixture(golden, tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from combat_b200 import config
    from combat_b200 import train_generator as tg
    g = golden("eval_b64x2.npz")
    seed, B, nb = int(g["seed"]), int(g["B"]), int(g["n_batches"])
    opt = config.get_arguments().parse_args(["--device", "cuda", "--post_transform_option", "no_use", "--dtype", "fp32"])
    opt.input_height = opt.input_width = 32
    opt.input_channel = 3
    opt.ckpt_path = os.path.join(str(tmp_path), "ck", "ckpt.pth.tar")
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    netC, optC, schC, netG, optG, schG, netF, clean = tg.get_model(opt)
    batches = [(torch.rand(B, 3, 32, 32) * 2 - 1, torch.randint(0, 10, (B,))) for _ in range(nb)]
    assert np.array_equal(batches[0][1].numpy(), g["y_0"])

    class W:
        def __init__(self):
            self.s = []

        def add_scalars(self, tag, d, epoch):
            self.s.append((tag, d))

    w = W()
    best = tg.eval(netC, optC, schC, netG, optG, schG, netF, clean, batches, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, w, 1, opt)
    n_bd = sum(int((y != 0).sum()) for _, y in batches)
    one = [100.0 / (B * nb), 100.0 / n_bd, 100.0 / n_bd, 100.0 / (B * nb), 100.0 / n_bd, 100.0 / n_bd]
    for got, ref, q in zip(best, g["best"], one):
        assert abs(float(got) - float(ref)) <= q * 1.001, (best, g["best"])   # at most one near-tied sample apart
    assert w.s and w.s[0][0] == "Test Accuracy"
    ck = torch.load(opt.ckpt_path, weights_only=False)
    assert sorted(ck.keys()) == list(g["ckpt_keys"])
    assert list(ck["netC"].keys()) == list(g["ckpt_netC_keys"])
    # a second call with the bests just reached saves nothing and returns them unchanged
    os.remove(opt.ckpt_path)
    best2 = tg.eval(netC, optC, schC, netG, optG, schG, netF, clean, batches, *best, w, 2, opt)
    assert best2 == best and not os.path.exists(opt.ckpt_path)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_standalone_evaluator_vs_oracle(dtype, capsys):
    """combat_b200.eval (reference eval.py:83-152): get_model / eval on three batches against the oracle's gathered-sub-batch
    restatement; all2one and all2all targets."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from combat_b200 import config
    from combat_b200 import eval as ev
    for attack in ("all2one", "all2all"):
        opt = config.get_arguments().parse_args(["--device", "cuda", "--dtype", dtype, "--attack_mode", attack])
        opt.input_height = opt.input_width = 32
        opt.input_channel = 3
        torch.manual_seed(8)
        netC, netG = ev.get_model(opt)
        sdC = {k: v.detach().cpu().clone() for k, v in netC.state_dict().items()}
        sdG = {k: v.detach().cpu().clone() for k, v in netG.state_dict().items()}
        pC, bC = O.split_state(sdC)
        oopt = O.default_opt(attack_mode=attack)
        g = torch.Generator().manual_seed(2)
        batches = [(torch.rand(40, 3, 32, 32, generator=g) * 2 - 1, torch.randint(0, 10, (40,), generator=g)) for _ in range(3)]
        torch.manual_seed(50)
        refs = [O.victim_eval_batch(pC, bC, sdG, x, y, oopt) for x, y in batches]
        torch.manual_seed(50)

        class W:
            def add_scalars(self, tag, d, step):
                self.last = (tag, d)

        w = W()
        acc = ev.eval(netC, netG, batches, w, opt)
        n, nb = sum(r["n_clean"] for r in refs), sum(r["n_bd"] for r in refs)
        want = (sum(r["clean_correct"] for r in refs) * 100.0 / n, sum(r["bd_ba"] for r in refs) * 100.0 / nb,
                sum(r["bd_asr"] for r in refs) * 100.0 / nb)
        slack = (0 if dtype == "fp32" else 3)   # bf16 logits of a random-init net flip a few near-tied argmaxes
        for got, ref, q in zip(acc, want, (100.0 / n, 100.0 / nb, 100.0 / nb)):
            assert abs(got - ref) <= slack * q + 1e-9, (attack, acc, want)
        assert w.last[0] == "Test Accuracy" and set(w.last[1]) == {"Clean", "Bd BA", "Bd ASR"}
        # tensors of one batch, same sigma
        torch.manual_seed(51)
        r = O.victim_eval_batch(pC, bC, sdG, *batches[0], oopt)
        torch.manual_seed(51)
        _, nbd, d = ev.eval_batch(netC, netG, *batches[0], opt)
        assert nbd == r["n_bd"] and d["sigma"] == r["sigma"]
        nt = r["ntrg"].cuda()
        tol = 2e-4 if dtype == "fp32" else 6e-2
        assert rel(d["preds_clean"], r["preds_clean"]) < tol and rel(d["preds_bd"][nt], r["preds_bd"]) < tol
        assert rel(d["x_bd"][nt], r["x_bd"]) < (1e-5 if dtype == "fp32" else 3e-2)
    assert "Clean Acc" in capsys.readouterr().out
