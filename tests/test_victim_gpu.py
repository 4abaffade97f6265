"""GPU parity of the victim trainer (train_victim.py:94-226) and the clean-classifier trainer (train_clean_classifier.py:75-120):
the engine's victim_step against the oracle restatement, and the public train() / eval() / main() surface."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import combat_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _seed(s):
    torch.manual_seed(s); np.random.seed(s); random.seed(s)


@pytest.mark.parametrize("mode", ["victim", "victim_use", "victim_no_poison", "clean_classifier"])
def test_victim_step_vs_oracle(mode):
    from test_step_gpu import seeded_state
    from combat_b200.engine import AlternatedStep, default_opt, make_plan_victim
    tf = "use" if mode == "victim_use" else "no_use"
    state = seeded_state(12)
    B = 40
    g = torch.Generator().manual_seed(3)
    with_g = mode != "clean_classifier"
    eng = AlternatedStep(default_opt(post_transform_option=tf), device="cuda", dtype=torch.float32, with_metrics=False)
    eng.load_state(netC={**state["netC_p"], **state["netC_b"]}, netG=state["netG_p"])
    if not with_g:
        eng.netG = None
    oopt = O.default_opt(post_transform_option=tf)
    momC = {}
    for it in range(2):
        x = torch.rand(B, 3, 32, 32, generator=g) * 2 - 1
        y = torch.randint(0, 10, (B,), generator=g)
        if mode == "clean_classifier":
            pz = None
        elif mode == "victim_no_poison":
            pz = np.zeros(B, dtype=bool)
        else:
            pz = (y.numpy() == 0) & (torch.rand(B, generator=g).numpy() < 0.7)
            pz[int(np.nonzero(y.numpy() == 0)[0][0]) if (y == 0).any() else 0] = True
        _seed(20 + it)
        r = O.victim_train_step(state["netC_p"], state["netC_b"], state["netG_p"] if with_g else None, momC, x, y, pz, oopt)
        _seed(20 + it)
        plan = make_plan_victim(y.numpy(), pz, eng.opt)
        assert plan.num_bd == r["num_bd"] and np.array_equal(plan.total_targets, r["total_y"].numpy())
        if r["num_bd"]:
            assert plan.sigma_c == r["sigma"]
        out = eng.victim_step(x.cuda(), y.numpy(), pz, plan, keep_debug=True)
        torch.cuda.synchronize()
        d = out["debug"]
        assert torch.equal(d["total_x"][plan.num_bd:].cpu(), r["total_x"][plan.num_bd:])       # pass-through rows: bit-exact
        tol = 1e-4 if it == 0 else 3e-3
        assert rel(d["total_x"], r["total_x"]) < tol and rel(d["logits_c"], r["logits"]) < tol
        assert abs(float(out["losses"][0]) - r["loss"]) < (2e-5 if it == 0 else 1e-3) * max(1.0, abs(r["loss"]))
        if it == 0:
            assert int(out["counts"][0]) == r["n_correct"]
    sd = eng.netC.state_dict()
    num = sum(float(((sd[n].cpu() - state["netC_p"][n]) ** 2).sum()) for n in state["netC_p"])
    den = sum(float((state["netC_p"][n] ** 2).sum()) for n in state["netC_p"])
    assert (num / den) ** 0.5 < 1e-3


def test_victim_and_clean_trainers_public_api(tmp_path, capsys):
    from combat_b200 import train_clean_classifier as tc
    from combat_b200 import train_victim as tv
    args = ["--synthetic_data", "--debug", "--bs", "32", "--n_iters", "2", "--log_every", "4", "--saving_prefix", "v",
            "--checkpoints", str(tmp_path), "--load_checkpoint", "none"]
    _seed(0)
    best = tv.main(args)
    out = capsys.readouterr().out
    assert "CE Loss" in out and "Bd Acc" in out and " Saving..." in out and len(best) == 2
    tv.main(args + ["--continue_training", "--n_iters", "3"])
    assert "Continue training!!" in capsys.readouterr().out
    # clean classifier: (inputs, targets) batches, graph replay, momentum exposed through the torch optimiser
    from combat_b200 import config
    opt = config.get_arguments().parse_args(["--device", "cuda", "--log_every", "2"])
    opt.input_height = opt.input_width = 32
    opt.input_channel = 3
    _seed(1)
    netC, optC, schC = tc.get_model(opt)
    g = torch.Generator().manual_seed(0)
    data = [(torch.rand(32, 3, 32, 32, generator=g) * 2 - 1, torch.randint(0, 10, (32,), generator=g)) for _ in range(4)]

    class W:
        def __init__(self):
            self.s = []

        def add_scalars(self, tag, d, e):
            self.s.append(tag)

        def add_scalar(self, tag, v, e):
            self.s.append(tag)

    w = W()
    tc.train(netC, optC, schC, data, w, 0, opt)
    torch.cuda.synchronize()
    assert w.s == ["Clean Accuracy", "CE Loss"] and schC.last_epoch == 1
    p0 = next(iter(netC.parameters()))
    assert float(optC.state[p0]["momentum_buffer"].abs().sum()) > 0
    assert int(netC.state_dict()["layer1.0.bn1.num_batches_tracked"]) == 4


def test_inputaware_victim_eval_vs_oracle_and_public_main(tmp_path, capsys):
    """train_victim_inputaware.py: eval_batch (clean / attack / cross-trigger accuracy, two sigma draws) against the oracle
    restatement of :187-223, then main() on synthetic data (three loaders, seven-key checkpoint dict)."""
    from combat_b200 import config
    from combat_b200 import train_victim_inputaware as tvi
    opt = config.get_arguments().parse_args(["--device", "cuda", "--dtype", "fp32"])
    opt.input_height = opt.input_width = 32
    opt.input_channel = 3
    _seed(4)
    netC, optC, schC, netG = tvi.get_model(opt)
    sdC = {k: v.detach().cpu().clone() for k, v in netC.state_dict().items()}
    sdG = {k: v.detach().cpu().clone() for k, v in netG.state_dict().items()}
    netC_p, netC_b = O.split_state(sdC)
    o = O.default_opt()
    g = torch.Generator().manual_seed(9)
    for it in range(2):
        x = torch.rand(24, 3, 32, 32, generator=g) * 2 - 1
        x2 = torch.rand(24, 3, 32, 32, generator=g) * 2 - 1
        y = torch.randint(0, 10, (24,), generator=g)
        torch.manual_seed(70 + it)
        r = O.victim_eval_batch(netC_p, netC_b, sdG, x, y, o, x2=x2)
        torch.manual_seed(70 + it)
        counts, nb, d = tvi.eval_batch(netC, netG, x, x2, y, tvi._variant(opt))
        c = counts.cpu().numpy()
        assert d["sigma"] == r["sigma"] and d["sigma2"] == r["sigma2"] and nb == r["n_bd"]
        assert int(c[0]) == r["clean_correct"] and int(c[2]) == r["bd_asr"] and int(c[4]) == r["cross_correct"]
        nt = r["ntrg"].cuda()
        assert rel(d["x_bd"][nt], r["x_bd"]) < 1e-5 and rel(d["x_bd2"], r["x_bd2"]) < 1e-5
        assert rel(d["preds_cross"], r["preds_cross"]) < 2e-4 and rel(d["preds_bd"][nt], r["preds_bd"]) < 2e-4
    args = ["--synthetic_data", "--debug", "--bs", "32", "--n_iters", "1", "--log_every", "4", "--saving_prefix", "vi",
            "--checkpoints", str(tmp_path), "--load_checkpoint", "none"]
    _seed(0)
    best = tvi.main(args)
    out = capsys.readouterr().out
    assert "Cross Acc" in out and len(best) == 3


def test_clean_classifier_main_writes_the_checkpoint_the_generator_trainer_loads(tmp_path, capsys):
    """train_clean_classifier.main() (reference :163-236): train + eval on synthetic data, checkpoint at
    <checkpoints>/<prefix>/<dataset>/<dataset>_<prefix>.pth.tar -- the path train_generator.main() resolves from
    --load_checkpoint_clean (:513-527) -- then the generator trainer's main() starts from it."""
    import os
    from combat_b200 import train_clean_classifier as tc
    from combat_b200 import train_generator as tg
    _seed(0)
    best = tc.main(["--synthetic_data", "--debug", "--bs", "32", "--n_iters", "1", "--log_every", "4", "--saving_prefix", "cc",
                    "--checkpoints", str(tmp_path)])
    out = capsys.readouterr().out
    path = tmp_path / "cc" / "cifar10" / "cifar10_cc.pth.tar"
    assert "Clean Acc" in out and " Saving..." in out and os.path.exists(path) and best >= 0.0
    ck = torch.load(str(path), map_location="cpu", weights_only=False)
    assert {"netC", "optimizerC", "schedulerC", "best_clean_acc", "epoch_current"} == set(ck)
    bests = tg.main(["--synthetic_data", "--debug", "--bs", "32", "--n_iters", "1", "--log_every", "4", "--saving_prefix", "g",
                     "--checkpoints", str(tmp_path), "--load_checkpoint_clean", "cc", "--post_transform_option", "no_use"])
    assert len(bests) == 6 and "Clean Model Acc" in capsys.readouterr().out
