"""debug: the wanet fixture iterations through the engine, one at a time, against the fixture's recorded tensors"""
import random, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from combat_b200 import config
from combat_b200 import train_generator_wanet as tw
from combat_b200.train_generator import _engine_for
from combat_b200.engine import make_plan, AlternatedStep
g = np.load("tests/golden/step_wanet_b32x2.npz")
seed, B, nb = int(g["seed"]), int(g["B"]), int(g["n_batches"])
opt = config.get_arguments().parse_args(["--device", "cuda", "--post_transform_option", "no_use", "--dtype", "fp32", "--no_graph"])
opt.input_height = opt.input_width = 32; opt.input_channel = 3
torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
netC, optC, schC, netG, optG, schG, netF, clean = tw.get_model(opt)
batches = [(torch.rand(B, 3, 32, 32) * 2 - 1, torch.randint(0, 10, (B,))) for _ in range(nb)]
sd0 = {k: v.detach().clone().cpu() for k, v in netC.state_dict().items()}
sdG0 = {k: v.detach().clone().cpu() for k, v in netG.state_dict().items()}
opt = tw._variant(opt)
eng = _engine_for(netC, clean, netG, netF, opt)
eng.set_lr(optC.param_groups[0]["lr"], optG.param_groups[0]["lr"])
def rel(a, b):
    a = a.detach().float().cpu().double(); b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
vals = g["loss_values"]
for i, (x, y) in enumerate(batches):
    plan = make_plan(y.numpy(), opt, True)
    print("it", i, "num_bd", plan.num_bd, int(g["num_bd_%d" % i]), "trg", plan.trg_ind[:plan.num_bd], g["poison_idx_%d" % i])
    out = eng.step(x.cuda(), y.numpy(), plan, keep_debug=True)
    s = AlternatedStep.unpack(out); d = out["debug"]
    print("  losses", s["loss_c"], vals[6 * i], "|", s["loss_ce"], vals[6 * i + 1], "|", s["loss_l2"], vals[6 * i + 2], "|", s["clean_model_loss"], vals[6 * i + 5])
    for k, dk in (("logits_c", "logits_c"), ("pred_clean", "pred_clean"), ("pred_bd", "pred_bd"), ("clean_preds", "clean_preds"), ("clean_model_preds", "clean_model_preds"), ("flow", "noise_raw")):
        print("  ", k, rel(d[dk], g["%s_%d" % (k, i)]))
    print("   x_bd head", rel(d["x_bd"][:8], g["x_bd_head_%d" % i]), "x_bd_c", rel(d["total_x"][:plan.num_bd], g["x_bd_c_%d" % i]))
sd = netC.state_dict()
print("per-tensor two-iteration update norm, engine / fixture:")
for pre, sd, s0 in (("netC_", netC.state_dict(), sd0), ("netG_", netG.state_dict(), sdG0)):
    for n, v0 in s0.items():
        if torch.is_floating_point(v0) and (pre + "dnorm_" + n) in g.files:
            d = float((sd[n].detach().cpu() - v0).double().norm()); ref = float(g[pre + "dnorm_" + n][0])
            flag = "" if abs(d - ref) <= 3e-2 * ref else "   <-----"
            print("  %-40s %.6e %.6e ratio %.4f%s" % (pre + n, d, ref, d / max(ref, 1e-30), flag))
