"""Layer-by-layer comparison of the bf16 CUDA generator / classifier with the quantisation-aware oracle (measurement tooling):
for every conv layer the relative L2 distance of its input and of its float32 output, so that a rounding point that the
oracle places differently from combat_b200/nets.py shows up as a jump at that layer instead of as a diffuse end-to-end error."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import combat_oracle as O  # noqa: E402
from combat_b200 import nets  # noqa: E402


def rel2(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nchw(t):
    return t.permute(0, 3, 1, 2)


torch.manual_seed(0)
gen = torch.default_generator
netG_p = O.init_unet_state(gen)
B = 64
x = torch.rand(B, 3, 32, 32) * 2 - 1
G = nets.Generator(device="cuda", dtype=torch.bfloat16)
G.load_state_dict({k: v.cuda() for k, v in netG_p.items()})
out, ctx = G.forward(x.cuda(), None, save=True)
taps = {}
with O.quantised(), torch.no_grad():
    ref = O.unet_forward(netG_p, x, taps=taps)
with torch.no_grad():
    ref32 = O.unet_forward(netG_p, x)
acts = ctx["acts"]
print("%-12s %12s %12s" % ("layer", "input", "conv out"))
print("%-12s %12s %12.3e" % ("conv0_0", "-", rel2(nchw(acts["c00"]), O._r(taps["conv0_0"][1]))))
print("%-12s %12.3e" % ("a00", rel2(nchw(acts["a00"]), taps["conv0_1"][0])))
for name, _ in nets.Generator.LAYERS[1:-1]:
    xin, c, st = acts[name]
    print("%-12s %12.3e %12.3e" % (name, rel2(nchw(xin), taps[name][0]), rel2(nchw(c), taps[name][1])))
print("%-12s %12.3e %12.3e" % ("upconv0_0", rel2(nchw(acts["a01"]), taps["upconv0_0"][0]), rel2(out, ref)))
print("quantised oracle vs float32 oracle (output): %.3e" % rel2(ref, ref32))
# the same layer fed with the CUDA path's OWN input: isolates each kernel from the accumulated difference
import torch.nn.functional as F
print("\nper-layer kernel error (torch conv of the CUDA layer's own input, bf16-rounded weights, float32 accumulate):")
for name, stride in nets.Generator.LAYERS[1:-1]:
    xin, c, st = acts[name]
    w = O._r(netG_p[name + ".weight"]).cuda()
    want = F.conv2d(nchw(xin).float(), w, netG_p[name + ".bias"].cuda(), stride=stride, padding=1)
    print("%-12s %12.3e" % (name, rel2(nchw(c), want)))
