"""HBM-bound kernels of the path at BASELINE configs[4] sizes: batched 2-D DCT / IDCT / fused low_freq over 65,536
CIFAR-shape images (fp32 and uint8 input), DCT at 64x64 (CelebA), the fused poison-blend forward/backward, the
flat nesterov-SGD update.  Reports algorithmic GB/s (SURVEY.md 8d byte counts) against MEASURED_PEAKS.json hbm_gbs.
Correctness of the full-size transforms is checked through size-independent properties (round trip, Parseval, idempotence
of the low-pass projection).  One JSON line per kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from combat_b200 import ops  # noqa: E402

dev = torch.device("cuda")
peak = 6538.3
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk)).get("hbm_gbs", peak)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def report(name, nbytes, sec, extra=None):
    gbs = nbytes / sec / 1e9
    line = {"kernel": name, "algorithmic_MB": nbytes / 1e6, "us": sec * 1e6, "GB/s": gbs, "frac_of_measured_hbm": gbs / peak,
            "peak_GB/s": peak}
    if extra:
        line.update(extra)
    print(json.dumps(line))


NI = 65536
x = torch.rand(NI, 3, 32, 32, device=dev) * 2 - 1                  # 805 MB: far larger than the 126 MB L2
out = torch.empty_like(x)
nb = 2 * x.numel() * 4
for kind in ("dct", "idct", "lowfreq"):
    t = timed(lambda: ops.plane_op(x, kind, keep=20, out=out))
    report("dct32 %s fp32 65536x3x32x32" % kind, nb, t, {"images_per_s": NI / t})
xu = (torch.rand(NI, 3, 32, 32, device=dev) * 255).to(torch.uint8)
t = timed(lambda: ops.plane_op(xu, "dct", in_mode=1, out=out))
report("dct32 dct uint8-in 65536x3x32x32", x.numel() * 5, t, {"images_per_s": NI / t})
# properties at full size
X = ops.plane_op(x, "dct")
back = ops.plane_op(X, "idct")
rt = float((back - x).abs().max())
pars = float(((X.double() ** 2).sum() - (x.double() ** 2).sum()).abs() / (x.double() ** 2).sum())
lf = ops.plane_op(x, "lowfreq", keep=20)
idem = float((ops.plane_op(lf, "lowfreq", keep=20) - lf).abs().max())
print(json.dumps({"properties_65536": {"idct(dct(x)) max abs err": rt, "Parseval rel err": pars, "lowfreq idempotence max abs err": idem}}))
assert rt < 5e-6 and pars < 1e-6 and idem < 5e-6
del X, back, lf, xu
# CelebA plane size
x64 = torch.rand(16384, 3, 64, 64, device=dev) * 2 - 1
o64 = torch.empty_like(x64)
for kind in ("dct", "lowfreq"):
    t = timed(lambda: ops.plane_op(x64, kind, keep=41, out=o64))
    report("plane_transform %s fp32 16384x3x64x64" % kind, 2 * x64.numel() * 4, t, {"images_per_s": 16384 / t})
del x64, o64
# fused poison blend (G-step form: every row poisoned) and its backward
B = 65536
noise = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
taps = ops.gaussian_taps(0.6)
sq = torch.empty(B * 3, device=dev)
t = timed(lambda: ops.poison_blend_fwd(x, noise, None, B, 0.08, taps, out=out, sq_partial=sq))
report("poison_blend_fwd (blend+clamp+blur+MSE partials) 65536 rows", 3 * x.numel() * 4, t, {"images_per_s": B / t})
g1 = torch.randn(B, 3, 32, 32, device=dev)
dn = torch.empty_like(x)
t = timed(lambda: ops.poison_blend_bwd(x, noise, out, g1, None, 1e-6, 0.08, taps, out=dn))
report("poison_blend_bwd 65536 rows", 5 * x.numel() * 4, t, {"images_per_s": B / t})
del noise, g1, dn, out, x
# nesterov SGD over both networks' parameters (20.5 M floats, 20 B/param)
n = 20541389
p_, g_, m_ = (torch.randn(n, device=dev) for _ in range(3))
lr = torch.full((1,), 1e-2, device=dev)
t = timed(lambda: ops.sgd_nesterov(p_, g_, m_, lr, 0.9, 5e-4, False))
report("sgd_nesterov 20.5M params", 20 * n, t)
