"""Batched 2-D DCT / IDCT / low_freq kernels (csrc/dct32.cu) at BASELINE configs[4] sizes: 65,536 CIFAR-shape images
(fp32 and uint8 input) and 16,384 CelebA-shape images.  Algorithmic GB/s (one read + one write of every plane) against
MEASURED_PEAKS.json hbm_gbs; full-size correctness through size-independent properties.  One JSON line per kernel."""
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


def report(name, nbytes, sec, images):
    gbs = nbytes / sec / 1e9
    print(json.dumps({"kernel": name, "algorithmic_MB": nbytes / 1e6, "us": round(sec * 1e6, 1), "GB/s": round(gbs, 1),
                      "frac_of_measured_hbm": round(gbs / peak, 4), "peak_GB/s": peak, "images_per_s": images / sec}), flush=True)


def props(x, keep, name):
    X = ops.plane_op(x, "dct")
    rt = float((ops.plane_op(X, "idct") - x).abs().max())
    e = (x.double() ** 2).sum()
    pars = float(((X.double() ** 2).sum() - e).abs() / e)
    lf = ops.plane_op(x, "lowfreq", keep=keep)
    idem = float((ops.plane_op(lf, "lowfreq", keep=keep) - lf).abs().max())
    # the low-pass projection is the masked transform: idct(mask * dct(x))
    X[..., keep:, :] = 0
    X[..., :, keep:] = 0
    proj = float((ops.plane_op(X, "idct") - lf).abs().max())
    print(json.dumps({name: {"idct(dct(x)) max abs err": rt, "Parseval rel err": pars, "lowfreq idempotence max abs err": idem,
                             "lowfreq vs idct(mask*dct) max abs err": proj}}), flush=True)
    assert rt < 5e-6 and pars < 1e-6 and idem < 5e-6 and proj < 5e-6


NI = 65536
x = torch.rand(NI, 3, 32, 32, device=dev) * 2 - 1                  # 805 MB: far larger than the 126 MB L2
out = torch.empty_like(x)
xu = (torch.rand(NI, 3, 32, 32, device=dev) * 255).to(torch.uint8)
for kind in ("dct", "idct", "lowfreq"):
    t = timed(lambda: ops.plane_op(x, kind, keep=20, out=out))
    report("dct32 %s fp32 65536x3x32x32" % kind, 2 * x.numel() * 4, t, NI)
t = timed(lambda: ops.plane_op(xu, "dct", in_mode=1, out=out))
report("dct32 dct uint8-in 65536x3x32x32", x.numel() * 5, t, NI)
del xu, out
props(x, 20, "properties_65536x3x32x32")
del x
NI = 16384
x64 = torch.rand(NI, 3, 64, 64, device=dev) * 2 - 1
o64 = torch.empty_like(x64)
xu = (torch.rand(NI, 3, 64, 64, device=dev) * 255).to(torch.uint8)
for kind in ("dct", "idct", "lowfreq"):
    t = timed(lambda: ops.plane_op(x64, kind, keep=41, out=o64))
    report("dct64 %s fp32 16384x3x64x64" % kind, 2 * x64.numel() * 4, t, NI)
t = timed(lambda: ops.plane_op(xu, "dct", in_mode=1, out=o64))
report("dct64 dct uint8-in 16384x3x64x64", x64.numel() * 5, t, NI)
del xu, o64
props(x64, 41, "properties_16384x3x64x64")
