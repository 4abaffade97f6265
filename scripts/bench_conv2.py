"""Micro-benchmark of the strided / 1x1 tcgen05 convolution launches of the step (forward and input-gradient forms)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from combat_b200 import ops  # noqa: E402
from combat_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda")
N = 512
NBUF = 3
# name, Ci (conv input ch), Co, H (conv input size), k, stride
LAYERS = [("l2.0.conv1", 64, 128, 32, 3, 2), ("l2.0.sc", 64, 128, 32, 1, 2), ("l3.0.conv1", 128, 256, 16, 3, 2),
          ("l3.0.sc", 128, 256, 16, 1, 2), ("l4.0.conv1", 256, 512, 8, 3, 2), ("l4.0.sc", 256, 512, 8, 1, 2)]


def timeit(descs, iters=20):
    for d in descs:
        check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(iters):
        check(lib.combat_conv_tc(C.byref(descs[it % NBUF]), ops._s()), "conv_tc")
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for name, Ci, Co, H, k, s in LAYERS:
    p = 1 if k == 3 else 0
    Ho = H // s
    xs = [torch.randn(N, H, H, Ci, device=dev).bfloat16() for _ in range(NBUF)]
    dys = [torch.randn(N, Ho, Ho, Co, device=dev).bfloat16() for _ in range(NBUF)]
    w = (torch.randn(Co, k, k, Ci, device=dev) * 0.05).bfloat16()
    wd = (torch.randn(Ci, k, k, Co, device=dev) * 0.05).bfloat16()
    o32 = [torch.empty(N, Ho, Ho, Co, device=dev) for _ in range(NBUF)]
    o16 = [torch.empty(N, Ho, Ho, Co, device=dev, dtype=torch.bfloat16) for _ in range(NBUF)]
    dx = [torch.empty(N, H, H, Ci, device=dev, dtype=torch.bfloat16) for _ in range(NBUF)]
    res = [torch.randn(N, H, H, Ci, device=dev).bfloat16() for _ in range(NBUF)]
    msk = [torch.randn(N, H, H, Ci, device=dev).bfloat16() for _ in range(NBUF)]
    sc, sh = torch.rand(Co, device=dev) + 0.5, torch.randn(Co, device=dev)
    msc = torch.rand(Ci, device=dev) + 0.5
    flops = 2.0 * N * Ho * Ho * Co * Ci * k * k
    variants = {
        "fwd f32 out": [ops.conv_tc_desc(xs[i], w.data_ptr(), o32[i], N, H, H, Ci, Ho, Ho, Co, k, k, s, p, 1) for i in range(NBUF)],
        "fwd bn only": [ops.conv_tc_desc(xs[i], w.data_ptr(), None, N, H, H, Ci, Ho, Ho, Co, k, k, s, p, 1, out2=o16[i], scale2=sc,
                                         shift2=sh) for i in range(NBUF)],
        "dgrad plain": [ops.conv_tc_desc(dys[i], wd.data_ptr(), dx[i], N, Ho, Ho, Co, H, H, Ci, k, k, 1, k - 1 - p, s) for i in range(NBUF)],
        "dgrad res+mask": [ops.conv_tc_desc(dys[i], wd.data_ptr(), dx[i], N, Ho, Ho, Co, H, H, Ci, k, k, 1, k - 1 - p, s, residual=res[i],
                                            mask=msk[i], mask_scale=msc) for i in range(NBUF)],
    }
    for vn, descs in variants.items():
        us = timeit(descs)
        io_mb = (N * H * H * Ci * 2 + N * Ho * Ho * Co * 2) / 1e6
        print("%-11s k%d s%d %4d->%-4d @%-2d %-15s %8.1f us %7.1f TFLOP/s  (min in+out %.0f MB = %.1f us at 6.5 TB/s)"
              % (name, k, s, Ci, Co, H, vn, us, flops / us / 1e6, io_mb, io_mb / 6.5e3 * 1e3 / 1e3))
