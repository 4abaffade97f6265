"""Micro-benchmark of the tcgen05 convolution entry points at the step's layer shapes (CUDA events, L2 flushed by
rotating over several distinct input/output buffers larger than L2 in total).
  python scripts/bench_conv.py [--only 64x32] [--iters 20]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from combat_b200 import ops  # noqa: E402
from combat_b200._lib import check, lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--variant", default="")
ap.add_argument("--batch", type=int, default=512)
args = ap.parse_args()
dev = torch.device("cuda")
N = args.batch
SHAPES = [("64x32", 64, 64, 32), ("128x16", 128, 128, 16), ("256x8", 256, 256, 8), ("512x4", 512, 512, 4)]
NBUF = 4
for name, Ci, Co, H in SHAPES:
    if args.only and args.only != name:
        continue
    xs = [torch.randn(N, H, H, Ci, device=dev).bfloat16() for _ in range(NBUF)]
    w = (torch.randn(Co, 3, 3, Ci, device=dev) * 0.05).bfloat16()
    res = [torch.randn(N, H, H, Co, device=dev) for _ in range(NBUF)]
    o32 = [torch.empty(N, H, H, Co, device=dev) for _ in range(NBUF)]
    o16 = [torch.empty(N, H, H, Co, device=dev, dtype=torch.bfloat16) for _ in range(NBUF)]
    m16 = [torch.randn(N, H, H, Co, device=dev).bfloat16() for _ in range(NBUF)]
    pad_ = torch.empty(1234 * 1024 + 512, device=dev, dtype=torch.uint8)  # break the power-of-two spacing of the buffers
    a16 = [torch.randn(N, H, H, Co, device=dev).bfloat16() for _ in range(NBUF)]
    sc, sh = torch.rand(Co, device=dev) + 0.5, torch.randn(Co, device=dev)
    flops = 2.0 * N * H * H * Co * Ci * 9
    variants = {
        "plain f32 out": lambda i: ops.conv_tc_desc(xs[i], w.data_ptr(), o32[i], N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1),
        "plain bf16 out": lambda i: ops.conv_tc_desc(xs[i], w.data_ptr(), o16[i], N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1),
        "fwd res+out+bn": lambda i: ops.conv_tc_desc(xs[i], w.data_ptr(), o32[i], N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, residual=res[i],
                                                     out2=o16[i], scale2=sc, shift2=sh),
        "fwd bn only": lambda i: ops.conv_tc_desc(xs[i], w.data_ptr(), None, N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, out2=o16[i],
                                                  scale2=sc, shift2=sh),
        "bwd mask+add": lambda i: ops.conv_tc_desc(xs[i], w.data_ptr(), o16[i], N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, mask=m16[i],
                                                   mask_scale=sc, post_add=m16[(i + 1) % NBUF]),
        "bwd mask only": lambda i: ops.conv_tc_desc(xs[i], w.data_ptr(), o16[i], N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, mask=m16[i],
                                                    mask_scale=sc),
        "bwd mask+pre": lambda i: ops.conv_tc_desc(xs[i], w.data_ptr(), o16[i], N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, mask=m16[i],
                                                   mask_scale=sc, residual=a16[i]),
        "bwd mask+add2": lambda i: ops.conv_tc_desc(xs[i], w.data_ptr(), o16[i], N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, mask=m16[i],
                                                    mask_scale=sc, post_add=a16[i]),
        "fwd bf16 res": lambda i: ops.conv_tc_desc(xs[i], w.data_ptr(), o16[i], N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, residual=a16[i]),
    }
    for vn, mk in variants.items():
        if args.variant and args.variant != vn:
            continue
        descs = [mk(i) for i in range(NBUF)]
        dbg = None
        if os.environ.get("COMBAT_TC_DBG"):
            dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
            for d in descs:
                d.stats = dbg.data_ptr()
        for d in descs:
            check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(args.iters):
            check(lib.combat_conv_tc(C.byref(descs[it % NBUF]), ops._s()), "conv_tc")
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        print("%-8s %-16s %8.1f us  %7.1f TFLOP/s" % (name, vn, us, flops / us / 1e6))
        if dbg is not None:
            torch.cuda.synchronize()
            dv = dbg.view(148, 8).double()
            used = dv[:, 4] + dv[:, 5] > 0
            m = dv[used].mean(0) / (args.iters + NBUF)
            tiles = N * H * H / 128 * max(1, Co // 256 if Co % 256 == 0 else Co // 128 if Co % 128 == 0 else Co // 64) / max(1, int(used.sum()))
            print("   %d CTAs, %.2f tiles per CTA; per tile (cycles): producer wait-empty %.0f | mma wait-tempty %.0f wait-full %.0f "
                  "[mma loop total/launch %.0f] | epi wait-acc %.0f body %.0f"
                  % ((int(used.sum()), tiles) + tuple(float(v) / tiles for v in m[:3]) + (float(dv[used][:, 3].mean()),)
                     + tuple(float(v) / tiles for v in m[4:6])))
            if float(m[7]) > 0:
                print("   whole kernel per CTA: %.0f cycles in %.2f us (SM clock %.0f MHz); event time per launch %.2f us"
                      % (float(m[6]), float(m[7]) / 1e3, float(m[6]) / float(m[7]) * 1e3, us))
    if args.variant:
        continue
    # wgrad
    dw = torch.zeros(Co, 3, 3, Ci, device=dev)
    d3 = [ops.conv_tc_desc(xs[i], None, None, N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1) for i in range(NBUF)]
    for i in range(NBUF):
        check(lib.combat_conv_tc_wgrad(C.byref(d3[i]), m16[i].data_ptr(), dw.data_ptr(), ops._s()), "wgrad")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(args.iters):
        check(lib.combat_conv_tc_wgrad(C.byref(d3[it % NBUF]), m16[it % NBUF].data_ptr(), dw.data_ptr(), ops._s()), "wgrad")
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / args.iters
    print("%-8s %-16s %8.1f us  %7.1f TFLOP/s" % (name, "wgrad", us, flops / us / 1e6))
