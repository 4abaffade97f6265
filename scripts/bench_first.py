"""Micro-benchmark / pipeline counters of the tensor-pipe first conv (conv_tc_first_kernel) against the CUDA-core conv_cin3 kernel.
  [COMBAT_TC_DBG=1] python scripts/bench_first.py [--batch 1024]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from combat_b200 import ops  # noqa: E402
from combat_b200._lib import check, lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda")
N, H = args.batch, 32
NBUF = 4
xs = [torch.randn(N, 3, H, H, device=dev) for _ in range(NBUF)]
w = (torch.randn(64, 3, 3, 3, device=dev) * 0.2).bfloat16()
w64 = torch.zeros(64, 64, dtype=torch.bfloat16, device=dev)
w64[:, :27] = w.permute(0, 2, 3, 1).reshape(64, 27)
w64[:, 32:59] = w64[:, :27]
wf = w.permute(0, 2, 3, 1).contiguous()
o1 = [torch.empty(N, H, H, 64, device=dev, dtype=torch.bfloat16) for _ in range(NBUF)]
o2 = [torch.empty(N, H, H, 64, device=dev, dtype=torch.bfloat16) for _ in range(NBUF)]
sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)


def timed(fn, name, bytes_):
    for i in range(NBUF):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(args.iters):
        fn(it % NBUF)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / args.iters
    print("%-34s %8.1f us  %6.0f GB/s of output+input" % (name, us, bytes_ / us / 1e3))
    return us


io2 = N * H * H * (3 * 4 + 2 * 64 * 2)
io1 = N * H * H * (3 * 4 + 64 * 2)
dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev) if os.environ.get("COMBAT_TC_DBG") else None


def tc(i, two):
    d = ops.conv_tc_desc(xs[i], w64.data_ptr(), o1[i], N, H, H, 64, H, H, 64, 3, 3, 1, 1, 1, out2=o2[i] if two else None,
                         scale2=sc if two else None, shift2=sh if two else None, in_nchw3=True)
    if dbg is not None:
        d.stats = dbg.data_ptr()
    check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc first")


for two in (True, False):
    if dbg is not None:
        dbg.zero_()
    us = timed(lambda i: tc(i, two), "tensor pipe, out%s" % (" + out2" if two else ""), io2 if two else io1)
    if dbg is not None:
        dv = dbg.view(148, 8).double()
        m = dv.mean(0) / (args.iters + NBUF)
        tiles = N * H * H / 128 / 148
        print("   per tile (cycles): build wait-free %.0f gather %.0f store+publish %.0f | epi prepare %.0f wait-acc %.0f body %.0f | whole kernel %.0f cycles "
              "= %.0f per tile, %.2f us (SM clock %.0f MHz)"
              % (m[0] / tiles, m[1] / tiles, m[2] / tiles, m[3] / tiles, m[4] / tiles, m[5] / tiles, m[6], m[6] / tiles, m[7] / 1e3,
                 m[6] / max(float(m[7]), 1.0) * 1e3))
    timed(lambda i: ops.conv_cin3(xs[i], wf.data_ptr(), ops.BF16, o1[i], 64, 1, out2=o2[i] if two else None, scale2=sc if two else None,
                                  shift2=sh if two else None), "CUDA cores (conv_cin3), out%s" % (" + out2" if two else ""),
          io2 if two else io1)
