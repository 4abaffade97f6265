"""Micro-benchmark of the 3-channel boundary convolutions at B=512 CIFAR shape (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from combat_b200 import ops  # noqa: E402

dev = torch.device("cuda")
N, H = 512, 32
BF = ops.dt_code(torch.bfloat16)
x = torch.rand(N, 3, H, H, device=dev) * 2 - 1
w_ci = (torch.randn(64, 3, 3, 3, device=dev) * 0.2).bfloat16()      # [co][kh][kw][ci]
w_co = (torch.randn(3, 3, 3, 64, device=dev) * 0.05).bfloat16()     # [co][kh][kw][ci]
o32 = torch.empty(N, H, H, 64, device=dev)
o16 = torch.empty(N, H, H, 64, device=dev, dtype=torch.bfloat16)
a16 = torch.randn(N, H, H, 64, device=dev).bfloat16()
dz = torch.randn(N, 3, H, H, device=dev)
dx = torch.empty(N, 3, H, H, device=dev)
sc, sh = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
dw1, db1 = torch.zeros(64, 3, 3, 3, device=dev), torch.zeros(64, device=dev)
dw2, db2 = torch.zeros(3, 3, 3, 64, device=dev), torch.zeros(3, device=dev)
cases = {
    "conv_cin3 f32 out": lambda: ops.conv_cin3(x, w_ci.data_ptr(), BF, o32, 64, 1),
    "conv_cin3 f32+bn out": lambda: ops.conv_cin3(x, w_ci.data_ptr(), BF, o32, 64, 1, out2=o16, scale2=sc, shift2=sh),
    "conv_cin3 bf16 out": lambda: ops.conv_cin3(x, w_ci.data_ptr(), BF, o16, 64, 1),
    "conv_cout3 (bf16 in)": lambda: ops.conv_cout3(a16, w_co.data_ptr(), BF, dx),
    "wgrad_cin3 (bf16 dy)": lambda: ops.wgrad_cin3(x, a16, dw1, db1, 64, 1),
    "wgrad_cout3 (bf16 a)": lambda: ops.wgrad_cout3(a16, dz, dw2, db2),
}
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, fn in cases.items():
    if only and only not in name:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print("%-24s %8.1f us   (%.1f GFMA/s of 36400 peak)" % (name, us, 906e6 / us / 1e3))
