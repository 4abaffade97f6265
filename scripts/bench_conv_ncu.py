"""Three launches per pass for `ncu --set full` (scripts/gpu_r02_ncu_conv.sh): the 64->64 @32x32 row-reuse kernel, the single-CTA
128->128 @16x16 kernel and the CTA-pair 256->256 @8x8 kernel, each in its 'fwd bn only' form at batch 512 (the shapes and
epilogue the step runs most).  Two warm-up passes, then one measured pass (ncu: -s 6 -c 3)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from combat_b200 import ops  # noqa: E402
from combat_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda")
N = 512
cases = []
for Ci, Co, H in ((64, 64, 32), (128, 128, 16), (256, 256, 8)):
    x = torch.randn(N, H, H, Ci, device=dev).bfloat16()
    w = (torch.randn(Co, 3, 3, Ci, device=dev) * 0.05).bfloat16()
    o16 = torch.empty(N, H, H, Co, device=dev, dtype=torch.bfloat16)
    sc, sh = torch.rand(Co, device=dev) + 0.5, torch.randn(Co, device=dev)
    cases.append((x, w, o16, sc, sh, ops.conv_tc_desc(x, w.data_ptr(), None, N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, out2=o16, scale2=sc, shift2=sh)))
for _ in range(3):
    for c in cases:
        check(lib.combat_conv_tc(C.byref(c[-1]), ops._s()), "conv_tc")
    torch.cuda.synchronize()
print("ok")
