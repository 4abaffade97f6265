"""One launch each of the 64x64 DCT, the 64x64 low_freq projection and the uint8-input 32x32 DCT at bench size, for
`ncu --set full -k regex:dct` (scripts/gpu_dct.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from combat_b200 import ops  # noqa: E402

x64 = torch.rand(16384, 3, 64, 64, device="cuda") * 2 - 1
o64 = torch.empty_like(x64)
ops.plane_op(x64, "dct", out=o64)
ops.plane_op(x64, "lowfreq", keep=41, out=o64)
xu = (torch.rand(65536, 3, 32, 32, device="cuda") * 255).to(torch.uint8)
ops.plane_op(xu, "dct", in_mode=1)
torch.cuda.synchronize()
