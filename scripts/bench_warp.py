"""WaNet warp kernels (csrc/warp.cu) timed alone at BASELINE configs[4]-like sizes: 65,536 CIFAR-shape and 16,384 CelebA-shape images."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from combat_b200 import ops  # noqa: E402

dev = torch.device("cuda")
for B, H in ((65536, 32), (16384, 64)):
    x = torch.rand(B, 3, H, H, device=dev) * 2 - 1
    o = torch.empty_like(x)
    sq = torch.empty(B * 3, device=dev)
    g1 = torch.randn(B, 3, H, H, device=dev)
    flow = torch.tanh(torch.randn(B, 2, 2, 2, device=dev))
    ident = torch.linspace(-1, 1, steps=H).to(dev)
    t = bench._timed_kernel(lambda: ops.wanet_warp_fwd(x, flow, ident, None, B, 0.15, 2, out=o, sq_partial=sq))
    print("H %d fwd us %.1f GB/s %.0f" % (H, t * 1e6, 2 * x.numel() * 4 / t / 1e9))
    t = bench._timed_kernel(lambda: ops.wanet_warp_bwd(x, flow, ident, g1, None, 0.15, 1e-6, 2))
    print("H %d bwd us %.1f GB/s %.0f" % (H, t * 1e6, 2 * x.numel() * 4 / t / 1e9))
    del x, o, g1
