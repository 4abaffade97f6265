"""'PyTorch eager on the same B200' comparator: the reference's algorithm as restated by
oracle/combat_oracle.py, run with every tensor on cuda:0 -- i.e. cuDNN convolutions, cuFFT DCTs, aten elementwise kernels,
autograd, fp32 (optionally TF32), one alternated step of batch 512 at CIFAR-10 shape.  It is neither the product path nor
the CPU baseline of bench.py; it answers "what would the reference's own PyTorch code reach on this GPU".
NOT YET RUN ON A GPU (written after the round's GPU budget was spent); one JSON line.
Usage: python scripts/bench_eager_torch.py [--batch 512] [--steps 10] [--tf32]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import combat_oracle as O  # noqa: E402  (scripts/ are measurement tooling, not product code)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tf32", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = a.tf32
    gen = torch.Generator().manual_seed(0)
    netC_p, netC_b = O.init_preact_resnet18_state(gen)
    clean_p, clean_b = O.init_preact_resnet18_state(gen)
    netG_p = O.init_unet_state(gen)
    netF_p, netF_b = O.init_frequency_model_state(gen)
    to = lambda d: {k: v.to(dev) for k, v in d.items()}
    state = dict(netC_p=to(netC_p), netC_b=to(netC_b), clean_p=to(clean_p), clean_b=to(clean_b), netG_p=to(netG_p), netF_p=to(netF_p),
                 netF_b=to(netF_b), momC={}, momG={})
    opt = O.default_opt()
    np.random.seed(0)
    torch.manual_seed(0)
    xs = [(torch.rand(a.batch, 3, 32, 32, generator=gen) * 2 - 1).to(dev) for _ in range(2)]
    ys = [torch.randint(0, 10, (a.batch,), generator=gen).to(dev) for _ in range(2)]
    times = []
    for i in range(a.warmup + a.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = O.alternated_step(state, xs[i % 2], ys[i % 2], opt)
        torch.cuda.synchronize()
        if i >= a.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    print(json.dumps({"impl": "torch-eager (oracle restatement on cuda:0, cuDNN/cuFFT)", "metric": "alternated-step images/sec at CIFAR-10 shape",
                      "value": a.batch / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "batch": a.batch, "steps": a.steps,
                      "dtype": "tf32" if a.tf32 else "f32", "loss_c": out["loss_c"], "loss_g": out["loss_g"]}))


if __name__ == "__main__":
    main()
