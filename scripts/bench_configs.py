"""Throughput of the alternated step at the OTHER shapes BASELINE.json names (the bench line itself is configs[1]):
  configs[0]  CIFAR-10 shape, batch 128 (the reference's CPU-runnable case), on the GPU path
  configs[3]  ImageNet-10 shape 3x224x224, ResNet18 (scaler-49 extension) + UnetGenerator, batch 64 and 256 per GPU
  configs[4]  CelebA 3x64x64 multilabel step (train_generator_multilabel.py:160-242), ResNet18(8) + CUnetGeneratorv1
Device-resident inputs, CUDA-graph replay, CUDA events, bf16 tcgen05 path, one GPU.  `--eager` runs one un-graphed iteration
per selected case instead (what `scripts/gpu_launchlist_configs.sh` puts under ncu).  One JSON line per case; a failing
case prints its exception instead of a number.  FLOPs per image from SURVEY.md section 8(d)."""
import json
import os
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from combat_b200 import config  # noqa: E402
from combat_b200 import train_generator as tg  # noqa: E402
from combat_b200 import train_generator_multilabel as tgm  # noqa: E402
from combat_b200.engine import AlternatedStep  # noqa: E402

dev = torch.device("cuda", 0)
peak_tf = 1359.8
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak_tf = json.load(open(pk)).get("bf16_tflops_sustained", peak_tf)

CASES = [
    # name, dataset, size, classes, batch, multilabel, update FLOPs per image (7 F_C + 3 F_G), learning rate
    ("configs[0] CIFAR-10 shape, batch 128", "cifar10", 32, 10, 128, False, 9.35e9, 1e-2),
    ("configs[4] CelebA 64x64 multilabel step, batch 256", "celeba", 64, 8, 256, True, 37.4e9, 1e-2),
    # 224x224: the scaler-49 head (25,088 pooled features into the linear layer) diverges within a few iterations at the
    # reference's lr 1e-2 on random data (loss_ce 1.8e7 after 8 steps, measured) - the reference itself never ran this
    # size (KeyError).  Throughput does not depend on lr; 1e-4 keeps the timed iterations numerically sane.
    ("configs[3] ImageNet-10 shape 224x224, batch 64", "imagenet10", 224, 10, 64, False, 458e9, 1e-4),
    ("configs[3] ImageNet-10 shape 224x224, batch 256", "imagenet10", 224, 10, 256, False, 458e9, 1e-4),
]
ARGS = [a for a in sys.argv[1:] if not a.startswith("--")]
EAGER = "--eager" in sys.argv       # one eager (no CUDA graph) iteration after one warm-up: for `ncu` launch lists
if ARGS:
    CASES = [c for c in CASES if any(a in c[0] for a in ARGS)]


def run(name, dataset, S, ncls, B, multilabel, flops, lr, steps=5, warmup=3):
    mod = tgm if multilabel else tg
    opt = config.get_arguments().parse_args(["--device", str(dev), "--post_transform_option", "no_use", "--dataset", dataset])
    opt.input_height = opt.input_width = S
    opt.input_channel = 3
    opt.num_classes = ncls
    opt.lr_C = opt.lr_G = lr
    torch.manual_seed(0)
    np.random.seed(0)
    netC, _, _, netG, _, _, netF, clean_model = mod.get_model(opt)
    eng = AlternatedStep(opt, device=dev, with_metrics=True, multilabel=multilabel,
                         nets=(netC.net, clean_model.net, netG.net, netF.net if netF is not None else None))
    g = torch.Generator().manual_seed(7)
    xs = [(torch.rand(B, 3, S, S, generator=g) * 2 - 1).to(dev) for _ in range(2)]
    ys = [torch.randint(0, ncls, (B,), generator=g).numpy() for _ in range(2)]
    if EAGER:
        steps, warmup = 1, 1
    for i in range(warmup):
        out = eng.step(xs[i % 2], ys[i % 2], use_graph=not EAGER)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = eng.step(xs[i % 2], ys[i % 2], use_graph=not EAGER)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    losses = [float(v) for v in out["losses"].cpu()]
    assert all(np.isfinite(losses)), losses
    rate = B / (ms * 1e-3)
    return {"case": name, "images_per_s": rate, "ms_per_step": ms, "batch": B, "cuda_graph": not EAGER, "steps": steps, "warmup": warmup, "dtype": "bf16",
            "lr": lr, "update_flops_per_image": flops, "step_frac_of_tensor_peak": rate * flops / 1e12 / peak_tf, "peak_TFLOP/s": peak_tf,
            "max_memory_GB": torch.cuda.max_memory_allocated() / 1e9, "losses": losses[:4]}


for case in CASES:
    try:
        torch.cuda.reset_peak_memory_stats()
        print(json.dumps(run(*case)), flush=True)
    except Exception as e:  # report and go on to the next shape
        print(json.dumps({"case": case[0], "error": "%s: %s" % (type(e).__name__, str(e)[:300]),
                          "trace": traceback.format_exc().splitlines()[-6:]}), flush=True)
    torch.cuda.empty_cache()
