"""defenses/frequency_based/test.py of the reference (:36-167): how well the frequency detector separates clean images from the
generator's poisoned ones -- get_model(opt) (the shipped 'original' detector), test(netC, netG, test_dl, opt), main().

Per batch (:76-103): trigger on EVERY image (netG -> low_freq -> blend + clamp -> GaussianBlur, one sigma per batch), uint8 DCT of
the clean and the poisoned images, detector logits over [clean ; poisoned] with labels [0 ; 1]; accuracy over both halves and the
detection rate over the poisoned half.  Here: the generator, the fused blend and ONE launch of the uint8-input DCT kernel over the
2*bs images (the call site pinned by tests/test_z_detector_dct_gpu.py), the detector forward, two counter launches -- no host
round trip per image.  As shipped the per-plane loop (:88-94) calls the TORCH `utils.dct.dct_2d` on a numpy uint8 plane, which
raises (ndarray has no `.contiguous()`); the evident intent -- the scipy `dct2` defined at :19-20 and used by train.py:195-197, i.e.
the orthonormal 2-D DCT of the uint8 plane -- is what is computed.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from ... import ops
from ...modules import FrequencyModel, UnetGenerator
from . import config


def get_model(opt):
    """test.py:36-64 for --model original / original_holdout."""
    if opt.model not in ("original", "original_holdout"):
        raise NotImplementedError("--model %s is outside the built path (the shipped detector is 'original')" % opt.model)
    netC = FrequencyModel(num_classes=2, n_input=opt.input_channel, input_size=opt.input_height, device=opt.device,
                          dtype=torch.float32)
    optimizerC = torch.optim.Adadelta(netC.parameters(), lr=0.05, weight_decay=1e-4)
    return netC, optimizerC


def test_batch(netC, netG, x, opt, sigma=None):
    """One iteration of :76-103.  Returns (device int32 counts [correct over 2*bs, -, detected over the poisoned half, -], debug)."""
    F_, G_ = netC.net, netG.net
    dev = G_.device
    x = x.to(dev, non_blocking=True).float().contiguous()
    bs = x.shape[0]
    if sigma is None:
        sigma = torch.empty(1).uniform_(float(opt.sigma[0]), float(opt.sigma[1])).item()      # gauss_smooth, :74,83
    noise_raw, _ = G_.forward(x, None, save=False)                                               # :80
    noise = ops.plane_op(noise_raw, "lowfreq", keep=int(opt.input_height * opt.ratio))           # :81
    both = torch.empty((2 * bs,) + tuple(x.shape[1:]), dtype=torch.float32, device=dev)
    both[:bs].copy_(x)
    ops.poison_blend_fwd(x, noise, None, bs, opt.noise_rate, ops.gaussian_taps(sigma), out=both[bs:])   # :82-83
    coef = ops.plane_op(both, "dct", in_mode=2)                                                  # :88-94: uint8((v + 1) / 2 * 255), DCT
    preds = F_.forward(coef)                                                                     # :97
    labels = torch.cat([torch.zeros(bs, dtype=torch.int64), torch.ones(bs, dtype=torch.int64)]).to(dev)
    counts = torch.zeros(4, dtype=torch.int32, device=dev)
    ops.cross_entropy(preds, labels, 1.0, False, counts_out=counts[0:2])                         # :101
    ops.cross_entropy(preds[bs:], labels[bs:], 1.0, False, counts_out=counts[2:4])               # :99
    return counts, dict(preds=preds, poi_x=both[bs:], coef=coef, sigma=sigma)


def test(netC, netG, test_dl, opt):
    """test.py:67-110; prints and returns (accuracy, detection rate)."""
    netC.eval()
    dev = netG.net.device
    tot = torch.zeros(4, dtype=torch.int64, device=dev)
    total_poi_sample = 0
    for x, _y in test_dl:
        counts, _ = test_batch(netC, netG, x, opt)
        tot += counts.long()
        total_poi_sample += x.shape[0]
    c = tot.cpu().numpy()
    acc = c[0] * 100.0 / max(2 * total_poi_sample, 1)
    detection_rate = c[2] * 100.0 / max(total_poi_sample, 1)
    print("Acc: {:.4f} - Detection rate: {:.4f}".format(acc, detection_rate))
    return acc, detection_rate


def main(argv=None):
    """test.py:113-163: detector checkpoint `<checkpoints>/<dataset>/<model>/<dataset>_<model>_detector.pth.tar["netC"]`, generator
    checkpoint `<load_checkpoint>/<prefix>_clean/<dataset>/<dataset>_<prefix>_clean.pth.tar["netG"]` (--synthetic_data: random
    weights where a checkpoint is absent)."""
    opt = config.get_arguments().parse_args(argv)
    sizes = {"cifar10": (32, 10), "celeba": (64, 8), "imagenet10": (224, 10)}
    if opt.dataset not in sizes:
        raise Exception("Invalid Dataset")
    opt.input_height = opt.input_width = sizes[opt.dataset][0]
    opt.input_channel, opt.num_classes = 3, sizes[opt.dataset][1]
    opt.post_transform_option = "no_use"
    from ...utils.dataloader import get_dataloader
    test_dl = get_dataloader(opt, False)
    netC, _ = get_model(opt)
    opt.ckpt_folder = os.path.join(opt.checkpoints, opt.dataset, opt.model)
    opt.ckpt_path = os.path.join(opt.ckpt_folder, "{}_{}_detector.pth.tar".format(opt.dataset, opt.model))
    if os.path.exists(opt.ckpt_path) or not opt.synthetic_data:
        netC.load_state_dict(torch.load(opt.ckpt_path, map_location=opt.device, weights_only=False)["netC"])
    netG = UnetGenerator(opt, device=opt.device)
    load_path = os.path.join(opt.load_checkpoint, "{}_clean".format(opt.saving_prefix), opt.dataset,
                             "{}_{}_clean.pth.tar".format(opt.dataset, opt.saving_prefix))
    if os.path.exists(load_path):
        netG.load_state_dict(torch.load(load_path, map_location=opt.device, weights_only=False)["netG"])
    elif not opt.synthetic_data:
        print("Error: {} not found".format(load_path))
        sys.exit()
    netG.eval()
    netG.requires_grad_(False)
    return test(netC, netG, test_dl, opt)


if __name__ == "__main__":
    main()
