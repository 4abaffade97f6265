"""eval.py of the reference (:83-217): evaluation of a trained victim classifier against the trained trigger generator --
clean accuracy, and on every NON-TARGET sample the benign accuracy (Bd BA) and the attack success rate (Bd ASR) of the
triggered image.  Same functions and signatures: get_model(opt) -> (netC, netG), eval(netC, netG, test_dl, tf_writer, opt),
main().

Per batch the reference runs netC(inputs), netG + low_freq + clamp + GaussianBlur on the gathered non-target rows, and
netC on those (:119-133).  Here the trigger is built for EVERY row of the fixed-shape batch (eval-mode networks, the DCT
projection and the blur are per-sample independent, so the non-target rows are bit-identical to the gathered sub-batch) and
the target rows are masked out of the counters with a negative label -- no device-to-host sync, no data-dependent shapes.
One blur sigma per batch from the torch CPU generator, as torchvision's GaussianBlur draws it (:113,131)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import config, ops
from .engine import create_targets_bd_np
from .modules import PreActResNet18, ResNet18, UnetGenerator
from .train_generator import _dataset_shape, _dtype, create_targets_bd, low_freq  # noqa: F401


def get_model(opt):
    """eval.py:83-105 (the 'default' classifiers; --model vgg13 / mobilenetv2 / vit* are outside the built path)."""
    kw = dict(device=opt.device, dtype=_dtype(opt))
    if opt.dataset == "cifar10":
        netC = PreActResNet18(**kw)
    elif opt.dataset == "celeba":
        netC = ResNet18(num_classes=opt.num_classes, **kw)
    elif opt.dataset == "imagenet10":
        netC = ResNet18(num_classes=opt.num_classes, n_input=opt.input_channel, input_size=opt.input_height, **kw)
    else:
        return None, None
    netG = UnetGenerator(opt, **kw)
    if opt.model != "default":
        raise NotImplementedError("--model %s is outside the built hot path" % opt.model)
    return netC, netG


def eval_batch(netC, netG, inputs, targets, opt, sigma=None, counts=None):
    """One iteration of eval.py:115-141.  Returns (device int32 counts [clean, -, bd_ba, bd_asr], n_bd, debug tensors)."""
    C_, G_ = netC.net, netG.net
    dev = C_.device
    y = targets.cpu().numpy().astype(np.int64) if torch.is_tensor(targets) else np.asarray(targets, dtype=np.int64)
    if sigma is None:
        sigma = torch.empty(1).uniform_(opt.sigma[0], opt.sigma[1]).item()
    ntrg = y != opt.target_label
    bd = create_targets_bd_np(y, opt)
    t = torch.from_numpy(np.stack([y, np.where(ntrg, y, -1), np.where(ntrg, bd, -1)])).to(dev, non_blocking=True)
    x = inputs.to(dev, non_blocking=True).float().contiguous()
    if counts is None:
        counts = torch.zeros(4, dtype=torch.int32, device=dev)
    preds_clean, _ = C_.forward(x, train=False, save=False)                                     # :119
    ops.cross_entropy(preds_clean, t[0], 1.0, False, counts_out=counts[0:2])
    noise_raw, _ = G_.forward(x, None, save=False)                                              # :128
    noise = ops.plane_op(noise_raw, "lowfreq", keep=int(opt.input_height * opt.ratio))          # :129
    x_bd = ops.poison_blend_fwd(x, noise, None, x.shape[0], opt.noise_rate, ops.gaussian_taps(sigma))   # :130-131
    preds_bd, _ = C_.forward(x_bd, train=False, save=False)                                     # :133
    ops.cross_entropy(preds_bd, t[1], 1.0, False, targets2=t[2], counts_out=counts[2:4])        # BA vs targets, ASR vs bd
    return counts, int(ntrg.sum()), dict(preds_clean=preds_clean, preds_bd=preds_bd, x_bd=x_bd, sigma=sigma)


def eval(netC, netG, test_dl, tf_writer, opt):
    """eval.py:108-152"""
    print(" Eval:")
    netC.eval()
    netG.eval()
    dev = netC.net.device
    tot = torch.zeros(4, dtype=torch.int64, device=dev)
    n_clean = n_bd = 0
    for inputs, targets in test_dl:
        if not inputs.is_cuda and not inputs.is_pinned() and torch.cuda.is_available():
            inputs = inputs.pin_memory()
        counts, nb, _ = eval_batch(netC, netG, inputs, targets, opt)
        tot += counts.long()
        n_clean += len(targets)
        n_bd += nb
    c = tot.cpu().numpy()
    acc_clean = c[0] * 100.0 / max(n_clean, 1)
    acc_bd_ba, acc_bd_asr = c[2] * 100.0 / max(n_bd, 1), c[3] * 100.0 / max(n_bd, 1)
    print("Clean Acc: {:.4f} | Bd BA: {:.4f} | Bd ASR: {:.4f}".format(acc_clean, acc_bd_ba, acc_bd_asr))
    tf_writer.add_scalars("Test Accuracy", {"Clean": acc_clean, "Bd BA": acc_bd_ba, "Bd ASR": acc_bd_asr}, 0)
    return acc_clean, acc_bd_ba, acc_bd_asr


class _NullWriter:
    def add_scalars(self, *a, **k):
        pass


def main(argv=None):
    """eval.py:155-217: the victim checkpoint <checkpoints>/<load_checkpoint_clean>/<dataset>/... and the generator checkpoint
    <checkpoints>/<load_checkpoint>/<dataset>/...; --synthetic_data (build-only) replaces the test set."""
    opt = config.get_arguments().parse_args(argv)
    _dataset_shape(opt)
    if opt.dataset == "imagenet10":
        opt.num_workers = 40
    from .utils.dataloader import get_dataloader
    test_dl = get_dataloader(opt, False)
    netC, netG = get_model(opt)
    mode = opt.saving_prefix
    opt.ckpt_folder = os.path.join(opt.checkpoints, "{}_clean".format(mode), opt.dataset)
    opt.ckpt_path = os.path.join(opt.ckpt_folder, "{}_{}_clean.pth.tar".format(opt.dataset, mode))
    opt.log_dir = os.path.join(opt.ckpt_folder, "log_dir")
    os.makedirs(opt.log_dir, exist_ok=True)
    for tag, net, key in ((opt.load_checkpoint_clean, netC, "netC"), (opt.load_checkpoint, netG, "netG")):
        load_path = os.path.join(opt.checkpoints, str(tag), opt.dataset, "{}_{}.pth.tar".format(opt.dataset, tag))
        if not os.path.exists(load_path):
            print("Error: {} not found".format(load_path))
            sys.exit()
        net.load_state_dict(torch.load(load_path, map_location=opt.device, weights_only=False)[key])
        net.eval()
    try:
        from torch.utils.tensorboard import SummaryWriter
        tf_writer = SummaryWriter(log_dir=opt.log_dir)
    except Exception:
        tf_writer = _NullWriter()
    return eval(netC, netG, test_dl, tf_writer, opt)


if __name__ == "__main__":
    main()
