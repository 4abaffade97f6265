"""train_generator_wanet.py of the reference: the alternated step of train_generator.py with a WARP trigger instead of the additive
one (train_generator_wanet.py:150-158,196-203) --

    noise_grid = bicubic_upsample(GridGenerator(inputs), H x H, align_corners=True).permute(0, 2, 3, 1)
    inputs_bd  = grid_sample(inputs, clamp(identity_grid * (1 - grid_rescale) + noise_grid * grid_rescale, -1, 1), align_corners=True)
    loss       = CE(netC(T(inputs_bd)), bd_targets) + L2_weight * MSE(noise_grid, 0) + clean_model_weight * CE(clean_model(T(inputs_bd)), targets)

(--s 2: side of the flow's control grid, --grid_rescale 0.15), no DCT low-pass and no GaussianBlur (so no sigma draws), and the
"Grad L2 Loss" scalar is the finite-difference term of the noise grid (:213-222).  `identity_grid` keeps its place in the
signatures; the kernels rebuild it from `torch.linspace(-1, 1, input_height)` (:560-562), which is what the reference passes.
Same engine, captured graphs and classifier kernels as the base step (engine.AlternatedStep, `variant = "wanet"`); the trigger
is ONE fused kernel and its backward (csrc/warp.cu), the generator is nets.GridGenerator.  The reference raises when an
iteration poisons no row (reshape of an empty tensor, networks/models.py:383); here such an iteration simply has no poisoned row.
Pinned by tests/golden/step_wanet_b32x2.npz, recorded from the unmodified reference variant.
"""
from __future__ import annotations

import torch

from . import train_generator as _base
from .modules import FrequencyModel, GridGenerator, PreActResNet18, ResNet18
from .train_generator import _dtype, create_targets_bd  # noqa: F401


def _variant(opt):
    opt.variant = "wanet"
    return opt


def _check_grid(identity_grid, opt):
    """the step takes the identity grid the reference's main() builds (:560-562); anything else is not this path"""
    if identity_grid is None:
        return
    a = torch.linspace(-1, 1, steps=opt.input_height)
    gx, gy = torch.meshgrid(a, a, indexing="ij")
    ref = torch.stack((gy, gx), 2)[None, ...]
    if tuple(identity_grid.shape) != tuple(ref.shape) or not torch.equal(identity_grid.detach().cpu().float(), ref):
        raise NotImplementedError("identity_grid other than the linspace(-1, 1, input_height) mesh of the reference's main()")


def get_model(opt):
    """train_generator_wanet.py:53-92 (construction order netC, clean_model, netG, netF)."""
    kw = dict(device=opt.device, dtype=_dtype(opt))
    if opt.dataset == "cifar10":
        netC, clean_model = PreActResNet18(**kw), PreActResNet18(**kw)
    elif opt.dataset == "celeba":
        netC, clean_model = ResNet18(num_classes=opt.num_classes, **kw), ResNet18(num_classes=opt.num_classes, **kw)
    elif opt.dataset == "imagenet10":
        netC = ResNet18(num_classes=opt.num_classes, input_size=opt.input_height, **kw)
        clean_model = ResNet18(num_classes=opt.num_classes, input_size=opt.input_height, **kw)
    else:
        raise Exception("Invalid Dataset")
    netG = GridGenerator(opt, **kw)
    if opt.F_model not in ("original", "original_holdout"):
        raise NotImplementedError("--F_model %s is outside the built hot path" % opt.F_model)
    netF = FrequencyModel(num_classes=2, n_input=opt.input_channel, input_size=opt.input_height, **kw) \
        if opt.input_height in (32, 64) else None
    optimizerC = torch.optim.SGD(netC.parameters(), opt.lr_C, momentum=0.9, weight_decay=5e-4, nesterov=True)
    schedulerC = torch.optim.lr_scheduler.MultiStepLR(optimizerC, opt.schedulerC_milestones, opt.schedulerC_lambda)
    optimizerG = torch.optim.SGD(netG.parameters(), opt.lr_G, momentum=0.9, weight_decay=5e-4, nesterov=True)
    schedulerG = torch.optim.lr_scheduler.MultiStepLR(optimizerG, opt.schedulerG_milestones, opt.schedulerG_lambda)
    return netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model


def train(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, train_dl, identity_grid, tf_writer, epoch,
          opt):
    """train_generator_wanet.py:95-300"""
    _check_grid(identity_grid, opt)
    return _base.train(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, train_dl, tf_writer, epoch,
                       _variant(opt))


def eval(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, test_dl, identity_grid, best_clean_acc,
         best_bd_acc, best_F_acc, best_clean_model_acc, best_clean_model_bd_ba, best_clean_model_bd_asr, tf_writer, epoch, opt):
    """train_generator_wanet.py:303-457 (same counters, save rule and checkpoint dict as the base trainer's eval)"""
    _check_grid(identity_grid, opt)
    return _base.eval(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, test_dl, best_clean_acc,
                      best_bd_acc, best_F_acc, best_clean_model_acc, best_clean_model_bd_ba, best_clean_model_bd_asr, tf_writer,
                      epoch, _variant(opt))


def main(argv=None):
    """train_generator_wanet.py:460-611: the base driver with this module's get_model / train / eval and the identity grid"""
    def with_grid(fn):
        def run(*a):
            opt = a[-1]
            aa = torch.linspace(-1, 1, steps=opt.input_height)                                  # :560-562
            gx, gy = torch.meshgrid(aa, aa, indexing="ij")
            grid = torch.stack((gy, gx), 2)[None, ...].to(opt.device)
            return fn(*a[:9], grid, *a[9:])
        return run
    return _base.main(argv, train_fn=with_grid(train), eval_fn=with_grid(eval), get_model_fn=get_model)


if __name__ == "__main__":
    main()
