"""train_clean_classifier.py of the reference (:51-120): the clean classifier that train_generator.py loads as `clean_model`
(--load_checkpoint_clean).  get_model(opt) -> (netC, optimizerC, schedulerC); train(netC, optimizerC, schedulerC, train_dl,
tf_writer, epoch, opt): PostTensorTransform, netC train-mode forward/backward, SGD per batch -- the same captured graph as the
victim trainer with no poisoned rows and no generator."""
from __future__ import annotations

import torch

from .modules import PreActResNet18, ResNet18
from .train_generator import _dtype
from .train_victim import _train_epoch


def get_model(opt):
    """train_clean_classifier.py:51-72"""
    kw = dict(device=opt.device, dtype=_dtype(opt))
    if opt.dataset == "cifar10":
        netC = PreActResNet18(**kw)
    elif opt.dataset == "celeba":
        netC = ResNet18(num_classes=opt.num_classes, **kw)
    elif opt.dataset == "imagenet10":
        netC = ResNet18(num_classes=opt.num_classes, n_input=opt.input_channel, input_size=opt.input_height, **kw)
    else:
        raise Exception("Invalid Dataset")
    if opt.model != "default":
        raise NotImplementedError("--model %s is outside the built hot path" % opt.model)
    optimizerC = torch.optim.SGD(netC.parameters(), opt.lr_C, momentum=0.9, weight_decay=5e-4, nesterov=True)
    schedulerC = torch.optim.lr_scheduler.MultiStepLR(optimizerC, opt.schedulerC_milestones, opt.schedulerC_lambda)
    return netC, optimizerC, schedulerC


def train(netC, optimizerC, schedulerC, train_dl, tf_writer, epoch, opt):
    """train_clean_classifier.py:75-120; train_dl yields (inputs, targets)."""
    _train_epoch(netC, optimizerC, schedulerC, None, train_dl, tf_writer, epoch, opt, False)
