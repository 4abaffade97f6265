"""train_generator_multilabel.py of the reference, hot-path surface: get_model and train (reference :78-137, :142-318).

Differences from train_generator.py that the engine's multilabel mode implements (engine.AlternatedStep(multilabel=True)):
conditional generator `CUnetGeneratorv1(x, y)`; the C-step poisons the FIRST num_bd rows (num_bd from `np.random.rand(bs)`,
:171) with the trigger conditioned on the true labels, labels unchanged; the G-step splits the batch into num_classes
contiguous chunks, chunk ci is pushed towards class ci with its own blur sigma (:203-221); netG's optimiser runs at
`lr_C * 0.1` (:115).  `train()` keeps the reference's signature, including the unused `mask` / `pattern` arguments.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import N_LOSSES, AlternatedStep, make_plan_multilabel
from .modules import CUnetGeneratorv1, FrequencyModel, PreActResNet18, ResNet18
from .train_generator import _adopt_momentum, _bind_momentum, _dtype, _engine_for, create_targets_bd, low_freq  # noqa: F401
from .utils.dataloader import PostTensorTransform


def get_model(opt):
    """reference :78-137 (construction order netC, clean_model, netG, netF).  The reference passes an unknown keyword
    to CUnetGeneratorv1 for celeba (TypeError as shipped); the class is built from `opt.num_classes` as its signature says."""
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
    netG = CUnetGeneratorv1(opt, **kw)
    if opt.F_model not in ("original", "original_holdout"):
        raise NotImplementedError("--F_model %s is outside the built hot path" % opt.F_model)
    netF = FrequencyModel(num_classes=2, n_input=opt.input_channel, input_size=opt.input_height, **kw) \
        if opt.input_height in (32, 64) else None
    optimizerC = torch.optim.SGD(netC.parameters(), opt.lr_C, momentum=0.9, weight_decay=5e-4, nesterov=True)
    schedulerC = torch.optim.lr_scheduler.MultiStepLR(optimizerC, opt.schedulerC_milestones, opt.schedulerC_lambda)
    optimizerG = torch.optim.SGD(netG.parameters(), opt.lr_C * 0.1, momentum=0.9, weight_decay=5e-4, nesterov=True)
    schedulerG = torch.optim.lr_scheduler.MultiStepLR(optimizerG, opt.schedulerC_milestones, opt.schedulerC_lambda)
    return netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model


def train(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, train_dl, mask, pattern, tf_writer,
          epoch, opt):
    """reference :142-318 (one epoch of alternated multilabel C/G steps), every iteration one captured-graph replay."""
    print(" Train:")
    netC.train()
    eng = _engine_for(netC, clean_model, netG, netF, opt, multilabel=True)
    _adopt_momentum(optimizerC, netC)
    _adopt_momentum(optimizerG, netG)
    eng.set_lr(optimizerC.param_groups[0]["lr"], optimizerG.param_groups[0]["lr"])
    use_graph = not getattr(opt, "no_graph", False)
    log_every = max(1, int(getattr(opt, "log_every", 50)))
    dev = netC.net.device
    tot = torch.zeros(16, dtype=torch.int64, device=dev)
    lsum = torch.zeros(N_LOSSES, dtype=torch.float64, device=dev)
    total_sample, n_batches, acc = 0, len(train_dl), {}
    for batch_idx, (inputs, targets) in enumerate(train_dl):
        y_host = targets.cpu().numpy() if torch.is_tensor(targets) else np.asarray(targets)
        plan = make_plan_multilabel(y_host, opt, eng.with_metrics)
        if not inputs.is_cuda:
            inputs = inputs.pin_memory()
        out = eng.step(inputs, y_host, plan, use_graph=use_graph)
        tot += out["counts"].long()
        lsum += out["losses"].double()
        total_sample += len(y_host)
        if (batch_idx + 1) % log_every == 0 or batch_idx + 1 == n_batches:
            c, l = tot.cpu().numpy(), lsum.cpu().numpy()
            acc = dict(avg_acc_clean=c[4] * 100.0 / total_sample, avg_acc_bd=c[6] * 100.0 / total_sample,
                       avg_acc_F=c[10] * 100.0 / total_sample, avg_clean_model_acc=c[2] * 100.0 / total_sample,
                       avg_clean_model_bd_ba=c[8] * 100.0 / total_sample, avg_clean_model_bd_asr=c[9] * 100.0 / total_sample,
                       avg_loss_l2=l[2] / total_sample, avg_clean_model_loss=l[3] / total_sample)
            print("[%d/%d] Clean Acc: %.4f | Bd Acc: %.4f | F Acc: %.4f | Clean Model Acc: %.4f | Clean Model Bd BA: %.4f | "
                  "Clean Model Bd ASR: %.4f" % (batch_idx + 1, n_batches, acc["avg_acc_clean"], acc["avg_acc_bd"], acc["avg_acc_F"],
                                               acc["avg_clean_model_acc"], acc["avg_clean_model_bd_ba"],
                                               acc["avg_clean_model_bd_asr"]))
    if acc:
        tf_writer.add_scalars("Clean Accuracy", {
            "Clean": acc["avg_acc_clean"], "Bd": acc["avg_acc_bd"], "F": acc["avg_acc_F"],
            "CleanModel Acc": acc["avg_clean_model_acc"], "CleanModel Bd BA": acc["avg_clean_model_bd_ba"],
            "CleanModel Bd ASR": acc["avg_clean_model_bd_asr"], "L2 Loss": acc["avg_loss_l2"],
            "CleanModel Loss": acc["avg_clean_model_loss"]}, epoch)
    _bind_momentum(optimizerC, netC)
    _bind_momentum(optimizerG, netG)
    for n, b in netC.named_buffers():
        if n.endswith("num_batches_tracked"):
            b.fill_(netC.net.num_batches_tracked[n[: -len(".num_batches_tracked")]])
    schedulerC.step()
    schedulerG.step()
