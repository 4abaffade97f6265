"""train_generator.py of the reference, hot-path surface: low_freq, create_targets_bd, get_model, train, main.

`train()` keeps the reference's signature (train_generator.py:131-144) and side effects: it updates netC / netG and
their BatchNorm buffers in place, steps both LR schedulers once, writes the same scalars to `tf_writer`, and prints
the same six running accuracies -- but every iteration is ONE replay of a captured CUDA graph of hand-written
kernels (combat_b200.engine.AlternatedStep) instead of ~1500 eager library launches, and the metric counters are
accumulated on the device and read back every `--log_every` iterations instead of forcing a sync per iteration.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import config, ops
from .engine import AlternatedStep, make_plan
from .modules import CUnetGeneratorv1, FrequencyModel, PreActResNet18, ResNet18, UnetGenerator  # noqa: F401
from .networks.models import Denormalizer
from .utils.dataloader import PostTensorTransform


def low_freq(x, opt):
    """train_generator.py:47-55 -- idct_2d(mask * dct_2d((x+1)/2*255))/255*2-1 == P x P^T (one fused kernel).
    Differentiable: P is symmetric, so the backward is the same projection."""
    keep = int(opt.input_height * opt.ratio)
    return _LowFreq.apply(x, keep)


class _LowFreq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, keep):
        ctx.keep = keep
        return ops.plane_op(x.contiguous().float(), "lowfreq", keep=keep)

    @staticmethod
    def backward(ctx, g):
        return ops.plane_op(g.contiguous().float(), "lowfreq", keep=ctx.keep), None


def create_targets_bd(targets, opt):
    """train_generator.py:70-77"""
    if opt.attack_mode == "all2one":
        bd_targets = torch.ones_like(targets) * opt.target_label
    elif opt.attack_mode == "all2all":
        bd_targets = torch.tensor([(label + 1) % opt.num_classes for label in targets])
    else:
        raise Exception("{} attack mode is not implemented".format(opt.attack_mode))
    return bd_targets.to(opt.device)


def _dtype(opt):
    return torch.float32 if getattr(opt, "dtype", "bf16") == "fp32" else torch.bfloat16


def get_model(opt):
    """train_generator.py:80-128: (netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model).
    Construction order netC, clean_model, netG, netF (it fixes the RNG stream of the initialisation)."""
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
    netG = UnetGenerator(opt, **kw)
    if opt.model != "default" or opt.model_clean != "default":
        raise NotImplementedError("--model/--model_clean other than 'default' are outside the built hot path")
    if opt.F_model not in ("original", "original_holdout"):
        raise NotImplementedError("--F_model %s is outside the built hot path" % opt.F_model)
    netF = FrequencyModel(num_classes=2, n_input=opt.input_channel, input_size=opt.input_height, **kw) \
        if opt.input_height in (32, 64) else None
    optimizerC = torch.optim.SGD(netC.parameters(), opt.lr_C, momentum=0.9, weight_decay=5e-4, nesterov=True)
    schedulerC = torch.optim.lr_scheduler.MultiStepLR(optimizerC, opt.schedulerC_milestones, opt.schedulerC_lambda)
    optimizerG = torch.optim.SGD(netG.parameters(), opt.lr_G, momentum=0.9, weight_decay=5e-4, nesterov=True)
    schedulerG = torch.optim.lr_scheduler.MultiStepLR(optimizerG, opt.schedulerG_milestones, opt.schedulerG_lambda)
    return netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model


_ENGINES: dict = {}


def _engine_for(netC, clean_model, netG, netF, opt):
    key = (id(netC), id(clean_model), id(netG), id(netF))
    eng = _ENGINES.get(key)
    if eng is None:
        eng = AlternatedStep(opt, device=netC.net.device, with_metrics=True,
                             nets=(netC.net, clean_model.net, netG.net, netF.net if netF is not None else None))
        _ENGINES[key] = eng
    return eng


def _bind_momentum(optimizer, module):
    """expose the fused optimiser's momentum buffers through the torch optimiser's state (checkpoint round-trip)."""
    for name, p in module._plist:
        optimizer.state[p]["momentum_buffer"] = module.net.store.m(name)


def train(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, train_dl, tf_writer, epoch, opt):
    """train_generator.py:131-318 (one epoch of alternated C/G steps)."""
    print(" Train:")
    netC.train()
    PostTensorTransform(opt)  # raises for the options that are not built
    eng = _engine_for(netC, clean_model, netG, netF, opt)
    for pg_c, pg_g in zip(optimizerC.param_groups, optimizerG.param_groups):
        for pg in (pg_c, pg_g):
            if not (pg["momentum"] == 0.9 and pg["weight_decay"] == 5e-4 and pg["nesterov"]):
                raise NotImplementedError("the fused optimiser implements the reference's SGD(0.9, 5e-4, nesterov) only")
    eng.set_lr(optimizerC.param_groups[0]["lr"], optimizerG.param_groups[0]["lr"])
    use_graph = not getattr(opt, "no_graph", False)
    log_every = max(1, int(getattr(opt, "log_every", 50)))
    dev = netC.net.device
    tot = torch.zeros(16, dtype=torch.int64, device=dev)
    lsum = torch.zeros(8, dtype=torch.float64, device=dev)
    total_sample = 0
    n_batches = len(train_dl)
    acc = {}
    inputs = None
    it = iter(train_dl)
    nxt = next(it, None)
    if nxt is not None and not nxt[0].is_cuda:
        nxt = (nxt[0].pin_memory(), nxt[1])
    batch_idx = -1
    while nxt is not None:
        batch_idx += 1
        inputs, targets = nxt
        y_host = targets.cpu().numpy() if torch.is_tensor(targets) else np.asarray(targets)
        plan = make_plan(y_host, opt)
        nxt = next(it, None)  # overlap the next batch's host->device copy with this iteration
        if nxt is not None and not nxt[0].is_cuda:
            nxt = (nxt[0].pin_memory(), nxt[1])
        out = eng.step(inputs, y_host, plan, use_graph=use_graph)
        if nxt is not None:
            eng.prefetch(nxt[0])
        tot += out["counts"].long()
        lsum += out["losses"].double()
        total_sample += len(y_host)
        if (batch_idx + 1) % log_every == 0 or batch_idx + 1 == n_batches:
            c = tot.cpu().numpy()
            l = lsum.cpu().numpy()
            acc = dict(avg_acc_clean=c[4] * 100.0 / total_sample, avg_acc_bd=c[6] * 100.0 / total_sample,
                       avg_acc_F=c[10] * 100.0 / total_sample, avg_clean_model_acc=c[2] * 100.0 / total_sample,
                       avg_clean_model_bd_ba=c[8] * 100.0 / total_sample, avg_clean_model_bd_asr=c[9] * 100.0 / total_sample,
                       avg_loss_l2=l[2] / total_sample, avg_clean_model_loss=l[3] / total_sample)
            print("[%d/%d] Clean Acc: %.4f | Bd Acc: %.4f | F Acc: %.4f | Clean Model Acc: %.4f | Clean Model Bd BA: %.4f | "
                  "Clean Model Bd ASR: %.4f" % (batch_idx + 1, n_batches, acc["avg_acc_clean"], acc["avg_acc_bd"], acc["avg_acc_F"],
                                               acc["avg_clean_model_acc"], acc["avg_clean_model_bd_ba"],
                                               acc["avg_clean_model_bd_asr"]))
    if acc and not epoch % 1:
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


def eval(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, test_dl, best_clean_acc, best_bd_acc,
         best_F_acc, best_clean_model_acc, best_clean_model_bd_ba, best_clean_model_bd_asr, tf_writer, epoch, opt):
    """train_generator.py:321-465: clean accuracy, attack success on every non-target sample, detector and clean-model
    legs; saves the reference's checkpoint dict to opt.ckpt_path when the clean accuracy improves.  Returns the six bests."""
    print(" Eval:")
    netC.eval()
    eng = _engine_for(netC, clean_model, netG, netF, opt)
    use_graph = not getattr(opt, "no_graph", False)
    dev = netC.net.device
    tot = torch.zeros(16, dtype=torch.int64, device=dev)
    n_clean = n_bd = 0
    for inputs, targets in test_dl:
        y_host = targets.cpu().numpy() if torch.is_tensor(targets) else np.asarray(targets)
        if not inputs.is_cuda:
            inputs = inputs.pin_memory()
        out = eng.eval_step(inputs, y_host, use_graph=use_graph)
        tot += out["counts"].long()
        n_clean += out["n"]
        n_bd += out["n_bd"]
    c = tot.cpu().numpy()
    n_bd_ = max(n_bd, 1)
    acc_clean, acc_bd, acc_F = c[0] * 100.0 / n_clean, c[2] * 100.0 / n_bd_, c[4] * 100.0 / n_bd_
    acc_clean_model, bd_ba_clean_model, bd_asr_clean_model = c[6] * 100.0 / n_clean, c[8] * 100.0 / n_bd_, c[9] * 100.0 / n_bd_
    print("Clean Acc: {:.4f} - Best: {:.4f} | Bd Acc: {:.4f} - Best: {:.4f} | F Acc: {:.4f} - Best: {:.4f} | Clean Model Acc: {:.4f} - "
          "Best: {:.4f} | Clean Model Bd BA: {:.4f} - Best: {:.4f} | Clean Model Bd ASR: {:.4f} - Best: {:.4f}".format(
              acc_clean, best_clean_acc, acc_bd, best_bd_acc, acc_F, best_F_acc, acc_clean_model, best_clean_model_acc,
              bd_ba_clean_model, best_clean_model_bd_ba, bd_asr_clean_model, best_clean_model_bd_asr))
    if not epoch % 1:
        tf_writer.add_scalars("Test Accuracy", {"Clean": acc_clean, "Bd": acc_bd, "F": acc_F, "Clean Model Acc": acc_clean_model,
                                                "Clean Model Bd BA": bd_ba_clean_model, "Clean Model Bd ASR": bd_asr_clean_model}, epoch)
    if acc_clean > best_clean_acc or (acc_clean == best_clean_acc and acc_bd > best_bd_acc):  # :433
        print(" Saving...")
        best_clean_acc, best_bd_acc, best_F_acc = acc_clean, acc_bd, acc_F
        best_clean_model_acc, best_clean_model_bd_ba, best_clean_model_bd_asr = acc_clean_model, bd_ba_clean_model, bd_asr_clean_model
        state_dict = {
            "netC": netC.state_dict(), "schedulerC": schedulerC.state_dict(), "optimizerC": optimizerC.state_dict(),
            "netG": netG.state_dict(), "schedulerG": schedulerG.state_dict(), "optimizerG": optimizerG.state_dict(),
            "clean_model": clean_model.state_dict(), "best_clean_acc": acc_clean, "best_bd_acc": acc_bd, "best_F_acc": acc_F,
            "best_clean_model_acc": best_clean_model_acc, "best_clean_model_bd_ba": best_clean_model_bd_ba,
            "best_clean_model_bd_asr": best_clean_model_bd_asr, "epoch_current": epoch,
        }
        ckpt_dir = os.path.dirname(opt.ckpt_path)
        if ckpt_dir:
            os.makedirs(ckpt_dir, exist_ok=True)
        torch.save(state_dict, opt.ckpt_path)
    return (best_clean_acc, best_bd_acc, best_F_acc, best_clean_model_acc, best_clean_model_bd_ba, best_clean_model_bd_asr)


class _NullWriter:
    def add_scalars(self, *a, **k):
        pass

    def add_image(self, *a, **k):
        pass


def main(argv=None):
    """Synthetic-data driver of the training loop (dataset loading is outside the built path)."""
    opt = config.get_arguments().parse_args(argv)
    if opt.dataset == "cifar10":
        opt.input_height = opt.input_width = 32
        opt.input_channel = 3
    elif opt.dataset == "celeba":
        opt.input_height = opt.input_width = 64
        opt.input_channel, opt.num_classes = 3, 8
    else:
        raise Exception("Invalid Dataset")
    netC, optC, schC, netG, optG, schG, netF, clean = get_model(opt)
    g = torch.Generator().manual_seed(0)
    n_it = 8 if opt.debug else 32
    data = [(torch.rand(opt.bs, 3, opt.input_height, opt.input_width, generator=g) * 2 - 1,
             torch.randint(0, opt.num_classes, (opt.bs,), generator=g)) for _ in range(n_it)]
    for epoch in range(1, 3):
        print("Epoch {} - {} | noise_rate: {} pc: {}".format(epoch, opt.dataset, opt.noise_rate, opt.pc))
        train(netC, optC, schC, netG, optG, schG, netF, clean, data, _NullWriter(), epoch, opt)


if __name__ == "__main__":
    main()
