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
from .engine import N_LOSSES, AlternatedStep, make_plan
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


_HOT_SCALARS = ("noise_rate", "ratio", "L2_weight", "clean_model_weight", "target_label", "attack_mode", "num_classes",
                "post_transform_option", "random_crop", "random_rotation", "dataset", "variant", "tv_weight", "cross_weight", "s",
                "grid_rescale")


def _engine_for(netC, clean_model, netG, netF, opt, multilabel=False):
    """The engine of a (netC, clean_model, netG, netF) quadruple lives ON the netC module (no id()-keyed registry: ids are
    reused after garbage collection).  Flags that are baked into captured graphs as kernel arguments are compared on every
    call; when one changed the engine is rebuilt (the networks and their optimiser state are untouched)."""
    sig = tuple(getattr(opt, k, None) for k in _HOT_SCALARS) + (bool(multilabel),)
    rec = getattr(netC, "_combat_engine", None)
    if rec is not None:
        eng, others, old_sig = rec
        if others[0] is clean_model and others[1] is netG and others[2] is netF and old_sig == sig:
            eng.opt = opt
            return eng
    eng = AlternatedStep(opt, device=netC.net.device, with_metrics=True, multilabel=multilabel,
                         nets=(netC.net, clean_model.net, netG.net, netF.net if netF is not None else None))
    object.__setattr__(netC, "_combat_engine", (eng, (clean_model, netG, netF), sig))
    return eng


def _bind_momentum(optimizer, module):
    """expose the fused optimiser's momentum buffers through the torch optimiser's state (checkpoint round-trip)."""
    if module.net.store.first_step:
        return  # no step taken yet: torch's SGD has no momentum_buffer either
    for name, p in module._plist:
        optimizer.state[p]["momentum_buffer"] = module.net.store.m(name)


def _adopt_momentum(optimizer, module):
    """`optimizer.load_state_dict(ckpt)` (train_generator.py:536-539, --continue_training) leaves fresh copies of the momentum
    buffers in optimizer.state; the fused SGD reads the flat store.  Copy any buffer that does not alias the store into it
    and clear `first_step`, so that a resumed run continues with the saved Nesterov momentum instead of re-initialising it."""
    st = module.net.store
    found = False
    for name, p in module._plist:
        buf = optimizer.state.get(p, {}).get("momentum_buffer")
        if buf is None:
            continue
        found = True
        m = st.m(name)
        if buf.data_ptr() != m.data_ptr():
            m.copy_(buf.to(m.device, torch.float32))
            optimizer.state[p]["momentum_buffer"] = m
    if found:
        st.first_step = False


def train(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, train_dl, tf_writer, epoch, opt):
    """train_generator.py:131-318 (one epoch of alternated C/G steps)."""
    print(" Train:")
    netC.train()
    eng = _engine_for(netC, clean_model, netG, netF, opt)   # PostTensorTransform(opt) lives inside the engine (:168,:196...)
    _adopt_momentum(optimizerC, netC)
    _adopt_momentum(optimizerG, netG)
    for pg_c, pg_g in zip(optimizerC.param_groups, optimizerG.param_groups):
        for pg in (pg_c, pg_g):
            if not (pg["momentum"] == 0.9 and pg["weight_decay"] == 5e-4 and pg["nesterov"]):
                raise NotImplementedError("the fused optimiser implements the reference's SGD(0.9, 5e-4, nesterov) only")
    eng.set_lr(optimizerC.param_groups[0]["lr"], optimizerG.param_groups[0]["lr"])
    use_graph = not getattr(opt, "no_graph", False)
    log_every = max(1, int(getattr(opt, "log_every", 50)))
    dev = netC.net.device
    tot = torch.zeros(16, dtype=torch.int64, device=dev)
    lsum = torch.zeros(N_LOSSES, dtype=torch.float64, device=dev)
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
        plan = make_plan(y_host, opt, eng.with_metrics)
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
                       avg_loss_l2=l[2] / total_sample, avg_clean_model_loss=l[3] / total_sample,
                       avg_loss_grad_l2=l[7] / total_sample, avg_loss_tv=l[8] / total_sample)
            print("[%d/%d] Clean Acc: %.4f | Bd Acc: %.4f | F Acc: %.4f | Clean Model Acc: %.4f | Clean Model Bd BA: %.4f | "
                  "Clean Model Bd ASR: %.4f" % (batch_idx + 1, n_batches, acc["avg_acc_clean"], acc["avg_acc_bd"], acc["avg_acc_F"],
                                               acc["avg_clean_model_acc"], acc["avg_clean_model_bd_ba"],
                                               acc["avg_clean_model_bd_asr"]))
    if acc and not epoch % 1:
        scalars = {
            "Clean": acc["avg_acc_clean"], "Bd": acc["avg_acc_bd"], "F": acc["avg_acc_F"],
            "CleanModel Acc": acc["avg_clean_model_acc"], "CleanModel Bd BA": acc["avg_clean_model_bd_ba"],
            "CleanModel Bd ASR": acc["avg_clean_model_bd_asr"], "L2 Loss": acc["avg_loss_l2"],
            "Grad L2 Loss": acc["avg_loss_grad_l2"], "CleanModel Loss": acc["avg_clean_model_loss"]}
        if getattr(opt, "variant", "") == "imperceptible":   # train_generator_imperceptible.py:304
            scalars["TV Loss"] = acc["avg_loss_tv"]
        tf_writer.add_scalars("Clean Accuracy", scalars, epoch)
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


def _dataset_shape(opt):
    """train_generator.py:468-487"""
    if opt.dataset == "cifar10":
        opt.input_height = opt.input_width = 32
        opt.input_channel = 3
    elif opt.dataset == "celeba":
        opt.input_height = opt.input_width = 64
        opt.input_channel = 3
        opt.num_workers = 40
        opt.num_classes = 8
    elif opt.dataset == "imagenet10":
        opt.input_height = opt.input_width = 224
        opt.input_channel = 3
        opt.num_classes = 10
        opt.bs = 32
    else:
        raise Exception("Invalid Dataset")


def main(argv=None, train_fn=None, eval_fn=None, get_model_fn=None):
    """(train_fn / eval_fn / get_model_fn: the variants -- train_generator_imperceptible.py, train_generator_wanet.py -- run this
    driver with their own train / eval / get_model)
    train_generator.py:466-609: dataset shape, loaders, get_model, detector / clean-model checkpoints, --continue_training
    resume, then n_iters epochs of train() + eval().  Build-only flag --synthetic_data replaces the dataset (there is no
    network here for torchvision's download) and makes the two pretrained checkpoints optional; everything else -- paths,
    checkpoint dict keys, prints -- is the reference's."""
    import shutil
    opt = config.get_arguments().parse_args(argv)
    _dataset_shape(opt)
    from .utils.dataloader import get_dataloader
    train_dl = get_dataloader(opt, True)
    test_dl = get_dataloader(opt, False)
    netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model = (get_model_fn or get_model)(opt)

    mode = opt.saving_prefix
    opt.ckpt_folder = os.path.join(opt.checkpoints, "{}_clean".format(mode), opt.dataset)
    opt.ckpt_path = os.path.join(opt.ckpt_folder, "{}_{}_clean.pth.tar".format(opt.dataset, mode))
    opt.log_dir = os.path.join(opt.ckpt_folder, "log_dir")
    os.makedirs(opt.log_dir, exist_ok=True)

    # pretrained frequency detector (:503-511)
    if netF is not None:
        opt.F_ckpt_folder = os.path.join(opt.F_checkpoints, opt.dataset)
        opt.F_ckpt_path = os.path.join(opt.F_ckpt_folder, opt.F_model, "{}_{}_detector.pth.tar".format(opt.dataset, opt.F_model))
        if os.path.exists(opt.F_ckpt_path) or not opt.synthetic_data:
            print(f"Loading {opt.F_model} at {opt.F_ckpt_path}")
            netF.load_state_dict(torch.load(opt.F_ckpt_path, map_location=opt.device, weights_only=False)["netC"])
            print("Done")
        netF.eval()
    # pretrained clean model (:513-527)
    if opt.load_checkpoint_clean is not None or not opt.synthetic_data:
        load_path = os.path.join(opt.checkpoints, str(opt.load_checkpoint_clean), opt.dataset,
                                 "{}_{}.pth.tar".format(opt.dataset, opt.load_checkpoint_clean))
        if not os.path.exists(load_path):
            print("Error: {} not found".format(load_path))
            sys.exit()
        clean_model.load_state_dict(torch.load(load_path, map_location=opt.device, weights_only=False)["netC"])
    clean_model.eval()

    bests = [0.0] * 6
    epoch_current = 0
    if opt.continue_training:                                                        # :529-552
        if not os.path.exists(opt.ckpt_path):
            print("Pretrained model doesnt exist")
            sys.exit()
        print("Continue training!!")
        sd = torch.load(opt.ckpt_path, map_location=opt.device, weights_only=False)  # our own checkpoint (holds numpy scalars, as the reference's)
        netC.load_state_dict(sd["netC"])
        optimizerC.load_state_dict(sd["optimizerC"])
        schedulerC.load_state_dict(sd["schedulerC"])
        netG.load_state_dict(sd["netG"])
        optimizerG.load_state_dict(sd["optimizerG"])
        schedulerG.load_state_dict(sd["schedulerG"])
        clean_model.load_state_dict(sd["clean_model"])
        bests = [sd[k] for k in ("best_clean_acc", "best_bd_acc", "best_F_acc", "best_clean_model_acc", "best_clean_model_bd_ba",
                                 "best_clean_model_bd_asr")]
        epoch_current = sd["epoch_current"]
    else:
        print("Train from scratch!!!")
        shutil.rmtree(opt.ckpt_folder, ignore_errors=True)
        os.makedirs(opt.log_dir, exist_ok=True)
    try:
        from torch.utils.tensorboard import SummaryWriter
        tf_writer = SummaryWriter(log_dir=opt.log_dir)
    except Exception:   # tensorboard is a logging nicety, not part of the path
        tf_writer = _NullWriter()
    for epoch in range(epoch_current, opt.n_iters):
        print("Epoch {}:".format(epoch + 1))
        (train_fn or train)(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, train_dl, tf_writer,
                            epoch, opt)
        bests = list((eval_fn or eval)(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, test_dl, *bests,
                                       tf_writer, epoch, opt))
    return bests


if __name__ == "__main__":
    main()
