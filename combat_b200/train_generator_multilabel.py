"""train_generator_multilabel.py of the reference: get_model, train, eval, main (reference :78-137, :142-318, :320-455, :457-614).
train_victim_multilabel.py of the reference is the same file up to two comments (`diff`), see combat_b200/train_victim_multilabel.py.

Differences from train_generator.py that the engine's multilabel mode implements (engine.AlternatedStep(multilabel=True)):
conditional generator `CUnetGeneratorv1(x, y)`; the C-step poisons the FIRST num_bd rows (num_bd from `np.random.rand(bs)`,
:171) with the trigger conditioned on the true labels, labels unchanged; the G-step splits the batch into num_classes
contiguous chunks, chunk ci is pushed towards class ci with its own blur sigma (:203-221); netG's optimiser runs at
`lr_C * 0.1` (:115).  `train()` keeps the reference's signature, including the unused `mask` / `pattern` arguments.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops
from . import train_generator as _base
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


def eval_batch(netC, clean_model, netG, netF, inputs, targets, opt, sigmas=None):
    """One iteration of reference :343-378.  Returns (device int32 counts [num_classes + 1, 8], n_bd, debug): row 0 = clean
    accuracy of netC [0] and clean_model [2]; row 1 + ci = class ci: attack success [0], clean-model BA [2] / ASR [3], detector
    [4].  Fixed-shape batches: the rows whose label IS ci are masked out of the counters with a negative label.  One blur sigma
    per class from the torch CPU generator (module-level GaussianBlur(3, (0.1, 1)), :53)."""
    C_, K_, G_ = netC.net, clean_model.net, netG.net
    F_ = netF.net if netF is not None else None
    dev = C_.device
    nc = opt.num_classes
    y = targets.cpu().numpy().astype(np.int64) if torch.is_tensor(targets) else np.asarray(targets, dtype=np.int64)
    B = len(y)
    x = inputs.to(dev, non_blocking=True).float().contiguous()
    rows = [y]
    for ci in range(nc):
        keep = y != ci
        rows += [np.full(B, ci, dtype=np.int64), np.where(keep, ci, -1), np.where(keep, y, -1)]   # labels | bd masked | y masked
    t = torch.from_numpy(np.stack(rows)).to(dev, non_blocking=True)
    ones = torch.ones(B, dtype=torch.int64, device=dev)
    counts = torch.zeros((nc + 1, 8), dtype=torch.int32, device=dev)
    keep_px = int(opt.input_height * opt.ratio)
    preds_clean, _ = C_.forward(x, train=False, save=False)                                      # :347
    ops.cross_entropy(preds_clean, t[0], 1.0, False, counts_out=counts[0, 0:2])
    cm_clean, _ = K_.forward(x, train=False, save=False)                                         # :351
    ops.cross_entropy(cm_clean, t[0], 1.0, False, counts_out=counts[0, 2:4])
    dbg = dict(sigmas=[], x_bd=[], preds_bd=[])
    n_bd = 0
    for ci in range(nc):                                                                         # :355-378
        sigma = sigmas[ci] if sigmas is not None else torch.empty(1).uniform_(0.1, 1.0).item()
        noise_raw, _ = G_.forward(x, t[1 + 3 * ci], save=False)
        noise = ops.plane_op(noise_raw, "lowfreq", keep=keep_px)
        x_bd = ops.poison_blend_fwd(x, noise, None, B, opt.noise_rate, ops.gaussian_taps(sigma))
        preds_bd, _ = C_.forward(x_bd, train=False, save=False)
        ops.cross_entropy(preds_bd, t[2 + 3 * ci], 1.0, False, counts_out=counts[1 + ci, 0:2])
        cm_bd, _ = K_.forward(x_bd, train=False, save=False)
        ops.cross_entropy(cm_bd, t[3 + 3 * ci], 1.0, False, targets2=t[2 + 3 * ci], counts_out=counts[1 + ci, 2:4])
        if F_ is not None:
            preds_F = F_.forward(ops.plane_op(x_bd, "dct", in_mode=2))
            ops.cross_entropy(preds_F, ones, 1.0, False, counts_out=counts[1 + ci, 4:6])
        n_bd += int((y != ci).sum())
        dbg["sigmas"].append(sigma)
        dbg["x_bd"].append(x_bd)
        dbg["preds_bd"].append(preds_bd)
    return counts, n_bd, dbg


def eval(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, test_dl, mask, pattern, best_clean_acc,
         best_bd_acc, best_F_acc, best_clean_model_acc, best_clean_model_bd_ba, best_clean_model_bd_asr, tf_writer, epoch, opt):
    """reference :320-455: every class in turn is the attack target; saves the checkpoint dict (+ mask, pattern) on improvement."""
    print(" Eval:")
    netC.eval()
    dev = netC.net.device
    nc = opt.num_classes
    tot = torch.zeros((nc + 1, 8), dtype=torch.int64, device=dev)
    n_clean = n_bd = 0
    for inputs, targets in test_dl:
        if not inputs.is_cuda and not inputs.is_pinned():
            inputs = inputs.pin_memory()
        counts, nb, _ = eval_batch(netC, clean_model, netG, netF, inputs, targets, opt)
        tot += counts.long()
        n_clean += len(targets)
        n_bd += nb
    c = tot.cpu().numpy()
    n_bd_ = max(n_bd, 1)
    acc_clean, acc_clean_model = c[0, 0] * 100.0 / n_clean, c[0, 2] * 100.0 / n_clean
    acc_bd = c[1:, 0].sum() * 100.0 / n_bd_
    acc_F = c[1:, 4].sum() * 100.0 / (n_clean * nc)                                              # :381
    bd_ba_clean_model, bd_asr_clean_model = c[1:, 2].sum() * 100.0 / n_bd_, c[1:, 3].sum() * 100.0 / n_bd_
    print("Clean Acc: {:.4f} - Best: {:.4f} | Bd Acc: {:.4f} - Best: {:.4f} | F Acc: {:.4f} - Best: {:.4f} | Clean Model Acc: {:.4f} - "
          "Best: {:.4f} | Clean Model Bd BA: {:.4f} - Best: {:.4f} | Clean Model Bd ASR: {:.4f} - Best: {:.4f}".format(
              acc_clean, best_clean_acc, acc_bd, best_bd_acc, acc_F, best_F_acc, acc_clean_model, best_clean_model_acc,
              bd_ba_clean_model, best_clean_model_bd_ba, bd_asr_clean_model, best_clean_model_bd_asr))
    if not epoch % 1:
        tf_writer.add_scalars("Test Accuracy", {"Clean": acc_clean, "Bd": acc_bd, "F": acc_F, "Clean Model Acc": acc_clean_model,
                                                "Clean Model Bd BA": bd_ba_clean_model, "Clean Model Bd ASR": bd_asr_clean_model}, epoch)
    if acc_clean > best_clean_acc or (acc_clean == best_clean_acc and acc_bd > best_bd_acc):     # :420
        print(" Saving...")
        best_clean_acc, best_bd_acc, best_F_acc = acc_clean, acc_bd, acc_F
        best_clean_model_acc, best_clean_model_bd_ba, best_clean_model_bd_asr = acc_clean_model, bd_ba_clean_model, bd_asr_clean_model
        state_dict = {
            "netC": netC.state_dict(), "schedulerC": schedulerC.state_dict(), "optimizerC": optimizerC.state_dict(),
            "netG": netG.state_dict(), "schedulerG": schedulerG.state_dict(), "optimizerG": optimizerG.state_dict(),
            "clean_model": clean_model.state_dict(), "best_clean_acc": acc_clean, "best_bd_acc": acc_bd, "best_F_acc": acc_F,
            "best_clean_model_acc": best_clean_model_acc, "best_clean_model_bd_ba": best_clean_model_bd_ba,
            "best_clean_model_bd_asr": best_clean_model_bd_asr, "epoch_current": epoch, "mask": mask, "pattern": pattern,
        }
        d = os.path.dirname(opt.ckpt_path)
        if d:
            os.makedirs(d, exist_ok=True)
        torch.save(state_dict, opt.ckpt_path)
    return (best_clean_acc, best_bd_acc, best_F_acc, best_clean_model_acc, best_clean_model_bd_ba, best_clean_model_bd_asr)


def main(argv=None):
    """reference :457-614: the base driver with this module's get_model / train / eval; `mask` / `pattern` (unused by the step,
    :573-576) are created on the first epoch, or restored from the checkpoint under --continue_training (:559-560)."""
    extra = {}

    def mp(opt):
        if not extra:
            if opt.continue_training and os.path.exists(opt.ckpt_path):
                sd = torch.load(opt.ckpt_path, map_location=opt.device, weights_only=False)
                extra.update(mask=sd["mask"], pattern=sd["pattern"])
            else:
                m = torch.zeros(opt.input_height, opt.input_width).to(opt.device)
                m[2:6, 2:6] = 0.1
                extra.update(mask=m, pattern=torch.rand(opt.input_channel, opt.input_height, opt.input_width).to(opt.device))
        return extra["mask"], extra["pattern"]

    def with_mp(fn):
        def run(*a):
            return fn(*a[:9], *mp(a[-1]), *a[9:])
        return run
    return _base.main(argv, train_fn=with_mp(train), eval_fn=with_mp(eval), get_model_fn=get_model)


if __name__ == "__main__":
    main()
