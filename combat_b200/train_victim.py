"""train_victim.py of the reference (:63-317): train a victim classifier on a clean-label poisoned dataset with the FROZEN
trigger generator -- get_model(opt) -> (netC, optimizerC, schedulerC, netG), train(...), eval(...), main().

Every iteration is the C-step half of the alternated step on the same kernels (engine.AlternatedStep.victim_step): the
poisoned rows are the dataset's per-sample flags (utils/dataloader_cleanbd.py PoisonedDataset), their triggers come from
netG + low_freq + clamp + GaussianBlur (:123-129), the batch is re-ordered [poisoned ; rest] (:130), PostTensorTransform
(:131), netC train-mode forward/backward, SGD -- one CUDA-graph replay.
As shipped, :121 `ntrg_ind = (poisoned is False).nonzero()` raises AttributeError (a Python bool has no nonzero); the evident
intent (rows whose flag is False) is implemented, and stated here."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import config
from .engine import N_LOSSES, AlternatedStep, make_plan_victim
from .eval import eval_batch
from .modules import PreActResNet18, ResNet18, UnetGenerator
from .train_generator import _adopt_momentum, _bind_momentum, _dataset_shape, _dtype, _HOT_SCALARS, create_targets_bd, low_freq  # noqa: F401


def get_model(opt):
    """train_victim.py:63-91 ('default' classifiers)."""
    kw = dict(device=opt.device, dtype=_dtype(opt))
    if opt.dataset == "cifar10":
        netC = PreActResNet18(**kw)
    elif opt.dataset == "celeba":
        netC = ResNet18(num_classes=opt.num_classes, **kw)
    elif opt.dataset == "imagenet10":
        netC = ResNet18(num_classes=opt.num_classes, n_input=opt.input_channel, input_size=opt.input_height, **kw)
    else:
        raise Exception("Invalid Dataset")
    netG = UnetGenerator(opt, **kw)
    if opt.model != "default":
        raise NotImplementedError("--model %s is outside the built hot path" % opt.model)
    optimizerC = torch.optim.SGD(netC.parameters(), opt.lr_C, momentum=0.9, weight_decay=5e-4, nesterov=True)
    schedulerC = torch.optim.lr_scheduler.MultiStepLR(optimizerC, opt.schedulerC_milestones, opt.schedulerC_lambda)
    return netC, optimizerC, schedulerC, netG


def _victim_engine(netC, netG, opt):
    sig = tuple(getattr(opt, k, None) for k in _HOT_SCALARS)
    rec = getattr(netC, "_combat_victim_engine", None)
    if rec is not None and rec[1] is netG and rec[2] == sig:
        rec[0].opt = opt
        return rec[0]
    eng = AlternatedStep(opt, device=netC.net.device, with_metrics=False,
                         nets=(netC.net, None, netG.net if netG is not None else None, None))
    object.__setattr__(netC, "_combat_victim_engine", (eng, netG, sig))
    return eng


def _train_epoch(netC, optimizerC, schedulerC, netG, train_dl, tf_writer, epoch, opt, with_flags):
    print(" Train:")
    netC.train()
    eng = _victim_engine(netC, netG, opt)
    _adopt_momentum(optimizerC, netC)
    for pg in optimizerC.param_groups:
        if not (pg["momentum"] == 0.9 and pg["weight_decay"] == 5e-4 and pg["nesterov"]):
            raise NotImplementedError("the fused optimiser implements the reference's SGD(0.9, 5e-4, nesterov) only")
    eng.set_lr(optimizerC.param_groups[0]["lr"])
    use_graph = not getattr(opt, "no_graph", False)
    log_every = max(1, int(getattr(opt, "log_every", 50)))
    dev = netC.net.device
    tot = torch.zeros(16, dtype=torch.int64, device=dev)
    lsum = torch.zeros(N_LOSSES, dtype=torch.float64, device=dev)
    total_sample, n_batches = 0, len(train_dl)
    avg_acc_clean = avg_loss_ce = 0.0
    for batch_idx, batch in enumerate(train_dl):
        inputs, targets = batch[0], batch[1]
        poisoned = batch[2] if with_flags else None
        y_host = targets.cpu().numpy() if torch.is_tensor(targets) else np.asarray(targets)
        pz = None if poisoned is None else (poisoned.cpu().numpy() if torch.is_tensor(poisoned) else np.asarray(poisoned))
        plan = make_plan_victim(y_host, pz, opt)
        if not inputs.is_cuda and not inputs.is_pinned():
            inputs = inputs.pin_memory()
        out = eng.victim_step(inputs, y_host, pz, plan, use_graph=use_graph)
        tot += out["counts"].long()
        lsum += out["losses"].double()
        total_sample += len(y_host)
        if (batch_idx + 1) % log_every == 0 or batch_idx + 1 == n_batches:
            avg_acc_clean = float(tot[0]) * 100.0 / total_sample
            avg_loss_ce = float(lsum[0]) / total_sample
            print("[%d/%d] CE Loss: %.4f | Clean Acc: %.4f" % (batch_idx + 1, n_batches, avg_loss_ce, avg_acc_clean))
    if total_sample and not epoch % 1:
        tf_writer.add_scalars("Clean Accuracy", {"Clean": avg_acc_clean}, epoch)
        if not with_flags and hasattr(tf_writer, "add_scalar"):
            tf_writer.add_scalar("CE Loss", avg_loss_ce, epoch)                  # train_clean_classifier.py:118
    _bind_momentum(optimizerC, netC)
    for n, b in netC.named_buffers():
        if n.endswith("num_batches_tracked"):
            b.fill_(netC.net.num_batches_tracked[n[: -len(".num_batches_tracked")]])
    schedulerC.step()


def train(netC, optimizerC, schedulerC, netG, train_dl, tf_writer, epoch, opt):
    """train_victim.py:94-165; train_dl yields (inputs, targets, poisoned)."""
    _train_epoch(netC, optimizerC, schedulerC, netG, train_dl, tf_writer, epoch, opt, True)


def eval(netC, optimizerC, schedulerC, netG, test_dl, best_clean_acc, best_bd_acc, tf_writer, epoch, opt):
    """train_victim.py:168-226: clean accuracy, attack success on the non-target samples; checkpoint on improvement."""
    print(" Eval:")
    netC.eval()
    dev = netC.net.device
    tot = torch.zeros(4, dtype=torch.int64, device=dev)
    n_clean = n_bd = 0
    for batch in test_dl:
        inputs, targets = batch[0], batch[1]
        counts, nb, _ = eval_batch(netC, netG, inputs, targets, opt)
        tot += counts.long()
        n_clean += len(targets)
        n_bd += nb
    c = tot.cpu().numpy()
    acc_clean, acc_bd = c[0] * 100.0 / max(n_clean, 1), c[3] * 100.0 / max(n_bd, 1)
    print("Clean Acc: {:.4f} - Best: {:.4f} | Bd Acc: {:.4f} - Best: {:.4f}".format(acc_clean, best_clean_acc, acc_bd, best_bd_acc))
    if not epoch % 1:
        tf_writer.add_scalars("Test Accuracy", {"Clean": acc_clean, "Bd": acc_bd}, epoch)
    if acc_clean > best_clean_acc:
        print(" Saving...")
        best_clean_acc, best_bd_acc = acc_clean, acc_bd
        state_dict = {"netC": netC.state_dict(), "schedulerC": schedulerC.state_dict(), "optimizerC": optimizerC.state_dict(),
                      "netG": netG.state_dict(), "best_clean_acc": acc_clean, "best_bd_acc": acc_bd, "epoch_current": epoch}
        d = os.path.dirname(opt.ckpt_path)
        if d:
            os.makedirs(d, exist_ok=True)
        torch.save(state_dict, opt.ckpt_path)
    return best_clean_acc, best_bd_acc


class _NullWriter:
    def add_scalars(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass


def main(argv=None):
    """train_victim.py:229-317."""
    import shutil
    from .utils.dataloader_cleanbd import get_dataloader
    opt = config.get_arguments().parse_args(argv)
    _dataset_shape(opt)
    if opt.dataset == "imagenet10":
        opt.num_workers = 40
    train_dl, test_dl = get_dataloader(opt, True), get_dataloader(opt, False)
    netC, optimizerC, schedulerC, netG = get_model(opt)
    mode = opt.saving_prefix
    opt.ckpt_folder = os.path.join(opt.checkpoints, "{}_clean".format(mode), opt.dataset)
    opt.ckpt_path = os.path.join(opt.ckpt_folder, "{}_{}_clean.pth.tar".format(opt.dataset, mode))
    opt.log_dir = os.path.join(opt.ckpt_folder, "log_dir")
    os.makedirs(opt.log_dir, exist_ok=True)
    load_path = os.path.join(opt.checkpoints, opt.load_checkpoint, opt.dataset, "{}_{}.pth.tar".format(opt.dataset, opt.load_checkpoint))
    if os.path.exists(load_path):
        netG.load_state_dict(torch.load(load_path, map_location=opt.device, weights_only=False)["netG"])
    elif not opt.synthetic_data:
        print("Error: {} not found".format(load_path))
        sys.exit()
    netG.eval()
    best_clean_acc = best_bd_acc = 0.0
    epoch_current = 0
    if opt.continue_training:
        if not os.path.exists(opt.ckpt_path):
            print("Pretrained model doesnt exist")
            sys.exit()
        print("Continue training!!")
        sd = torch.load(opt.ckpt_path, map_location=opt.device, weights_only=False)
        netC.load_state_dict(sd["netC"])
        optimizerC.load_state_dict(sd["optimizerC"])
        schedulerC.load_state_dict(sd["schedulerC"])
        best_clean_acc, best_bd_acc, epoch_current = sd["best_clean_acc"], sd["best_bd_acc"], sd["epoch_current"]
    else:
        print("Train from scratch!!!")
        shutil.rmtree(opt.ckpt_folder, ignore_errors=True)
        os.makedirs(opt.log_dir, exist_ok=True)
    try:
        from torch.utils.tensorboard import SummaryWriter
        tf_writer = SummaryWriter(log_dir=opt.log_dir)
    except Exception:
        tf_writer = _NullWriter()
    for epoch in range(epoch_current, opt.n_iters):
        print("Epoch {}:".format(epoch + 1))
        train(netC, optimizerC, schedulerC, netG, train_dl, tf_writer, epoch, opt)
        best_clean_acc, best_bd_acc = eval(netC, optimizerC, schedulerC, netG, test_dl, best_clean_acc, best_bd_acc, tf_writer,
                                           epoch, opt)
    return best_clean_acc, best_bd_acc


if __name__ == "__main__":
    main()
