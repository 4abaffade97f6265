"""train_victim_wanet.py of the reference (:37-285): the victim trainer of train_victim.py with the WaNet warp trigger -- the
poisoned rows (the dataset's flags) are warped by the frozen GridGenerator's flow instead of receiving an additive trigger
(:86-97), evaluation likewise (:164-172), `grid_rescale` is added to the checkpoint dict (:204), no --continue_training branch.
`identity_grid` keeps its place in the signatures (the kernels rebuild it from torch.linspace(-1, 1, input_height), :261-263).
As in train_victim.py, `(poisoned is False).nonzero()` (:84) raises as shipped; the rows whose flag is False are taken."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import config, ops
from . import train_victim as _base
from .engine import create_targets_bd_np
from .modules import GridGenerator, PreActResNet18, ResNet18
from .train_generator import _dataset_shape, _dtype, create_targets_bd  # noqa: F401
from .train_generator_wanet import _check_grid, _variant


def get_model(opt):
    """train_victim_wanet.py:37-55"""
    kw = dict(device=opt.device, dtype=_dtype(opt))
    if opt.dataset == "cifar10":
        netC = PreActResNet18(**kw)
    elif opt.dataset == "celeba":
        netC = ResNet18(num_classes=opt.num_classes, **kw)
    elif opt.dataset == "imagenet10":
        netC = ResNet18(num_classes=opt.num_classes, input_size=opt.input_height, **kw)
    else:
        raise Exception("Invalid Dataset")
    netG = GridGenerator(opt, **kw)
    optimizerC = torch.optim.SGD(netC.parameters(), opt.lr_C, momentum=0.9, weight_decay=5e-4, nesterov=True)
    schedulerC = torch.optim.lr_scheduler.MultiStepLR(optimizerC, opt.schedulerC_milestones, opt.schedulerC_lambda)
    return netC, optimizerC, schedulerC, netG


def train(netC, optimizerC, schedulerC, netG, train_dl, identity_grid, tf_writer, epoch, opt):
    """train_victim_wanet.py:58-135; train_dl yields (inputs, targets, poisoned)."""
    _check_grid(identity_grid, opt)
    return _base.train(netC, optimizerC, schedulerC, netG, train_dl, tf_writer, epoch, _variant(opt))


def eval_batch(netC, netG, inputs, targets, opt, counts=None):
    """One iteration of :153-181.  Returns (device int32 counts [clean, -, bd, -], n_bd, debug tensors); fixed-shape batch, the
    target rows masked with a negative label (see combat_b200/eval.py); nothing is drawn (no blur)."""
    C_, G_ = netC.net, netG.net
    dev = C_.device
    y = targets.cpu().numpy().astype(np.int64) if torch.is_tensor(targets) else np.asarray(targets, dtype=np.int64)
    ntrg = y != opt.target_label
    bd = create_targets_bd_np(y, opt)
    t = torch.from_numpy(np.stack([y, np.where(ntrg, bd, -1)])).to(dev, non_blocking=True)
    x = inputs.to(dev, non_blocking=True).float().contiguous()
    B = x.shape[0]
    if counts is None:
        counts = torch.zeros(4, dtype=torch.int32, device=dev)
    preds_clean, _ = C_.forward(x, train=False, save=False)                                      # :156
    ops.cross_entropy(preds_clean, t[0], 1.0, False, counts_out=counts[0:2])
    flow, _ = G_.forward(x, None, save=False)                                                    # :165
    ident = torch.linspace(-1, 1, steps=opt.input_height).to(dev)
    x_bd = ops.wanet_warp_fwd(x, flow, ident, None, B, opt.grid_rescale, opt.s)                  # :166-171
    preds_bd, _ = C_.forward(x_bd, train=False, save=False)                                      # :173
    ops.cross_entropy(preds_bd, t[1], 1.0, False, counts_out=counts[2:4])
    return counts, int(ntrg.sum()), dict(preds_clean=preds_clean, preds_bd=preds_bd, x_bd=x_bd, flow=flow)


def eval(netC, optimizerC, schedulerC, netG, test_dl, identity_grid, best_clean_acc, best_bd_acc, tf_writer, epoch, opt):
    """train_victim_wanet.py:138-208"""
    print(" Eval:")
    _check_grid(identity_grid, opt)
    netC.eval()
    dev = netC.net.device
    tot = torch.zeros(4, dtype=torch.int64, device=dev)
    n_clean = n_bd = 0
    for batch in test_dl:
        counts, nb, _ = eval_batch(netC, netG, batch[0], batch[1], opt)
        tot += counts.long()
        n_clean += len(batch[1])
        n_bd += nb
    c = tot.cpu().numpy()
    acc_clean, acc_bd = c[0] * 100.0 / max(n_clean, 1), c[2] * 100.0 / max(n_bd, 1)
    print("Clean Acc: {:.4f} - Best: {:.4f} | Bd Acc: {:.4f} - Best: {:.4f}".format(acc_clean, best_clean_acc, acc_bd, best_bd_acc))
    if not epoch % 1:
        tf_writer.add_scalars("Test Accuracy", {"Clean": acc_clean, "Bd": acc_bd}, epoch)
    if acc_clean > best_clean_acc:
        print(" Saving...")
        best_clean_acc, best_bd_acc = acc_clean, acc_bd
        state_dict = {"netC": netC.state_dict(), "schedulerC": schedulerC.state_dict(), "optimizerC": optimizerC.state_dict(),
                      "netG": netG.state_dict(), "best_clean_acc": acc_clean, "best_bd_acc": acc_bd, "epoch_current": epoch,
                      "grid_rescale": opt.grid_rescale}
        d = os.path.dirname(opt.ckpt_path)
        if d:
            os.makedirs(d, exist_ok=True)
        torch.save(state_dict, opt.ckpt_path)
    return best_clean_acc, best_bd_acc


def main(argv=None):
    """train_victim_wanet.py:211-285"""
    import shutil
    from .utils.dataloader_cleanbd import get_dataloader
    opt = config.get_arguments().parse_args(argv)
    _dataset_shape(opt)
    train_dl, test_dl = get_dataloader(opt, True), get_dataloader(opt, False)
    netC, optimizerC, schedulerC, netG = get_model(opt)
    mode = opt.saving_prefix
    opt.ckpt_folder = os.path.join(opt.checkpoints, "{}_clean".format(mode), opt.dataset)
    opt.ckpt_path = os.path.join(opt.ckpt_folder, "{}_{}_clean.pth.tar".format(opt.dataset, mode))
    opt.log_dir = os.path.join(opt.ckpt_folder, "log_dir")
    shutil.rmtree(opt.ckpt_folder, ignore_errors=True)
    os.makedirs(opt.log_dir, exist_ok=True)
    load_path = os.path.join(opt.checkpoints, opt.load_checkpoint, opt.dataset, "{}_{}.pth.tar".format(opt.dataset, opt.load_checkpoint))
    if os.path.exists(load_path):
        netG.load_state_dict(torch.load(load_path, map_location=opt.device, weights_only=False)["netG"])
    elif not opt.synthetic_data:
        print("Error: {} not found".format(load_path))
        sys.exit()
    netG.eval()
    try:
        from torch.utils.tensorboard import SummaryWriter
        tf_writer = SummaryWriter(log_dir=opt.log_dir)
    except Exception:
        tf_writer = _base._NullWriter()
    a = torch.linspace(-1, 1, steps=opt.input_height)                                           # :261-263
    gx, gy = torch.meshgrid(a, a, indexing="ij")
    identity_grid = torch.stack((gy, gx), 2)[None, ...].to(opt.device)
    best_clean_acc = best_bd_acc = 0.0
    for epoch in range(opt.n_iters):
        print("Epoch {}:".format(epoch + 1))
        train(netC, optimizerC, schedulerC, netG, train_dl, identity_grid, tf_writer, epoch, opt)
        best_clean_acc, best_bd_acc = eval(netC, optimizerC, schedulerC, netG, test_dl, identity_grid, best_clean_acc, best_bd_acc,
                                           tf_writer, epoch, opt)
    return best_clean_acc, best_bd_acc


if __name__ == "__main__":
    main()
