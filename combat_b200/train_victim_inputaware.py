"""train_victim_inputaware.py of the reference (:66-330): the victim trainer of train_victim.py (same train(): C-step half of the
alternated step on a clean-label poisoned dataset with the frozen generator; `gauss_smooth` fixed at T.GaussianBlur(3, (0.1, 1)),
:37) whose eval() takes a SECOND test loader and adds the cross-trigger accuracy -- netC on `inputs + trigger of inputs2`, counted
on the non-target rows against their TRUE labels (:213-223) -- and `best_cross_acc` to the returned bests and the checkpoint dict.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import config, ops
from . import train_victim as _base
from .engine import create_targets_bd_np
from .train_generator import _dataset_shape, create_targets_bd, low_freq  # noqa: F401
from .train_victim import get_model  # noqa: F401  (:66-90, 'default' classifiers)


def _variant(opt):
    opt.kernel_size, opt.sigma = 3, (0.1, 1.0)   # module-level gauss_smooth (:37)
    return opt


def train(netC, optimizerC, schedulerC, netG, train_dl, tf_writer, epoch, opt):
    """train_victim_inputaware.py:93-159 (the base victim iteration)."""
    return _base.train(netC, optimizerC, schedulerC, netG, train_dl, tf_writer, epoch, _variant(opt))


def eval_batch(netC, netG, inputs, inputs2, targets, opt, sigma=None, sigma2=None, counts=None):
    """One iteration of :187-223.  Returns (device int32 counts [clean, -, bd, -, cross, -], n_bd, debug tensors).  Fixed-shape
    batch: the trigger is built for every row and the target rows are masked with a negative label (see combat_b200/eval.py)."""
    C_, G_ = netC.net, netG.net
    dev = C_.device
    y = targets.cpu().numpy().astype(np.int64) if torch.is_tensor(targets) else np.asarray(targets, dtype=np.int64)
    if sigma is None:
        sigma = torch.empty(1).uniform_(opt.sigma[0], opt.sigma[1]).item()       # inputs_bd (:205)
    if sigma2 is None:
        sigma2 = torch.empty(1).uniform_(opt.sigma[0], opt.sigma[1]).item()      # inputs_bd2 (:216)
    ntrg = y != opt.target_label
    bd = create_targets_bd_np(y, opt)
    t = torch.from_numpy(np.stack([y, np.where(ntrg, bd, -1), np.where(ntrg, y, -1)])).to(dev, non_blocking=True)
    x = inputs.to(dev, non_blocking=True).float().contiguous()
    x2 = inputs2.to(dev, non_blocking=True).float().contiguous()
    B = x.shape[0]
    keep = int(opt.input_height * opt.ratio)
    if counts is None:
        counts = torch.zeros(6, dtype=torch.int32, device=dev)
    preds_clean, _ = C_.forward(x, train=False, save=False)                                      # :194
    ops.cross_entropy(preds_clean, t[0], 1.0, False, counts_out=counts[0:2])
    noise_raw, _ = G_.forward(torch.cat([x, x2]), None, save=False)                              # :203 and :214, one launch set
    noise = ops.plane_op(noise_raw, "lowfreq", keep=keep)
    x_bd = ops.poison_blend_fwd(x, noise[:B], None, B, opt.noise_rate, ops.gaussian_taps(sigma))       # :205
    x_bd2 = ops.poison_blend_fwd(x, noise[B:], None, B, opt.noise_rate, ops.gaussian_taps(sigma2))     # :216
    preds, _ = C_.forward(torch.cat([x_bd, x_bd2]), train=False, save=False)                     # :207 and :217
    preds_bd, preds_cross = preds[:B], preds[B:]
    ops.cross_entropy(preds_bd, t[1], 1.0, False, counts_out=counts[2:4])
    ops.cross_entropy(preds_cross, t[2], 1.0, False, counts_out=counts[4:6])                     # :220-223
    return counts, int(ntrg.sum()), dict(preds_clean=preds_clean, preds_bd=preds_bd, preds_cross=preds_cross, x_bd=x_bd,
                                         x_bd2=x_bd2, sigma=sigma, sigma2=sigma2)


def eval(netC, optimizerC, schedulerC, netG, test_dl, test_dl2, best_clean_acc, best_bd_acc, best_cross_acc, tf_writer, epoch, opt):
    """train_victim_inputaware.py:162-254."""
    print(" Eval:")
    opt = _variant(opt)
    netC.eval()
    dev = netC.net.device
    tot = torch.zeros(6, dtype=torch.int64, device=dev)
    n_clean = n_bd = 0
    for _, batch1, batch2 in zip(range(len(test_dl)), test_dl, test_dl2):
        counts, nb, _ = eval_batch(netC, netG, batch1[0], batch2[0], batch1[1], opt)
        tot += counts.long()
        n_clean += len(batch1[1])
        n_bd += nb
    c = tot.cpu().numpy()
    acc_clean, acc_bd, acc_cross = c[0] * 100.0 / max(n_clean, 1), c[2] * 100.0 / max(n_bd, 1), c[4] * 100.0 / max(n_bd, 1)
    print("Clean Acc: {:.4f} - Best: {:.4f} | Bd Acc: {:.4f} - Best: {:.4f} | Cross Acc: {:.4f} - Best: {:.4f}".format(
        acc_clean, best_clean_acc, acc_bd, best_bd_acc, acc_cross, best_cross_acc))
    if not epoch % 1:
        tf_writer.add_scalars("Test Accuracy", {"Clean": acc_clean, "Bd": acc_bd, "Cross": acc_cross}, epoch)
    if acc_clean > best_clean_acc:
        print(" Saving...")
        best_clean_acc, best_bd_acc, best_cross_acc = acc_clean, acc_bd, acc_cross
        state_dict = {"netC": netC.state_dict(), "schedulerC": schedulerC.state_dict(), "optimizerC": optimizerC.state_dict(),
                      "netG": netG.state_dict(), "best_clean_acc": acc_clean, "best_bd_acc": acc_bd, "best_cross_acc": acc_cross,
                      "epoch_current": epoch}
        d = os.path.dirname(opt.ckpt_path)
        if d:
            os.makedirs(d, exist_ok=True)
        torch.save(state_dict, opt.ckpt_path)
    return best_clean_acc, best_bd_acc, best_cross_acc


def main(argv=None):
    """train_victim_inputaware.py:257-330 (no --continue_training branch in this variant)."""
    import shutil
    from .utils.dataloader_cleanbd import get_dataloader
    opt = config.get_arguments().parse_args(argv)
    _dataset_shape(opt)
    opt.num_workers = 0                                                                        # :279
    train_dl, test_dl, test_dl2 = get_dataloader(opt, True), get_dataloader(opt, False), get_dataloader(opt, False)
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
    bests = (0.0, 0.0, 0.0)
    for epoch in range(opt.n_iters):
        print("Epoch {}:".format(epoch + 1))
        train(netC, optimizerC, schedulerC, netG, train_dl, tf_writer, epoch, opt)
        bests = eval(netC, optimizerC, schedulerC, netG, test_dl, test_dl2, *bests, tf_writer, epoch, opt)
    return bests


if __name__ == "__main__":
    main()
