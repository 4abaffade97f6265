"""train_clean_classifier.py of the reference (:38-236): the clean classifier that train_generator.py loads as `clean_model`
(--load_checkpoint_clean).  get_model(opt) -> (netC, optimizerC, schedulerC); train(netC, optimizerC, schedulerC, train_dl,
tf_writer, epoch, opt): PostTensorTransform, netC train-mode forward/backward, SGD per batch -- the same captured graph as the
victim trainer with no poisoned rows and no generator; eval(...) -> clean accuracy + the checkpoint that train_generator.py's
main() reads (`<checkpoints>/<saving_prefix>/<dataset>/<dataset>_<saving_prefix>.pth.tar`, key "netC"); main()."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import config, ops
from .modules import PreActResNet18, ResNet18
from .train_generator import _dataset_shape, _dtype, create_targets_bd  # noqa: F401
from .train_victim import _NullWriter, _train_epoch


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


def eval(netC, optimizerC, schedulerC, test_dl, best_clean_acc, tf_writer, epoch, opt):
    """train_clean_classifier.py:122-160: clean accuracy over the test loader; checkpoint when it improves."""
    print(" Eval:")
    netC.eval()
    dev = netC.net.device
    tot = torch.zeros(2, dtype=torch.int64, device=dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    total_sample = 0
    for inputs, targets in test_dl:
        y = targets.cpu().numpy().astype(np.int64) if torch.is_tensor(targets) else np.asarray(targets, dtype=np.int64)
        x = inputs.to(dev, non_blocking=True).float().contiguous()
        preds_clean, _ = netC.net.forward(x, train=False, save=False)                              # :134
        ops.cross_entropy(preds_clean, torch.from_numpy(y).to(dev, non_blocking=True), 1.0, False, counts_out=counts)
        tot += counts.long()
        total_sample += len(y)
    acc_clean = float(tot[0]) * 100.0 / max(total_sample, 1)
    print("Clean Acc: {:.4f} - Best: {:.4f}".format(acc_clean, best_clean_acc))
    if not epoch % 1:
        tf_writer.add_scalars("Accuracy", {"Test": acc_clean}, epoch)
    if acc_clean > best_clean_acc:
        print(" Saving...")
        best_clean_acc = acc_clean
        state_dict = {"netC": netC.state_dict(), "schedulerC": schedulerC.state_dict(), "optimizerC": optimizerC.state_dict(),
                      "best_clean_acc": acc_clean, "epoch_current": epoch}
        d = os.path.dirname(opt.ckpt_path)
        if d:
            os.makedirs(d, exist_ok=True)
        torch.save(state_dict, opt.ckpt_path)
    return best_clean_acc


def main(argv=None):
    """train_clean_classifier.py:163-236 (checkpoint folder WITHOUT the "_clean" suffix: the path train_generator.py's main()
    builds from --load_checkpoint_clean, :513-517)."""
    import shutil
    from .utils.dataloader import get_dataloader
    opt = config.get_arguments().parse_args(argv)
    _dataset_shape(opt)
    train_dl, test_dl = get_dataloader(opt, True), get_dataloader(opt, False)
    netC, optimizerC, schedulerC = get_model(opt)
    mode = opt.saving_prefix
    opt.ckpt_folder = os.path.join(opt.checkpoints, "{}".format(mode), opt.dataset)
    opt.ckpt_path = os.path.join(opt.ckpt_folder, "{}_{}.pth.tar".format(opt.dataset, mode))
    opt.log_dir = os.path.join(opt.ckpt_folder, "log_dir")
    os.makedirs(opt.log_dir, exist_ok=True)
    best_clean_acc, epoch_current = 0.0, 0
    if opt.continue_training:
        if not os.path.exists(opt.ckpt_path):
            print("Pretrained model doesnt exist")
            sys.exit()
        print("Continue training!!")
        sd = torch.load(opt.ckpt_path, map_location=opt.device, weights_only=False)
        netC.load_state_dict(sd["netC"])
        optimizerC.load_state_dict(sd["optimizerC"])
        schedulerC.load_state_dict(sd["schedulerC"])
        best_clean_acc, epoch_current = sd["best_clean_acc"], sd["epoch_current"]
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
        train(netC, optimizerC, schedulerC, train_dl, tf_writer, epoch, opt)
        best_clean_acc = eval(netC, optimizerC, schedulerC, test_dl, best_clean_acc, tf_writer, epoch, opt)
    return best_clean_acc


if __name__ == "__main__":
    main()
