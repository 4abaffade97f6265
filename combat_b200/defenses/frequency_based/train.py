"""defenses/frequency_based/train.py of the reference, hot-path part: `get_model` (:146-176, the shipped 'original' detector),
`train` (:178-221) and `eval` (:224-272) with the reference's signatures.

What runs where: the synthetic triggers (`patching_train`, :106-143) stay on the host -- they are per-image numpy draws;
the per-plane scipy DCT loops (:195-197, :242-246) become ONE launch of the uint8-input DCT kernel over the whole
(clean, patched) batch; the FrequencyModel iteration (train-mode forward with host-drawn dropout masks, cross entropy,
backward, Adadelta) runs on the CUDA kernels of nets.FrequencyDetector.  Checker: oracle/detector_oracle.py, pinned to the
unmodified reference (tests/golden/detector_b8x2.npz).

STATUS: GPU-validated in round 2 -- the DCT launch (tests/test_z_detector_dct_gpu.py) and the training iteration / evaluation
against the fixture of the unmodified reference (tests/test_detector_train_gpu.py, un-gated).
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch

from ... import ops
from ...modules import FrequencyModel
from ...utils.dct import dct_2d


class AlbumentationsAugment:
    """addnoise / randshadow of train.py:49-62 through albumentations, as the reference does.  The package is imported on
    first use: it is not needed for the three block / blend triggers."""

    def __init__(self):
        self._alb = None

    def _a(self):
        if self._alb is None:
            import albumentations
            self._alb = albumentations
        return self._alb

    def addnoise(self, img):
        aug = self._a().GaussNoise(p=1, mean=25, var_limit=(10, 70))
        return aug(image=(img * 255).astype(np.uint8))["image"] / 255

    def randshadow(self, img, input_size=32):
        import cv2
        aug = self._a().RandomShadow(p=1)
        return aug(image=cv2.resize((img * 255).astype(np.uint8), (input_size, input_size)))["image"] / 255


def _hwc(t):
    return t.detach().cpu().numpy().transpose(1, 2, 0)


def patching_train(sample, train_data, n_input=3, input_size=32, augment=None):
    """train.py:106-143: one patched image (HWC, values in [0, 1]).  Trigger type, patch size and placement come from the
    numpy global RNG in the reference's order."""
    augment = augment or AlbumentationsAugment()
    kind = np.random.randint(0, 5)
    ph, pw = np.random.randint(2, 8), np.random.randint(2, 8)
    img = _hwc(sample).copy()
    if kind == 2:
        return augment.addnoise(img)
    if kind == 3:
        return augment.randshadow(img, input_size)
    if kind == 4:
        other = _hwc(train_data[np.random.randint(train_data.shape[0])])
        return np.minimum(img + 0.3 * other, 1)
    patch = np.ones((ph, pw, n_input)) if kind == 0 else np.random.rand(ph, pw, n_input)
    margin, corner = np.random.randint(0, 6), np.random.randint(0, 4)
    top = margin if corner < 2 else input_size - margin - ph
    left = margin if corner % 2 == 0 else input_size - margin - pw
    img[top:top + ph, left:left + pw, :] = patch
    return np.minimum(img, 1)


def make_batch(x, opt, shuffle, augment=None):
    """train.py:188-200 / :236-250: (clean, patched) rows, quantised to uint8 on the host, ONE DCT launch on the device.
    Returns (coefficients float32 [2B,C,H,W] on opt.device, labels int64 [2B] on opt.device)."""
    B = x.shape[0]
    xc = x.detach().cpu()
    patched = np.zeros((B, opt.input_channel, opt.input_height, opt.input_width))
    for i in range(B):
        patched[i] = np.transpose(patching_train(xc[i], xc, opt.input_channel, opt.input_height, augment), (2, 0, 1))
    planes = (np.vstack((xc.numpy(), patched)) * 255).astype(np.uint8)     # :197, truncation
    labels = np.concatenate((np.zeros(B, dtype=np.int64), np.ones(B, dtype=np.int64)))
    idx = np.arange(2 * B)
    if shuffle:
        random.shuffle(idx)                                                 # :199
    coef = dct_2d(torch.from_numpy(planes[idx]).to(opt.device))
    return coef, torch.from_numpy(labels[idx]).to(opt.device)


def get_model(opt):
    """train.py:146-176 for --model original / original_holdout."""
    if opt.model not in ("original", "original_holdout"):
        raise NotImplementedError("--model %s is outside the built path (the shipped detector is 'original')" % opt.model)
    netC = FrequencyModel(num_classes=2, n_input=opt.input_channel, input_size=opt.input_height, device=opt.device,
                          dtype=torch.float32, trainable=True)
    optimizerC = torch.optim.Adadelta(netC.parameters(), lr=0.05, weight_decay=1e-4)
    return netC, optimizerC


def train_iteration(netC, x, opt, lr_dev, augment=None, weight_decay=1e-4):
    """One iteration of train.py:185-209.  Returns (loss [1] device float32, counts [2] device int32: correct, -)."""
    net = netC.net
    coef, y = make_batch(x, opt, True, augment)
    masks = net.draw_dropout_masks(coef.shape[0], coef.shape[2], coef.shape[3])
    net.zero_grad()
    logits, ctx = net.train_forward(coef, masks)
    loss, dlogits, counts = ops.cross_entropy(logits, y, 1.0, True)
    net.train_backward(ctx, dlogits)
    net.adadelta_step(lr_dev, wd=weight_decay)
    return loss, counts, logits


def _bind_adadelta_state(optimizerC, netC):
    """expose the fused optimiser's accumulators through the torch optimiser's state (checkpoint round trip, train.py:262-268)"""
    st = netC.net.store
    if not hasattr(st, "acc_delta"):
        return
    for name, p in netC._plist:
        state = optimizerC.state[p]
        state["square_avg"], state["acc_delta"] = st._view(st.mom, name), st._view(st.acc_delta, name)
        state["step"] = state.get("step", torch.zeros((), dtype=torch.float32)) + 0


def _adopt_adadelta_state(optimizerC, netC):
    """`optimizerC.load_state_dict(ckpt)` (train.py:318, --continue_training) leaves fresh copies of square_avg / acc_delta in
    optimizer.state; the fused Adadelta reads the flat store.  Copy any accumulator that does not alias the store into it, so that a
    resumed run continues with the saved accumulators instead of zeros."""
    st = netC.net.store
    for name, p in netC._plist:
        state = optimizerC.state.get(p, {})
        sq, ad = state.get("square_avg"), state.get("acc_delta")
        if sq is None or ad is None:
            continue
        if not hasattr(st, "acc_delta"):
            st.acc_delta = torch.zeros_like(st.flat)
        v_sq, v_ad = st._view(st.mom, name), st._view(st.acc_delta, name)
        if sq.data_ptr() != v_sq.data_ptr():
            v_sq.copy_(sq.to(v_sq.device, torch.float32))
        if ad.data_ptr() != v_ad.data_ptr():
            v_ad.copy_(ad.to(v_ad.device, torch.float32))


def train(netC, optimizerC, train_dl, tf_writer, epoch, opt, augment=None):
    """train.py:178-221"""
    print(" Train:")
    netC.train()
    _adopt_adadelta_state(optimizerC, netC)
    dev = torch.device(opt.device)
    group = optimizerC.param_groups[0]
    lr_dev = torch.full((1,), float(group["lr"]), dtype=torch.float32, device=dev)
    total_loss = torch.zeros(1, dtype=torch.float64, device=dev)
    total_correct = torch.zeros(1, dtype=torch.int64, device=dev)
    total_sample, batch_idx = 0, -1
    for batch_idx, (x, _) in enumerate(train_dl):
        loss, counts, logits = train_iteration(netC, x, opt, lr_dev, augment, float(group.get("weight_decay", 1e-4)))
        total_loss += loss.double()
        total_correct += counts[0].long()
        total_sample += logits.shape[0]
    for name, buf in netC.named_buffers():      # BatchNorm2d.num_batches_tracked (checkpoint round trip)
        if name.endswith("num_batches_tracked"):
            buf += batch_idx + 1 if total_sample else 0
    _bind_adadelta_state(optimizerC, netC)
    avg_acc = float(total_correct) * 100.0 / max(total_sample, 1)
    avg_loss = float(total_loss) / max(total_sample, 1)
    print("CE Loss: {:.4f} | Acc: {:.4f}".format(avg_loss, avg_acc))
    if not epoch % 1:
        tf_writer.add_scalars("Accuracy", {"Train": avg_acc}, epoch)
        tf_writer.add_scalar("CE_Loss", avg_loss, epoch)


def eval(netC, optimizerC, test_dl, best_acc, tf_writer, epoch, opt, augment=None):
    """train.py:224-272: accuracy on (clean, patched) batches without shuffling; checkpoint on improvement."""
    print(" Eval:")
    netC.eval()
    total_sample, total_correct = 0, 0
    acc = 0.0
    for batch_idx, (x, _) in enumerate(test_dl):
        coef, y = make_batch(x, opt, False, augment)
        preds = netC(coef)
        total_correct += int((preds.argmax(1) == y).sum())
        total_sample += coef.shape[0]
        acc = total_correct * 100.0 / total_sample
    print("Acc: {:.4f} - Best: {:.4f}".format(acc, best_acc))
    if not epoch % 1:
        tf_writer.add_scalars("Accuracy", {"Test": acc}, epoch)
    if acc > best_acc:
        print(" Saving...")
        best_acc = acc
        os.makedirs(os.path.dirname(os.path.abspath(opt.ckpt_path)), exist_ok=True)
        torch.save({"netC": netC.state_dict(), "optimizerC": optimizerC.state_dict(), "best_acc": acc, "epoch_current": epoch},
                   opt.ckpt_path)
    return best_acc


class _NullWriter:
    def add_scalars(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass


class SyntheticUnitBatches:
    """--synthetic_data: batches of uniform [0, 1] images (the range defenses/frequency_based/dataloader.py yields: ToTensor
    without normalisation) with random labels -- there is no network here for torchvision's download=True."""

    def __init__(self, opt, train, n_batches=None, seed=0):
        g = torch.Generator().manual_seed(seed + (0 if train else 1))
        if n_batches is None:
            n_batches = (4 if train else 2) if getattr(opt, "debug", False) else (16 if train else 4)
        self.batches = [(torch.rand(opt.bs, opt.input_channel, opt.input_height, opt.input_width, generator=g),
                         torch.randint(0, opt.num_classes, (opt.bs,), generator=g)) for _ in range(n_batches)]

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return iter(self.batches)


def get_dataloader(opt, train=True, shuffle=True):
    """defenses/frequency_based/dataloader.py:103-122: CIFAR-10 / CelebA through torchvision, images in [0, 1] (ToTensor only,
    :12-28; CelebA resized to the input size).  GTSRB / MNIST (1 or 3 channels at 32 x 32 from local files) are not built."""
    if getattr(opt, "synthetic_data", False):
        return SyntheticUnitBatches(opt, train)
    import torchvision
    import torchvision.transforms as T
    tf = T.Compose([T.Resize((opt.input_height, opt.input_width)), T.ToTensor()])
    if opt.dataset == "cifar10":
        ds = torchvision.datasets.CIFAR10(opt.data_root, train, transform=tf, download=True)
    elif opt.dataset == "celeba":
        from ...utils.dataloader import CelebA_attr
        ds = CelebA_attr(opt, "train" if train else "test", tf)
    else:
        raise NotImplementedError("dataset %s is outside the built path" % opt.dataset)
    return torch.utils.data.DataLoader(ds, batch_size=opt.bs, num_workers=opt.num_workers, shuffle=shuffle)


def main(argv=None, augment=None):
    """train.py:275-344: dataset shape, loaders (images in [0, 1]), get_model, --continue_training, n_iters epochs of train() +
    eval(); checkpoint `<checkpoints>/<dataset>/<model>/<dataset>_<model>_detector.pth.tar` -- the file train_generator.py's main()
    loads as the frequency detector (F_ckpt_path)."""
    import shutil

    from . import config
    opt = config.get_arguments().parse_args(argv)
    sizes = {"cifar10": (32, 3, 10), "celeba": (64, 3, 8)}
    if opt.dataset in ("gtsrb", "mnist"):
        raise NotImplementedError("dataset %s is outside the built path (local GTSRB / MNIST files)" % opt.dataset)
    if opt.dataset not in sizes:
        raise Exception("Invalid Dataset")
    opt.input_height = opt.input_width = sizes[opt.dataset][0]
    opt.input_channel, opt.num_classes = sizes[opt.dataset][1], sizes[opt.dataset][2]
    train_dl, test_dl = get_dataloader(opt, True), get_dataloader(opt, False)
    netC, optimizerC = get_model(opt)
    opt.ckpt_folder = os.path.join(opt.checkpoints, opt.dataset, opt.model)
    opt.ckpt_path = os.path.join(opt.ckpt_folder, "{}_{}_detector.pth.tar".format(opt.dataset, opt.model))
    opt.log_dir = os.path.join(opt.ckpt_folder, "log_dir")
    os.makedirs(opt.log_dir, exist_ok=True)
    best_acc, epoch_current = 0.0, 0
    if opt.continue_training:
        if not os.path.exists(opt.ckpt_path):
            print("Pretrained model doesnt exist")
            raise SystemExit
        print("Continue training!!")
        sd = torch.load(opt.ckpt_path, map_location=opt.device, weights_only=False)
        netC.load_state_dict(sd["netC"])
        optimizerC.load_state_dict(sd["optimizerC"])
        best_acc, epoch_current = sd["best_acc"], sd["epoch_current"]
    else:
        print("Train from scratch!!!")
        shutil.rmtree(opt.ckpt_folder, ignore_errors=True)
        os.makedirs(opt.log_dir, exist_ok=True)
    try:
        from torch.utils.tensorboard import SummaryWriter
        tf_writer = SummaryWriter(log_dir=opt.log_dir)
    except Exception:
        tf_writer = _NullWriter()
    augment = augment or AlbumentationsAugment()
    for epoch in range(epoch_current, opt.n_iters):
        print("Epoch {}:".format(epoch + 1))
        train(netC, optimizerC, train_dl, tf_writer, epoch, opt, augment)
        best_acc = eval(netC, optimizerC, test_dl, best_acc, tf_writer, epoch, opt, augment)
    return best_acc


if __name__ == "__main__":
    main()
