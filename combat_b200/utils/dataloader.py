"""utils/dataloader.py of the reference, hot-path part: PostTensorTransform (:45-60) and ProbTransform (:11-22).

The reference builds the transform from kornia 0.6.6 augmentations (requirements.txt:12), which are absent from this image
and from /root/reference, so their PUBLISHED algorithm is restated here (and, independently, by the CPU checker used in tests/):

  RandomCrop((H, W), padding=opt.random_crop)  zero-pad by `padding`, per-sample window start (xs, ys) = floor(U[0, 2*pad+1))
  RandomRotation(opt.random_rotation)          per-sample angle ~ U(-deg, deg); warp_affine about ((W-1)/2, (H-1)/2) with
                                               get_rotation_matrix2d(center, angle, 1), bilinear, zeros, align_corners=True
  RandomHorizontalFlip(p=0.5)                  per-sample Bernoulli(0.5) mirror (cifar10 only, :54-55)
  ProbTransform(f, p)                          ONE `random.random() < p` per call gates the whole batch (:17)

Where the random numbers come from (the part of kornia that nothing in the reference pins -- "parity unpinned w.r.t. the kornia
RNG stream"): the gates use Python's `random` exactly as the reference does; the per-sample parameters are
drawn from the torch CPU default generator with `torch.rand(B)` per parameter in the order xs, ys (crop), angle (rotation), flip
-- the order kornia's generators sample them in.  What IS pinned: given the parameters, the pixels (and the gradient through
them) equal the restated kornia pipeline (`tests/test_post_transform_gpu.py` against the torch F.pad / F.affine_grid /
F.grid_sample restatement in the oracle).

The pixels are produced by ONE fused gather kernel (csrc/augment.cu) instead of three library passes; `forward` is
differentiable (the G-step back-propagates through transforms(inputs_bd), train_generator.py:228,250)."""
import math
import random

import numpy as np
import torch

from .. import ops

PARAM_WIDTH = 8


def identity_params(rows: int) -> np.ndarray:
    p = np.zeros((rows, PARAM_WIDTH), dtype=np.float32)
    p[:, 2] = 1.0
    return p


def draw_params(rows: int, opt) -> np.ndarray:
    """One PostTensorTransform call's random decisions for a batch of `rows` images -> float32 [rows, 8] (csrc/augment.cu).
    Consumes `random` (gates) and the torch CPU generator (per-sample parameters) in the reference's module order:
    random_crop, random_rotation, random_horizontal_flip (utils/dataloader.py:48-55,58-59)."""
    option = opt.post_transform_option
    p = identity_params(rows)
    if option == "no_use":
        return p
    pad = int(opt.random_crop)
    if option != "use_modified":                      # :49-52
        if random.random() < 0.8:                     # ProbTransform(p=0.8), :17
            span = float(2 * pad + 1)                 # (H + 2*pad) - H + 1 window starts
            xs = torch.floor(torch.rand(rows) * span)
            ys = torch.floor(torch.rand(rows) * span)
            p[:, 0] = xs.numpy() - pad
            p[:, 1] = ys.numpy() - pad
    if random.random() < 0.5:                         # ProbTransform(RandomRotation, p=0.5), :53
        deg = float(opt.random_rotation)
        ang = (torch.rand(rows) * (2.0 * deg) - deg).to(torch.float32)
        rad = ang * (math.pi / 180.0)                 # kornia deg2rad in float32
        p[:, 2] = torch.cos(rad).numpy()
        p[:, 3] = torch.sin(rad).numpy()
        p[:, 4] = 1.0
    if opt.dataset == "cifar10":                      # :54-55
        p[:, 5] = (torch.rand(rows) < 0.5).to(torch.float32).numpy()
    return p


class _TransformFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params):
        ctx.save_for_backward(params)
        return ops.post_transform_fwd(x.contiguous().float(), params)

    @staticmethod
    def backward(ctx, g):
        (params,) = ctx.saved_tensors
        return ops.post_transform_bwd(g.contiguous().float(), params), None


class PostTensorTransform(torch.nn.Module):
    """Drop-in for the reference's module: `transforms = PostTensorTransform(opt); y = transforms(x)` on CUDA tensors.
    `opt` needs post_transform_option, random_crop, random_rotation, dataset (config.py:75-79)."""

    def __init__(self, opt):
        super().__init__()
        if opt.post_transform_option not in ("use", "no_use", "use_modified"):
            raise ValueError("post_transform_option must be one of use / no_use / use_modified")
        self.opt = opt
        self.option = opt.post_transform_option
        self.last_params = None   # float32 [rows, 8] host array of the most recent call (tests, logging)

    def forward(self, x):
        if self.option == "no_use":
            return x
        if not x.is_cuda:
            raise RuntimeError("combat_b200.PostTensorTransform runs on CUDA tensors only (no CPU fallback)")
        self.last_params = draw_params(x.shape[0], self.opt)
        return _TransformFn.apply(x, torch.from_numpy(self.last_params).to(x.device))


# ------------------------------------------------------------------ data (utils/dataloader.py:24-43,98-123)
def get_transform(opt, train=True, pretensor_transform=False):
    """utils/dataloader.py:24-43: Resize -> [RandomCrop, RandomRotation, (cifar10) RandomHorizontalFlip] -> ToTensor ->
    Normalize(0.5, 0.5), i.e. images in [-1, 1]."""
    import torchvision.transforms as T
    tl = [T.Resize((opt.input_height, opt.input_width))]
    if pretensor_transform and train:
        tl.append(T.RandomCrop((opt.input_height, opt.input_width), padding=opt.random_crop))
        tl.append(T.RandomRotation(opt.random_rotation))
        if opt.dataset == "cifar10":
            tl.append(T.RandomHorizontalFlip(p=0.5))
    tl.append(T.ToTensor())
    if opt.dataset in ("cifar10", "celeba", "imagenet10"):
        tl.append(T.Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5]))
    else:
        raise Exception("Invalid Dataset")
    return T.Compose(tl)


class SyntheticBatches:
    """--synthetic_data: a fixed list of uniform [-1, 1] image batches with random labels in PINNED host memory (what the
    benchmarks and the GPU box use: there is no network for torchvision's download=True)."""

    def __init__(self, opt, train, bs, n_batches=None, seed=0):
        g = torch.Generator().manual_seed(seed + (0 if train else 1))
        if n_batches is None:
            n_batches = (8 if train else 2) if getattr(opt, "debug", False) else (32 if train else 8)
        pin = torch.cuda.is_available()
        self.batches = []
        for _ in range(n_batches):
            x = torch.rand(bs, opt.input_channel, opt.input_height, opt.input_width, generator=g) * 2 - 1
            self.batches.append((x.pin_memory() if pin else x, torch.randint(0, opt.num_classes, (bs,), generator=g)))

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return iter(self.batches)


class CelebA_attr(torch.utils.data.Dataset):
    """utils/dataloader.py:63-81: three binary attributes (18, 31, 21) -> 8 classes."""

    def __init__(self, opt, split, transforms):
        import torchvision
        self.dataset = torchvision.datasets.CelebA(root=opt.data_root, split=split, target_type="attr", download=False)
        self.list_attributes = [18, 31, 21]
        self.transforms = transforms
        self.split = split

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, index):
        x, target = self.dataset[index]
        a = target[self.list_attributes]
        return self.transforms(x), (a[0] << 2) + (a[1] << 1) + a[2]


def get_dataloader(opt, train=True, pretensor_transform=False, target_label=None, bs=None, shuffle=True):
    """utils/dataloader.py:98-123 (same signature).  The reference asks torchvision to download the dataset; there is no
    network here, so the dataset has to be present under opt.data_root -- or pass --synthetic_data."""
    import os
    import torchvision
    if bs is None:
        bs = opt.bs
    if getattr(opt, "synthetic_data", False):
        return SyntheticBatches(opt, train, bs)
    transform = get_transform(opt, train, pretensor_transform)
    if opt.dataset == "cifar10":
        dataset = torchvision.datasets.CIFAR10(opt.data_root, train, transform, download=False)
        if target_label is not None:
            pairs = [(x, y) for x, y in zip(dataset.data, dataset.targets) if int(y) == target_label]
            dataset.data, dataset.targets = [x[0] for x in pairs], [x[1] for x in pairs]
    elif opt.dataset == "celeba":
        dataset = CelebA_attr(opt, "train" if train else "test", transform)
    elif opt.dataset == "imagenet10":
        dataset = torchvision.datasets.ImageNet(root=os.path.join(opt.data_root, "imagenet10"), split="train" if train else "val",
                                                transform=transform)
    else:
        raise Exception("Invalid dataset")
    if opt.debug:
        dataset = torch.utils.data.Subset(dataset, range(min(len(dataset), 1000)))
    return torch.utils.data.DataLoader(dataset, batch_size=bs, num_workers=opt.num_workers, shuffle=shuffle, pin_memory=True)
