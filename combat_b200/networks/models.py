"""The names train_generator.py imports from networks/models.py of the reference: the trigger generators (additive U-Nets and the WaNet flow generator; kernel-backed
modules) and the `Denormalizer` helper (reference networks/models.py:38-86) that maps normalised images back to [0, 1]."""
import torch

from ..modules import CUnetGeneratorv1, GridGenerator, UnetGenerator  # noqa: F401

# per-dataset (mean, std) of the reference's input normalisation; None = the dataset is not normalised
_STATS = {"cifar10": (0.5, 0.5), "celeba": (0.5, 0.5), "imagenet10": (0.5, 0.5), "mnist": (0.5, 0.5), "gtsrb": None}


class Denormalize:
    """x * variance[c] + expected_values[c] per channel, on a copy (one broadcast multiply-add instead of a channel loop)."""

    def __init__(self, opt, expected_values, variance):
        self.n_channels = opt.input_channel
        self.expected_values, self.variance = list(expected_values), list(variance)
        assert self.n_channels == len(self.expected_values)

    def __call__(self, x):
        shape = (1, self.n_channels) + (1,) * (x.dim() - 2)
        scale = torch.as_tensor(self.variance, dtype=x.dtype, device=x.device).view(shape)
        shift = torch.as_tensor(self.expected_values, dtype=x.dtype, device=x.device).view(shape)
        return x * scale + shift


class Denormalizer:
    def __init__(self, opt):
        if opt.dataset not in _STATS:
            raise Exception("Invalid dataset")
        stats = _STATS[opt.dataset]
        n = 1 if opt.dataset == "mnist" else opt.input_channel
        self.denormalizer = None if stats is None else Denormalize(opt, [stats[0]] * n, [stats[1]] * n)

    def __call__(self, x):
        return self.denormalizer(x) if self.denormalizer else x
