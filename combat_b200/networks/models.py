"""networks/models.py of the reference: the generators on the hot path plus the (de)normalisers train() touches."""
import torch

from ..modules import CUnetGeneratorv1, UnetGenerator  # noqa: F401


class Denormalize:
    """networks/models.py:38-52"""

    def __init__(self, opt, expected_values, variance):
        self.n_channels = opt.input_channel
        self.expected_values, self.variance = expected_values, variance
        assert self.n_channels == len(self.expected_values)

    def __call__(self, x):
        x_clone = x.clone()
        for channel in range(self.n_channels):
            x_clone[:, channel] = x[:, channel] * self.variance[channel] + self.expected_values[channel]
        return x_clone


class Denormalizer:
    """networks/models.py:71-86: only gtsrb/celeba-free datasets denormalise; cifar10 uses mean/std 0.5 in [-1,1] data."""

    def __init__(self, opt):
        self.denormalizer = self._get_denormalizer(opt)

    def _get_denormalizer(self, opt):
        if opt.dataset in ("cifar10", "celeba", "imagenet10"):
            return Denormalize(opt, [0.5] * opt.input_channel, [0.5] * opt.input_channel)
        if opt.dataset == "mnist":
            return Denormalize(opt, [0.5], [0.5])
        if opt.dataset == "gtsrb":
            return None
        raise Exception("Invalid dataset")

    def __call__(self, x):
        if self.denormalizer:
            x = self.denormalizer(x)
        return x
