"""utils/dataloader.py of the reference, hot-path part: PostTensorTransform.

Only `--post_transform_option no_use` (identity, utils/dataloader.py:48) is implemented: the random crop / rotation /
flip of the default option are kornia 0.6.6 ops whose parameter sampling is not pinned by anything in the reference
("next" row 1)."""
import torch


class PostTensorTransform(torch.nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.option = opt.post_transform_option
        if self.option != "no_use":
            raise NotImplementedError(
                "combat_b200: --post_transform_option %s (kornia RandomCrop/RandomRotation/RandomHorizontalFlip) is not "
                "part of the built hot path yet; run with --post_transform_option no_use" % self.option)

    def forward(self, x):
        return x
