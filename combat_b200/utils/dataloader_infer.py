"""utils/dataloader_infer.py of the reference (:100-138): the inference-side variant of the poisoned loader -- every item is
(input, target, index): the running index of the sample instead of a poison flag (the poison bookkeeping of
utils/dataloader_cleanbd.py is commented out in this file, :101-107).  Same transforms and dataset branches."""
import torch

from .dataloader import CelebA_attr, PostTensorTransform, SyntheticBatches, get_transform  # noqa: F401


class PoisonedDataset(torch.utils.data.Dataset):
    """utils/dataloader_infer.py:100-113"""

    def __init__(self, refdata, n_classes, opt):
        self.dataset = refdata

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, index):
        x, target = self.dataset[index]
        return x, target, index


class SyntheticIndexedBatches(SyntheticBatches):
    """--synthetic_data: the synthetic batches with the running sample index as third element."""

    def __init__(self, opt, train, bs):
        super().__init__(opt, train, bs)
        self.batches = [(x, y, torch.arange(k * bs, k * bs + len(y))) for k, (x, y) in enumerate(self.batches)]


def get_dataloader(opt, train=True, pretensor_transform=False, bs=None, shuffle=True):
    """utils/dataloader_infer.py:116-138 (no network here: the dataset must be present under opt.data_root, or pass
    --synthetic_data)."""
    import os
    import torchvision
    if bs is None:
        bs = opt.bs
    if getattr(opt, "synthetic_data", False):
        return SyntheticIndexedBatches(opt, train, bs)
    transform = get_transform(opt, train, pretensor_transform)
    if opt.dataset == "cifar10":
        dataset = PoisonedDataset(torchvision.datasets.CIFAR10(opt.data_root, train, transform, download=False), opt.num_classes, opt)
    elif opt.dataset == "celeba":
        dataset = PoisonedDataset(CelebA_attr(opt, "train" if train else "test", transform), opt.num_classes, opt)
    elif opt.dataset == "imagenet10":
        dataset = PoisonedDataset(torchvision.datasets.ImageNet(root=os.path.join(opt.data_root, "imagenet10"),
                                                                split="train" if train else "val", transform=transform),
                                  opt.num_classes, opt)
    else:
        raise Exception("Invalid dataset")
    return torch.utils.data.DataLoader(dataset, batch_size=bs, num_workers=opt.num_workers, shuffle=shuffle, pin_memory=True)
