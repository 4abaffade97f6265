"""utils/dataloader_cleanbd.py of the reference (:124-175): the clean-label poisoned dataset of the victim trainers -- every
item is (input, target, poisoned), where `poisoned` marks a fixed random subset (fraction opt.pc, drawn ONCE with Python's
`random.sample`) of the target-class images (all classes for all2all)."""
import random

import torch

from .dataloader import CelebA_attr, PostTensorTransform, SyntheticBatches, get_transform  # noqa: F401


class PoisonedDataset(torch.utils.data.Dataset):
    """utils/dataloader_cleanbd.py:124-153"""

    def __init__(self, refdata, n_classes, opt):
        self.dataset = refdata
        if opt.debug:
            self.dataset = torch.utils.data.Subset(self.dataset, range(min(len(self.dataset), 1000)))
        target_label = {opt.target_label} if opt.attack_mode == "all2one" else set(range(0, n_classes))
        self.poisoned = self._poison_flags(target_label, opt.pc)

    def _poison_flags(self, target_label, pc):
        ids = [idx for idx, (_, label) in enumerate(self.dataset) if int(label) in target_label]
        num_poisoned = max(0, int(pc * len(ids)))
        print(f"Poison {num_poisoned} images ({pc * len(ids)})")
        return set(random.sample(ids, num_poisoned))

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, index):
        x, target = self.dataset[index]
        return x, target, index in self.poisoned


class SyntheticPoisonedBatches(SyntheticBatches):
    """--synthetic_data: the synthetic batches with the dataset-level flags of PoisonedDataset (same rule, same RNG)."""

    def __init__(self, opt, train, bs):
        super().__init__(opt, train, bs)
        labels = torch.cat([y for _, y in self.batches]).tolist()
        target_label = {opt.target_label} if opt.attack_mode == "all2one" else set(range(0, opt.num_classes))
        ids = [i for i, l in enumerate(labels) if int(l) in target_label]
        chosen = set(random.sample(ids, max(0, int(opt.pc * len(ids)))))
        flags = torch.tensor([i in chosen for i in range(len(labels))])
        self.batches = [(x, y, flags[k * bs:(k + 1) * bs]) for k, (x, y) in enumerate(self.batches)]


def get_dataloader(opt, train=True, pretensor_transform=False, bs=None, shuffle=True):
    """utils/dataloader_cleanbd.py:156-175 (no network here: the dataset must be present under opt.data_root, or pass
    --synthetic_data)."""
    import os
    import torchvision
    if bs is None:
        bs = opt.bs
    if getattr(opt, "synthetic_data", False):
        return SyntheticPoisonedBatches(opt, train, bs)
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
    if opt.debug:
        dataset = torch.utils.data.Subset(dataset, range(min(len(dataset), 1000)))
    return torch.utils.data.DataLoader(dataset, batch_size=bs, num_workers=opt.num_workers, shuffle=shuffle, pin_memory=True)
