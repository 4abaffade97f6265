"""Loader that makes the UNMODIFIED reference importable in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/combat_oracle.py header).  Used solely by
tests/golden/make_golden.py to generate the committed golden fixtures; it is
never imported on the GPU box (``/root/reference`` does not exist there).

The reference cannot be imported as shipped: ``kornia`` and
``vit_pytorch`` are absent and ``classifier_models/__init__.py`` is empty.  This
module installs inert ``sys.modules`` stand-ins for the two absent third-party
packages (only touched with ``--post_transform_option use``) and populates the
empty package namespace.  No reference source is copied or modified.
"""
import sys
import types

import torch.nn as nn

REF = "/root/reference"


def load_reference(ref: str = REF):
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if ref not in sys.path:
        sys.path.insert(0, ref)
    if "kornia" not in sys.modules:
        k, ka = types.ModuleType("kornia"), types.ModuleType("kornia.augmentation")

        def _absent(*a, **kw):
            raise RuntimeError("kornia is not installed (stand-in)")

        for n in ("RandomCrop", "RandomRotation", "RandomHorizontalFlip"):
            setattr(ka, n, _absent)
        k.augmentation = ka
        # kornia.losses.total_variation (train_generator_imperceptible.py:13): kornia is absent, so the stand-in is the oracle's
        # restatement of the kornia 0.6.6 formula -- fixtures of the imperceptible variant pin everything AROUND that formula,
        # not the formula itself ("parity unpinned w.r.t. kornia")
        kl = types.ModuleType("kornia.losses")

        def total_variation(img):
            from oracle.combat_oracle import total_variation as tv
            return tv(img)

        kl.total_variation = total_variation
        k.losses = kl
        sys.modules["kornia.losses"] = kl
        sys.modules["kornia"], sys.modules["kornia.augmentation"] = k, ka
    if "vit_pytorch" not in sys.modules:
        v = types.ModuleType("vit_pytorch")
        v.SimpleViT = type("SimpleViT", (nn.Module,), {})
        sys.modules["vit_pytorch"] = v
    import classifier_models
    from classifier_models.densenet import DenseNet121
    from classifier_models.mobilenetv2 import MobileNetV2
    from classifier_models.preact_resnet import PreActResNet18
    from classifier_models.resnet import ResNet18
    from classifier_models.vgg import VGG

    for n, o in dict(VGG=VGG, DenseNet121=DenseNet121, MobileNetV2=MobileNetV2,
                     PreActResNet18=PreActResNet18, ResNet18=ResNet18).items():
        setattr(classifier_models, n, o)
    import config
    import train_generator

    return train_generator, config


def load_reference_multilabel(ref: str = REF):
    """train_generator_multilabel.py of the reference (same stand-ins)."""
    load_reference(ref)
    import train_generator_multilabel

    return train_generator_multilabel


def load_reference_imperceptible(ref: str = REF):
    """train_generator_imperceptible.py of the reference (same stand-ins + kornia.losses.total_variation, see above)."""
    load_reference(ref)
    import train_generator_imperceptible

    return train_generator_imperceptible


def load_reference_inputaware(ref: str = REF):
    """train_generator_inputaware.py of the reference (same stand-ins)."""
    load_reference(ref)
    import train_generator_inputaware

    return train_generator_inputaware


def load_reference_wanet(ref: str = REF):
    """train_generator_wanet.py of the reference (same stand-ins)."""
    load_reference(ref)
    import train_generator_wanet

    return train_generator_wanet


class NullWriter:
    """tf_writer stand-in for train()."""

    def add_scalars(self, *a, **k):
        pass

    def add_image(self, *a, **k):
        pass
