"""classifier_models of the reference (classifier_models/preact_resnet.py, resnet.py), B200-native."""
from ..modules import PreActResNet18, ResNet18  # noqa: F401
