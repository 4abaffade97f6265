"""Host half of combat_b200.defenses.frequency_based.train (the detector trainer's synthetic triggers, quantisation, labels
and shuffle -- numpy / Python RNG work that stays on the host) against oracle/detector_oracle.py, which is pinned to the
unmodified reference.  The one device call of `make_batch` (the uint8 DCT launch) is replaced by the oracle's transform."""
import random
import types

import numpy as np
import torch

from oracle import combat_oracle as O
from oracle import detector_oracle as D


class StandInAugment:
    """the oracle's stand-ins for the two albumentations transforms (absent from the container)"""

    def addnoise(self, img):
        return D.addnoise(img)

    def randshadow(self, img, input_size=32):
        return D.randshadow(img, input_size)


def test_patching_and_batch_match_the_oracle(monkeypatch):
    import combat_b200.defenses.frequency_based.train as T
    x = torch.rand(24, 3, 32, 32, generator=torch.Generator().manual_seed(4))
    np.random.seed(7)
    want = [D.patching_train(x[i], x) for i in range(24)]
    np.random.seed(7)
    got = [T.patching_train(x[i], x, 3, 32, StandInAugment()) for i in range(24)]
    assert all(np.array_equal(a, b) for a, b in zip(want, got))          # every trigger type occurs in 24 draws
    monkeypatch.setattr(T, "dct_2d", lambda q: O.dct_2d(q))               # CPU stand-in for the device launch
    opt = types.SimpleNamespace(input_channel=3, input_height=32, input_width=32, device="cpu")
    for shuffle in (True, False):
        np.random.seed(7)
        random.seed(7)
        _, coef, y = D.make_detector_batch(x[:8], shuffle=shuffle)
        np.random.seed(7)
        random.seed(7)
        coef2, y2 = T.make_batch(x[:8], opt, shuffle, StandInAugment())
        assert torch.equal(y, y2) and y2.dtype == torch.int64
        assert float((coef2.double() - coef.double()).abs().max() / coef.abs().max()) < 1e-6


def test_get_model_rejects_other_detectors():
    import pytest

    import combat_b200.defenses.frequency_based.train as T
    with pytest.raises(NotImplementedError):
        T.get_model(types.SimpleNamespace(model="vgg13", input_channel=3, input_height=32, device="cuda"))
