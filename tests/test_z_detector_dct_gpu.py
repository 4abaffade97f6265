"""The detector-defense DCT call site (defenses/frequency_based/train.py:37-38, 195-197: per-plane scipy dct2 of
`(plane * 255).astype(uint8)`) on the CUDA path: the whole (clean, patched) batch in ONE launch of the uint8-input DCT
kernel through the public `combat_b200.utils.dct.dct_2d`, against the coefficients the UNMODIFIED reference fed to its
network (tests/golden/detector_b8x2.npz).  The uint8 planes are rebuilt by the oracle from the fixture's seed (integer work,
bit-exact; the CPU suite checks that rebuild against the same fixture)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import detector_oracle as D  # noqa: E402


def test_detector_batch_dct_vs_reference_fixture(golden):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from combat_b200.utils.dct import dct_2d
    fx = golden("detector_b8x2.npz")
    seed = int(fx["seed"])
    xs = torch.from_numpy(fx["x"])
    np.random.seed(seed)
    random.seed(seed)
    for i in range(2):
        q, coef, y = D.make_detector_batch(xs[i], shuffle=True)
        assert torch.equal(y, torch.from_numpy(fx["y_final%d" % i]))
        got = dct_2d(torch.from_numpy(q).cuda())
        assert got.dtype == torch.float32 and got.shape == (16, 3, 32, 32)
        ref = torch.from_numpy(fx["x_final%d" % i]).double()
        # float32 butterfly network against scipy's float64 transform rounded to float32; the DC term is up to 255 * 32
        assert float((got.cpu().double() - ref).abs().max() / ref.abs().max()) < 5e-6
