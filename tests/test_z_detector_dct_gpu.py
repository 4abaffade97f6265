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


def test_detector_test_script_vs_oracle(tmp_path, capsys):
    """defenses/frequency_based/test.py mirror: test_batch against the oracle restatement of :76-103 (sigma draw bit-exact, uint8
    DCT coefficients exact up to float32 rounding, logits, both counters), then main() on synthetic data."""
    import random

    import numpy as np

    from combat_b200.defenses.frequency_based import config as fconfig
    from combat_b200.defenses.frequency_based import test as ftest
    from combat_b200.modules import UnetGenerator
    from oracle import combat_oracle as O
    opt = fconfig.get_arguments().parse_args(["--device", "cuda"])
    torch.manual_seed(3); np.random.seed(3); random.seed(3)
    netC, _ = ftest.get_model(opt)
    netG = UnetGenerator(opt, device="cuda", dtype=torch.float32)
    sd = lambda m: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    netF_p, netF_b = O.split_state(sd(netC))
    o = O.default_opt()
    g = torch.Generator().manual_seed(5)
    for it in range(2):
        x = torch.rand(24, 3, 32, 32, generator=g) * 2 - 1
        torch.manual_seed(60 + it)
        r = O.detector_test_batch(netF_p, netF_b, sd(netG), x, o)
        torch.manual_seed(60 + it)
        counts, d = ftest.test_batch(netC, netG, x, opt)
        c = counts.cpu().numpy()
        assert d["sigma"] == r["sigma"]
        # a poisoned pixel within float32 rounding of an integer boundary moves one uint8 step: compare coefficients on the
        # clean half exactly (same uint8 input), the poisoned half and the logits at the level of such flips
        assert float((d["coef"][:24].cpu() - r["coef"][:24]).abs().max()) < 2e-2
        assert float((d["poi_x"].cpu() - r["poi_x"]).abs().max()) < 2e-5
        assert float((d["preds"].cpu() - r["preds"]).abs().max() / r["preds"].abs().max()) < 2e-3
        assert abs(int(c[0]) - r["correct"]) <= 1 and abs(int(c[2]) - r["detected"]) <= 1
    acc, det = ftest.main(["--device", "cuda", "--synthetic_data", "--debug", "--bs", "32", "--saving_prefix", "none",
                           "--checkpoints", str(tmp_path), "--load_checkpoint", str(tmp_path)])
    assert "Detection rate" in capsys.readouterr().out and 0.0 <= acc <= 100.0 and 0.0 <= det <= 100.0
