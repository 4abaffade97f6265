"""oracle/detector_oracle.py (CPU restatement of defenses/frequency_based/train.py) against
tests/golden/detector_b8x2.npz, recorded from the UNMODIFIED reference train() / eval() by
tests/golden/make_golden_detector.py.  Integer work (trigger selection, quantisation, labels, shuffle) bit-exact; floats to
1e-5 relative (they are bit-equal on the build container's torch / thread count)."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import combat_oracle as O
from oracle import detector_oracle as D

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(HERE, "golden", "detector_b8x2.npz"))


def rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _state(seed):
    gen = torch.Generator().manual_seed(seed)
    p, b = O.init_frequency_model_state(gen)
    xs = [torch.rand(8, 3, 32, 32, generator=gen) for _ in range(3)]
    return {"p": p, "b": b, "opt": {}}, xs


def test_two_training_iterations_match_reference(fx):
    torch.set_num_threads(8)
    seed = int(fx["seed"])
    state, xs = _state(seed)
    assert torch.equal(torch.stack(xs), torch.from_numpy(fx["x"]))
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed + 1)
    for i in range(2):
        r = D.detector_train_step(state, xs[i])
        assert torch.equal(r["y_final"], torch.from_numpy(fx["y_final%d" % i])), "labels / shuffle order"
        # DCT coefficients of the quantised (clean, patched) batch: exact integers in, float64 transform, float32 out
        assert rel(r["x_final"], fx["x_final%d" % i]) < 1e-6
        assert rel(r["preds"], fx["preds%d" % i]) < 1e-5
        assert abs(r["loss"] - float(fx["loss%d" % i])) < 1e-5
    sd = {**state["p"], **state["b"]}
    for name, s, l2 in zip(fx["final_names"], fx["final_sum"], fx["final_l2"]):
        t = sd[str(name)].double()
        assert abs(float(t.norm()) - l2) < 1e-5 * max(1.0, l2), name
        assert abs(float(t.sum()) - s) < 1e-4 * max(1.0, abs(s), l2), name
    for k in ("conv1.weight", "conv1.bias", "bn1.weight", "bn1.running_mean", "bn1.running_var", "bn6.running_var", "linear6.weight",
              "linear6.bias"):
        assert rel(sd[k], fx["final." + k]) < 1e-5, k
    st = state["opt"]["linear6.weight"]
    assert rel(st["square_avg"], fx["adadelta.linear6.weight.square_avg"]) < 1e-5
    assert rel(st["acc_delta"], fx["adadelta.linear6.weight.acc_delta"]) < 1e-5
    assert int(state["b"]["bn1.num_batches_tracked"]) == int(fx["num_batches_tracked"]) == 2
    # the eval() batch that followed in the reference run (same RNG streams, no shuffle, eval-mode network)
    e = D.detector_eval_batch(state, xs[2])
    assert rel(e["x_final"], fx["eval_x_final"]) < 1e-6
    assert rel(e["preds"], fx["eval_preds"]) < 1e-5
    assert e["y_final"].tolist() == [0] * 8 + [1] * 8


def test_batch_construction_properties():
    """What the CUDA path for this row has to reproduce from the uint8 planes alone: the coefficients are the orthonormal
    2-D DCT of the quantised planes (combat_oracle.dct_2d of the uint8 tensor is the same transform), clean rows first."""
    np.random.seed(3)
    random.seed(3)
    x = torch.rand(6, 3, 32, 32, generator=torch.Generator().manual_seed(3))
    q, coef, y = D.make_detector_batch(x, shuffle=False)
    assert q.dtype == np.uint8 and q.shape == (12, 3, 32, 32) and y.tolist() == [0] * 6 + [1] * 6
    assert np.array_equal(q[:6], (x.numpy().astype(np.float64) * 255).astype(np.uint8))      # truncation, not rounding
    assert rel(O.dct_2d(torch.from_numpy(q)), coef) < 5e-6
    assert rel(coef[0, 0], D.dct2(q[0, 0])) < 1e-6
    # Parseval on the float64 transform of one plane
    assert abs(float((D.dct2(q[3, 1]) ** 2).sum()) - float((q[3, 1].astype(np.float64) ** 2).sum())) < 1e-6 * float((q[3, 1].astype(np.float64) ** 2).sum())


def test_patching_rng_order_and_ranges():
    """train.py:106-143: every synthetic trigger stays in [0, 1] (white / noise blocks, blends clipped at 1) and differs from
    its clean image; the number of numpy draws depends on the attack type, so two runs from one seed are identical."""
    x = torch.rand(16, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    outs = []
    for _ in range(2):
        np.random.seed(9)
        outs.append([D.patching_train(x[i], x) for i in range(16)])
    for a, b, i in zip(outs[0], outs[1], range(16)):
        assert np.array_equal(a, b)
        assert a.shape == (32, 32, 3) and a.min() >= 0.0 and a.max() <= 1.0
        assert not np.array_equal(a, x[i].numpy().transpose(1, 2, 0))


def test_adadelta_matches_torch_optimizer():
    g = torch.Generator().manual_seed(2)
    w = torch.randn(7, 5, generator=g)
    ref = torch.nn.Parameter(w.clone())
    optim = torch.optim.Adadelta([ref], lr=0.05, weight_decay=1e-4)
    params, st = {"w": w.clone()}, {}
    for _ in range(3):
        grad = torch.randn(7, 5, generator=g)
        ref.grad = grad.clone()
        optim.step()
        D.adadelta_step(params, {"w": grad}, st)
    assert torch.equal(params["w"], ref.detach())
