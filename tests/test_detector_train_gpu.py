"""Frequency-detector TRAINING iteration on the CUDA path (combat_b200.defenses.frequency_based.train, float32 CUDA-core
kernels) against oracle/detector_oracle.py and the fixture recorded from the unmodified reference train().

"""
import random
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import combat_oracle as O  # noqa: E402
from oracle import detector_oracle as D  # noqa: E402


class StandInAugment:
    def addnoise(self, img):
        return D.addnoise(img)

    def randshadow(self, img, input_size=32):
        return D.randshadow(img, input_size)


def rel(a, b):
    a, b = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a)).double(), torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_training_primitives(golden):
    """elu_bwd, maxpool2_bwd, mask_scale, adadelta against torch-CPU on seeded inputs."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import torch.nn.functional as F

    from combat_b200 import ops
    g = torch.Generator().manual_seed(3)
    z = torch.randn(4, 6, 6, 32, generator=g, requires_grad=True)
    a = F.elu(z)
    da = torch.randn(4, 6, 6, 32, generator=g)
    a.backward(da)
    assert rel(ops.elu_bwd(da.cuda(), a.detach().cuda()), z.grad) < 1e-6
    x = torch.randn(3, 8, 8, 16, generator=g)
    x[0, 0, 0, 0] = x[0, 0, 1, 0] = 5.0                                   # a tie: the first position takes the gradient
    xn = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    y = F.max_pool2d(xn, 2)
    dy = torch.randn(3, 4, 4, 16, generator=g)
    y.backward(dy.permute(0, 3, 1, 2))
    assert torch.equal(ops.maxpool2_bwd(dy.cuda(), x.cuda()).cpu(), xn.grad.permute(0, 2, 3, 1).contiguous())
    keep = (torch.rand(4, 6, 6, 32, generator=g) > 0.2).to(torch.uint8)
    assert torch.equal(ops.mask_scale(da.cuda(), keep.cuda(), 1.25).cpu(), da * keep * 1.25)
    w = torch.randn(1000, generator=g)
    ref = torch.nn.Parameter(w.clone())
    optim = torch.optim.Adadelta([ref], lr=0.05, weight_decay=1e-4)
    p, v, u = w.clone().cuda(), torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda")
    lr = torch.full((1,), 0.05, device="cuda")
    for _ in range(3):
        grad = torch.randn(1000, generator=g)
        ref.grad = grad.clone()
        optim.step()
        ops.adadelta(p, grad.cuda(), v, u, lr)
    assert rel(p, ref.detach()) < 1e-6


def test_two_training_iterations_vs_oracle_and_reference_fixture(golden):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import combat_b200.defenses.frequency_based.train as T
    fx = golden("detector_b8x2.npz")
    seed = int(fx["seed"])
    gen = torch.Generator().manual_seed(seed)
    p0, b0 = O.init_frequency_model_state(gen)
    xs = [torch.rand(8, 3, 32, 32, generator=gen) for _ in range(3)]
    opt = types.SimpleNamespace(model="original", input_channel=3, input_height=32, input_width=32, device="cuda:0")
    netC, optimizerC = T.get_model(opt)
    netC.load_state_dict({**p0, **b0}, strict=False)
    lr_dev = torch.full((1,), 0.05, device="cuda:0")
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed + 1)
    netC.train()
    for i in range(2):
        loss, counts, logits = T.train_iteration(netC, xs[i], opt, lr_dev, StandInAugment())
        assert rel(logits, fx["preds%d" % i]) < 2e-4, i
        assert abs(float(loss) - float(fx["loss%d" % i])) < 1e-4
        assert int(counts[0]) == int((torch.from_numpy(fx["preds%d" % i]).argmax(1) == torch.from_numpy(fx["y_final%d" % i])).sum())
    sd = netC.state_dict()
    for k in ("conv1.weight", "conv1.bias", "bn1.weight", "bn1.running_mean", "bn1.running_var", "bn6.running_var", "linear6.weight",
              "linear6.bias"):
        assert rel(sd[k], fx["final." + k]) < 1e-3, k
    for name, l2 in zip(fx["final_names"], fx["final_l2"]):
        assert abs(float(sd[str(name)].double().norm()) - l2) < 1e-4 * max(1.0, l2), name
    netC.eval()
    coef, y = T.make_batch(xs[2], opt, False, StandInAugment())
    assert rel(coef, fx["eval_x_final"]) < 5e-6
    assert rel(netC(coef), fx["eval_preds"]) < 5e-4


def test_detector_trainer_main_on_synthetic_data(tmp_path, capsys):
    """train.py:275-344 through the mirror's main(): two epochs of train() + eval() on synthetic [0, 1] images, the detector
    checkpoint at the path train_generator.py's main() loads (F_ckpt_path), then --continue_training from it."""
    import os
    from combat_b200.defenses.frequency_based import train as ftrain
    args = ["--device", "cuda", "--synthetic_data", "--debug", "--bs", "16", "--n_iters", "2", "--checkpoints", str(tmp_path)]
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    best = ftrain.main(args, augment=StandInAugment())
    out = capsys.readouterr().out
    path = tmp_path / "cifar10" / "original" / "cifar10_original_detector.pth.tar"
    assert "CE Loss" in out and "Acc:" in out and os.path.exists(path) and 0.0 <= best <= 100.0
    ck = torch.load(str(path), map_location="cpu", weights_only=False)
    assert {"netC", "optimizerC", "best_acc", "epoch_current"} == set(ck) and "linear6.weight" in ck["netC"]
    ftrain.main(args + ["--continue_training", "--n_iters", "3"], augment=StandInAugment())
    assert "Continue training!!" in capsys.readouterr().out
    # the fused Adadelta adopts the accumulators a checkpoint restores into the torch optimiser's state
    opt = types.SimpleNamespace(device="cuda", model="original", input_channel=3, input_height=32)
    netC, optC = ftrain.get_model(opt)
    optC.load_state_dict(ck["optimizerC"])
    ftrain._adopt_adadelta_state(optC, netC)
    st = netC.net.store
    name, p0 = netC._plist[0]
    saved = ck["optimizerC"]["state"][0]
    assert float(saved["square_avg"].abs().sum()) > 0
    assert torch.allclose(st._view(st.mom, name).cpu(), saved["square_avg"].cpu().float())
    assert torch.allclose(st._view(st.acc_delta, name).cpu(), saved["acc_delta"].cpu().float())
