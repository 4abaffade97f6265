"""GPU parity of whole networks (explicit forward/backward graphs over the C-ABI kernels) against the CPU oracle
(oracle/combat_oracle.py, itself pinned to the unmodified reference) on identical seeded weights and inputs.

float32 path (CUDA-core convs): outputs 2e-5, gradients 2e-4 relative (max-abs over max-abs).
bf16 path (tcgen05 convs, bf16 activations, fp32 statistics/accumulators/master weights): outputs 3e-2,
gradients 6e-2 -- the measured bf16 rounding floor of 17-layer nets; see DESIGN.md "tolerances"."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import combat_oracle as O  # noqa: E402

MODES = [("fp32", torch.float32, 2e-5, 2e-4), ("bf16", torch.bfloat16, 3e-2, 6e-2)]


def rel(a, b):
    a = a.detach().float().cpu().double()
    b = b.detach().float().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")


@pytest.mark.parametrize("mode", MODES, ids=[m[0] for m in MODES])
@pytest.mark.parametrize("arch,size,ncls", [("preact_resnet18", 32, 10), ("resnet18", 64, 8)])
def test_classifier_train_and_eval(mode, arch, size, ncls):
    need_gpu()
    from combat_b200.nets import Classifier
    _, dtype, tol_out, tol_grad = mode
    gen = torch.Generator().manual_seed(11)
    scaler = {32: 1, 64: 4}[size]
    init = O.init_preact_resnet18_state if arch == "preact_resnet18" else O.init_resnet18_state
    fwd = O.CLASSIFIERS[arch]
    p, b = init(gen, num_classes=ncls, scaler=scaler)
    # non-trivial BN affine + running stats so every term of the BN backward is exercised
    for k in p:
        if ".bn" in k or k.startswith("bn") or "shortcut.1" in k:
            p[k] = p[k] + 0.2 * torch.randn(p[k].shape, generator=gen)
    B = 8 if size == 32 else 4
    x = torch.rand(B, 3, size, size, generator=gen) * 2 - 1
    t = torch.randint(0, ncls, (B,), generator=gen)
    net = Classifier(arch, ncls, 3, size, device="cuda", dtype=dtype)
    sd = dict(p)
    sd.update(b)
    net.load_state_dict(sd)
    # ---- train mode: logits, loss, all parameter gradients, input gradient, running statistics
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    br = {k: v.clone() for k, v in b.items()}
    xr = x.clone().requires_grad_(True)
    logits_ref = fwd(pr, br, xr, True)
    loss_ref = F.cross_entropy(logits_ref, t)
    loss_ref.backward()
    from combat_b200 import ops
    xd = x.cuda()
    logits, ctx = net.forward(xd, train=True, save=True)
    assert rel(logits, logits_ref) < tol_out
    loss, dl, _ = ops.cross_entropy(logits, t.cuda(), 1.0, True)
    assert abs(float(loss) - float(loss_ref)) < tol_out * 3
    net.zero_grad()
    dx = net.backward(ctx, dl, need_wgrad=True, need_dx=True)
    assert rel(dx, xr.grad) < tol_grad
    worst = 0.0
    for k in p:
        e = rel(net.store.g(k), pr[k].grad)
        worst = max(worst, e)
        assert e < tol_grad, (k, e)
    sd2 = net.state_dict()
    for k in br:
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert rel(sd2[k], br[k]) < max(tol_out, 1e-5), k
    # ---- eval mode on the updated running statistics, dgrad-only backward (the G-step use)
    xr2 = x.clone().requires_grad_(True)
    with torch.no_grad():
        pe = {k: v.detach() for k, v in pr.items()}
    logits_e_ref = fwd(pe, br, xr2, False)
    F.cross_entropy(logits_e_ref, t).backward()
    logits_e, ctx_e = net.forward(xd, train=False, save=True)
    assert rel(logits_e, logits_e_ref) < tol_out
    _, dl_e, _ = ops.cross_entropy(logits_e, t.cuda(), 1.0, True)
    dx_e = net.backward(ctx_e, dl_e, need_wgrad=False, need_dx=True)
    assert rel(dx_e, xr2.grad) < tol_grad


@pytest.mark.parametrize("mode", MODES, ids=[m[0] for m in MODES])
@pytest.mark.parametrize("size,cond", [(32, 0), (64, 8)])
def test_generator_forward_backward(mode, size, cond):
    need_gpu()
    from combat_b200.nets import Generator
    _, dtype, tol_out, tol_grad = mode
    gen = torch.Generator().manual_seed(12 + size)
    p = O.init_unet_state(gen, num_classes=cond)
    B = 4 if size == 32 else 2
    x = torch.rand(B, 3, size, size, generator=gen) * 2 - 1
    lab = torch.randint(0, max(cond, 1), (B,), generator=gen) if cond else None
    w = torch.rand(B, 3, size, size, generator=gen) - 0.5
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    y_ref = O.unet_forward(pr, x, lab, cond if cond else None)
    (y_ref * w).sum().backward()
    net = Generator(3, 64, cond, device="cuda", dtype=dtype)
    net.load_state_dict(p)
    y, ctx = net.forward(x.cuda(), lab.cuda() if cond else None, save=True)
    assert rel(y, y_ref) < tol_out
    net.zero_grad()
    net.backward(ctx, w.cuda())
    for k in p:
        g_ref = pr[k].grad
        if k.endswith(".bias") and k not in ("conv0_0.bias", "upconv0_0.bias"):
            # bias feeding a non-affine InstanceNorm: the true gradient is 0, both sides hold fp noise
            scale = float(pr[k.replace(".bias", ".weight")].grad.abs().max())
            assert float(net.store.g(k).abs().max()) < 1e-3 * max(scale, 1e-6) * (100 if dtype == torch.bfloat16 else 1), k
            continue
        assert rel(net.store.g(k), g_ref) < tol_grad, k


def test_generator_golden_fixture(golden):
    """UnetGenerator under seed 2: weights re-derived by the oracle's bit-identical init, output compared with the
    output of the reference module recorded in tests/golden/modules.npz."""
    need_gpu()
    from combat_b200.nets import Generator
    g = golden("modules.npz")
    p = O.init_unet_state(torch.Generator().manual_seed(2))
    net = Generator(3, 64, 0, device="cuda", dtype=torch.float32)
    net.load_state_dict(p)
    y, ctx = net.forward(torch.from_numpy(g["unet_x"]).cuda(), save=True)
    assert rel(y, torch.from_numpy(g["unet_y"])) < 2e-5
    net.zero_grad()
    net.backward(ctx, torch.from_numpy(g["unet_w"]).cuda())
    for k in ("conv0_0.weight", "upconv0_0.weight"):
        assert rel(net.store.g(k), torch.from_numpy(g["unet_gfull_" + k])) < 2e-4


def test_preact_golden_fixture(golden):
    need_gpu()
    from combat_b200 import ops
    from combat_b200.nets import Classifier
    g = golden("modules.npz")
    p, b = O.init_preact_resnet18_state(torch.Generator().manual_seed(4))
    net = Classifier("preact_resnet18", 10, 3, 32, device="cuda", dtype=torch.float32)
    sd = dict(p)
    sd.update(b)
    net.load_state_dict(sd)
    x = torch.from_numpy(g["preact_x"]).cuda()
    logits, ctx = net.forward(x, train=True, save=True)
    assert rel(logits, torch.from_numpy(g["preact_logits_train"])) < 2e-5
    loss, dl, _ = ops.cross_entropy(logits, torch.from_numpy(g["preact_t"]).cuda(), 1.0, True)
    assert abs(float(loss) - float(g["preact_loss"])) < 2e-5
    net.zero_grad()
    dx = net.backward(ctx, dl, True, True)
    assert rel(dx, torch.from_numpy(g["preact_dx"])) < 2e-4
    for k in ("conv1.weight", "linear.weight", "layer2.0.shortcut.0.weight"):
        assert rel(net.store.g(k), torch.from_numpy(g["preact_gfull_" + k])) < 2e-4
    le, _ = net.forward(x, train=False, save=False)
    assert rel(le, torch.from_numpy(g["preact_logits_eval"])) < 2e-5


def test_frequency_detector_shipped_weights(golden):
    """dct_2d(uint8) -> FrequencyModel with the weights the reference ships: known-answer test (SURVEY 8c pin 3)."""
    need_gpu()
    from combat_b200 import ops
    from combat_b200.nets import FrequencyDetector
    g = golden("modules.npz")
    sd = {k[len("freq_sd_"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("freq_sd_")}
    net = FrequencyDetector(2, 3, 32, device="cuda")
    net.load_state_dict({k: v.cuda() for k, v in sd.items()})
    xu = torch.from_numpy(g["freq_xu"]).cuda()
    logits = net.forward(ops.plane_op(xu, "dct", in_mode=1))
    assert rel(logits, torch.from_numpy(g["freq_logits"])) < 5e-5
