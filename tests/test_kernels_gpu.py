"""GPU parity tests of the individual kernels, called through the C-ABI (combat_b200.ops -> libcombat_b200.so),
against the CPU oracle / torch-CPU fp32 on identical seeded inputs.  Integer results bit-exact; float32 kernels
1e-5 relative; bf16-storage kernels 1e-2 (bf16 has 8 mantissa bits; the statistic compared is max-abs error over
max-abs reference)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import combat_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from combat_b200 import ops as _ops
    return _ops


def rel(a, b):
    a = a.detach().float().cpu().double() if torch.is_tensor(a) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().float().cpu().double() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def dev(t):
    return t.cuda().contiguous()


# ------------------------------------------------------------------ DCT family
@pytest.mark.parametrize("N,fast", [(32, True), (32, False), (64, True), (64, False), (28, False), (8, False)])
def test_dct_idct_lowfreq(ops, N, fast):
    g = torch.Generator().manual_seed(N)
    x = torch.rand(5, 3, N, N, generator=g) * 255
    d = ops.plane_op(dev(x), "dct", fast=fast)
    assert rel(d, O.dct_2d_exact(x.numpy())) < 2e-6
    assert rel(d, O.dct_2d(x)) < 5e-6
    i = ops.plane_op(dev(x), "idct", fast=fast)
    assert rel(i, O.idct_2d_exact(x.numpy())) < 2e-6
    # round trip property
    assert rel(ops.plane_op(d, "idct", fast=fast), x) < 2e-6
    xn = torch.rand(5, 3, N, N, generator=g) * 2 - 1
    keep = int(N * 0.65)
    lf = ops.plane_op(dev(xn), "lowfreq", keep=keep, fast=fast)
    assert float((lf.cpu().double() - torch.from_numpy(O.low_freq_exact(xn.numpy(), N, 0.65))).abs().max()) < 2e-6
    assert float((lf.cpu() - O.low_freq(xn, N, 0.65)).abs().max()) < 2e-5
    # idempotence of the projection
    assert float((ops.plane_op(lf, "lowfreq", keep=keep, fast=fast) - lf).abs().max()) < 2e-6
    # other retained block sizes (the pruned instantiation is chosen by `keep`): below, at and above the default
    for k2 in (1, keep - 3, keep + 1, N):
        ratio = (k2 + 0.5) / N
        lf2 = ops.plane_op(dev(xn), "lowfreq", keep=k2, fast=fast)
        assert int(N * ratio) == k2
        assert float((lf2.cpu().double() - torch.from_numpy(O.low_freq_exact(xn.numpy(), N, ratio))).abs().max()) < 2e-6, k2


@pytest.mark.parametrize("N,fast", [(32, True), (32, False), (64, True), (64, False)])
def test_dct_uint8_paths(ops, N, fast):
    g = torch.Generator().manual_seed(100 + N)
    xu = (torch.rand(4, 3, N, N, generator=g) * 256).clamp(0, 255).byte()
    d = ops.plane_op(dev(xu), "dct", in_mode=1, fast=fast)
    assert d.dtype == torch.float32
    assert rel(d, O.dct_2d(xu)) < 5e-6
    xf = torch.rand(4, 3, N, N, generator=g) * 2 - 1
    xf[0, 0, 0, :4] = torch.tensor([1.0, -1.0, 0.999, -0.999])  # truncation edge cases (SURVEY trap 5)
    q = ((xf + 1) / 2 * 255).byte()
    d2 = ops.plane_op(dev(xf), "dct", in_mode=2, fast=fast)
    assert rel(d2, O.dct_2d(q)) < 5e-6


def test_dct_golden_fixture(ops, golden):
    g = golden("dct.npz")
    for N in (32, 64):
        x = torch.from_numpy(g["x%d" % N])
        assert rel(ops.plane_op(dev(x), "dct"), g["dct%d" % N]) < 5e-6
        assert rel(ops.plane_op(dev(x), "idct"), g["idct%d" % N]) < 5e-6
        assert rel(ops.plane_op(dev(torch.from_numpy(g["xu%d" % N])), "dct", in_mode=1), g["dctu%d" % N]) < 5e-6
        lf = ops.plane_op(dev(torch.from_numpy(g["xn%d" % N])), "lowfreq", keep=int(N * 0.65))
        assert float((lf.cpu() - torch.from_numpy(g["lowfreq%d" % N])).abs().max()) < 2e-5
    assert rel(ops.plane_op(dev(torch.from_numpy(g["x28"])), "dct"), g["dct28"]) < 5e-6


def test_dct_empty_and_large(ops):
    e = ops.plane_op(torch.empty(0, 3, 32, 32, device="cuda"), "dct")
    assert e.numel() == 0
    # size-independent property at scale: Parseval (orthonormal transform preserves the energy) and round trip
    x = torch.rand(4096, 3, 32, 32, device="cuda")
    d = ops.plane_op(x, "dct")
    assert abs(float((d.double() ** 2).sum() / (x.double() ** 2).sum()) - 1) < 1e-6
    assert float((ops.plane_op(d, "idct") - x).abs().max()) < 5e-6


# ------------------------------------------------------------------ poisoned-batch builder
@pytest.mark.parametrize("H", [32, 64, 12])
def test_poison_blend_fwd_bwd(ops, H):
    g = torch.Generator().manual_seed(H)
    B, Cc = 10, 3
    x = torch.rand(B, Cc, H, H, generator=g) * 2 - 1
    noise = (torch.rand(B, Cc, H, H, generator=g) * 2 - 1) * 8  # large so that the clamp is active
    sigma = 0.37
    taps = ops.gaussian_taps(sigma)
    # G-step form: all rows, identity order
    nz = noise.clone().requires_grad_(True)
    ref = O.gaussian_blur(torch.clamp(x + nz * 0.08, -1, 1), sigma)
    sq = torch.empty(B * Cc, device="cuda")
    out = ops.poison_blend_fwd(dev(x), dev(noise), None, B, 0.08, taps, sq_partial=sq)
    assert rel(out, ref) < 1e-6
    assert abs(float(sq.double().sum()) / x.numel() - float(F.mse_loss(ref, x))) < 1e-6
    # backward: d/dnoise of  sum(g1*x_bd) + sum(g2*x_bd) + w*MSE(x_bd, x)
    g1 = torch.rand(B, Cc, H, H, generator=g) - 0.5
    g2 = torch.rand(B, Cc, H, H, generator=g) - 0.5
    w = 0.02
    ((ref * (g1 + g2)).sum() + w * F.mse_loss(ref, x)).backward()
    dn = ops.poison_blend_bwd(dev(x), dev(noise), out, dev(g1), dev(g2), 2.0 * w / x.numel(), 0.08, taps)
    assert rel(dn, nz.grad) < 1e-5
    # C-step form: gather + pass-through rows, device-resident dynamic parameters
    perm = torch.tensor([7, 2, 9, 0, 1, 3, 4, 5, 6, 8], dtype=torch.int32)
    num_bd = 3
    exp = torch.cat([O.gaussian_blur(torch.clamp(x[perm[:num_bd].long()] + noise[perm[:num_bd].long()] * 0.08, -1, 1), sigma),
                     x[perm[num_bd:].long()]])
    out2 = ops.poison_blend_fwd(dev(x), dev(noise), dev(perm), 0, 0.08, None,
                                taps_dev=torch.tensor(taps, device="cuda"), num_bd_dev=torch.tensor([num_bd], dtype=torch.int32, device="cuda"))
    assert rel(out2, exp) < 1e-6
    assert torch.equal(out2[num_bd:].cpu(), x[perm[num_bd:].long()])  # pass-through rows are bit-exact copies
    # num_bd == 0 (empty poison set): pure permutation
    out3 = ops.poison_blend_fwd(dev(x), None, dev(perm), 0, 0.08, taps)
    assert torch.equal(out3.cpu(), x[perm.long()])


# ------------------------------------------------------------------ losses / optimiser
def test_cross_entropy(ops):
    g = torch.Generator().manual_seed(3)
    for B, Cn in [(128, 10), (37, 8), (512, 2)]:
        logits = (torch.randn(B, Cn, generator=g) * 3).requires_grad_(True)
        t = torch.randint(0, Cn, (B,), generator=g)
        t2 = torch.randint(0, Cn, (B,), generator=g)
        loss = F.cross_entropy(logits, t)
        (0.8 * loss).backward()
        lo, dl, cnt = ops.cross_entropy(dev(logits.detach()), dev(t), 0.8, True, targets2=dev(t2))
        assert abs(float(lo) - float(loss)) < 1e-5 * max(1.0, abs(float(loss)))
        assert rel(dl, logits.grad) < 1e-5
        am = logits.argmax(1)
        assert cnt.cpu().tolist() == [int((am == t).sum()), int((am == t2).sum())]  # bit-exact counters


def test_sgd_nesterov(ops):
    g = torch.Generator().manual_seed(4)
    n = 10007
    p = torch.randn(n + 1, generator=g)[:n]
    n4 = (n + 3) // 4 * 4
    pd, gd, bd = torch.zeros(n4, device="cuda"), torch.zeros(n4, device="cuda"), torch.zeros(n4, device="cuda")
    pd[:n] = p.cuda()
    lr = torch.tensor([1e-2], device="cuda")
    params, bufs = {"w": p.clone()}, {}
    for it in range(3):
        gr = torch.randn(n, generator=g)
        gd[:n] = gr.cuda()
        ops.sgd_nesterov(pd, gd, bd, lr, 0.9, 5e-4, it == 0)
        O.sgd_nesterov_step(params, {"w": gr}, bufs, 1e-2)
        assert rel(pd[:n], params["w"]) < 1e-6
        assert rel(bd[:n], bufs["w"]) < 1e-6


# ------------------------------------------------------------------ convolutions
CONV_CASES = [
    # N, Ci, Co, H, k, stride, pad
    (4, 3, 64, 32, 3, 1, 1),      # first conv of the classifiers (K = 27)
    (4, 3, 64, 32, 3, 2, 1),      # conv0_0 of the generator
    (3, 64, 3, 16, 3, 1, 1),      # upconv0_0 (Cout = 3)
    (2, 64, 128, 16, 3, 2, 1),    # strided 3x3
    (2, 64, 128, 16, 1, 2, 0),    # 1x1 stride-2 shortcut
    (5, 72, 64, 8, 3, 1, 1),      # CUnet conv0_1 (72 input channels)
    (2, 128, 128, 8, 3, 1, 1),
]


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_simt_fp32(ops, case):
    N, Ci, Co, H, k, s, p = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, Ci, H, H, generator=g, requires_grad=True)
    w = (torch.randn(Co, Ci, k, k, generator=g) * 0.1).requires_grad_(True)
    bias = torch.randn(Co, generator=g)
    y = F.conv2d(x, w, bias, s, p)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    Ho = y.shape[2]
    w_ohwi = dev(w.detach().permute(0, 2, 3, 1))
    xd = dev(_nhwc(x.detach()))
    out = torch.empty(N, Ho, Ho, Co, device="cuda")
    ops.conv_simt(xd, (N, H, H), ops.nhwc_strides(H, H, Ci), w_ohwi, 0, out, (Ho, Ho), ops.nhwc_strides(Ho, Ho, Co),
                  Ci=Ci, Co=Co, KH=k, KW=k, stride=s, pad=p, bias=dev(bias))
    assert rel(out.permute(0, 3, 1, 2), y) < 1e-5
    # NCHW input variant
    out2 = torch.empty(N, Ho, Ho, Co, device="cuda")
    ops.conv_simt(dev(x.detach()), (N, H, H), ops.nchw_strides(Ci, H, H), w_ohwi, 0, out2, (Ho, Ho),
                  ops.nhwc_strides(Ho, Ho, Co), Ci=Ci, Co=Co, KH=k, KW=k, stride=s, pad=p, bias=dev(bias))
    assert rel(out2, out) < 1e-6
    # dgrad: conv over dy with flipped/transposed weights, up = stride
    w_d = dev(w.detach().flip(2, 3).permute(1, 2, 3, 0))  # [ci][kh'][kw'][co]
    dx = torch.empty(N, H, H, Ci, device="cuda")
    dyd = dev(_nhwc(dy))
    ops.conv_simt(dyd, (N, Ho, Ho), ops.nhwc_strides(Ho, Ho, Co), w_d, 0, dx, (H, H), ops.nhwc_strides(H, H, Ci),
                  Ci=Co, Co=Ci, KH=k, KW=k, stride=1, pad=k - 1 - p, up=s)
    assert rel(dx.permute(0, 3, 1, 2), x.grad) < 1e-5
    # wgrad (+ bias gradient), channels-last gradient layout
    dw = torch.zeros(Co, k, k, Ci, device="cuda")
    db = torch.zeros(Co, device="cuda")
    ops.conv_wgrad_simt(xd, (N, H, H), ops.nhwc_strides(H, H, Ci), dyd, (Ho, Ho), ops.nhwc_strides(Ho, Ho, Co), dw,
                        Ci=Ci, Co=Co, KH=k, KW=k, stride=s, pad=p, db=db)
    assert rel(dw.permute(0, 3, 1, 2), w.grad) < 2e-5
    assert rel(db, dy.sum((0, 2, 3))) < 2e-5


TC_CASES = [
    # N, Ci, Co, H, k, stride, pad
    (8, 64, 64, 32, 3, 1, 1),     # layer1: 128-pixel tile = 4 rows of one image
    (8, 64, 128, 32, 3, 2, 1),    # layer2.0.conv1 (parity views)
    (8, 64, 128, 32, 1, 2, 0),    # layer2.0.shortcut
    (8, 128, 128, 16, 3, 1, 1),
    (8, 256, 256, 8, 3, 1, 1),    # tile spans 2 images
    (16, 512, 512, 4, 3, 1, 1),   # tile spans 8 images
    (40, 512, 512, 2, 3, 1, 1),   # UNet bottleneck 2x2: tile spans 32 images, ragged last tile
    (3, 128, 64, 16, 3, 1, 1),    # ragged batch (3 images of 256 pixels -> 6 tiles)
    (2, 64, 64, 28, 3, 1, 1),     # non power-of-two plane (masked tile columns)
    (5, 64, 64, 16, 3, 1, 1),     # resident-filter / row-reuse kernel, tile = 8 rows of 16
    (300, 64, 64, 32, 3, 1, 1),   # same kernel, more tiles than SMs (persistent loop, pipeline wrap-around)
    (2, 64, 64, 24, 3, 1, 1),     # 8 | W but H not a multiple of the tile height
    # CTA-pair kernels (cta_group::2, Co >= 128) and the shapes bench.py actually runs
    (5, 256, 256, 8, 3, 1, 1),    # 3 pixel tiles: the last pair's second half is padding
    (9, 128, 384, 8, 3, 1, 1),    # three 128-wide channel tiles per pixel-tile pair
    (512, 64, 64, 32, 3, 1, 1),   # layer1 at the benchmark batch
    (1024, 256, 256, 8, 3, 1, 1),  # the batched [x ; x_bd] forward: 256-wide tiles, 512 tiles on 74 clusters
    (512, 128, 256, 16, 3, 2, 1),  # stride 2 at the benchmark batch (parity views; dgrad = 4 parity classes, paired per class)
    (512, 256, 512, 8, 1, 2, 0),   # 1x1 stride-2 shortcut: dgrad classes without taps
    # row-reuse kernel for 128 output channels (conv_tc_rr_kernel: super-tiles of two pixel tiles sharing each weight tile)
    (150, 128, 128, 16, 3, 1, 1),  # 150 super-tiles on 148 CTAs: the 2 left over are split into 4 single tiles
    (223, 128, 128, 16, 3, 1, 1),  # 75 left over: too many to split, 75 CTAs run a second full item
    (512, 128, 128, 16, 3, 1, 1),  # layer2 at the benchmark batch: 3 full rounds + 68 split super-tiles
    (4, 192, 128, 32, 3, 1, 1),    # three K chunks, 32-wide rows (tile = 4 rows, box = 10 rows), 4 super-tiles per image
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_bf16(ops, case):
    import ctypes as C

    from combat_b200._lib import check, lib
    N, Ci, Co, H, k, s, p = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, Ci, H, H, generator=g).bfloat16().float().requires_grad_(True)
    w = (torch.randn(Co, Ci, k, k, generator=g) * 0.05).bfloat16().float().requires_grad_(True)
    bias = torch.randn(Co, generator=g)
    y = F.conv2d(x, w, bias, s, p)
    dy = torch.randn(y.shape, generator=g).bfloat16().float()
    res = torch.randn(y.shape, generator=g).bfloat16().float()
    y.backward(dy)
    Ho = y.shape[2]
    xd = dev(_nhwc(x.detach()).bfloat16())
    w_f = dev(w.detach().permute(0, 2, 3, 1).bfloat16())
    out = torch.empty(N, Ho, Ho, Co, device="cuda", dtype=torch.bfloat16)
    resd = dev(_nhwc(res).bfloat16())
    d = ops.conv_tc_desc(xd, w_f.data_ptr(), out, N, H, H, Ci, Ho, Ho, Co, k, k, s, p, 1, bias=dev(bias), residual=resd)
    assert lib.combat_conv_tc_supported(C.byref(d))
    check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc")
    torch.cuda.synchronize()
    # inputs are exactly representable in bf16, accumulation is fp32: only the bf16 rounding of the OUTPUT remains
    assert rel(out.float().permute(0, 3, 1, 2), y + res) < 6e-3
    # dgrad
    w_d = dev(w.detach().flip(2, 3).permute(1, 2, 3, 0).bfloat16())
    dyd = dev(_nhwc(dy).bfloat16())
    dx = torch.empty(N, H, H, Ci, device="cuda", dtype=torch.bfloat16)
    d2 = ops.conv_tc_desc(dyd, w_d.data_ptr(), dx, N, Ho, Ho, Co, H, H, Ci, k, k, 1, k - 1 - p, s)
    assert lib.combat_conv_tc_supported(C.byref(d2))
    check(lib.combat_conv_tc(C.byref(d2), ops._s()), "conv_tc dgrad")
    torch.cuda.synchronize()
    assert rel(dx.float().permute(0, 3, 1, 2), x.grad) < 6e-3
    # wgrad
    dw = torch.zeros(Co, k, k, Ci, device="cuda")
    d3 = ops.conv_tc_desc(xd, None, None, N, H, H, Ci, Ho, Ho, Co, k, k, s, p, 1)
    check(lib.combat_conv_tc_wgrad(C.byref(d3), dyd.data_ptr(), dw.data_ptr(), ops._s()), "conv_tc_wgrad")
    torch.cuda.synchronize()
    assert rel(dw.permute(0, 3, 1, 2), w.grad) < 2e-5 * max(1.0, (N * Ho * Ho) ** 0.5 / 30)


# ------------------------------------------------------------------ normalisation / activation / pooling
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1.5e-2)])
def test_batchnorm_relu_train_eval(ops, dtype, tol):
    g = torch.Generator().manual_seed(5)
    N, Cc, H = 6, 64, 8
    x = (torch.randn(N, Cc, H, H, generator=g) * 2 + 0.5)
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    x.requires_grad_(True)
    gamma = (torch.rand(Cc, generator=g) + 0.5).requires_grad_(True)
    beta = torch.randn(Cc, generator=g).requires_grad_(True)
    rm, rv = torch.randn(Cc, generator=g) * 0.1, torch.rand(Cc, generator=g) + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    res = torch.randn(N, Cc, H, H, generator=g)
    if dtype == torch.bfloat16:
        res = res.bfloat16().float()
    y = F.relu(F.batch_norm(x, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5) + res)
    dy = torch.randn(N, Cc, H, H, generator=g)
    if dtype == torch.bfloat16:
        dy = dy.bfloat16().float()
    y.backward(dy)
    xd, resd, dyd = dev(_nhwc(x.detach()).to(dtype)), dev(_nhwc(res).to(dtype)), dev(_nhwc(dy).to(dtype))
    rmd, rvd = dev(rm), dev(rv)
    R = N * H * H
    sc, sh, mean, invstd = ops.bn_train_prepare(xd, R, Cc, dev(gamma.detach()), dev(beta.detach()), rmd, rvd, 0.1, 1e-5)
    yd = ops.affine_act(xd, sc, sh, True, residual=resd)
    assert rel(yd.float().permute(0, 3, 1, 2), y) < tol
    assert rel(rmd, rm_ref) < 1e-5 and rel(rvd, rv_ref) < 1e-5
    dg, db = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    dx, dres = ops.bn_bwd_train(dyd, xd, yd, dev(gamma.detach()), mean, invstd, True, dg, db, want_dres=True)
    assert rel(dx.float().permute(0, 3, 1, 2), x.grad) < tol
    assert rel(dg, gamma.grad) < tol and rel(db, beta.grad) < tol
    mask = (y > 0).float()
    assert rel(dres.float().permute(0, 3, 1, 2), dy * mask) < tol
    # eval mode
    x2 = x.detach().clone().requires_grad_(True)
    y2 = F.relu(F.batch_norm(x2, rm_ref, rv_ref, gamma, beta, False, 0.1, 1e-5))
    y2.backward(dy)
    sc2, sh2 = ops.bn_eval_prepare(Cc, dev(gamma.detach()), dev(beta.detach()), rmd, rvd, 1e-5)
    y2d = ops.affine_act(xd, sc2, sh2, True)
    assert rel(y2d.float().permute(0, 3, 1, 2), y2) < tol
    dadd = torch.randn(N, H, H, Cc, generator=g).to(dtype)
    dx2, _ = ops.bn_bwd_eval(dyd, y2d, sc2, True, dadd=dev(dadd))
    assert rel(dx2.float().permute(0, 3, 1, 2), x2.grad + dadd.float().permute(0, 3, 1, 2)) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize("H", [2, 4, 8, 12, 16])
def test_instancenorm_lrelu(ops, dtype, tol, H):
    g = torch.Generator().manual_seed(6 + H)
    N, Cc = 5, 128
    x = torch.randn(N, Cc, H, H, generator=g) * 3 + 1
    skip = torch.randn(N, Cc, H, H, generator=g)
    dy = torch.randn(N, Cc, H, H, generator=g)
    dy2 = torch.randn(N, Cc, H, H, generator=g)
    if dtype == torch.bfloat16:
        x, skip, dy, dy2 = [t.bfloat16().float() for t in (x, skip, dy, dy2)]
    for act, use_skip in [(True, False), (False, True)]:
        xr = x.clone().requires_grad_(True)
        y = F.instance_norm(xr, eps=1e-5)
        if act:
            y = F.leaky_relu(y, 0.2)
        if use_skip:
            y = y + skip
        y.backward(dy + dy2)
        xd = dev(_nhwc(x).to(dtype))
        yd, st = ops.instnorm_fwd(xd, act, skip=dev(_nhwc(skip).to(dtype)) if use_skip else None)
        assert rel(yd.float().permute(0, 3, 1, 2), y) < tol
        dx = ops.instnorm_bwd(dev(_nhwc(dy).to(dtype)), dev(_nhwc(dy2).to(dtype)), xd, st, act)
        # IN backward over 4 elements is ill-conditioned in bf16 storage: compare against the gradient scale
        assert float((dx.float().permute(0, 3, 1, 2).cpu() - xr.grad).abs().max() / xr.grad.abs().max()) < tol * (4 if H == 2 else 2 if H == 4 else 1)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize("N,Cc,H", [(3, 64, 112), (2, 128, 72), (5, 66, 64)])
def test_instancenorm_lrelu_large_planes_split_path(ops, dtype, tol, N, Cc, H):
    """[r2] planes of >= 4096 pixels with few (sample, channel-group) pairs (the ImageNet-10 shape) take the split kernels: chunk
    partials merged with Chan's formula.  Same bars as the single-kernel path; a large mean (+40) checks the centred merge; ragged
    last chunk (72 * 72 = 5184 rows over K chunks) and a channel count that is not a multiple of 64."""
    from combat_b200._lib import lib
    K = lib.combat_instnorm_splits(N, H * H, Cc)
    assert K > 1, K
    g = torch.Generator().manual_seed(16 + H)
    x = torch.randn(N, Cc, H, H, generator=g) * 3 + 40
    skip = torch.randn(N, Cc, H, H, generator=g)
    dy = torch.randn(N, Cc, H, H, generator=g)
    if dtype == torch.bfloat16:
        x = (x - 40).bfloat16().float()   # bf16 storage cannot carry a large offset; the float32 case does
        skip, dy = skip.bfloat16().float(), dy.bfloat16().float()
    x_all = x
    for act, use_skip in [(True, False), (False, True)]:
        # with the LeakyReLU the large offset stays out: a value within float32 rounding of the plane mean flips its mask, and
        # one flipped element moves the max-norm of the gradient by 16 % in torch's own float32 path (measured against float64)
        x = x_all - 39 if (act and dtype == torch.float32) else x_all
        xr = x.clone().requires_grad_(True)
        y = F.instance_norm(xr, eps=1e-5)
        if act:
            y = F.leaky_relu(y, 0.2)
        if use_skip:
            y = y + skip
        y.backward(dy)
        xd = dev(_nhwc(x).to(dtype))
        yd, st = ops.instnorm_fwd(xd, act, skip=dev(_nhwc(skip).to(dtype)) if use_skip else None)
        assert rel(yd.float().permute(0, 3, 1, 2), y) < tol
        assert rel(st[0].view(N, Cc), x.mean(dim=(2, 3))) < 1e-5
        dx = ops.instnorm_bwd(dev(_nhwc(dy).to(dtype)), None, xd, st, act)
        assert float((dx.float().permute(0, 3, 1, 2).cpu() - xr.grad).abs().max() / xr.grad.abs().max()) < tol
    assert lib.combat_instnorm_splits(512, 1024, 64) == 1 and lib.combat_instnorm_splits(512, 16384, 64) == 1


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 1e-2)])
def test_upsample_lrelu(ops, dtype, tol):
    g = torch.Generator().manual_seed(8)
    for H in (2, 4, 16):
        x = torch.randn(3, 64, H, H, generator=g)
        dy = torch.randn(3, 64, 2 * H, 2 * H, generator=g)
        if dtype == torch.bfloat16:
            x, dy = x.bfloat16().float(), dy.bfloat16().float()
        xr = x.clone().requires_grad_(True)
        y = F.leaky_relu(F.interpolate(xr, scale_factor=(2, 2), mode="bilinear"), 0.2)
        y.backward(dy)
        yd = ops.upsample2x_act(dev(_nhwc(x).to(dtype)))
        assert rel(yd.float().permute(0, 3, 1, 2), y) < tol
        dx = ops.upsample2x_act_bwd(dev(_nhwc(dy).to(dtype)), yd)
        assert rel(dx.float().permute(0, 3, 1, 2), xr.grad) < max(tol, 2e-6)


@pytest.mark.parametrize("Hf,ncls", [(4, 10), (8, 8)])
def test_pool_linear(ops, Hf, ncls):
    g = torch.Generator().manual_seed(9)
    B, Cc = 7, 512
    x = torch.randn(B, Cc, Hf, Hf, generator=g, requires_grad=True)
    Fdim = Cc * (Hf // 4) ** 2
    W = (torch.randn(ncls, Fdim, generator=g) * 0.05).requires_grad_(True)
    bb = torch.randn(ncls, generator=g).requires_grad_(True)
    logits = F.linear(F.avg_pool2d(x, 4).flatten(1), W, bb)
    dl = torch.randn(B, ncls, generator=g)
    logits.backward(dl)
    xd = dev(_nhwc(x.detach()))
    lo, pooled = ops.pool_linear_fwd(xd, 4, dev(W.detach()), dev(bb.detach()))
    assert rel(lo, logits) < 1e-5
    dW, db = torch.empty(ncls, Fdim, device="cuda"), torch.empty(ncls, device="cuda")
    dx = ops.pool_linear_bwd(dev(dl), pooled, dev(W.detach()), tuple(xd.shape), torch.float32, 4, dW=dW, db=db)
    assert rel(dx.permute(0, 3, 1, 2), x.grad) < 1e-5
    assert rel(dW, W.grad) < 1e-5 and rel(db, bb.grad) < 1e-5


def test_small_ops(ops):
    g = torch.Generator().manual_seed(10)
    x = torch.randn(3, 5, 6, 64, generator=g)
    assert rel(ops.leaky_relu(dev(x)), F.leaky_relu(x, 0.2)) < 1e-7
    dy = torch.randn(x.shape, generator=g)
    assert rel(ops.leaky_relu_bwd(dev(dy), dev(x)), dy * torch.where(x > 0, 1.0, 0.2)) < 1e-7
    y = torch.tanh(torch.randn(4, 3, 8, 8, generator=g))
    d = torch.randn(4, 3, 8, 8, generator=g)
    assert rel(ops.tanh_bwd(dev(d), dev(y)), d * (1 - y * y)) < 1e-6
    xn = torch.randn(2, 4, 6, 32, generator=g)
    assert rel(ops.maxpool2(dev(xn)).permute(0, 3, 1, 2), F.max_pool2d(xn.permute(0, 3, 1, 2), 2)) < 1e-7
    img = torch.randn(3, 3, 8, 8, generator=g)
    assert torch.equal(ops.nhwc_to_nchw(ops.nchw_to_nhwc(dev(img), torch.float32)).cpu(), img)
    lab = torch.tensor([2, 0, 7])
    t = torch.zeros(3, 4, 4, 72, device="cuda")
    ops.onehot_planes(t, dev(lab), 64, 8)
    exp = F.one_hot(lab, 8).float()[:, None, None, :].expand(-1, 4, 4, -1)
    assert torch.equal(t[..., 64:].cpu(), exp)


# ------------------------------------------------------------------ direct kernels for the 3-channel boundary layers
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("stride,Co,N,H", [(1, 64, 5, 16), (2, 64, 5, 16), (1, 32, 5, 16),
                                           (1, 64, 64, 32), (2, 64, 160, 32)])  # large cases: 8-pixels-per-thread variants
def test_conv_small_cin3(ops, dtype, tol, stride, Co, N, H):
    g = torch.Generator().manual_seed(20 + stride + Co)
    x = torch.randn(N, 3, H, H, generator=g, requires_grad=True)
    w = (torch.randn(Co, 3, 3, 3, generator=g) * 0.2).requires_grad_(True)
    bias = torch.randn(Co, generator=g)
    wq = w.detach().to(dtype).float()
    y = F.conv2d(x, wq, bias, stride, 1)
    Ho = y.shape[2]
    w_cl = dev(w.detach().permute(0, 2, 3, 1).to(dtype))
    out = torch.empty(N, Ho, Ho, Co, device="cuda", dtype=dtype)
    ops.conv_cin3(dev(x.detach()), w_cl.data_ptr(), ops.dt_code(dtype), out, Co, stride, bias=dev(bias))
    assert rel(out.float().permute(0, 3, 1, 2), y) < tol
    # ELU + affine epilogue (FrequencyModel conv1)
    sc, sh = torch.rand(Co, generator=g) + 0.5, torch.randn(Co, generator=g)
    out2 = torch.empty(N, Ho, Ho, Co, device="cuda", dtype=torch.float32)
    ops.conv_cin3(dev(x.detach()), w_cl.data_ptr(), ops.dt_code(dtype), out2, Co, stride, bias=dev(bias), act=2,
                  post_scale=dev(sc), post_shift=dev(sh))
    assert rel(out2.permute(0, 3, 1, 2), F.elu(y) * sc[None, :, None, None] + sh[None, :, None, None]) < 1e-5
    # weight / bias gradient
    dy = torch.randn(y.shape, generator=g).to(dtype).float()
    F.conv2d(x, w, bias, stride, 1).backward(dy)
    dw = torch.zeros(Co, 3, 3, 3, device="cuda")  # [co][kh][kw][ci]
    db = torch.zeros(Co, device="cuda")
    ops.wgrad_cin3(dev(x.detach()), dev(_nhwc(dy).to(dtype)), dw, db, Co, stride)
    assert rel(dw.permute(0, 3, 1, 2), w.grad) < 2e-5 * max(1.0, (N * Ho * Ho) ** 0.5 / 30)
    assert rel(db, dy.sum((0, 2, 3))) < 2e-5 * max(1.0, (N * Ho * Ho) ** 0.5 / 30)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("N,H", [(3, 16), (160, 32)])  # the large case takes the 8-pixels-per-thread kernels
def test_conv_small_cout3(ops, dtype, tol, N, H):
    g = torch.Generator().manual_seed(31)
    x = torch.randn(N, 64, H, H, generator=g).to(dtype).float().requires_grad_(True)
    w = (torch.randn(3, 64, 3, 3, generator=g) * 0.05).to(dtype).float().requires_grad_(True)
    bias = torch.randn(3, generator=g)
    y = torch.tanh(F.conv2d(x, w, bias, 1, 1))
    out = torch.empty(N, 3, H, H, device="cuda")
    w_cl = dev(w.detach().permute(0, 2, 3, 1).to(dtype))
    xd = dev(_nhwc(x.detach()).to(dtype))
    ops.conv_cout3(xd, w_cl.data_ptr(), ops.dt_code(dtype), out, bias=dev(bias), act=1)
    assert rel(out, y) < 1e-5
    dz = torch.randn(N, 3, H, H, generator=g)
    z = F.conv2d(x, w, bias, 1, 1)
    z.backward(dz)
    dw = torch.zeros(3, 3, 3, 64, device="cuda")
    db = torch.zeros(3, device="cuda")
    ops.wgrad_cout3(xd, dev(dz), dw, db)
    assert rel(dw.permute(0, 3, 1, 2), w.grad) < 2e-5 * max(1.0, (N * H * H) ** 0.5 / 30)
    assert rel(db, dz.sum((0, 2, 3))) < 2e-5 * max(1.0, (N * H * H) ** 0.5 / 30)
    # input gradient of a 3 -> 64 conv == 64 -> 3 conv with the flipped, transposed weights (classifier conv1 dgrad)
    w2 = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).to(dtype).float()
    xin = torch.randn(N, 3, H, H, generator=g, requires_grad=True)
    dy2 = torch.randn(N, 64, H, H, generator=g).to(dtype).float()
    F.conv2d(xin, w2, None, 1, 1).backward(dy2)
    w_d = dev(w2.flip(2, 3).permute(1, 2, 3, 0).to(dtype))  # [ci][kh'][kw'][co]
    dx = torch.empty(N, 3, H, H, device="cuda")
    ops.conv_cout3(dev(_nhwc(dy2).to(dtype)), w_d.data_ptr(), ops.dt_code(dtype), dx)
    assert rel(dx, xin.grad) < 1e-5


def test_conv_tc_elu_affine_epilogue(ops):
    import ctypes as C

    from combat_b200._lib import check, lib
    g = torch.Generator().manual_seed(41)
    N, Ci, Co, H = 4, 64, 128, 8
    x = torch.randn(N, Ci, H, H, generator=g).bfloat16().float()
    w = (torch.randn(Co, Ci, 3, 3, generator=g) * 0.05).bfloat16().float()
    bias, sc, sh = torch.randn(Co, generator=g), torch.rand(Co, generator=g) + 0.5, torch.randn(Co, generator=g)
    y = F.elu(F.conv2d(x, w, bias, 1, 1)) * sc[None, :, None, None] + sh[None, :, None, None]
    out = torch.empty(N, H, H, Co, device="cuda", dtype=torch.bfloat16)
    w_f = dev(w.permute(0, 2, 3, 1).bfloat16())
    d = ops.conv_tc_desc(dev(_nhwc(x).bfloat16()), w_f.data_ptr(), out, N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, bias=dev(bias),
                         act=2, post_scale=dev(sc), post_shift=dev(sh))
    check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc")
    torch.cuda.synchronize()
    assert rel(out.float().permute(0, 3, 1, 2), y) < 6e-3
    # float32 output + float32 residual (pre-normalisation tensors)
    res = torch.randn(N, Co, H, H, generator=g)
    out32 = torch.empty(N, H, H, Co, device="cuda", dtype=torch.float32)
    d2 = ops.conv_tc_desc(dev(_nhwc(x).bfloat16()), w_f.data_ptr(), out32, N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, residual=dev(_nhwc(res)))
    check(lib.combat_conv_tc(C.byref(d2), ops._s()), "conv_tc")
    torch.cuda.synchronize()
    assert rel(out32.permute(0, 3, 1, 2), F.conv2d(x, w, None, 1, 1) + res) < 2e-5


@pytest.mark.parametrize("shape", [(3, 64, 128, 8), (5, 128, 128, 16)])   # the second runs conv_tc_rr_kernel (stride_up (1, 1))
@pytest.mark.parametrize("stride_up", [(1, 1), (1, 2)])
def test_conv_tc_fused_batchnorm_epilogues(ops, stride_up, shape):
    """out2 = relu(v*scale2+shift2) next to out (eval BatchNorm+ReLU of the consumer); mask/mask_scale/post_add (its
    backward) -- the epilogues behind nets.Classifier._forward_eval_fused / _backward_eval_fused."""
    import ctypes as C

    from combat_b200._lib import check, lib
    _, up = stride_up
    g = torch.Generator().manual_seed(77 + up)
    N, Ci, Co, H = shape
    x = torch.randn(N, Ci, H, H, generator=g).bfloat16().float()
    w = (torch.randn(Co, Ci, 3, 3, generator=g) * 0.05).bfloat16().float()
    sc, sh = torch.rand(Co, generator=g) + 0.5, torch.randn(Co, generator=g) * 0.3
    xd = dev(_nhwc(x).bfloat16())
    w_f = dev(w.permute(0, 2, 3, 1).bfloat16())
    if up == 1:
        res = torch.randn(N, Co, H, H, generator=g)
        y = F.conv2d(x, w, None, 1, 1) + res
        y2 = F.relu(y * sc[None, :, None, None] + sh[None, :, None, None])
        out = torch.empty(N, H, H, Co, device="cuda", dtype=torch.float32)
        out2 = torch.empty(N, H, H, Co, device="cuda", dtype=torch.bfloat16)
        d = ops.conv_tc_desc(xd, w_f.data_ptr(), out, N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, residual=dev(_nhwc(res)), out2=out2,
                             scale2=dev(sc), shift2=dev(sh))
        check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc")
        torch.cuda.synchronize()
        assert rel(out.permute(0, 3, 1, 2), y) < 2e-5
        assert rel(out2.float().permute(0, 3, 1, 2), y2) < 6e-3
        # fused train-mode BatchNorm statistics of `out`: per-CTA partial sums / sums of squares
        part = torch.full((256 * 2 * Co,), 7.0, device="cuda")
        d = ops.conv_tc_desc(xd, w_f.data_ptr(), out, N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, residual=dev(_nhwc(res)), stats=part)
        check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc")
        nblk = lib.combat_conv_tc_last_grid()
        torch.cuda.synchronize()
        ps = part[: nblk * 2 * Co].view(nblk, 2, Co).double().sum(0).cpu()
        assert rel(ps[0], y.double().sum((0, 2, 3))) < 1e-5 and rel(ps[1], (y.double() ** 2).sum((0, 2, 3))) < 1e-5
        # out2 only
        out2b = torch.zeros_like(out2)
        d = ops.conv_tc_desc(xd, w_f.data_ptr(), None, N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 1, residual=dev(_nhwc(res)), out2=out2b,
                             scale2=dev(sc), shift2=dev(sh))
        d.res_f32 = 1
        check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc")
        torch.cuda.synchronize()
        assert torch.equal(out2b, out2)
    # masked backward epilogue on a (possibly stride-2) input-gradient problem
    s = up
    Ho = H // s
    dy = torch.randn(N, Co, Ho, Ho, generator=g).bfloat16().float()
    xin = x.clone().requires_grad_(True)
    F.conv2d(xin, w, None, s, 1).backward(dy)
    mask = (torch.randn(N, Ci, H, H, generator=g) > 0).float() * torch.rand(N, Ci, H, H, generator=g)
    msc = torch.rand(Ci, generator=g) + 0.5
    pre = torch.randn(N, Ci, H, H, generator=g).bfloat16().float()
    post = torch.randn(N, Ci, H, H, generator=g).bfloat16().float()
    w_d = dev(w.flip(2, 3).permute(1, 2, 3, 0).bfloat16())
    for use_pre in (True, False):  # the shortcut gradient arrives either before the mask (conv shortcut) or after it (identity)
        exp = (xin.grad + pre) * (mask > 0).float() * msc[None, :, None, None] if use_pre else \
            xin.grad * (mask > 0).float() * msc[None, :, None, None] + post
        dx = torch.empty(N, H, H, Ci, device="cuda", dtype=torch.bfloat16)
        d2 = ops.conv_tc_desc(dev(_nhwc(dy).bfloat16()), w_d.data_ptr(), dx, N, Ho, Ho, Co, H, H, Ci, 3, 3, 1, 1, s,
                              residual=dev(_nhwc(pre).bfloat16()) if use_pre else None, mask=dev(_nhwc(mask).bfloat16()),
                              mask_scale=dev(msc), post_add=None if use_pre else dev(_nhwc(post).bfloat16()))
        check(lib.combat_conv_tc(C.byref(d2), ops._s()), "conv_tc dgrad")
        torch.cuda.synchronize()
        assert rel(dx.float().permute(0, 3, 1, 2), exp) < 6e-3


@pytest.mark.parametrize("shape", [(6, 64, 64, 32, 1), (5, 128, 128, 16, 1), (9, 256, 256, 8, 1), (6, 64, 128, 32, 2), (40, 512, 512, 4, 1)])
def test_conv_tc_fused_batchnorm_backward_reduction(ops, shape):
    """MODE 3 epilogue of an input-gradient launch: g = (x * scale + shift > 0) ? dgrad : 0 in bf16 plus the per-CTA partial sums
    of g and g * (x - mean) * invstd -- the reduction half of the train-mode relu(bn(x)) backward (preact_resnet.py:20-23) that
    combat_bn_bwd_reduce otherwise computes in a pass of its own.  Shapes hit conv_tc64 / conv_tc_rr / the CTA-pair kernel / the
    four parity classes of a stride-2 input gradient / tiles spanning several images."""
    import ctypes as C

    from combat_b200._lib import check, lib
    N, Ci, Co, H, s = shape
    g = torch.Generator().manual_seed(sum(shape))
    Ho = H // s
    w = (torch.randn(Co, Ci, 3, 3, generator=g) * 0.05).bfloat16().float()
    dy = torch.randn(N, Co, Ho, Ho, generator=g).bfloat16().float()
    xin = torch.zeros(N, Ci, H, H, requires_grad=True)
    F.conv2d(xin, w, None, s, 1).backward(dy)
    dgrad = xin.grad                                                  # [N, Ci, H, H]
    x = (torch.randn(N, Ci, H, H, generator=g) * 1.5 + 0.3).bfloat16().float()    # the saved pre-normalisation tensor
    mean, var = x.mean((0, 2, 3)), x.var((0, 2, 3), unbiased=False)
    invstd = (var + 1e-5).rsqrt()
    gamma, beta = torch.rand(Ci, generator=g) + 0.5, torch.randn(Ci, generator=g) * 0.3
    scale = gamma * invstd
    shift = beta - mean * scale
    pred = torch.addcmul(shift[None, :, None, None], x, scale[None, :, None, None]) > 0   # not fused, but only ties differ
    want_g = dgrad * pred
    xhat = (x - mean[None, :, None, None]) * invstd[None, :, None, None]
    w_d = dev(w.flip(2, 3).permute(1, 2, 3, 0).bfloat16())
    dx = torch.empty(N, H, H, Ci, device="cuda", dtype=torch.bfloat16)
    part = torch.full((148 * 4 * 2 * Ci,), 7.0, device="cuda")
    d = ops.conv_tc_desc(dev(_nhwc(dy).bfloat16()), w_d.data_ptr(), dx, N, Ho, Ho, Co, H, H, Ci, 3, 3, 1, 1, s, stats=part,
                         bnb=(dev(_nhwc(x).bfloat16()), dev(scale), dev(shift), dev(mean), dev(invstd)))
    check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc dgrad + bn backward reduction")
    nblk = lib.combat_conv_tc_last_grid()
    torch.cuda.synchronize()
    got = dx.float().permute(0, 3, 1, 2).cpu()
    assert rel(got, want_g) < 6e-3
    ps = part[: nblk * 2 * Ci].view(nblk, 2, Ci).double().sum(0).cpu()
    # the sums are taken over the UNROUNDED float32 g (like bn_bwd_reduce over a float32 dy would): compare with float64 torch
    assert rel(ps[0], want_g.double().sum((0, 2, 3))) < 1e-4
    assert rel(ps[1], (want_g.double() * xhat.double()).sum((0, 2, 3))) < 1e-4


@pytest.mark.parametrize("N,H", [(5, 32), (300, 32), (3, 16), (2, 64)])
def test_conv_tc_first(ops, N, H):
    """3 -> 64 first conv on the tensor pipe (conv_tc_first_kernel): float32 NCHW image split into bf16 hi + lo halves by the
    kernel, operand built in shared memory, filter [64][27 | 0 | 27 | 0]; forward epilogue features (second output with the fused
    eval BatchNorm+ReLU, train-mode statistics).  Only the bf16 rounding of the FILTER separates it from float32 torch."""
    import ctypes as C

    from combat_b200._lib import check, lib
    g = torch.Generator().manual_seed(70 + N + H)
    x = torch.randn(N, 3, H, H, generator=g) * 1.5
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).bfloat16().float()
    y = F.conv2d(x, w, None, 1, 1)
    sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.3
    y2 = F.relu(y * sc[None, :, None, None] + sh[None, :, None, None])
    w64 = torch.zeros(64, 64, dtype=torch.bfloat16)
    w64[:, :27] = w.permute(0, 2, 3, 1).reshape(64, 27).bfloat16()
    w64[:, 32:59] = w64[:, :27]
    w64 = dev(w64)
    xd = dev(x)
    out = torch.full((N, H, H, 64), 7.0, device="cuda")
    out2 = torch.zeros(N, H, H, 64, device="cuda", dtype=torch.bfloat16)
    d = ops.conv_tc_desc(xd, w64.data_ptr(), out, N, H, H, 64, H, H, 64, 3, 3, 1, 1, 1, out2=out2, scale2=dev(sc), shift2=dev(sh),
                         in_nchw3=True)
    check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc first")
    torch.cuda.synchronize()
    assert rel(out.permute(0, 3, 1, 2), y) < 2e-5
    assert rel(out2.float().permute(0, 3, 1, 2), y2) < 6e-3
    # bf16 out + train-mode statistics (taken from the float32 accumulators)
    outb = torch.zeros(N, H, H, 64, device="cuda", dtype=torch.bfloat16)
    part = torch.full((148 * 4 * 2 * 64,), 7.0, device="cuda")
    d = ops.conv_tc_desc(xd, w64.data_ptr(), outb, N, H, H, 64, H, H, 64, 3, 3, 1, 1, 1, stats=part, in_nchw3=True)
    check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc first + stats")
    nblk = lib.combat_conv_tc_last_grid()
    torch.cuda.synchronize()
    assert rel(outb.float().permute(0, 3, 1, 2), y) < 6e-3
    ps = part[: nblk * 2 * 64].view(nblk, 2, 64).double().sum(0).cpu()
    assert rel(ps[0], y.double().sum((0, 2, 3))) < 1e-4 and rel(ps[1], (y.double() ** 2).sum((0, 2, 3))) < 1e-4


@pytest.mark.parametrize("shape", [(5, 3, 32, 32), (2, 3, 64, 64), (3, 1, 7, 9)])
def test_tv_loss_fwd_bwd(ops, shape):
    """total-variation term of train_generator_imperceptible.py:228 (kornia 0.6.6 total_variation(.).mean()): loss and the gradient
    ADDED to an existing gradient buffer, against torch autograd on the oracle's restatement (ties: sign(0) = 0 on both sides)."""
    from oracle import combat_oracle as O
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.rand(shape, generator=g) * 2 - 1)
    x[0, 0, 1, :] = x[0, 0, 0, :]          # exact ties: zero differences
    xr = x.clone().requires_grad_(True)
    loss = O.total_variation(xr).mean()
    w = 0.37
    (w * loss).backward()
    g0 = torch.randn(shape, generator=g)
    gd = dev(g0.clone())
    out = torch.zeros(1, device="cuda")
    ops.tv_loss(dev(x), out, grad=gd, grad_weight=w / shape[0])
    torch.cuda.synchronize()
    assert abs(float(out) - float(loss)) < 2e-6 * abs(float(loss))
    assert rel(gd, g0 + xr.grad) < 1e-6
    out2 = torch.zeros(1, device="cuda")
    ops.tv_loss(dev(x), out2)              # loss only
    assert float(out2) == float(out)


@pytest.mark.parametrize("N,H,act", [(5, 32, 1), (300, 32, 0), (7, 16, 1), (3, 64, 0)])
def test_conv_tc_cout3(ops, N, H, act):
    """64 -> 3 conv on the tensor pipe (conv_tc_cout3_kernel): the generator's last conv (bias + tanh, networks/models.py:316,341)
    and the input gradient of the classifiers' first conv, float32 NCHW out, against torch on the same bf16 operands."""
    from combat_b200._lib import lib
    g = torch.Generator().manual_seed(1000 + N + H)
    x = torch.randn(N, 64, H, H, generator=g).bfloat16().float()
    w = (torch.randn(3, 64, 3, 3, generator=g) * 0.1).bfloat16().float()
    bias = torch.randn(3, generator=g) if act else None
    want = F.conv2d(x, w, bias, 1, 1)
    if act:
        want = torch.tanh(want)
    xd = dev(_nhwc(x).bfloat16())
    assert ops.conv_tc_cout3_ok(xd) == bool(lib.combat_conv_tc_cout3_supported(N, H, H))
    assert ops.conv_tc_cout3_ok(xd)
    wd = dev(w.permute(0, 2, 3, 1).bfloat16())   # [3][9][64]
    out = torch.full((N, 3, H, H), 7.0, device="cuda")
    ops.conv_tc_cout3(xd, wd.data_ptr(), out, bias=dev(bias) if act else None, act=act)
    torch.cuda.synchronize()
    assert rel(out, want) < 2e-5


@pytest.mark.parametrize("stride,N,H", [(1, 5, 16), (2, 6, 32), (1, 3, 24)])
def test_im2col3_tensor_core_conv(ops, stride, N, H):
    """3 -> 64 conv as im2col3 ([hi | lo] bf16 halves of the float32 image) + 1x1 tcgen05 conv with the filter stored twice:
    the image enters with ~16 mantissa bits, so only the bf16 rounding of the FILTER separates it from float32."""
    import ctypes as C

    from combat_b200._lib import check, lib
    g = torch.Generator().manual_seed(50 + stride)
    Co = 64
    x = torch.randn(N, 3, H, H, generator=g) * 3.0
    w = (torch.randn(Co, 3, 3, 3, generator=g) * 0.2).bfloat16().float()
    bias = torch.randn(Co, generator=g)
    y = F.conv2d(x, w, bias, stride, 1)
    Ho = y.shape[2]
    A = ops.im2col3(dev(x), stride)
    assert A.shape == (N, Ho, Ho, 64) and A.dtype == torch.bfloat16
    # the operand itself: hi + lo reproduces the patch to 2^-16 relative, padding columns are zero
    patches = F.unfold(x, 3, padding=1, stride=stride).view(N, 3, 9, Ho, Ho).permute(0, 3, 4, 2, 1).reshape(N, Ho, Ho, 27)
    Af = A.float().cpu()
    assert float((Af[..., :27] + Af[..., 32:59] - patches).abs().max()) <= 2e-5 * float(patches.abs().max())
    assert float(Af[..., 27:32].abs().max()) == 0 and float(Af[..., 59:].abs().max()) == 0
    w64 = torch.zeros(Co, 64, dtype=torch.bfloat16)
    w64[:, :27] = w.permute(0, 2, 3, 1).reshape(Co, 27).bfloat16()
    w64[:, 32:59] = w64[:, :27]
    w64 = dev(w64)
    out = torch.empty(N, Ho, Ho, Co, device="cuda")
    d = ops.conv_tc_desc(A, w64.data_ptr(), out, N, Ho, Ho, 64, Ho, Ho, Co, 1, 1, 1, 0, 1, bias=dev(bias))
    check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc")
    torch.cuda.synchronize()
    assert rel(out.permute(0, 3, 1, 2), y) < 2e-5


@pytest.mark.parametrize("case", [(6, 64, 128, 32), (16, 128, 256, 16), (33, 256, 512, 8)])
@pytest.mark.parametrize("masked", [False, True])
def test_conv_tc_dgrad_with_fused_shortcut(ops, case, masked):
    """Input gradient of a PreAct / ResNet block with a projection shortcut (preact_resnet.py:26-35): conv1 (3x3, stride 2) and
    shortcut (1x1, stride 2) read the same tensor, so dx = conv1^T(d_c1) + shortcut^T(dh) -- ONE launch, the shortcut's
    gradient being one more tap of output-parity class (0,0) from a second (tensor, filter) pair.  256- and 512-channel cases
    run the CTA-pair kernel; `masked` adds the fused eval-mode relu(bn(.)) backward epilogue."""
    import ctypes as C

    from combat_b200._lib import check, lib
    N, Ci, Co, H = case
    g = torch.Generator().manual_seed(sum(case) + int(masked))
    x = torch.randn(N, Ci, H, H, generator=g).bfloat16().float().requires_grad_(True)
    w1 = (torch.randn(Co, Ci, 3, 3, generator=g) * 0.05).bfloat16().float()
    wsc = (torch.randn(Co, Ci, 1, 1, generator=g) * 0.1).bfloat16().float()
    y = F.conv2d(x, w1, None, 2, 1)
    ysc = F.conv2d(x, wsc, None, 2, 0)
    dy1 = torch.randn(y.shape, generator=g).bfloat16().float()
    dy2 = torch.randn(y.shape, generator=g).bfloat16().float()
    (y * dy1 + ysc * dy2).sum().backward()
    want = x.grad
    Ho = y.shape[2]
    w1_d = dev(w1.flip(2, 3).permute(1, 2, 3, 0).bfloat16())      # [ci][kh'][kw'][co]
    wsc_d = dev(wsc.permute(1, 2, 3, 0).bfloat16())               # [ci][1][1][co]
    d1, d2 = dev(_nhwc(dy1).bfloat16()), dev(_nhwc(dy2).bfloat16())
    dx = torch.empty(N, H, H, Ci, device="cuda", dtype=torch.bfloat16)
    kw = {}
    if masked:
        mask = torch.randn(N, H, H, Ci, generator=g).bfloat16()
        ms = torch.rand(Ci, generator=g) + 0.5
        kw = dict(mask=dev(mask), mask_scale=dev(ms))
        want = torch.where(mask.float().permute(0, 3, 1, 2) > 0, want * ms.view(1, -1, 1, 1), torch.zeros(()))
    d = ops.conv_tc_desc(d1, w1_d.data_ptr(), dx, N, Ho, Ho, Co, H, H, Ci, 3, 3, 1, 1, 2, in2=d2, w2=wsc_d.data_ptr(), **kw)
    assert lib.combat_conv_tc_supported(C.byref(d))
    check(lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc dgrad + shortcut")
    torch.cuda.synchronize()
    assert rel(dx.float().permute(0, 3, 1, 2), want) < 6e-3
