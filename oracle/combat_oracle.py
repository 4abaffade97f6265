"""CPU oracle for the COMBAT alternated generator/surrogate training step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``combat_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the CPU baseline -- never as the product path.

What it is: a functional (state-dict driven) restatement, on torch-CPU fp32, of
the algorithm the reference runs for the hot path.  The reference is pure
Python/PyTorch, so the arithmetic primitives (conv2d, batch_norm,
instance_norm, fft, autograd) are the same torch CPU ops the reference itself
dispatches; what is *restated* here is the reference's own control flow, layer
wiring, RNG consumption order, losses and optimiser.  It contains no reference
source and does not import ``/root/reference``.

Parity pins (see tests/golden/ and tests/test_oracle_golden.py): the fixtures in
``tests/golden/*.npz`` were produced by running the UNMODIFIED reference
(`/root/reference`, through `oracle/ref_loader.py`) in the build container with
`tests/golden/make_golden.py`; this oracle reproduces them (integer selections
bit-exact, floats to <=1e-6 relative -- they are bit-equal on the same torch
build and thread count).

Every function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# Quantisation-aware mode (checker for the bf16 tensor-core path)
#
# The reference computes in float32 throughout.  The CUDA path that is benchmarked stores activations and activation
# gradients in bfloat16 and feeds bfloat16 conv operands to the tensor cores (float32 accumulate, float32 master weights,
# float32 pre-normalisation tensors / residual stream, float32 statistics, losses and optimiser).  Comparing it with the
# float32 reference measures the QUANTISATION, not the implementation: bf16's epsilon (2^-8) is above the north star's 1e-3.
# Inside `with quantised():` the restated networks below round at exactly the points combat_b200/nets.py does -- and nowhere
# else -- so that (a) CUDA-bf16 vs this mode isolates implementation error and can be held to 1e-3 per tensor, and (b) this
# mode vs the float32 reference is the measured cost of the storage format (tests/test_bf16_parity_*.py state both).
#   qa(t): an activation stored in bf16 -- value rounded in the forward, its gradient rounded in the backward
#   qg(t): a float32 forward tensor whose GRADIENT is stored in bf16 (the generator's pre-normalisation tensors)
#   qp(t): a CLASSIFIER pre-normalisation tensor / the residual stream: stored in bf16 like qa (`quantised(pre_bf16=True)`, the
#          default since combat_b200 stores them in bf16), or float32 with a bf16 gradient like qg (`pre_bf16=False`, which is
#          COMBAT_PRE_F32=1 in combat_b200/nets.py)
#   qw(w): a conv weight rounded for the tensor cores; its gradient accumulates in float32 (straight through)
# --------------------------------------------------------------------------
#   cj(t): a float32 conv output; identity, or -- `quantised(jitter=e)` -- multiplied by (1 + e * N(0,1)) element-wise from a
#          PRIVATE generator: a model of float32 accumulation-order noise (the CUDA kernels sit 5e-7 .. 4e-6 from torch's conv
#          on identical inputs, scripts/diag_bf16_layers.py).  Two runs of the SAME quantised algorithm that differ only by this
#          jitter bound what any faithful bf16 implementation can agree to: a value within that noise of a rounding boundary
#          rounds the other way, and a random-init network amplifies every flip layer by layer.
_QUANT = {"on": False, "jitter": 0.0, "gen": None, "pre_bf16": True}


class quantised:
    def __init__(self, jitter: float = 0.0, seed: int = 1234, pre_bf16: bool = True):
        self.jitter, self.seed, self.pre_bf16 = float(jitter), seed, bool(pre_bf16)

    def __enter__(self):
        self.prev = dict(_QUANT)
        _QUANT["on"] = True
        _QUANT["jitter"] = self.jitter
        _QUANT["pre_bf16"] = self.pre_bf16
        _QUANT["gen"] = torch.Generator().manual_seed(self.seed) if self.jitter else None

    def __exit__(self, *a):
        _QUANT.update(self.prev)


class _CJ(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t * (1.0 + _QUANT["jitter"] * torch.randn(t.shape, generator=_QUANT["gen"]))

    @staticmethod
    def backward(ctx, g):
        return g * (1.0 + _QUANT["jitter"] * torch.randn(g.shape, generator=_QUANT["gen"]))


def cj(t):
    return _CJ.apply(t) if (_QUANT["on"] and _QUANT["jitter"]) else t


def _r(t):
    return t.to(torch.bfloat16).to(torch.float32)


class _QA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return _r(t)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class _QG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class _QW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return _r(t)

    @staticmethod
    def backward(ctx, g):
        return g


def qa(t):
    return _QA.apply(t) if _QUANT["on"] else t


def qg(t):
    return _QG.apply(t) if _QUANT["on"] else t


def qp(t):
    return qa(t) if _QUANT["pre_bf16"] else qg(t)


def qf(t):
    """stored in bf16, gradient passed through unchanged (it is rounded where the float32 value was produced)"""
    return _QW.apply(t) if _QUANT["on"] else t


def qw(w):
    return _QW.apply(w) if _QUANT["on"] else w


# --------------------------------------------------------------------------
# DCT  (utils/dct.py)
# --------------------------------------------------------------------------


def dct(x: torch.Tensor, norm: str = "ortho") -> torch.Tensor:
    """DCT-II along the last dim via one length-N complex FFT (utils/dct.py:13-42)."""
    shape = x.shape
    N = shape[-1]
    x = x.contiguous().view(-1, N)
    v = torch.cat([x[:, ::2], x[:, 1::2].flip([1])], dim=1)  # :26
    Vc = torch.view_as_real(torch.fft.fft(v, dim=1))  # :6,28
    # :30 -- note: for uint8 input the arange is uint8 and the negation wraps
    k = -torch.arange(N, dtype=x.dtype, device=x.device)[None, :] * np.pi / (2 * N)
    W_r, W_i = torch.cos(k), torch.sin(k)
    V = Vc[:, :, 0] * W_r - Vc[:, :, 1] * W_i  # :34
    if norm == "ortho":  # :36-38
        V[:, 0] /= np.sqrt(N) * 2
        V[:, 1:] /= np.sqrt(N / 2) * 2
    return 2 * V.view(*shape)  # :40


def idct(X: torch.Tensor, norm: str = "ortho") -> torch.Tensor:
    """DCT-III (inverse of `dct`) along the last dim via irfft (utils/dct.py:45-82)."""
    shape = X.shape
    N = shape[-1]
    X_v = X.contiguous().view(-1, N) / 2  # :59
    if norm == "ortho":  # :61-63
        X_v[:, 0] *= np.sqrt(N) * 2
        X_v[:, 1:] *= np.sqrt(N / 2) * 2
    k = torch.arange(N, dtype=X.dtype, device=X.device)[None, :] * np.pi / (2 * N)
    W_r, W_i = torch.cos(k), torch.sin(k)
    V_t_r = X_v
    V_t_i = torch.cat([X_v[:, :1] * 0, -X_v.flip([1])[:, :-1]], dim=1)  # :70
    V_r = V_t_r * W_r - V_t_i * W_i
    V_i = V_t_r * W_i + V_t_i * W_r
    V = torch.cat([V_r.unsqueeze(2), V_i.unsqueeze(2)], dim=2)
    v = torch.fft.irfft(torch.view_as_complex(V), n=N, dim=1)  # :10,77
    x = v.new_zeros(v.shape)
    x[:, ::2] += v[:, : N - (N // 2)]  # :79
    x[:, 1::2] += v.flip([1])[:, : N // 2]  # :80
    return x.view(*shape)


def dct_2d(x: torch.Tensor, norm: str = "ortho") -> torch.Tensor:
    """utils/dct.py:85-96."""
    X1 = dct(x, norm=norm)
    X2 = dct(X1.transpose(-1, -2), norm=norm)
    return X2.transpose(-1, -2)


def idct_2d(X: torch.Tensor, norm: str = "ortho") -> torch.Tensor:
    """utils/dct.py:99-111."""
    x1 = idct(X, norm=norm)
    x2 = idct(x1.transpose(-1, -2), norm=norm)
    return x2.transpose(-1, -2)


def dct_matrix(N: int) -> np.ndarray:
    """Exact float64 orthonormal DCT-II matrix D (rows = frequencies): the closed
    form the FFT route above evaluates; scipy.fft.dctn(norm='ortho') == D X D^T."""
    n = np.arange(N, dtype=np.float64)
    k = n[:, None]
    D = np.sqrt(2.0 / N) * np.cos(np.pi * (2 * n[None, :] + 1) * k / (2 * N))
    D[0, :] *= 1.0 / np.sqrt(2.0)
    return D


def dct_2d_exact(x: np.ndarray) -> np.ndarray:
    D = dct_matrix(x.shape[-1])
    return D @ x.astype(np.float64) @ D.T


def idct_2d_exact(X: np.ndarray) -> np.ndarray:
    D = dct_matrix(X.shape[-1])
    return D.T @ X.astype(np.float64) @ D


def low_freq(x: torch.Tensor, image_size: int, ratio: float) -> torch.Tensor:
    """train_generator.py:47-55."""
    k = int(image_size * ratio)
    mask = torch.zeros_like(x)
    mask[:, :, :k, :k] = 1
    x_dct = dct_2d((x + 1) / 2 * 255)
    x_dct = x_dct * mask
    return (idct_2d(x_dct) / 255 * 2) - 1


def low_freq_exact(x: np.ndarray, image_size: int, ratio: float) -> np.ndarray:
    """P X P^T with P = D^T diag(1_k) D -- the closed form of `low_freq`
    (the affine (x+1)/2*255 ... /255*2-1 cancels because the DC term is kept)."""
    k = int(image_size * ratio)
    D = dct_matrix(x.shape[-1])
    P = D[:k].T @ D[:k]
    return P @ x.astype(np.float64) @ P.T


# --------------------------------------------------------------------------
# poison selection / targets (train_generator.py:70-77,181-188)
# --------------------------------------------------------------------------


def create_targets_bd(targets: torch.Tensor, attack_mode: str, target_label: int, num_classes: int) -> torch.Tensor:
    """train_generator.py:70-77."""
    if attack_mode == "all2one":
        return torch.ones_like(targets) * target_label
    if attack_mode == "all2all":
        return torch.tensor([(int(l) + 1) % num_classes for l in targets])
    raise Exception("{} attack mode is not implemented".format(attack_mode))


def select_poison(targets: torch.Tensor, bd_targets: torch.Tensor, pc: float, rng=np.random):
    """train_generator.py:181-183.  Returns (trg_ind, ntrg_ind, num_bd).
    Consumes exactly one ``rng.rand(n_trg)`` from the numpy global RandomState."""
    trg_ind = (targets == bd_targets).nonzero()[:, 0]
    ntrg_ind = (targets != bd_targets).nonzero()[:, 0]
    num_bd = int(np.sum(rng.rand(trg_ind.shape[0]) < pc))
    return trg_ind, ntrg_ind, num_bd


# --------------------------------------------------------------------------
# Gaussian blur (torchvision GaussianBlur(3,(0.1,1.0)), train_generator.py:165,194,226)
# --------------------------------------------------------------------------


def draw_sigma(lo: float = 0.1, hi: float = 1.0) -> float:
    """One torch CPU-generator uniform per call (torchvision GaussianBlur.get_params)."""
    return torch.empty(1).uniform_(lo, hi).item()


def gaussian_kernel1d(sigma: float, ksize: int = 3) -> torch.Tensor:
    half = (ksize - 1) * 0.5
    x = torch.linspace(-half, half, steps=ksize, dtype=torch.float32)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    return pdf / pdf.sum()


def gaussian_blur(img: torch.Tensor, sigma: float, ksize: int = 3) -> torch.Tensor:
    """Separable Gaussian as a depthwise conv with reflect padding."""
    k1 = gaussian_kernel1d(sigma, ksize).to(img)  # dtype and device of the image (the taps are evaluated in float32 on the host)
    k2 = torch.mm(k1[:, None], k1[None, :])
    C = img.shape[-3]
    w = k2.expand(C, 1, ksize, ksize)
    pad = ksize // 2
    x = F.pad(img, [pad, pad, pad, pad], mode="reflect")
    return F.conv2d(x, w, groups=C)


# --------------------------------------------------------------------------
# networks -- functional forwards over reference state-dict names
# --------------------------------------------------------------------------


def unet_forward(p: dict, x: torch.Tensor, y: torch.Tensor | None = None, num_classes: int | None = None, taps: dict | None = None):
    """UnetGenerator.forward (networks/models.py:318-341) and, when `y` is given,
    CUnetGeneratorv1.forward (networks/models.py:523-555).

    `self.act` is LeakyReLU(0.2, inplace=True) (:273): every `self.act(t)` mutates `t`,
    so the skip tensors f0/f1/f2 that are added back (:331,334,337) are the ACTIVATED ones.
    InstanceNorm2d defaults: affine=False, eps=1e-5, no running stats.  Upsample is
    bilinear, align_corners=False (:274)."""

    # Rounding points of the bf16 CUDA path (combat_b200/nets.py Generator.forward / backward), active only inside
    # `with quantised():` -- conv weights qw; conv0_0's output, every normalised / activated / upsampled tensor qa; the
    # float32 conv outputs that feed an InstanceNorm get their GRADIENT stored in bf16 (qg); image and tanh output float32.
    def act(t):
        return F.leaky_relu(t, 0.2)

    def conv(name, t, stride=1):
        o = cj(F.conv2d(t, qw(p[name + ".weight"]), p[name + ".bias"], stride=stride, padding=1))
        if taps is not None:  # layer-by-layer diagnostics (scripts/diag_bf16_layers.py): conv input and output of every layer
            taps[name] = (t.detach(), o.detach())
        return o

    def inorm(t):
        return F.instance_norm(qg(t), eps=1e-5)

    def up(t):
        return F.interpolate(t, scale_factor=(2, 2), mode="bilinear")

    f0 = qa(conv("conv0_0", x, 2))
    if y is not None:  # models.py:525-530
        y_emb = F.one_hot(y, num_classes=num_classes).float()[:, :, None, None].expand(-1, -1, f0.shape[2], f0.shape[3])
        f0 = torch.cat((f0, y_emb), 1)
    f0 = qa(act(inorm(conv("conv0_1", qa(act(f0))))))  # f0 after the in-place act at :322
    f1 = qa(act(inorm(conv("conv1_0", f0, 2))))
    f1 = qa(act(inorm(conv("conv1_1", f1))))  # activated by :324
    f2 = qa(act(inorm(conv("conv2_0", f1, 2))))
    f2 = qa(act(inorm(conv("conv2_1", f2))))  # activated by :326
    f3 = qa(act(inorm(conv("conv3_0", f2, 2))))
    f3 = qa(inorm(conv("conv3_1", f3)))
    u3 = qa(act(inorm(conv("upconv3_1", qa(act(up(f3)))))))
    u3 = qa(inorm(conv("upconv3_0", u3)) + f2)
    u2 = qa(act(inorm(conv("upconv2_1", qa(act(up(u3)))))))
    u2 = qa(inorm(conv("upconv2_0", u2)) + f1)
    u1 = qa(act(inorm(conv("upconv1_1", qa(act(up(u2)))))))
    u1 = qa(inorm(conv("upconv1_0", u1)) + f0)
    u0 = qa(act(inorm(conv("upconv0_1", qa(act(up(u1)))))))
    return torch.tanh(conv("upconv0_0", u0))


def grid_generator_forward(p: dict, x: torch.Tensor, S: int) -> torch.Tensor:
    """GridGenerator.forward (networks/models.py:372-385): the encoder half of UnetGenerator (same in-place LeakyReLU rule, same
    InstanceNorm), global average pool, fc1 -> LeakyReLU -> fc2 -> reshape (-1, 2, S, S) -> tanh: the S x S control grid of a flow."""
    def act(t):
        return F.leaky_relu(t, 0.2)

    def conv(name, t, stride=1):
        return cj(F.conv2d(t, qw(p[name + ".weight"]), p[name + ".bias"], stride=stride, padding=1))

    def inorm(t):
        return F.instance_norm(qg(t), eps=1e-5)

    f0 = qa(conv("conv0_0", x, 2))
    f0 = qa(act(inorm(conv("conv0_1", qa(act(f0))))))
    f1 = qa(act(inorm(conv("conv1_0", f0, 2))))
    f1 = qa(act(inorm(conv("conv1_1", f1))))
    f2 = qa(act(inorm(conv("conv2_0", f1, 2))))
    f2 = qa(act(inorm(conv("conv2_1", f2))))
    f3 = qa(act(inorm(conv("conv3_0", f2, 2))))
    f3 = qa(inorm(conv("conv3_1", f3)))
    f = F.adaptive_avg_pool2d(f3, 1).flatten(1)   # `.squeeze()` at :381 (also drops a batch of one; restored by the reshape at :383)
    f = F.linear(f, p["fc1.weight"], p["fc1.bias"])
    f = F.linear(act(f), p["fc2.weight"], p["fc2.bias"]).reshape((-1, 2, S, S))
    return torch.tanh(f)


def identity_grid(H: int) -> torch.Tensor:
    """train_generator_wanet.py:560-562: linspace(-1, 1, H) mesh, [..., 0] = column (x) coordinate, [..., 1] = row (y)."""
    a = torch.linspace(-1, 1, steps=H)
    xx, yy = torch.meshgrid(a, a, indexing="ij")
    return torch.stack((yy, xx), 2)[None, ...]


def wanet_warp(x: torch.Tensor, flow: torch.Tensor, opt):
    """train_generator_wanet.py:152-158 / :197-203: bicubic (align_corners) upsampling of the S x S control grid to the image size,
    blend with the identity grid by grid_rescale, clamp, bilinear `grid_sample` (zeros padding, align_corners).
    Returns (inputs_bd, noise_grid [N, H, W, 2])."""
    H = opt.input_height
    noise_grid = F.interpolate(flow, size=H, mode="bicubic", align_corners=True).permute((0, 2, 3, 1))
    grid = torch.clamp(identity_grid(H) * (1 - opt.grid_rescale) + noise_grid * opt.grid_rescale, -1, 1)
    return F.grid_sample(x, grid, align_corners=True), noise_grid


def wanet_grad_l2(noise_grid: torch.Tensor) -> torch.Tensor:
    """the logged-only "gradient loss" of train_generator_wanet.py:213-222 (finite differences of the zero-padded noise grid)."""
    e, z = F.pad(noise_grid, (1, 1, 2, 1)), F.pad(noise_grid * 0, (1, 1, 2, 1))
    return (F.mse_loss(e[:, :, 1:] - e[:, :, :-1], z[:, :, 1:] - z[:, :, :-1])
            + F.mse_loss(e[:, :, :, 1:] - e[:, :, :, :-1], z[:, :, :, 1:] - z[:, :, :, :-1]))


def alternated_step_wanet(state: dict, x: torch.Tensor, y: torch.Tensor, opt, with_metrics: bool = True) -> dict:
    """One iteration of train_generator_wanet.train() (:133-234): the alternated step with a WARP trigger -- netG is a
    GridGenerator, inputs_bd = grid_sample(inputs, clamp(identity * (1 - grid_rescale) + bicubic(netG(inputs)) * grid_rescale)) --
    no DCT low-pass, no blur (hence no sigma draws), loss_l2 = MSE(noise_grid, 0).
    RNG order: numpy rand(n_trg), then the PostTensorTransform calls T1, T2, T3, T4, T5.
    The reference raises for num_bd == 0 (`reshape((-1, 2, S, S))` of an empty tensor, models.py:383); here no row is poisoned."""
    fwdC = CLASSIFIERS[opt.classifier]
    netC_p, netC_b, netG_p = state["netC_p"], state["netC_b"], state["netG_p"]
    clean_p, clean_b = state["clean_p"], state["clean_b"]
    out = {}
    bd_targets = create_targets_bd(y, opt.attack_mode, opt.target_label, opt.num_classes)
    trg_ind, ntrg_ind, num_bd = select_poison(y, bd_targets, opt.pc)
    out["trg_ind"], out["ntrg_ind"], out["num_bd"] = trg_ind.clone(), ntrg_ind.clone(), num_bd
    x_sel = x[trg_ind[:num_bd]]
    for t in netC_p.values():
        t.requires_grad_(True)
        t.grad = None
    with torch.no_grad():
        x_bd_c = wanet_warp(x_sel, grid_generator_forward(netG_p, x_sel, opt.s), opt)[0] if num_bd > 0 else x_sel
    total_x = torch.cat([x_bd_c, x[trg_ind[num_bd:]], x[ntrg_ind]], dim=0)
    total_y = torch.cat([bd_targets[trg_ind[:num_bd]], y[trg_ind[num_bd:]], y[ntrg_ind]], dim=0)
    out["tf"] = tf_log = []
    total_x = post_transform(total_x, opt, tf_log)  # :161
    logits_c = fwdC(netC_p, netC_b, total_x, True)
    loss_c = F.cross_entropy(logits_c, total_y)
    loss_c.backward()
    out["total_x"], out["total_y"] = total_x.detach(), total_y
    out["logits_c"], out["loss_c"] = logits_c.detach().clone(), float(loss_c.detach())
    gradsC = {k: v.grad for k, v in netC_p.items()}
    with torch.no_grad():
        for t in netC_p.values():
            t.requires_grad_(False)
        sgd_nesterov_step(netC_p, gradsC, state["momC"], opt.lr_C)
        out["clean_preds"] = fwdC(clean_p, clean_b, post_transform(x, opt, tf_log), False)  # :184
    for t in netG_p.values():
        t.requires_grad_(True)
        t.grad = None
    flow = grid_generator_forward(netG_p, x, opt.s)  # :196
    x_bd, noise_grid = wanet_warp(x, flow, opt)
    with torch.no_grad():
        out["pred_clean"] = fwdC(netC_p, netC_b, post_transform(x, opt, tf_log), False)  # :205
    pred_bd = fwdC(netC_p, netC_b, post_transform(x_bd, opt, tf_log), False)  # :206
    loss_ce = F.cross_entropy(pred_bd, bd_targets)
    loss_l2 = F.mse_loss(noise_grid, noise_grid * 0)  # :212
    out["loss_grad_l2"] = float(wanet_grad_l2(noise_grid.detach()))
    clean_model_preds = fwdC(clean_p, clean_b, post_transform(x_bd, opt, tf_log), False)  # :229
    clean_model_loss = F.cross_entropy(clean_model_preds, y)
    loss = loss_ce + opt.L2_weight * loss_l2 + opt.clean_model_weight * clean_model_loss  # :235
    loss.backward()
    gradsG = {k: v.grad for k, v in netG_p.items()}
    out["gradsG"] = {k: g.clone() for k, g in gradsG.items()}
    with torch.no_grad():
        for t in netG_p.values():
            t.requires_grad_(False)
        sgd_nesterov_step(netG_p, gradsG, state["momG"], opt.lr_G)
        if with_metrics and state.get("netF_p") is not None:
            inputs_F = dct_2d(((x_bd.detach() + 1) / 2 * 255).byte())  # :224
            out["inputs_F"] = inputs_F
            out["pred_F"] = frequency_model_forward(state["netF_p"], state["netF_b"], inputs_F)
    am = lambda t: torch.argmax(t, dim=1)
    out.update(x_bd=x_bd.detach(), flow=flow.detach(), noise_grid=noise_grid.detach(), pred_bd=pred_bd.detach(),
               clean_model_preds=clean_model_preds.detach(), loss_ce=float(loss_ce), loss_l2=float(loss_l2),
               clean_model_loss=float(clean_model_loss), loss_g=float(loss), bd_targets=bd_targets,
               n_clean_correct=int((am(out["pred_clean"]) == y).sum()), n_bd_correct=int((am(pred_bd) == bd_targets).sum()),
               n_clean_model_correct=int((am(out["clean_preds"]) == y).sum()),
               n_clean_model_bd_ba=int((am(clean_model_preds) == y).sum()),
               n_clean_model_bd_asr=int((am(clean_model_preds) == bd_targets).sum()))
    if "pred_F" in out:
        out["n_F_correct"] = int((am(out["pred_F"]) == 1).sum())
    return out


def _bn(p, b, name, x, training, momentum=0.1, eps=1e-5):
    """nn.BatchNorm2d: batch stats + running update (train) or running stats (eval)."""
    rm, rv = b[name + ".running_mean"], b[name + ".running_var"]
    out = F.batch_norm(x, rm, rv, p[name + ".weight"], p[name + ".bias"], training, momentum, eps)
    if training and (name + ".num_batches_tracked") in b:
        b[name + ".num_batches_tracked"] += 1
    return out


def preact_resnet18_forward(p: dict, b: dict, x: torch.Tensor, training: bool):
    """PreActResNet18 (classifier_models/preact_resnet.py:13-40,72-110).
    `p` = parameters, `b` = buffers (running stats; updated in place when training)."""
    # quantised(): bf16 conv weights and relu(bn(.)) tensors; the residual stream / pre-normalisation tensors are STORED in bf16
    # (`pre_bf16`) with bf16 gradients, as in combat_b200/nets.py Classifier.forward / backward; the image stays float32.
    # Which value a BatchNorm reads follows the CUDA schedule: in eval mode relu(bn(.)) is computed in the producing conv's
    # epilogue from the float32 accumulator (the unrounded value), in train mode by a separate pass over the stored tensor.
    def pre_t(t):  # -> (float32 value whose gradient is stored in bf16, the stored tensor)
        t = qg(t)
        return t, (qf(t) if _QUANT["pre_bf16"] else t)

    def bn_in(pair):
        return pair[1] if training else pair[0]

    out = pre_t(cj(F.conv2d(x, qw(p["conv1.weight"]), None, 1, 1)))  # :77,93
    in_planes = 64
    for li, (planes, stride0) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
        for bi, stride in enumerate([stride0, 1]):
            pre = "layer%d.%d." % (li, bi)
            o = qa(F.relu(_bn(p, b, pre + "bn1", bn_in(out), training)))  # :32
            if stride != 1 or in_planes != planes:  # :26-29,33
                sc = pre_t(cj(F.conv2d(o, qw(p[pre + "shortcut.0.weight"]), None, stride, 0)))[1]
            else:
                sc = out[1]
            c1 = pre_t(cj(F.conv2d(o, qw(p[pre + "conv1.weight"]), None, stride, 1)))  # :34
            o = cj(F.conv2d(qa(F.relu(_bn(p, b, pre + "bn2", bn_in(c1), training))), qw(p[pre + "conv2.weight"]), None, 1, 1))  # :35
            out = pre_t(o + sc)  # :39
            in_planes = planes
    out = F.avg_pool2d(out[1], 4)  # :99
    out = out.view(out.size(0), -1)
    return F.linear(out, p["linear.weight"], p["linear.bias"])  # :101


def resnet18_forward(p: dict, b: dict, x: torch.Tensor, training: bool):
    """ResNet18 (classifier_models/resnet.py:15-37,68-106), post-activation BasicBlocks."""
    out = qa(F.relu(_bn(p, b, "bn1", qp(F.conv2d(x, qw(p["conv1.weight"]), None, 1, 1)), training)))  # :90
    in_planes = 64
    for li, (planes, stride0) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
        for bi, stride in enumerate([stride0, 1]):
            pre = "layer%d.%d." % (li, bi)
            o = qa(F.relu(_bn(p, b, pre + "bn1", qp(F.conv2d(out, qw(p[pre + "conv1.weight"]), None, stride, 1)), training)))
            o = _bn(p, b, pre + "bn2", qp(F.conv2d(o, qw(p[pre + "conv2.weight"]), None, 1, 1)), training)
            if stride != 1 or in_planes != planes:  # :25-29
                sc = qa(_bn(p, b, pre + "shortcut.1", qp(F.conv2d(out, qw(p[pre + "shortcut.0.weight"]), None, stride, 0)), training))
            else:
                sc = out
            out = qa(F.relu(o + sc))  # :34-35
            in_planes = planes
    out = F.avg_pool2d(out, 4)  # :95
    out = out.view(out.size(0), -1)
    return F.linear(out, p["linear.weight"], p["linear.bias"])


def frequency_model_forward(p: dict, b: dict, x: torch.Tensor):
    """FrequencyModel in eval mode (defenses/frequency_based/model.py:8-52):
    conv -> ELU -> BN (x6), maxpool after 2/4/6, dropout = identity, flatten, linear."""
    for i in range(1, 7):  # quantised(): conv1 is float32 arithmetic (raw DCT coefficients), conv2..6 bf16 operands; bf16 storage
        x = F.conv2d(x, p["conv%d.weight" % i] if i == 1 else qw(p["conv%d.weight" % i]), p["conv%d.bias" % i], 1, 1)
        x = F.elu(x)
        x = qa(_bn(p, b, "bn%d" % i, x, False))
        if i % 2 == 0:
            x = F.max_pool2d(x, 2)
    x = x.flatten(1)
    return F.linear(x, p["linear6.weight"], p["linear6.bias"])


CLASSIFIERS = {"preact_resnet18": preact_resnet18_forward, "resnet18": resnet18_forward}


def split_state(sd: dict):
    """Split a state_dict into (parameters, buffers) by the reference's naming."""
    p, b = {}, {}
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
            b[k] = v
        else:
            p[k] = v
    return p, b


# --------------------------------------------------------------------------
# SGD (torch.optim.SGD(momentum=.9, weight_decay=5e-4, nesterov=True), train_generator.py:123-126)
# --------------------------------------------------------------------------


def sgd_nesterov_step(params: dict, grads: dict, bufs: dict, lr: float, momentum=0.9, wd=5e-4):
    """g <- g + wd*p; buf <- g (first step) or mu*buf + g; p <- p - lr*(g + mu*buf).
    Parameters with grad None are skipped (torch semantics)."""
    for k, pt in params.items():
        g = grads.get(k)
        if g is None:
            continue
        g = g + wd * pt
        if k not in bufs:
            bufs[k] = g.clone()
        else:
            bufs[k].mul_(momentum).add_(g)
        g = g + momentum * bufs[k]
        pt.add_(g, alpha=-lr)


# --------------------------------------------------------------------------
# PostTensorTransform (utils/dataloader.py:11-22,45-60) over kornia 0.6.6 (requirements.txt:12)
#
# kornia is a third-party dependency that is absent from /root/reference and from this image, so its published algorithm is
# restated with the torch ops kornia itself dispatches to:
#   A.RandomCrop(size, padding=p)   -> F.pad(x, [p,p,p,p], "constant", 0) then the window x[.., ys:ys+H, xs:xs+W] per sample,
#                                      (xs, ys) = floor(U[0, 2p+1))                         (kornia random_crop_generator)
#   A.RandomRotation(d)             -> angle ~ U(-d, d); M = get_rotation_matrix2d(((W-1)/2,(H-1)/2), angle, 1)
#                                      = [[a, b, (1-a)cx - b cy], [-b, a, b cx + (1-a) cy]], a = cos, b = sin;
#                                      warp_affine(x, M, (H,W), "bilinear", "zeros", align_corners=True), i.e.
#                                      F.grid_sample(x, F.affine_grid(normalised(M)^-1))     (kornia warp_affine)
#   A.RandomHorizontalFlip(p=0.5)   -> per-sample Bernoulli(p); flipped rows are x.flip(-1)
#   ProbTransform(f, p)             -> one random.random() < p per call gates the whole batch (dataloader.py:17)
# PARITY UNPINNED for the kornia RNG stream: the reference holds no test or vector for it and the package cannot be run here;
# given the parameters, pixels and gradients follow the algorithm above.
# --------------------------------------------------------------------------


def draw_post_transform(rows: int, opt):
    """The random decisions of ONE PostTensorTransform call, in the module order random_crop, random_rotation,
    random_horizontal_flip (utils/dataloader.py:48-55,58-59).  Gates: python `random`; per-sample parameters: torch CPU
    generator, one torch.rand(rows) per parameter (xs, ys, angle, flip).  Returns a dict of per-sample tensors."""
    import random as _random
    option = getattr(opt, "post_transform_option", "no_use")
    pad = int(getattr(opt, "random_crop", 5))
    out = {"xs": torch.full((rows,), pad, dtype=torch.long), "ys": torch.full((rows,), pad, dtype=torch.long), "crop": False,
           "angle": torch.zeros(rows), "rot": False, "flip": torch.zeros(rows, dtype=torch.bool), "pad": pad}
    if option == "no_use":
        return out
    if option != "use_modified":
        if _random.random() < 0.8:
            span = float(2 * pad + 1)
            out["xs"] = torch.floor(torch.rand(rows) * span).long()
            out["ys"] = torch.floor(torch.rand(rows) * span).long()
            out["crop"] = True
    if _random.random() < 0.5:
        deg = float(getattr(opt, "random_rotation", 10))
        out["angle"] = torch.rand(rows) * (2.0 * deg) - deg
        out["rot"] = True
    if getattr(opt, "dataset", "cifar10") == "cifar10":
        out["flip"] = torch.rand(rows) < 0.5
    return out


def _normal_transform_pixel(H: int, W: int) -> torch.Tensor:
    """kornia normal_transform_pixel: pixel coordinates -> [-1, 1] with the (size-1) convention (align_corners=True)."""
    return torch.tensor([[2.0 / (W - 1), 0.0, -1.0], [0.0, 2.0 / (H - 1), -1.0], [0.0, 0.0, 1.0]])


def apply_post_transform(x: torch.Tensor, prm: dict) -> torch.Tensor:
    """x: [rows, C, H, W] float32 (differentiable) -> the transformed batch."""
    rows, C, H, W = x.shape
    pad = prm["pad"]
    if prm["crop"]:
        xp = F.pad(x, [pad, pad, pad, pad], mode="constant", value=0.0)
        x = torch.stack([xp[i, :, int(prm["ys"][i]):int(prm["ys"][i]) + H, int(prm["xs"][i]):int(prm["xs"][i]) + W]
                         for i in range(rows)], 0) if rows else x
    if prm["rot"] and rows:
        rad = prm["angle"].to(torch.float32) * (math.pi / 180.0)
        a, b = torch.cos(rad), torch.sin(rad)
        cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
        M = torch.zeros(rows, 3, 3)
        M[:, 0, 0], M[:, 0, 1], M[:, 0, 2] = a, b, (1 - a) * cx - b * cy
        M[:, 1, 0], M[:, 1, 1], M[:, 1, 2] = -b, a, b * cx + (1 - a) * cy
        M[:, 2, 2] = 1.0
        Nrm = _normal_transform_pixel(H, W)
        dst_norm_trans_src_norm = Nrm @ M @ torch.inverse(Nrm)          # kornia normalize_homography
        src_norm_trans_dst_norm = torch.inverse(dst_norm_trans_src_norm)
        grid = F.affine_grid(src_norm_trans_dst_norm[:, :2, :], [rows, C, H, W], align_corners=True)
        x = F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    if bool(prm["flip"].any()):
        x = torch.where(prm["flip"].view(-1, 1, 1, 1), x.flip(-1), x)
    return x


def post_transform(x: torch.Tensor, opt, log=None) -> torch.Tensor:
    """transforms(x) of the training loops (train_generator.py:196,214,227,228,250).  `log`: list collecting the drawn
    parameters of every call (so that a test can hand the SAME decisions to the CUDA path)."""
    if getattr(opt, "post_transform_option", "no_use") == "no_use":
        return x
    prm = draw_post_transform(x.shape[0], opt)
    if log is not None:
        log.append(prm)
    return apply_post_transform(x, prm)


# --------------------------------------------------------------------------
# The alternated step (train_generator.py:170-255)
# --------------------------------------------------------------------------


def default_opt(**kw):
    """The hot-path flags of config.py:4-86 with their defaults."""
    o = SimpleNamespace(
        input_height=32, input_width=32, input_channel=3, num_classes=10, attack_mode="all2one",
        noise_rate=0.08, target_label=0, pc=0.5, ratio=0.65, kernel_size=3, sigma=(0.1, 1.0),
        L2_weight=0.02, clean_model_weight=0.8, lr_C=1e-2, lr_G=1e-2, classifier="preact_resnet18",
        post_transform_option="no_use", random_crop=5, random_rotation=10, dataset="cifar10",
        variant="", tv_weight=0.01,   # variant "imperceptible": train_generator_imperceptible.py (+ tv_weight * TV(x_bd).mean())
        s=2, grid_rescale=0.15,       # variant "wanet": train_generator_wanet.py (GridGenerator control grid size, flow strength)
        cross_weight=0.2,             # variant "inputaware": train_generator_inputaware.py (+ cross_weight * CE(netC(x + G(x2)), y))
    )
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def total_variation(img: torch.Tensor) -> torch.Tensor:
    """kornia.losses.total_variation of kornia 0.6.6 (requirements.txt:12; the package is absent here, restated from its published
    source): per image, the sum over (C, H, W) of |img[.., 1:, :] - img[.., :-1, :]| plus |img[.., :, 1:] - img[.., :, :-1]|."""
    d1 = img[..., 1:, :] - img[..., :-1, :]
    d2 = img[..., :, 1:] - img[..., :, :-1]
    return d1.abs().sum(dim=(-3, -2, -1)) + d2.abs().sum(dim=(-3, -2, -1))


def make_bd(netG_p, x, opt, sigma, y=None):
    """noise = low_freq(netG(x)); x_bd = blur(clamp(x + noise*noise_rate, -1, 1))
    (train_generator.py:189-194 / :223-226)."""
    raw = unet_forward(netG_p, x, y, opt.num_classes if y is not None else None)
    noise = low_freq(raw, opt.input_height, opt.ratio)
    x_bd = torch.clamp(x + noise * opt.noise_rate, -1, 1)
    return gaussian_blur(x_bd, sigma, opt.kernel_size), noise, raw


def alternated_step(state: dict, x: torch.Tensor, y: torch.Tensor, opt, with_metrics: bool = True, grad_hook=None,
                    buf_hook=None, x2: torch.Tensor | None = None) -> dict:
    """One iteration of train() (train_generator.py:170-255); PostTensorTransform per opt.post_transform_option (identity for
    "no_use", the default here; the five calls :196,:214,:227,:228,:250 draw their own parameters, logged in out["tf"]).

    `state` holds: netC_p/netC_b, clean_p/clean_b, netG_p (dicts of tensors, updated IN PLACE),
    netF_p/netF_b (optional), momC/momG (momentum buffer dicts, {} before the first step).
    RNG order: numpy rand(n_trg) -> torch uniform (if num_bd>0) -> torch uniform.
    grad_hook(name, grads_dict) / buf_hook(buffers_dict): data-parallel exchange points (local-BN policy) --
    called after each backward, before the optimiser step / after the C-step update; None for the single-process step.
    x2: the batch of the SECOND loader of train_generator_inputaware.py (:180-185), variant "inputaware" only: the G-step also
    builds inputs_bd2 = blur(clamp(x + low_freq(netG(x2)) * noise_rate)) -- the trigger of ANOTHER image on this image, its own
    sigma draw right after the G-step's (:234-239) -- and adds cross_weight * CE(netC(T(inputs_bd2)), y) (:241,246,259-264); the
    transform calls of the G-step run in the order x, inputs_bd2, inputs_bd (:240-242); no gradient-image loss in this variant.
    Returns a dict of everything observable (indices, losses, logits, metric counts)."""
    fwdC = CLASSIFIERS[opt.classifier]
    inputaware = getattr(opt, "variant", "") == "inputaware"
    if inputaware and x2 is None:
        raise ValueError("the inputaware step needs the second loader's batch")
    netC_p, netC_b, netG_p = state["netC_p"], state["netC_b"], state["netG_p"]
    clean_p, clean_b = state["clean_p"], state["clean_b"]
    out = {}
    bd_targets = create_targets_bd(y, opt.attack_mode, opt.target_label, opt.num_classes)

    # ---------------- C-step (:176-212) ----------------
    trg_ind, ntrg_ind, num_bd = select_poison(y, bd_targets, opt.pc)
    out["trg_ind"], out["ntrg_ind"], out["num_bd"] = trg_ind.clone(), ntrg_ind.clone(), num_bd
    x_sel = x[trg_ind[:num_bd]]
    for t in netC_p.values():
        t.requires_grad_(True)
        t.grad = None
    with torch.no_grad():  # G grads from this backward are discarded at :220
        if num_bd > 0:
            sigma_c = draw_sigma(*opt.sigma)
            x_bd_c, _, _ = make_bd(netG_p, x_sel, opt, sigma_c)
        else:
            sigma_c = None
            x_bd_c = x_sel
    out["sigma_c"] = sigma_c
    total_x = torch.cat([x_bd_c, x[trg_ind[num_bd:]], x[ntrg_ind]], dim=0)  # :195
    total_y = torch.cat([bd_targets[trg_ind[:num_bd]], y[trg_ind[num_bd:]], y[ntrg_ind]], dim=0)  # :197-204
    out["tf"] = tf_log = []
    total_x = post_transform(total_x, opt, tf_log)  # :196
    logits_c = fwdC(netC_p, netC_b, total_x, True)  # netC.train()
    loss_c = F.cross_entropy(logits_c, total_y)
    loss_c.backward()
    out["total_x"], out["total_y"] = total_x.detach(), total_y
    out["logits_c"], out["loss_c"] = logits_c.detach().clone(), float(loss_c)
    gradsC = {k: v.grad for k, v in netC_p.items()}
    out["gradsC"] = {k: g.clone() for k, g in gradsC.items()}
    if grad_hook is not None:
        grad_hook("netC", gradsC)
    with torch.no_grad():
        for t in netC_p.values():
            t.requires_grad_(False)
        sgd_nesterov_step(netC_p, gradsC, state["momC"], opt.lr_C)
        if buf_hook is not None:
            buf_hook(netC_b)
        if with_metrics:
            out["clean_preds"] = fwdC(clean_p, clean_b, post_transform(x, opt, tf_log), False)  # :214

    # ---------------- G-step (:217-255) ----------------
    for t in netG_p.values():
        t.requires_grad_(True)
        t.grad = None
    sigma_g = draw_sigma(*opt.sigma)
    out["sigma_g"] = sigma_g
    x_bd, noise, noise_raw = make_bd(netG_p, x, opt, sigma_g)
    if inputaware:  # train_generator_inputaware.py:238-239
        sigma_g2 = draw_sigma(*opt.sigma)
        noise2 = low_freq(unet_forward(netG_p, x2), opt.input_height, opt.ratio)
        x_bd2 = gaussian_blur(torch.clamp(x + noise2 * opt.noise_rate, -1, 1), sigma_g2, opt.kernel_size)
        out.update(sigma_g2=sigma_g2, x_bd2=x_bd2.detach())
    with torch.no_grad():
        if with_metrics or inputaware:
            out["pred_clean"] = fwdC(netC_p, netC_b, post_transform(x, opt, tf_log), False)  # :227
    if inputaware:
        pred_cross = fwdC(netC_p, netC_b, post_transform(x_bd2, opt, tf_log), False)  # inputaware :241
        loss_cross = F.cross_entropy(pred_cross, y)  # :246
        out.update(pred_cross=pred_cross.detach(), loss_cross=float(loss_cross), n_cross_correct=int((torch.argmax(pred_cross, 1) == y).sum()))
    pred_bd = fwdC(netC_p, netC_b, post_transform(x_bd, opt, tf_log), False)  # :228 netC.eval(), weights AFTER the C update
    loss_ce = F.cross_entropy(pred_bd, bd_targets)  # :231
    loss_l2 = F.mse_loss(x_bd, x)  # :234
    with torch.no_grad():  # :235-243, logged only
        xe, be = F.pad(x, (1, 1, 2, 1)), F.pad(x_bd.detach(), (1, 1, 2, 1))
        out["loss_grad_l2"] = float(F.mse_loss(xe[:, :, 1:] - xe[:, :, :-1], be[:, :, 1:] - be[:, :, :-1])
                                    + F.mse_loss(xe[:, :, :, 1:] - xe[:, :, :, :-1], be[:, :, :, 1:] - be[:, :, :, :-1]))
    clean_model_preds = fwdC(clean_p, clean_b, post_transform(x_bd, opt, tf_log), False)  # :250
    clean_model_loss = F.cross_entropy(clean_model_preds, y)  # :251
    loss = loss_ce + opt.L2_weight * loss_l2 + opt.clean_model_weight * clean_model_loss  # :253
    if getattr(opt, "variant", "") == "imperceptible":  # train_generator_imperceptible.py:228,235-237
        loss_tv = total_variation(x_bd).mean()
        loss = loss + opt.tv_weight * loss_tv
        out["loss_tv"] = float(loss_tv)
    if inputaware:  # :259-264
        loss = loss + opt.cross_weight * loss_cross
    loss.backward()
    gradsG = {k: v.grad for k, v in netG_p.items()}
    out["gradsG"] = {k: g.clone() for k, g in gradsG.items()}
    if grad_hook is not None:
        grad_hook("netG", gradsG)
    with torch.no_grad():
        for t in netG_p.values():
            t.requires_grad_(False)
        sgd_nesterov_step(netG_p, gradsG, state["momG"], opt.lr_G)
        if with_metrics and state.get("netF_p") is not None:
            inputs_F = dct_2d(((x_bd.detach() + 1) / 2 * 255).byte())  # :245 (uint8 -> float32)
            out["inputs_F"] = inputs_F
            out["pred_F"] = frequency_model_forward(state["netF_p"], state["netF_b"], inputs_F)  # :247
    out.update(
        x_bd=x_bd.detach(), noise=noise.detach(), noise_raw=noise_raw.detach(), pred_bd=pred_bd.detach(), clean_model_preds=clean_model_preds.detach(),
        loss_ce=float(loss_ce), loss_l2=float(loss_l2), clean_model_loss=float(clean_model_loss), loss_g=float(loss),
        bd_targets=bd_targets,
    )
    if with_metrics:  # :262-267
        am = lambda t: torch.argmax(t, dim=1)
        out["n_clean_correct"] = int((am(out["pred_clean"]) == y).sum())
        out["n_bd_correct"] = int((am(out["pred_bd"]) == bd_targets).sum())
        out["n_clean_model_correct"] = int((am(out["clean_preds"]) == y).sum())
        out["n_clean_model_bd_ba"] = int((am(out["clean_model_preds"]) == y).sum())
        out["n_clean_model_bd_asr"] = int((am(out["clean_model_preds"]) == bd_targets).sum())
        if "pred_F" in out:
            out["n_F_correct"] = int((am(out["pred_F"]) == 1).sum())
    return out


# --------------------------------------------------------------------------
# evaluation loop (train_generator.py:321-465), one batch
# --------------------------------------------------------------------------


def eval_batch(state: dict, x: torch.Tensor, y: torch.Tensor, opt, x2: torch.Tensor | None = None) -> dict:
    """One iteration of eval() (:355-391): clean accuracy of netC, attack success on ALL non-target samples, detector
    and clean-model legs.  RNG: one torch uniform per batch (GaussianBlur.get_params, :372).
    x2 (variant "inputaware", train_generator_inputaware.py:402-413): the second loader's batch; the trigger of its rows on x
    (a second sigma draw, right after the first), netC's accuracy on the non-target rows against their true labels."""
    fwdC = CLASSIFIERS[opt.classifier]
    netC_p, netC_b, netG_p = state["netC_p"], state["netC_b"], state["netG_p"]
    clean_p, clean_b = state["clean_p"], state["clean_b"]
    am = lambda t: torch.argmax(t, dim=1)
    out = {}
    with torch.no_grad():
        preds_clean = fwdC(netC_p, netC_b, x, False)  # :360
        ntrg = (y != opt.target_label).nonzero()[:, 0]  # :366
        x_sel, y_sel = x[ntrg], y[ntrg]
        if getattr(opt, "variant", "") == "wanet":  # train_generator_wanet.py:352-364: warp trigger, nothing drawn
            sigma = None
            x_bd = wanet_warp(x_sel, grid_generator_forward(netG_p, x_sel, opt.s), opt)[0]
        else:
            sigma = draw_sigma(*opt.sigma)
            x_bd, _, _ = make_bd(netG_p, x_sel, opt, sigma)  # :369-373
        bd_t = create_targets_bd(y_sel, opt.attack_mode, opt.target_label, opt.num_classes)
        preds_bd = fwdC(netC_p, netC_b, x_bd, False)
        if x2 is not None:
            sigma2 = draw_sigma(*opt.sigma)
            noise2 = low_freq(unet_forward(netG_p, x2), opt.input_height, opt.ratio)
            x_bd2 = gaussian_blur(torch.clamp(x + noise2 * opt.noise_rate, -1, 1), sigma2, opt.kernel_size)
            preds_cross = fwdC(netC_p, netC_b, x_bd2, False)
            out.update(sigma2=sigma2, x_bd2=x_bd2, preds_cross=preds_cross, cross_correct=int((am(preds_cross[ntrg]) == y_sel).sum()))
        inputs_F = dct_2d(((x_bd + 1) / 2 * 255).byte())  # :381
        preds_F = frequency_model_forward(state["netF_p"], state["netF_b"], inputs_F)
        cm_clean = fwdC(clean_p, clean_b, x, False)
        cm_bd = fwdC(clean_p, clean_b, x_bd, False)
    out.update(sigma=sigma, ntrg=ntrg, x_bd=x_bd, preds_clean=preds_clean, preds_bd=preds_bd, preds_F=preds_F, cm_clean=cm_clean,
               cm_bd=cm_bd, n_clean=len(x), n_bd=len(ntrg),
               clean_correct=int((am(preds_clean) == y).sum()), bd_correct=int((am(preds_bd) == bd_t).sum()),
               F_correct=int((am(preds_F) == 1).sum()), cm_correct=int((am(cm_clean) == y).sum()),
               cm_bd_ba=int((am(cm_bd) == y_sel).sum()), cm_bd_asr=int((am(cm_bd) == bd_t).sum()))
    return out


def detector_test_batch(netF_p, netF_b, netG_p, x: torch.Tensor, opt) -> dict:
    """One iteration of defenses/frequency_based/test.py:76-103: trigger on every image (one blur sigma per batch), uint8 DCT of
    [clean ; poisoned] (the per-plane scipy `dct2` the script defines at :19-20 -- its loop calls the torch dct_2d on a numpy plane
    and raises as shipped), detector logits, accuracy against labels [0 ; 1] and detection rate on the poisoned half."""
    am = lambda t: torch.argmax(t, dim=1)
    with torch.no_grad():
        bs = x.shape[0]
        sigma = draw_sigma(*opt.sigma)
        poi_x, _, _ = make_bd(netG_p, x, opt, sigma)
        both = torch.cat([x, poi_x])
        coef = dct_2d(((both + 1) / 2 * 255).byte())
        preds = frequency_model_forward(netF_p, netF_b, coef)
        labels = torch.cat([torch.zeros(bs, dtype=torch.long), torch.ones(bs, dtype=torch.long)])
    return dict(sigma=sigma, poi_x=poi_x, coef=coef, preds=preds, correct=int((am(preds) == labels).sum()),
                detected=int((am(preds[bs:]) == 1).sum()))


def victim_train_step(netC_p, netC_b, netG_p, momC: dict, x: torch.Tensor, y: torch.Tensor, poisoned, opt) -> dict:
    """One iteration of train_victim.py:110-140 (poisoned given: per-sample bool flags from the dataset) or of
    train_clean_classifier.py:88-104 (poisoned None, netG_p None).  `(poisoned is False).nonzero()` at train_victim.py:121 raises
    as shipped; the rows whose flag is False are taken.  RNG: one torch uniform for the blur when a poisoned row exists (:129),
    then the transform draws (:131)."""
    fwdC = CLASSIFIERS[opt.classifier]
    out = {}
    for t in netC_p.values():
        t.requires_grad_(True)
        t.grad = None
    if poisoned is None:
        total_x, total_y = x, y
        out["num_bd"] = 0
    else:
        pz = torch.as_tensor(poisoned).bool()
        bd_targets = create_targets_bd(y, opt.attack_mode, opt.target_label, opt.num_classes)
        trg_ind, ntrg_ind = pz.nonzero()[:, 0], (~pz).nonzero()[:, 0]
        out["num_bd"] = int(trg_ind.shape[0])
        with torch.no_grad():
            x_sel = x[trg_ind]
            if out["num_bd"] > 0:
                out["sigma"] = draw_sigma(*opt.sigma)
                x_bd, _, _ = make_bd(netG_p, x_sel, opt, out["sigma"])
            else:
                x_bd = x_sel
        total_x = torch.cat([x_bd, x[ntrg_ind]], dim=0)  # :130
        total_y = torch.cat([bd_targets[trg_ind], y[ntrg_ind]], dim=0)  # :132
    out["tf"] = tf_log = []
    total_in = post_transform(total_x, opt, tf_log)  # :131
    logits = fwdC(netC_p, netC_b, total_in, True)
    loss = F.cross_entropy(logits, total_y)
    loss.backward()
    grads = {k: v.grad for k, v in netC_p.items()}
    with torch.no_grad():
        for t in netC_p.values():
            t.requires_grad_(False)
        sgd_nesterov_step(netC_p, grads, momC, opt.lr_C)
    out.update(total_x=total_x.detach(), total_y=total_y, logits=logits.detach(), loss=float(loss.detach()),
               n_correct=int((torch.argmax(logits, 1) == total_y).sum()), grads={k: g.clone() for k, g in grads.items()})
    return out


def victim_eval_batch(netC_p, netC_b, netG_p, x: torch.Tensor, y: torch.Tensor, opt, x2: torch.Tensor | None = None) -> dict:
    """One iteration of eval.py:115-141 (the stand-alone evaluator): clean accuracy, and on the gathered non-target rows the
    benign accuracy and attack success of the triggered images.  RNG: one torch uniform per batch (GaussianBlur, :131).
    x2: train_victim_inputaware.py:213-223 -- the cross-trigger accuracy (trigger of the second loader's rows on x, second sigma
    draw, non-target rows against their true labels)."""
    fwdC = CLASSIFIERS[opt.classifier]
    am = lambda t: torch.argmax(t, dim=1)
    with torch.no_grad():
        preds_clean = fwdC(netC_p, netC_b, x, False)  # :119
        ntrg = (y != opt.target_label).nonzero()[:, 0]  # :125
        x_sel, y_sel = x[ntrg], y[ntrg]
        sigma = draw_sigma(*opt.sigma)
        x_bd, _, _ = make_bd(netG_p, x_sel, opt, sigma)  # :128-131
        bd_t = create_targets_bd(y_sel, opt.attack_mode, opt.target_label, opt.num_classes)
        preds_bd = fwdC(netC_p, netC_b, x_bd, False)  # :133
        extra = {}
        if x2 is not None:
            sigma2 = draw_sigma(*opt.sigma)
            noise2 = low_freq(unet_forward(netG_p, x2), opt.input_height, opt.ratio)
            x_bd2 = gaussian_blur(torch.clamp(x + noise2 * opt.noise_rate, -1, 1), sigma2, opt.kernel_size)
            preds_cross = fwdC(netC_p, netC_b, x_bd2, False)
            extra = dict(sigma2=sigma2, x_bd2=x_bd2, preds_cross=preds_cross, cross_correct=int((am(preds_cross[ntrg]) == y_sel).sum()))
    return dict(**extra, sigma=sigma, ntrg=ntrg, x_bd=x_bd, preds_clean=preds_clean, preds_bd=preds_bd, n_clean=len(x), n_bd=len(ntrg),
                clean_correct=int((am(preds_clean) == y).sum()), bd_ba=int((am(preds_bd) == y_sel).sum()),
                bd_asr=int((am(preds_bd) == bd_t).sum()))


# --------------------------------------------------------------------------
# multilabel variant (train_generator_multilabel.py:160-242): conditional generator, class-chunked G-step
# --------------------------------------------------------------------------


def make_bd_cond(netG_p, x, labels, opt, sigma):
    """create_inputs_bd (train_generator_multilabel.py:66-75) with the sigma drawn by the caller."""
    noise_raw = unet_forward(netG_p, x, labels, opt.num_classes)
    noise = low_freq(noise_raw, opt.input_height, opt.ratio)
    x_bd = torch.clamp(x + noise * opt.noise_rate, -1, 1)
    return gaussian_blur(x_bd, sigma), noise, noise_raw


def eval_batch_multilabel(state: dict, x: torch.Tensor, y: torch.Tensor, opt) -> dict:
    """One iteration of train_generator_multilabel.eval() (:343-378; train_victim_multilabel.py is the same file): clean accuracy
    of netC and clean_model; then for EVERY class ci the whole batch is triggered towards ci (one sigma draw per class, the
    module-level GaussianBlur :53), attack success / clean-model BA and ASR on the rows whose label is not ci, detector on all rows."""
    fwdC = CLASSIFIERS[opt.classifier]
    netC_p, netC_b, netG_p = state["netC_p"], state["netC_b"], state["netG_p"]
    clean_p, clean_b = state["clean_p"], state["clean_b"]
    am = lambda t: torch.argmax(t, dim=1)
    out = dict(sigmas=[], n_clean=len(x), n_bd=0, bd_correct=0, cm_bd_ba=0, cm_bd_asr=0, F_correct=0, x_bd=[], preds_bd=[])
    with torch.no_grad():
        out["clean_correct"] = int((am(fwdC(netC_p, netC_b, x, False)) == y).sum())
        out["cm_correct"] = int((am(fwdC(clean_p, clean_b, x, False)) == y).sum())
        for ci in range(opt.num_classes):
            tmp = y * 0 + ci
            sigma = draw_sigma(0.1, 1.0)
            x_bd, _, _ = make_bd_cond(netG_p, x, tmp, opt, sigma)
            preds_bd = fwdC(netC_p, netC_b, x_bd, False)
            cm_bd = fwdC(clean_p, clean_b, x_bd, False)
            ntrg = (y != tmp).nonzero()[:, 0]
            out["sigmas"].append(sigma)
            out["n_bd"] += len(ntrg)
            out["bd_correct"] += int((am(preds_bd[ntrg]) == ci).sum())
            out["cm_bd_ba"] += int((am(cm_bd[ntrg]) == y[ntrg]).sum())
            out["cm_bd_asr"] += int((am(cm_bd[ntrg]) == ci).sum())
            preds_F = frequency_model_forward(state["netF_p"], state["netF_b"], dct_2d(((x_bd + 1) / 2 * 255).byte()))
            out["F_correct"] += int((am(preds_F) == 1).sum())
            out["x_bd"].append(x_bd)
            out["preds_bd"].append(preds_bd)
    return out


def multilabel_chunks(bs: int, num_classes: int):
    """:203-211 -- contiguous chunks of ps rows, chunk ci is pushed towards class ci."""
    ps = int((bs - 1) / num_classes) + 1
    out = []
    for ci in range(num_classes):
        si, ei = ci * ps, min(ci * ps + ps, bs)
        if si >= ei:
            break
        out.append((ci, si, ei))
    return out


def alternated_step_multilabel(state: dict, x: torch.Tensor, y: torch.Tensor, opt, with_metrics: bool = True) -> dict:
    """One iteration of train_generator_multilabel.train() (:160-242), identity PostTensorTransform.
    RNG order: numpy rand(bs) (:171) -> torch uniform for the C-step blur (only if num_bd > 0, :74) -> one torch uniform
    per class chunk of the G-step (:219 via create_inputs_bd)."""
    fwdC = CLASSIFIERS[opt.classifier]
    netC_p, netC_b, netG_p = state["netC_p"], state["netC_b"], state["netG_p"]
    clean_p, clean_b = state["clean_p"], state["clean_b"]
    bs = x.shape[0]
    out = {}
    # ---------------- C-step (:163-191)
    num_bd = int(np.sum(np.random.rand(bs) < opt.pc))  # :171
    out["num_bd"] = num_bd
    for t in netC_p.values():
        t.requires_grad_(True)
        t.grad = None
    with torch.no_grad():
        if num_bd > 0:
            sigma_c = draw_sigma(*opt.sigma)
            x_bd_c, _, _ = make_bd_cond(netG_p, x[:num_bd], y[:num_bd], opt, sigma_c)  # conditioned on the TRUE labels
        else:
            sigma_c, x_bd_c = None, x[:0]
    out["sigma_c"] = sigma_c
    total_x = torch.cat([x_bd_c, x[num_bd:]], dim=0)  # :178
    out["tf"] = tf_log = []
    total_x = post_transform(total_x, opt, tf_log)  # :179
    logits_c = fwdC(netC_p, netC_b, total_x, True)
    loss_c = F.cross_entropy(logits_c, y)  # :180-183, labels unchanged
    loss_c.backward()
    out["total_x"], out["logits_c"], out["loss_c"] = total_x.detach(), logits_c.detach().clone(), float(loss_c.detach())
    gradsC = {k: v.grad for k, v in netC_p.items()}
    with torch.no_grad():
        for t in netC_p.values():
            t.requires_grad_(False)
        sgd_nesterov_step(netC_p, gradsC, state["momC"], opt.lr_C)
        if with_metrics:
            out["clean_preds"] = fwdC(clean_p, clean_b, post_transform(x, opt, tf_log), False)  # :190
    # ---------------- G-step (:193-232)
    for t in netG_p.values():
        t.requires_grad_(True)
        t.grad = None
    with torch.no_grad():
        if with_metrics:
            out["pred_clean"] = fwdC(netC_p, netC_b, post_transform(x, opt, tf_log), False)  # :199
    parts, bd_t, sigmas = [], [], []
    for ci, si, ei in multilabel_chunks(bs, opt.num_classes):
        tmp = y[si:ei] * 0 + ci
        sg = draw_sigma(*opt.sigma)
        sigmas.append(sg)
        xb, _, _ = make_bd_cond(netG_p, x[si:ei], tmp, opt, sg)
        parts.append(xb)
        bd_t.append(tmp)
    x_bd = torch.cat(parts, 0)
    bd_targets = torch.cat(bd_t, 0)
    out["sigmas_g"] = sigmas
    pred_bd = fwdC(netC_p, netC_b, post_transform(x_bd, opt, tf_log), False)  # :225
    loss_ce = F.cross_entropy(pred_bd, bd_targets)
    loss_l2 = F.mse_loss(x_bd, x)
    clean_model_preds = fwdC(clean_p, clean_b, post_transform(x_bd, opt, tf_log), False)  # :236
    clean_model_loss = F.cross_entropy(clean_model_preds, y)
    loss = loss_ce + opt.L2_weight * loss_l2 + opt.clean_model_weight * clean_model_loss  # :239
    loss.backward()
    gradsG = {k: v.grad for k, v in netG_p.items()}
    with torch.no_grad():
        for t in netG_p.values():
            t.requires_grad_(False)
        sgd_nesterov_step(netG_p, gradsG, state["momG"], opt.lr_G)
        if with_metrics and state.get("netF_p") is not None:
            inputs_F = dct_2d(((x_bd.detach() + 1) / 2 * 255).byte())  # :230
            out["inputs_F"] = inputs_F
            out["pred_F"] = frequency_model_forward(state["netF_p"], state["netF_b"], inputs_F)
    out.update(x_bd=x_bd.detach(), pred_bd=pred_bd.detach(), clean_model_preds=clean_model_preds.detach(),
               loss_ce=float(loss_ce.detach()), loss_l2=float(loss_l2.detach()), clean_model_loss=float(clean_model_loss.detach()),
               loss_g=float(loss.detach()), bd_targets=bd_targets)
    if with_metrics:  # :247-252
        am = lambda t: torch.argmax(t, dim=1)
        out["n_clean_correct"] = int((am(out["pred_clean"]) == y).sum())
        out["n_bd_correct"] = int((am(out["pred_bd"]) == bd_targets).sum())
        out["n_clean_model_correct"] = int((am(out["clean_preds"]) == y).sum())
        out["n_clean_model_bd_ba"] = int((am(out["clean_model_preds"]) == y).sum())
        out["n_clean_model_bd_asr"] = int((am(out["clean_model_preds"]) == bd_targets).sum())
        if "pred_F" in out:
            out["n_F_correct"] = int((am(out["pred_F"]) == 1).sum())
    return out


# --------------------------------------------------------------------------
# deterministic random-init state (shapes of the reference modules), no reference import
# --------------------------------------------------------------------------


def _uniform(shape, bound, gen):
    return torch.empty(shape).uniform_(-bound, bound, generator=gen)


def _kaiming_bound(fan_in):
    """torch.nn.init.kaiming_uniform_(w, a=sqrt(5)) bound, evaluated the way torch does so that the
    draw is bit-identical to constructing nn.Conv2d / nn.Linear under the same seed."""
    gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
    std = gain / math.sqrt(fan_in)
    return math.sqrt(3.0) * std


def _conv_init(cout, cin, k, bias, gen):
    """nn.Conv2d default init (weight: kaiming_uniform a=sqrt(5); bias: U(+-1/sqrt(fan_in)))."""
    fan_in = cin * k * k
    w = _uniform((cout, cin, k, k), _kaiming_bound(fan_in), gen)
    bb = _uniform((cout,), 1 / math.sqrt(fan_in), gen) if bias else None
    return w, bb


def _linear_init(cout, cin, gen):
    w = _uniform((cout, cin), _kaiming_bound(cin), gen)
    bb = _uniform((cout,), 1 / math.sqrt(cin), gen)
    return w, bb


def init_unet_state(gen, in_ch=3, nf=64, num_classes=0):
    """Parameter shapes of UnetGenerator / CUnetGeneratorv1 (networks/models.py:268-316,472-521)."""
    spec = [("conv0_0", in_ch, nf), ("conv0_1", nf + num_classes, nf), ("conv1_0", nf, nf * 2), ("conv1_1", nf * 2, nf * 2),
            ("conv2_0", nf * 2, nf * 4), ("conv2_1", nf * 4, nf * 4), ("conv3_0", nf * 4, nf * 8), ("conv3_1", nf * 8, nf * 8),
            ("upconv3_1", nf * 8, nf * 8), ("upconv3_0", nf * 8, nf * 4), ("upconv2_1", nf * 4, nf * 4),
            ("upconv2_0", nf * 4, nf * 2), ("upconv1_1", nf * 2, nf * 2), ("upconv1_0", nf * 2, nf),
            ("upconv0_1", nf, nf), ("upconv0_0", nf, in_ch)]
    p = {}
    for name, ci, co in spec:
        w, bb = _conv_init(co, ci, 3, True, gen)
        p[name + ".weight"], p[name + ".bias"] = w, bb
    return p


def init_grid_generator_state(gen, in_ch=3, nf=64, S=2):
    """Parameter shapes and construction order of GridGenerator (networks/models.py:344-370)."""
    spec = [("conv0_0", in_ch, nf), ("conv0_1", nf, nf), ("conv1_0", nf, nf * 2), ("conv1_1", nf * 2, nf * 2),
            ("conv2_0", nf * 2, nf * 4), ("conv2_1", nf * 4, nf * 4), ("conv3_0", nf * 4, nf * 8), ("conv3_1", nf * 8, nf * 8)]
    p = {}
    for name, ci, co in spec:
        p[name + ".weight"], p[name + ".bias"] = _conv_init(co, ci, 3, True, gen)
    p["fc1.weight"], p["fc1.bias"] = _linear_init(nf, nf * 8, gen)
    p["fc2.weight"], p["fc2.bias"] = _linear_init(S * S * 2, nf, gen)
    return p


def _bn_state(p, b, name, c):
    p[name + ".weight"], p[name + ".bias"] = torch.ones(c), torch.zeros(c)
    b[name + ".running_mean"], b[name + ".running_var"] = torch.zeros(c), torch.ones(c)
    b[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def init_preact_resnet18_state(gen, num_classes=10, n_input=3, scaler=1):
    p, b = {}, {}
    p["conv1.weight"], _ = _conv_init(64, n_input, 3, False, gen)
    in_planes = 64
    for li, (planes, stride0) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
        for bi, stride in enumerate([stride0, 1]):
            pre = "layer%d.%d." % (li, bi)
            _bn_state(p, b, pre + "bn1", in_planes)
            p[pre + "conv1.weight"], _ = _conv_init(planes, in_planes, 3, False, gen)
            _bn_state(p, b, pre + "bn2", planes)
            p[pre + "conv2.weight"], _ = _conv_init(planes, planes, 3, False, gen)
            if stride != 1 or in_planes != planes:
                p[pre + "shortcut.0.weight"], _ = _conv_init(planes, in_planes, 1, False, gen)
            in_planes = planes
    p["linear.weight"], p["linear.bias"] = _linear_init(num_classes, 512 * scaler, gen)
    return p, b


def init_resnet18_state(gen, num_classes=10, n_input=3, scaler=4):
    p, b = {}, {}
    p["conv1.weight"], _ = _conv_init(64, n_input, 3, False, gen)
    _bn_state(p, b, "bn1", 64)
    in_planes = 64
    for li, (planes, stride0) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
        for bi, stride in enumerate([stride0, 1]):
            pre = "layer%d.%d." % (li, bi)
            p[pre + "conv1.weight"], _ = _conv_init(planes, in_planes, 3, False, gen)
            _bn_state(p, b, pre + "bn1", planes)
            p[pre + "conv2.weight"], _ = _conv_init(planes, planes, 3, False, gen)
            _bn_state(p, b, pre + "bn2", planes)
            if stride != 1 or in_planes != planes:
                p[pre + "shortcut.0.weight"], _ = _conv_init(planes, in_planes, 1, False, gen)
                _bn_state(p, b, pre + "shortcut.1", planes)
            in_planes = planes
    p["linear.weight"], p["linear.bias"] = _linear_init(num_classes, 512 * scaler, gen)
    return p, b


def init_frequency_model_state(gen, num_classes=2, n_input=3, scaler=1, randomize_bn_stats=False):
    p, b = {}, {}
    chans = [n_input, 32, 32, 64, 64, 128, 128]
    for i in range(1, 7):
        w, bb = _conv_init(chans[i], chans[i - 1], 3, True, gen)
        p["conv%d.weight" % i], p["conv%d.bias" % i] = w, bb
        _bn_state(p, b, "bn%d" % i, chans[i])
    p["linear6.weight"], p["linear6.bias"] = _linear_init(num_classes, 2048 * scaler, gen)
    if randomize_bn_stats:  # non-trivial eval statistics so the BN leg is actually exercised (drawn last)
        for i in range(1, 7):
            b["bn%d.running_mean" % i] = torch.randn(chans[i], generator=gen) * 0.1
            b["bn%d.running_var" % i] = torch.rand(chans[i], generator=gen) + 0.5
    return p, b


def init_step_state(seed: int = 0, classifier: str = "preact_resnet18", num_classes: int = 10, scaler: int = 1,
                    cond_classes: int = 0):
    """A complete, seeded state for `alternated_step` (synthetic weights of the reference shapes)."""
    gen = torch.Generator().manual_seed(seed)
    initC = init_preact_resnet18_state if classifier == "preact_resnet18" else init_resnet18_state
    netC_p, netC_b = initC(gen, num_classes=num_classes, scaler=scaler)
    clean_p, clean_b = initC(gen, num_classes=num_classes, scaler=scaler)
    netG_p = init_unet_state(gen, num_classes=cond_classes)
    netF_p, netF_b = init_frequency_model_state(gen, scaler=scaler)
    return dict(netC_p=netC_p, netC_b=netC_b, clean_p=clean_p, clean_b=clean_b, netG_p=netG_p,
                netF_p=netF_p, netF_b=netF_b, momC={}, momG={})
