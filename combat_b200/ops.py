"""Thin torch-tensor wrappers over the C-ABI kernels (include/combat_b200.h).

torch is used only for device memory and the current CUDA stream; every computation is a
hand-written sm_100a kernel in csrc/.  Nothing here falls back to torch math: tensors that are
not on a CUDA device raise.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from ._lib import ConvDesc, ConvTcDesc, WPrepDesc, check, lib

F32, BF16 = 0, 1


def dt_code(t: torch.Tensor | torch.dtype) -> int:
    d = t.dtype if torch.is_tensor(t) else t
    if d == torch.float32:
        return F32
    if d == torch.bfloat16:
        return BF16
    raise TypeError("combat_b200: unsupported dtype %s" % d)


def _p(t):
    if t is None:
        return None
    if isinstance(t, int):  # raw device address (a slice of a larger device buffer)
        return t
    if not t.is_cuda:
        raise RuntimeError("combat_b200 kernels need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


def _contig(t):
    if not t.is_contiguous():
        raise RuntimeError("combat_b200: tensor must be contiguous")
    return t


# ------------------------------------------------------------------ DCT family
_MAT_CACHE: dict = {}


def dct_matrix_np(N: int) -> np.ndarray:
    """Orthonormal DCT-II matrix (float64): D[k,n] = c_k sqrt(2/N) cos(pi (2n+1) k / 2N), the closed form of
    the reference's FFT route (utils/dct.py:13-42)."""
    n = np.arange(N, dtype=np.float64)
    D = np.sqrt(2.0 / N) * np.cos(np.pi * (2 * n[None, :] + 1) * n[:, None] / (2 * N))
    D[0, :] *= 1.0 / np.sqrt(2.0)
    return D


def transform_matrix(kind: str, N: int, device, keep: int = 0) -> torch.Tensor:
    """Device-resident N x N float32 matrix M such that the transform is M X M^T (constants, built once)."""
    key = (kind, N, keep, str(device))
    m = _MAT_CACHE.get(key)
    if m is None:
        D = dct_matrix_np(N)
        if kind == "dct":
            M = D
        elif kind == "idct":
            M = D.T
        elif kind == "lowfreq":
            M = D[:keep].T @ D[:keep]
        elif kind == "eye":
            M = np.eye(N)
        else:
            raise ValueError(kind)
        m = torch.from_numpy(np.ascontiguousarray(M).astype(np.float32)).to(device)
        _MAT_CACHE[key] = m
    return m


def plane_transform(x: torch.Tensor, M: torch.Tensor, in_mode: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
    """out[p] = M x[p] M^T over the last two (square) dims."""
    x = _contig(x)
    N = x.shape[-1]
    assert x.shape[-2] == N and M.shape == (N, N)
    planes = x.numel() // (N * N)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    if planes == 0:
        return out
    if N <= 64 and N % 4 == 0:
        check(lib.combat_plane_transform(_p(x), _p(out), _p(M), _p(M), planes, N, in_mode, None, _s()), "plane_transform")
    else:
        esz = x.element_size()
        chunk = 4096
        ws = torch.empty(min(planes, chunk) * N * N, dtype=torch.float32, device=x.device)
        for p0 in range(0, planes, chunk):
            n = min(chunk, planes - p0)
            check(lib.combat_plane_transform(x.data_ptr() + p0 * N * N * esz, out.data_ptr() + p0 * N * N * 4, _p(M), _p(M),
                                             n, N, in_mode, _p(ws), _s()), "plane_transform")
    return out


def plane_transform_lr(x, L, R, in_mode=0):
    """out[p] = L x[p] R^T with distinct left / right matrices (1-D transforms: L = I)."""
    x = _contig(x)
    N = x.shape[-1]
    planes = x.numel() // (N * N)
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    if planes == 0:
        return out
    if N <= 64 and N % 4 == 0:
        check(lib.combat_plane_transform(_p(x), _p(out), _p(L), _p(R), planes, N, in_mode, None, _s()), "plane_transform")
        return out
    esz = x.element_size()
    chunk = 4096
    ws = torch.empty(min(planes, chunk) * N * N, dtype=torch.float32, device=x.device)
    for p0 in range(0, planes, chunk):
        n = min(chunk, planes - p0)
        check(lib.combat_plane_transform(x.data_ptr() + p0 * N * N * esz, out.data_ptr() + p0 * N * N * 4, _p(L), _p(R), n, N,
                                         in_mode, _p(ws), _s()), "plane_transform")
    return out


_KIND = {"dct": 1, "idct": 2, "lowfreq": 3}


def plane_op(x: torch.Tensor, kind: str, keep: int = 0, in_mode: int = 0, out: torch.Tensor | None = None, fast=True):
    """dct_2d / idct_2d / low-pass projection over the last two (square) dims of `x`.
    N == 32 / 64 run the register-butterfly kernels (csrc/dct32.cu); other sizes the generic M X M^T kernels."""
    x = _contig(x)
    N = x.shape[-1]
    assert x.shape[-2] == N
    if in_mode == 1 and x.dtype != torch.uint8:
        raise TypeError("in_mode 1 needs a uint8 tensor")
    if in_mode != 1 and x.dtype != torch.float32:
        raise TypeError("float32 input expected")
    if N in (32, 64) and fast:
        if out is None:
            out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        planes = x.numel() // (N * N)
        if planes:
            fn = lib.combat_dct32_fast if N == 32 else lib.combat_dct64_fast
            check(fn(_p(x), _p(out), planes, _KIND[kind], keep, in_mode, _s()), "dct%d_fast" % N)
        return out
    return plane_transform(x, transform_matrix(kind, N, x.device, keep), in_mode, out)


# ------------------------------------------------------------------ blend
def gaussian_taps(sigma: float) -> tuple[float, float]:
    """Normalised 3-tap Gaussian (centre, side) evaluated in float32 like torchvision's
    _get_gaussian_kernel1d (host-side parameter preparation for the fused blend kernel)."""
    x = torch.linspace(-1.0, 1.0, steps=3, dtype=torch.float32)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    k = pdf / pdf.sum()
    return float(k[1]), float(k[0])


def poison_blend_fwd(x, noise, perm, num_bd, noise_rate, taps, out=None, sq_partial=None, nperm=None, taps_dev=None,
                     num_bd_dev=None, taps_rows=None):
    """taps = (k0, k1) host floats, or pass taps_dev / num_bd_dev (device tensors) for graph-captured steps;
    taps_rows: device float32 [rows, 2] per-row taps (multilabel step: one sigma per class chunk)."""
    x = _contig(x)
    B, Cc, H, W = x.shape
    rows = B if perm is None else perm.numel()
    if out is None:
        out = torch.empty((rows, Cc, H, W), dtype=torch.float32, device=x.device)
    k0, k1 = taps if taps is not None else (1.0, 0.0)
    check(lib.combat_poison_blend_fwd(_p(x), _p(noise), _p(perm), _p(nperm), rows, num_bd, noise_rate, k0, k1,
                                      _p(out), _p(sq_partial), Cc, H, W, _p(taps_dev), _p(num_bd_dev), _p(taps_rows), _s()),
          "poison_blend_fwd")
    return out


def poison_blend_bwd(x, noise, x_bd, g1, g2, mse_scale, noise_rate, taps, out=None, taps_dev=None, taps_rows=None):
    B, Cc, H, W = x.shape
    if out is None:
        out = torch.empty_like(x)
    k0, k1 = taps if taps is not None else (1.0, 0.0)
    check(lib.combat_poison_blend_bwd(_p(x), _p(noise), _p(x_bd), _p(g1), _p(g2), mse_scale, noise_rate, k0, k1,
                                      _p(out), B, Cc, H, W, _p(taps_dev), _p(taps_rows), _s()), "poison_blend_bwd")
    return out


# ------------------------------------------------------------------ losses / optimiser
def cross_entropy(logits, targets, grad_scale=1.0, want_grad=True, targets2=None, loss_out=None, counts_out=None):
    B, Cn = logits.shape
    dev = logits.device
    if loss_out is None:
        loss_out = torch.empty(1, dtype=torch.float32, device=dev)
    if counts_out is None:
        counts_out = torch.empty(2, dtype=torch.int32, device=dev)
    dl = torch.empty_like(logits) if want_grad else None
    check(lib.combat_cross_entropy(_p(logits), _p(targets), _p(targets2), B, Cn, grad_scale, _p(loss_out), _p(dl),
                                   _p(counts_out), _s()), "cross_entropy")
    return loss_out, dl, counts_out


def sum_scale(partial, scale, out=None):
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=partial.device)
    check(lib.combat_sum_scale(_p(partial), partial.numel(), scale, _p(out), _s()), "sum_scale")
    return out


def tanh_fwd(x):
    y = torch.empty_like(x)
    check(lib.combat_tanh_fwd(_p(x), _p(y), x.numel(), _s()), "tanh_fwd")
    return y


def wanet_warp_fwd(x, z, ident, perm, num_bd, grid_rescale, S, out=None, noise_grid=None, sq_partial=None, gl_partial=None,
                   num_bd_dev=None):
    """train_generator_wanet.py:151-159 / :196-203 (csrc/warp.cu): rows i < num_bd of the output are sample perm[i] warped by
    its own flow z[perm[i]] (the GridGenerator's output), the others plain copies."""
    x = _contig(x)
    B, Cc, H, W = x.shape
    rows = B if perm is None else perm.numel()
    if out is None:
        out = torch.empty((rows, Cc, H, W), dtype=torch.float32, device=x.device)
    check(lib.combat_wanet_warp_fwd(_p(x), _p(z), _p(ident), _p(perm), rows, int(num_bd), _p(num_bd_dev), float(grid_rescale),
                                    _p(out), _p(noise_grid), _p(sq_partial), _p(gl_partial), Cc, H, W, int(S), _s()),
          "wanet_warp_fwd")
    return out


def wanet_warp_bwd(x, z, ident, g1, g2, grid_rescale, l2_scale, S):
    B, Cc, H, W = x.shape
    dz = torch.empty((B, 2 * S * S), dtype=torch.float32, device=x.device)
    check(lib.combat_wanet_warp_bwd(_p(x), _p(z), _p(ident), _p(g1), _p(g2), float(grid_rescale), float(l2_scale), _p(dz), B, Cc,
                                    H, W, int(S), _s()), "wanet_warp_bwd")
    return dz


def grad_l2(x, x_bd, out, partial=None):
    """the logged 'Grad L2 Loss' of train_generator.py:235-243 -> out[0]"""
    B, Cc, H, W = x.shape
    if partial is None:
        partial = torch.empty(2 * B * Cc, dtype=torch.float32, device=x.device)
    check(lib.combat_grad_l2(_p(x), _p(x_bd), _p(partial), _p(out), B, Cc, H, W, _s()), "grad_l2")
    return out


def tv_loss(x, out, grad=None, grad_weight=0.0, partial=None):
    """total-variation loss of train_generator_imperceptible.py:228 -> out[0]; grad += grad_weight * d(sum TV)/dx when given"""
    B, Cc, H, W = x.shape
    if partial is None:
        partial = torch.empty(B * Cc, dtype=torch.float32, device=x.device)
    check(lib.combat_tv_loss(_p(x), _p(grad), float(grad_weight), _p(partial), _p(out), B, Cc, H, W, _s()), "tv_loss")
    return out


def sgd_nesterov(p, g, buf, lr_dev, momentum, wd, first):
    check(lib.combat_sgd_nesterov(_p(p), _p(g), _p(buf), p.numel(), _p(lr_dev), momentum, wd, int(first), _s()), "sgd_nesterov")


# ------------------------------------------------------------------ conv
def nhwc_strides(H, W, Cc):
    return (H * W * Cc, W * Cc, Cc, 1)


def nchw_strides(Cc, H, W):
    return (Cc * H * W, W, 1, H * W)  # (n, h, w, c)


def conv_simt(x, x_geom, x_strides, w, w_dt, out, out_geom, out_strides, *, Ci, Co, KH, KW, stride, pad, up=1,
              bias=None, residual=None, act=0, post_scale=None, post_shift=None):
    """x_geom = (N, Hi, Wi), out_geom = (Ho, Wo)."""
    d = ConvDesc()
    d.in_, d.w, d.out, d.bias, d.residual = _p(x), _p(w), _p(out), _p(bias), _p(residual)
    d.post_scale, d.post_shift = _p(post_scale), _p(post_shift)
    d.N, d.Hi, d.Wi = x_geom
    d.Ci, d.Co, d.KH, d.KW, d.stride, d.pad, d.up = Ci, Co, KH, KW, stride, pad, up
    d.Ho, d.Wo = out_geom
    d.in_sn, d.in_sh, d.in_sw, d.in_sc = x_strides
    d.out_sn, d.out_sh, d.out_sw, d.out_sc = out_strides
    d.in_dtype, d.w_dtype, d.out_dtype, d.act = dt_code(x), w_dt, dt_code(out), act
    check(lib.combat_conv_simt(C.byref(d), _s()), "conv_simt")
    return out


def conv_wgrad_simt(x, x_geom, x_strides, dy, dy_geom, dy_strides, dw, *, Ci, Co, KH, KW, stride, pad, db=None):
    d = ConvDesc()
    d.in_, d.w, d.out, d.bias = _p(x), None, None, _p(db)  # `bias` slot carries the bias-gradient accumulator
    d.N, d.Hi, d.Wi = x_geom
    d.Ci, d.Co, d.KH, d.KW, d.stride, d.pad, d.up = Ci, Co, KH, KW, stride, pad, 1
    d.Ho, d.Wo = dy_geom
    d.in_sn, d.in_sh, d.in_sw, d.in_sc = x_strides
    d.out_sn, d.out_sh, d.out_sw, d.out_sc = dy_strides
    d.in_dtype, d.w_dtype, d.out_dtype, d.act = dt_code(x), 0, dt_code(dy), 0
    check(lib.combat_conv_wgrad_simt(C.byref(d), _p(dy), dt_code(dy), _p(dw), _s()), "conv_wgrad_simt")


def im2col3(x_nchw, stride):
    """3-channel NCHW float32 image -> bf16 [N, Ho, Wo, 64] im2col operand ([hi | lo] halves) of the tensor-core path."""
    N, _, H, W = x_nchw.shape
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    A = torch.empty((N, Ho, Wo, 64), dtype=torch.bfloat16, device=x_nchw.device)
    check(lib.combat_im2col3(_p(x_nchw), _p(A), N, H, W, stride, _s()), "im2col3")
    return A


def fold_w64(dw64, dw, rows, mode, colsum=None, db=None):
    check(lib.combat_fold_w64(_p(dw64), _p(dw), rows, mode, _p(colsum), _p(db), _s()), "fold_w64")


def conv_cin3(x_nchw, w_ptr, w_dt, out, Co, stride, bias=None, act=0, post_scale=None, post_shift=None, out2=None,
              scale2=None, shift2=None):
    N, _, H, W = x_nchw.shape
    check(lib.combat_conv_cin3(_p(x_nchw), w_ptr, w_dt, _p(bias), _p(out), dt_code(out), N, H, W, Co, stride, act,
                               _p(post_scale), _p(post_shift), _p(out2), _p(scale2), _p(shift2), _s()), "conv_cin3")
    return out


def conv_cout3(x_nhwc, w_ptr, w_dt, out_nchw, bias=None, act=0):
    N, H, W, Ci = x_nhwc.shape
    check(lib.combat_conv_cout3(_p(x_nhwc), dt_code(x_nhwc), w_ptr, w_dt, _p(bias), _p(out_nchw), N, H, W, Ci, act, _s()),
          "conv_cout3")
    return out_nchw


def conv_tc_cout3_ok(x_nhwc):
    N, H, W, Ci = x_nhwc.shape
    return Ci == 64 and x_nhwc.dtype == torch.bfloat16 and bool(lib.combat_conv_tc_cout3_supported(N, H, W))


def conv_tc_cout3(x_nhwc, w_ptr, out_nchw, bias=None, act=0):
    """64 -> 3 conv on the tensor pipe (bf16 NHWC in, bf16 [3][9][64] filter, float32 NCHW out)."""
    N, H, W, _ = x_nhwc.shape
    check(lib.combat_conv_tc_cout3(_p(x_nhwc), w_ptr, _p(bias), _p(out_nchw), N, H, W, act, _s()), "conv_tc_cout3")
    return out_nchw


def wgrad_cin3(x_nchw, dy, dw, db, Co, stride):
    N, _, H, W = x_nchw.shape
    check(lib.combat_wgrad_cin3(_p(x_nchw), _p(dy), dt_code(dy), _p(dw), _p(db), N, H, W, Co, stride, _s()), "wgrad_cin3")


def wgrad_cout3(a_nhwc, dz_nchw, dw, db):
    N, H, W, Ci = a_nhwc.shape
    check(lib.combat_wgrad_cout3(_p(a_nhwc), dt_code(a_nhwc), _p(dz_nchw), _p(dw), _p(db), N, H, W, Ci, _s()), "wgrad_cout3")


def conv_tc_desc(x, w_ptr, out, N, Hi, Wi, Ci, Ho, Wo, Co, KH, KW, stride, pad, up, bias=None, residual=None, stats=None,
                 act=0, post_scale=None, post_shift=None, out2=None, scale2=None, shift2=None, mask=None, mask_scale=None,
                 post_add=None, in2=None, w2=None, bnb=None, in_nchw3=False):
    """bnb = (x, scale, shift, mean, invstd): fused reduction of a train-mode relu(bn(x)) backward (needs `stats`)."""
    d = ConvTcDesc()
    d.in2, d.w2 = _p(in2), w2
    d.in_nchw3 = int(in_nchw3)
    if bnb is not None:
        d.bnb_x, d.bnb_scale, d.bnb_shift, d.bnb_mean, d.bnb_invstd = (_p(t) for t in bnb)
    d.act, d.post_scale, d.post_shift = act, _p(post_scale), _p(post_shift)
    d.out2, d.scale2, d.shift2 = _p(out2), _p(scale2), _p(shift2)
    d.mask, d.mask_scale, d.post_add = _p(mask), _p(mask_scale), _p(post_add)
    d.in_, d.w, d.out, d.bias, d.residual, d.stats = _p(x), w_ptr, _p(out), _p(bias), _p(residual), _p(stats)
    d.out_f32 = int(out is not None and out.dtype == torch.float32)
    d.res_f32 = int(residual is not None and residual.dtype == torch.float32)
    d.N, d.Hi, d.Wi, d.Ci, d.Ho, d.Wo, d.Co = N, Hi, Wi, Ci, Ho, Wo, Co
    d.KH, d.KW, d.stride, d.pad, d.up = KH, KW, stride, pad, up
    return d


# ------------------------------------------------------------------ norm / activation
import os as _os

# train-mode BatchNorm backward as ONE cooperative launch (reduce -> grid sync -> finalize -> grid sync -> apply) instead of three
# launches.  Measured on B200 inside the step (batch 512, 16 BatchNorm layers): 12.36 ms with it, 11.67 ms without -- two
# grid-wide barriers over ~1200 resident CTAs cost more than the two launches and the DRAM re-read they remove.  Opt-in only.
FUSED_BN_BWD = bool(_os.environ.get("COMBAT_FUSED_BN_BWD"))
_PARTIAL_BLOCKS = 256
_PARTIAL_FLOATS = _PARTIAL_BLOCKS * 2 * 2048


def _max_partial_blocks(Cc):
    """Row blocks of a column reduction: enough CTAs to fill the GPU (8 per SM) as far as the scratch buffer allows --
    256 blocks left a 64-channel reduction at 1.7 CTAs per SM."""
    return max(1, min(148 * 8, _PARTIAL_FLOATS // (2 * Cc)))


class Scratch:
    """Per-device scratch for the column reductions (partials) -- allocated once."""

    _inst: dict = {}

    @classmethod
    def get(cls, device):
        k = str(device)
        if k not in cls._inst:
            cls._inst[k] = torch.empty(_PARTIAL_FLOATS, dtype=torch.float32, device=device)
        return cls._inst[k]


def bn_train_prepare(x2d, R, Cc, gamma, beta, rm, rv, momentum, eps):
    """Batch statistics -> (scale, shift, mean, invstd); updates running stats in place."""
    dev = x2d.device
    partial = Scratch.get(dev)
    nblk = C.c_int(0)
    check(lib.combat_bn_stats(_p(x2d), dt_code(x2d), R, Cc, _p(partial), _max_partial_blocks(Cc), C.byref(nblk), _s()), "bn_stats")
    st = torch.empty((4, Cc), dtype=torch.float32, device=dev)
    check(lib.combat_bn_finalize(_p(partial), nblk.value, R, Cc, _p(gamma), _p(beta), _p(rm), _p(rv), momentum, eps,
                                 _p(st[0]), _p(st[1]), _p(st[2]), _p(st[3]), _s()), "bn_finalize")
    return st[0], st[1], st[2], st[3]


def bn_train_finalize(nblk, R, Cc, gamma, beta, rm, rv, momentum, eps):
    """Same as bn_train_prepare when the producer conv already left its per-CTA partial sums in the scratch buffer
    (conv_tc_desc(stats=Scratch.get(dev)); nblk = lib.combat_conv_tc_last_grid())."""
    dev = rm.device
    partial = Scratch.get(dev)
    st = torch.empty((4, Cc), dtype=torch.float32, device=dev)
    check(lib.combat_bn_finalize(_p(partial), nblk, R, Cc, _p(gamma), _p(beta), _p(rm), _p(rv), momentum, eps,
                                 _p(st[0]), _p(st[1]), _p(st[2]), _p(st[3]), _s()), "bn_finalize")
    return st[0], st[1], st[2], st[3]


def bn_eval_prepare(Cc, gamma, beta, rm, rv, eps):
    st = torch.empty((2, Cc), dtype=torch.float32, device=rm.device)
    check(lib.combat_bn_finalize(None, 0, 1, Cc, _p(gamma), _p(beta), _p(rm), _p(rv), 0.0, eps, _p(st[0]), _p(st[1]),
                                 None, None, _s()), "bn_finalize(eval)")
    return st[0], st[1]


def bn_eval_affine(params, bufs, table4, n, eps, out):
    """out: float32 [2, n] (scale | shift) for every BatchNorm channel of a network, one launch."""
    check(lib.combat_bn_eval_affine(_p(params), _p(bufs), _p(table4), n, eps, _p(out[0]), _p(out[1]), _s()), "bn_eval_affine")
    return out


def affine_act(x, scale, shift, relu, residual=None, out=None, out_dtype=None):
    """x may be float32 (pre-normalisation tensor) while residual / output use the activation dtype."""
    Cc = x.shape[-1]
    R = x.numel() // Cc
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=x.device)
    check(lib.combat_affine_act(_p(x), dt_code(x), _p(residual), _p(out), dt_code(out), R, Cc, _p(scale), _p(shift),
                                int(relu), _s()), "affine_act")
    return out


def bn_bwd_train(dy, x, y, gamma, mean, invstd, relu, dgamma_out, dbeta_out, dadd=None, want_dres=False):
    """Returns (dx, dres).  dgamma/dbeta are WRITTEN into the given float32 views."""
    Cc = x.shape[-1]
    R = x.numel() // Cc
    partial = Scratch.get(x.device)
    if FUSED_BN_BWD and Cc % 8 == 0:   # one cooperative launch instead of reduce / finalize / apply
        dx = torch.empty_like(dy)
        dres = torch.empty_like(dy) if want_dres else None
        check(lib.combat_bn_bwd_fused(_p(dy), _p(x), dt_code(x), _p(y), _p(dadd), _p(dx), _p(dres), dt_code(dy), R, Cc, _p(gamma),
                                      _p(mean), _p(invstd), _p(partial), _max_partial_blocks(Cc), _p(dgamma_out), _p(dbeta_out),
                                      int(relu), _s()), "bn_bwd_fused")
        return dx, dres
    nblk = C.c_int(0)
    check(lib.combat_bn_bwd_reduce(_p(dy), _p(x), dt_code(x), _p(y), dt_code(dy), R, Cc, _p(mean), _p(invstd), _p(partial),
                                   _max_partial_blocks(Cc), C.byref(nblk), int(relu), _s()), "bn_bwd_reduce")
    check(lib.combat_bn_bwd_finalize(_p(partial), nblk.value, Cc, _p(dgamma_out), _p(dbeta_out), _s()), "bn_bwd_finalize")
    dx = torch.empty_like(dy)
    dres = torch.empty_like(dy) if want_dres else None
    check(lib.combat_bn_bwd_apply(_p(dy), _p(x), dt_code(x), _p(y), _p(dadd), _p(dx), _p(dres), dt_code(dy), R, Cc, _p(gamma),
                                  _p(mean), _p(invstd), _p(dgamma_out), _p(dbeta_out), None, int(relu), _s()), "bn_bwd_apply")
    return dx, dres


def bn_bwd_train_from_partials(g, x, gamma, mean, invstd, nblk, dgamma_out, dbeta_out, dadd=None):
    """Second half of bn_bwd_train when the producing input-gradient conv already masked the gradient (g) and left the
    per-CTA partial sums of (g, g * xhat) in Scratch (conv_tc_desc(bnb=...)): finalize + apply, no reduction pass."""
    Cc = x.shape[-1]
    R = x.numel() // Cc
    partial = Scratch.get(x.device)
    check(lib.combat_bn_bwd_finalize(_p(partial), nblk, Cc, _p(dgamma_out), _p(dbeta_out), _s()), "bn_bwd_finalize")
    dx = torch.empty_like(g)
    check(lib.combat_bn_bwd_apply(_p(g), _p(x), dt_code(x), None, _p(dadd), _p(dx), None, dt_code(g), R, Cc, _p(gamma),
                                  _p(mean), _p(invstd), _p(dgamma_out), _p(dbeta_out), None, 0, _s()), "bn_bwd_apply")
    return dx


def bn_bwd_eval(dy, y, eval_scale, relu, dadd=None, want_dres=False):
    Cc = dy.shape[-1]
    R = dy.numel() // Cc
    dx = torch.empty_like(dy)
    dres = torch.empty_like(dy) if want_dres else None
    check(lib.combat_bn_bwd_apply(_p(dy), None, dt_code(dy), _p(y), _p(dadd), _p(dx), _p(dres), dt_code(dy), R, Cc, None, None,
                                  None, None, None, _p(eval_scale), int(relu), _s()), "bn_bwd_apply(eval)")
    return dx, dres


def instnorm_fwd(x, act, skip=None, eps=1e-5, slope=0.2, out_dtype=None):
    N, H, W, Cc = x.shape
    y = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=x.device)
    st = torch.empty((2, N, Cc), dtype=torch.float32, device=x.device)
    K = lib.combat_instnorm_splits(N, H * W, Cc)
    if K > 1:   # large planes, few (sample, channel-group) pairs: every plane split over K CTAs (csrc/norm.cu)
        part = torch.empty((N, K, 2, Cc), dtype=torch.float32, device=x.device)
        check(lib.combat_instnorm_fwd_split(_p(x), dt_code(x), _p(skip), _p(y), dt_code(y), N, H * W, Cc, eps, slope, int(act),
                                            _p(st[0]), _p(st[1]), _p(part), K, _s()), "instnorm_fwd_split")
        return y, st
    check(lib.combat_instnorm_fwd(_p(x), dt_code(x), _p(skip), _p(y), dt_code(y), N, H * W, Cc, eps, slope, int(act),
                                  _p(st[0]), _p(st[1]), _s()), "instnorm_fwd")
    return y, st


def instnorm_bwd(dy1, dy2, x, st, act, slope=0.2):
    N, H, W, Cc = x.shape
    dx = torch.empty_like(dy1)
    K = lib.combat_instnorm_splits(N, H * W, Cc)
    if K > 1:
        part = torch.empty((N, K, 2, Cc), dtype=torch.float32, device=x.device)
        check(lib.combat_instnorm_bwd_split(_p(dy1), _p(dy2), _p(x), dt_code(x), _p(dx), dt_code(dy1), N, H * W, Cc, slope, int(act),
                                            _p(st[0]), _p(st[1]), _p(part), K, _s()), "instnorm_bwd_split")
        return dx
    check(lib.combat_instnorm_bwd(_p(dy1), _p(dy2), _p(x), dt_code(x), _p(dx), dt_code(dy1), N, H * W, Cc, slope, int(act),
                                  _p(st[0]), _p(st[1]), _s()), "instnorm_bwd")
    return dx


def upsample2x_act(x, slope=0.2):
    N, H, W, Cc = x.shape
    y = torch.empty((N, 2 * H, 2 * W, Cc), dtype=x.dtype, device=x.device)
    check(lib.combat_upsample2x_act(_p(x), _p(y), dt_code(x), N, H, W, Cc, slope, _s()), "upsample2x_act")
    return y


def upsample2x_act_bwd(dy, y, slope=0.2):
    N, H2, W2, Cc = dy.shape
    dx = torch.empty((N, H2 // 2, W2 // 2, Cc), dtype=dy.dtype, device=dy.device)
    check(lib.combat_upsample2x_act_bwd(_p(dy), _p(y), _p(dx), dt_code(dy), N, H2 // 2, W2 // 2, Cc, slope, _s()),
          "upsample2x_act_bwd")
    return dx


def leaky_relu(x, slope=0.2, out=None):
    if out is None:
        out = torch.empty_like(x)
    check(lib.combat_leaky_relu(_p(x), _p(out), dt_code(x), x.numel(), slope, _s()), "leaky_relu")
    return out


def leaky_relu_bwd(dy, x, slope=0.2):
    dx = torch.empty_like(dy)
    check(lib.combat_leaky_relu_bwd(_p(dy), _p(x), _p(dx), dt_code(dy), dy.numel(), slope, _s()), "leaky_relu_bwd")
    return dx


def tanh_bwd(dy, y):
    dz = torch.empty_like(dy)
    check(lib.combat_tanh_bwd(_p(dy), _p(y), _p(dz), dy.numel(), _s()), "tanh_bwd")
    return dz


def colsum(x2d, Cc, out):
    R = x2d.numel() // Cc
    check(lib.combat_colsum(_p(x2d), dt_code(x2d), R, Cc, _p(out), _s()), "colsum")


def pool_linear_fwd(x, P, W, b):
    B, Hf, Wf, Cc = x.shape
    ncls = W.shape[0]
    Fdim = Cc * (Hf // P) * (Wf // P)
    pooled = torch.empty((B, Fdim), dtype=torch.float32, device=x.device)
    logits = torch.empty((B, ncls), dtype=torch.float32, device=x.device)
    check(lib.combat_pool_linear_fwd(_p(x), dt_code(x), B, Hf, Wf, Cc, P, _p(W), _p(b), ncls, _p(pooled), _p(logits), _s()),
          "pool_linear_fwd")
    return logits, pooled


def pool_linear_bwd(dlogits, pooled, W, x_shape, dtype, P, dW=None, db=None, want_dx=True):
    B, Hf, Wf, Cc = x_shape
    ncls = W.shape[0]
    dx = torch.empty(x_shape, dtype=dtype, device=dlogits.device) if want_dx else None
    check(lib.combat_pool_linear_bwd(_p(dlogits), _p(pooled), _p(W), B, Hf, Wf, Cc, P, ncls, _p(dx), dt_code(dtype), _p(dW),
                                     _p(db), _s()), "pool_linear_bwd")
    return dx


def maxpool2(x):
    N, H, W, Cc = x.shape
    y = torch.empty((N, H // 2, W // 2, Cc), dtype=x.dtype, device=x.device)
    check(lib.combat_maxpool2(_p(x), _p(y), dt_code(x), N, H, W, Cc, _s()), "maxpool2")
    return y


def maxpool2_bwd(dy, x):
    """dy: NHWC gradient of maxpool2(x); x: the pooled tensor's input."""
    N, H, W, Cc = x.shape
    dx = torch.empty_like(x)
    check(lib.combat_maxpool2_bwd(_p(dy), _p(x), _p(dx), dt_code(x), N, H, W, Cc, _s()), "maxpool2_bwd")
    return dx


def elu_bwd(da, a):
    """da: gradient w.r.t. a = elu(z) (the kept output) -> gradient w.r.t. z."""
    dz = torch.empty_like(da)
    check(lib.combat_elu_bwd(_p(da), _p(a), _p(dz), dt_code(da), da.numel(), _s()), "elu_bwd")
    return dz


def mask_scale(x, keep, scale):
    """dropout with a given uint8 keep mask (same element order as x): keep ? x * scale : 0."""
    if keep.dtype != torch.uint8 or keep.numel() != x.numel():
        raise ValueError("keep mask: uint8, one byte per element of x")
    y = torch.empty_like(x)
    check(lib.combat_mask_scale(_p(x), _p(keep), _p(y), dt_code(x), x.numel(), float(scale), _s()), "mask_scale")
    return y


def adadelta(p, g, square_avg, acc_delta, lr_dev, rho=0.9, eps=1e-6, wd=1e-4):
    check(lib.combat_adadelta(_p(p), _p(g), _p(square_avg), _p(acc_delta), p.numel(), _p(lr_dev), rho, eps, wd, _s()), "adadelta")


def post_transform_fwd(x, params, out=None):
    """PostTensorTransform of an NCHW float32 batch with per-row parameters (device float32 [rows, 8], see csrc/augment.cu)."""
    x = _contig(x)
    rows, Cc, H, W = x.shape
    if params.dtype != torch.float32 or params.numel() != rows * 8 or not params.is_contiguous():
        raise ValueError("post_transform: params must be a contiguous float32 [rows, 8] tensor")
    if out is None:
        out = torch.empty_like(x)
    check(lib.combat_post_transform_fwd(_p(x), _p(out), _p(params), rows, Cc, H, W, _s()), "post_transform_fwd")
    return out


def post_transform_bwd(dout, params, out=None, accumulate=False):
    """adjoint of post_transform_fwd: gradient w.r.t. the untransformed batch (accumulate=True adds into `out`)."""
    dout = _contig(dout)
    rows, Cc, H, W = dout.shape
    if params.dtype != torch.float32 or params.numel() != rows * 8 or not params.is_contiguous():
        raise ValueError("post_transform: params must be a contiguous float32 [rows, 8] tensor")
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs an output buffer")
        out = torch.empty_like(dout)
    check(lib.combat_post_transform_bwd(_p(dout), _p(out), _p(params), rows, Cc, H, W, int(accumulate), _s()), "post_transform_bwd")
    return out


def nchw_to_nhwc(x, dtype):
    N, Cc, H, W = x.shape
    y = torch.empty((N, H, W, Cc), dtype=dtype, device=x.device)
    check(lib.combat_nchw_to_nhwc(_p(x), _p(y), dt_code(dtype), N, Cc, H, W, _s()), "nchw_to_nhwc")
    return y


def nhwc_to_nchw(x):
    N, H, W, Cc = x.shape
    y = torch.empty((N, Cc, H, W), dtype=torch.float32, device=x.device)
    check(lib.combat_nhwc_to_nchw(_p(x), dt_code(x), _p(y), N, Cc, H, W, _s()), "nhwc_to_nchw")
    return y


def onehot_planes(y, labels, c_off, ncls):
    N, H, W, Ctot = y.shape
    check(lib.combat_onehot_planes(_p(y), dt_code(y), _p(labels), N, H * W, Ctot, c_off, ncls, _s()), "onehot_planes")


def lrelu_into_slice(src, dst, c_off, slope=0.2):
    Cs, Cd = src.shape[-1], dst.shape[-1]
    check(lib.combat_lrelu_into_slice(_p(src), _p(dst), dt_code(src), src.numel() // Cs, Cs, Cd, c_off, slope, _s()),
          "lrelu_into_slice")


def make_wprep_table(descs, device):
    """descs: list of (src_off, fwd_off, dgrad_off, Cout, Cin, KH, KW) -> device byte tensor + max elems."""
    arr = (WPrepDesc * len(descs))()
    mx = 0
    for i, (s, f, g, co, ci, kh, kw) in enumerate(descs):
        arr[i].src_off, arr[i].fwd_off, arr[i].dgrad_off = s, f, g
        arr[i].Cout, arr[i].Cin, arr[i].KH, arr[i].KW = co, ci, kh, kw
        mx = max(mx, co * ci * kh * kw)
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    return torch.from_numpy(raw).to(device), mx


def prep_weights(params_flat, wbuf, table, n_desc, max_elems):
    check(lib.combat_prep_weights(_p(params_flat), _p(wbuf), dt_code(wbuf), _p(table), n_desc, max_elems, _s()), "prep_weights")
