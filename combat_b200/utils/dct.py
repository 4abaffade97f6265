"""utils/dct.py of the reference -- same function names and semantics (orthonormal DCT-II / DCT-III over the last
one or two dims, float or uint8 input, float32 output), evaluated by the CUDA kernels of csrc/dct.cu / dct32.cu.

Deviation kept on purpose (trap 4): for uint8 input the reference's `-arange(N, dtype=uint8)`
wraps, which is harmless for N = 32 / 64 and WRONG for N = 224; this implementation always evaluates the correct DCT."""
import torch

from .. import ops


def _check(x, norm):
    if norm != "ortho":
        raise NotImplementedError("only norm='ortho' is used by the reference's call sites")
    if not x.is_cuda:
        raise RuntimeError("combat_b200.utils.dct runs on CUDA tensors only (no CPU fallback)")


def _mode(x):
    if x.dtype == torch.uint8:
        return x.contiguous(), 1
    return x.contiguous().float(), 0


def dct_2d(x, norm="ortho"):
    """utils/dct.py:85-96"""
    _check(x, norm)
    x, m = _mode(x)
    return ops.plane_op(x, "dct", in_mode=m)


def idct_2d(X, norm="ortho"):
    """utils/dct.py:99-111"""
    _check(X, norm)
    X, m = _mode(X)
    return ops.plane_op(X, "idct", in_mode=m)


def _rows(x, kind):
    """1-D transform along the last dim: rows are packed into N x N planes and transformed as I * X * M^T."""
    x, m = _mode(x)
    shape = x.shape
    N = shape[-1]
    flat = x.reshape(-1, N)
    rows = flat.shape[0]
    pad = (-rows) % N
    if pad:
        flat = torch.cat([flat, flat.new_zeros(pad, N)], 0)
    planes = flat.reshape(-1, N, N).contiguous()
    M = ops.transform_matrix(kind, N, x.device)
    eye = ops.transform_matrix("eye", N, x.device)
    out = ops.plane_transform_lr(planes, eye, M, m)
    return out.reshape(-1, N)[:rows].reshape(shape)


def dct(x, norm="ortho"):
    """utils/dct.py:13-42"""
    _check(x, norm)
    return _rows(x, "dct")


def idct(X, norm="ortho"):
    """utils/dct.py:45-82"""
    _check(X, norm)
    return _rows(X, "idct")
