"""CPU check of the CONTRACT of csrc/warp.cu (no GPU here): a numpy model of exactly what the kernels compute -- separable
bicubic table (cubic convolution, A = -0.75, align_corners), blend with the identity grid, clamp, bilinear gather with zero
padding, and the hand-derived backward (bilinear derivative, clamp mask, table transposed) -- against torch's own
F.interpolate / F.grid_sample / autograd, which is what the reference calls (train_generator_wanet.py:151-158).  The GPU test
(tests/test_wanet_gpu.py) then holds the CUDA kernels to the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import combat_oracle as O


def cubic(t):
    A = -0.75
    x0, x3, u = t + 1, 2 - t, 1 - t
    return [((A * x0 - 5 * A) * x0 + 8 * A) * x0 - 4 * A, ((A + 2) * t - (A + 3)) * t * t + 1,
            ((A + 2) * u - (A + 3)) * u * u + 1, ((A * x3 - 5 * A) * x3 + 8 * A) * x3 - 4 * A]


def table(S, H):
    wt = np.zeros((S, H))
    scale = (S - 1) / (H - 1)
    for o in range(H):
        real = np.float32(scale) * np.float32(o)
        fl = np.floor(real)
        c = cubic(float(real - fl))
        for i in range(4):
            wt[min(max(int(fl) - 1 + i, 0), S - 1), o] += c[i]
    return wt


def model_fwd_bwd(x, z, ident, r, l2_scale, g):
    """returns (out, noise_grid, dz) as the kernels define them."""
    N, C, H, W = x.shape
    S = z.shape[-1]
    flow = z
    wt = table(S, H)
    noise = np.einsum("ph,qw,ncpq->nchw", wt, wt, flow)          # [N, 2, H, W]
    rx = ident[None, None, :] * (1 - r) + noise[:, 0] * r           # [N, H, W], x coordinate: ident over w
    ry = ident[None, :, None] * (1 - r) + noise[:, 1] * r
    gx, gy = np.clip(rx, -1, 1), np.clip(ry, -1, 1)
    ix, iy = (gx + 1) * 0.5 * (W - 1), (gy + 1) * 0.5 * (H - 1)
    x0, y0 = np.floor(ix).astype(int), np.floor(iy).astype(int)
    wx, wy = ix - x0, iy - y0

    def tap(n, c, yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        return np.where(ok, x[n, c][np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0.0)

    out = np.zeros_like(x)
    dix, diy = np.zeros((N, H, W)), np.zeros((N, H, W))
    for n in range(N):
        for c in range(C):
            v00, v01 = tap(n, c, y0[n], x0[n]), tap(n, c, y0[n], x0[n] + 1)
            v10, v11 = tap(n, c, y0[n] + 1, x0[n]), tap(n, c, y0[n] + 1, x0[n] + 1)
            out[n, c] = v00 * (1 - wx[n]) * (1 - wy[n]) + v01 * wx[n] * (1 - wy[n]) + v10 * (1 - wx[n]) * wy[n] + v11 * wx[n] * wy[n]
            dix[n] += g[n, c] * ((v01 - v00) * (1 - wy[n]) + (v11 - v10) * wy[n])
            diy[n] += g[n, c] * ((v10 - v00) * (1 - wx[n]) + (v11 - v01) * wx[n])
    dnx = np.where((rx >= -1) & (rx <= 1), dix * 0.5 * (W - 1) * r, 0.0) + l2_scale * noise[:, 0]
    dny = np.where((ry >= -1) & (ry <= 1), diy * 0.5 * (H - 1) * r, 0.0) + l2_scale * noise[:, 1]
    dflow = np.einsum("ph,qw,nchw->ncpq", wt, wt, np.stack([dnx, dny], 1))
    return out, noise.transpose(0, 2, 3, 1), dflow


@pytest.mark.parametrize("S,H,r", [(2, 32, 0.15), (4, 16, 0.9), (3, 20, 2.5)])
def test_kernel_contract_matches_torch(S, H, r):
    """r = 0.9 / 2.5 drive part of the grid into the clamp (mask path) and the taps onto the image border (zero padding)."""
    torch.manual_seed(S * 100 + H)
    N, C = 3, 3
    x = torch.rand(N, C, H, H, dtype=torch.float64) * 2 - 1
    z = torch.tanh(torch.randn(N, 2, S, S, dtype=torch.float64) * 1.5).requires_grad_(True)   # the generator's tanh output
    g = torch.randn(N, C, H, H, dtype=torch.float64)
    opt = O.default_opt(input_height=H, grid_rescale=r, s=S)
    ident = torch.linspace(-1, 1, steps=H, dtype=torch.float64)
    flow = z
    noise_grid = F.interpolate(flow, size=H, mode="bicubic", align_corners=True).permute((0, 2, 3, 1))
    ig = O.identity_grid(H).double()
    ref = F.grid_sample(x, torch.clamp(ig * (1 - r) + noise_grid * r, -1, 1), align_corners=True)
    l2_scale = 0.37
    ((ref * g).sum() + 0.5 * l2_scale * (noise_grid ** 2).sum()).backward()
    out, ng, dz = model_fwd_bwd(x.numpy(), z.detach().numpy(), ident.numpy(), r, l2_scale, g.numpy())
    assert np.abs(ng - noise_grid.detach().numpy()).max() < 1e-6
    assert np.abs(out - ref.detach().numpy()).max() < 1e-5
    assert np.abs(dz - z.grad.numpy()).max() < 1e-5 * max(1.0, np.abs(z.grad.numpy()).max())


def test_logged_gradient_term_formula():
    """gl_partial of the forward kernel: per-row closed form of train_generator_wanet.py:213-222."""
    torch.manual_seed(3)
    N, H = 4, 8
    ng = torch.randn(N, H, H, 2, dtype=torch.float64)
    ref = float(O.wanet_grad_l2(ng))
    v = ng.numpy()
    tot = 0.0
    for n in range(N):
        s1 = (v[n, :, 0] ** 2).sum() + ((v[n, :, 1:] - v[n, :, :-1]) ** 2).sum() + (v[n, :, -1] ** 2).sum()
        s2 = (v[n, ..., 0] ** 2 + (v[n, ..., 1] - v[n, ..., 0]) ** 2 + v[n, ..., 1] ** 2).sum()
        tot += s1 / (H * (H + 2) * 4) + s2 / (H * (H + 3) * 3)
    assert abs(tot / N - ref) < 1e-12 * max(1.0, abs(ref))
