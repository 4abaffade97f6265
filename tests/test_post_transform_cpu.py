"""PostTensorTransform (utils/dataloader.py:11-22,45-60): the host-side parameter draws of the product and the oracle's
restatement of the kornia pipeline agree decision for decision, and the restatement behaves like crop / rotate / flip."""
import random
from types import SimpleNamespace

import numpy as np
import torch

from oracle import combat_oracle as O


def _opt(option="use", dataset="cifar10"):
    return SimpleNamespace(post_transform_option=option, random_crop=5, random_rotation=10, dataset=dataset)


def _seed(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def test_product_draws_equal_oracle_draws():
    from combat_b200.utils.dataloader import draw_params
    for option, dataset in (("use", "cifar10"), ("use", "celeba"), ("use_modified", "cifar10"), ("no_use", "cifar10")):
        opt = _opt(option, dataset)
        for seed in range(12):
            _seed(seed)
            ref = [O.draw_post_transform(9, opt) for _ in range(3)]
            st_ref = (random.random(), float(torch.rand(1)))
            _seed(seed)
            got = [draw_params(9, opt) for _ in range(3)]
            st_got = (random.random(), float(torch.rand(1)))
            assert st_ref == st_got, "RNG streams must be left at the same position"
            for r, g in zip(ref, got):
                assert np.array_equal(g[:, 0], (r["xs"] - r["pad"]).numpy().astype(np.float32))
                assert np.array_equal(g[:, 1], (r["ys"] - r["pad"]).numpy().astype(np.float32))
                rad = r["angle"].float() * (np.pi / 180.0)
                if r["rot"]:
                    assert np.allclose(g[:, 2], torch.cos(rad).numpy(), atol=1e-7) and np.allclose(g[:, 3], torch.sin(rad).numpy(), atol=1e-7)
                    assert (g[:, 4] == 1).all()
                else:
                    assert (g[:, 2] == 1).all() and (g[:, 3] == 0).all() and (g[:, 4] == 0).all()
                assert np.array_equal(g[:, 5] != 0, r["flip"].numpy())


def test_restatement_is_crop_rotate_flip():
    x = torch.arange(2 * 3 * 8 * 8, dtype=torch.float32).view(2, 3, 8, 8)
    base = dict(xs=torch.tensor([5, 5]), ys=torch.tensor([5, 5]), crop=False, angle=torch.zeros(2), rot=False,
                flip=torch.zeros(2, dtype=torch.bool), pad=5)
    assert torch.equal(O.apply_post_transform(x, base), x)
    # crop: window start (xs, ys) in the zero-padded image -> out[y, x] = in[y + ys - 5, x + xs - 5]
    p = dict(base, crop=True, xs=torch.tensor([7, 3]), ys=torch.tensor([5, 10]))
    y = O.apply_post_transform(x, p)
    assert torch.equal(y[0, :, :, :6], x[0, :, :, 2:]) and (y[0, :, :, 6:] == 0).all()
    assert torch.equal(y[1, :, :3, 2:], x[1, :, 5:, :6]) and (y[1, :, 3:] == 0).all() and (y[1, :, :, :2] == 0).all()
    # flip only the flagged rows
    p = dict(base, flip=torch.tensor([True, False]))
    y = O.apply_post_transform(x, p)
    assert torch.equal(y[0], x[0].flip(-1)) and torch.equal(y[1], x[1])
    # rotation by 0 is the identity; by 90 degrees it is a quarter turn about the centre (square image, exact grid)
    p = dict(base, rot=True, angle=torch.tensor([0.0, 90.0]))
    y = O.apply_post_transform(x, p)
    assert torch.allclose(y[0], x[0], atol=1e-4)
    q = torch.rot90(x[1], 1, dims=(-2, -1))
    assert torch.allclose(y[1], q, atol=2e-3) or torch.allclose(y[1], torch.rot90(x[1], -1, dims=(-2, -1)), atol=2e-3)


def test_no_use_draws_nothing():
    opt = _opt("no_use")
    _seed(3)
    a = (random.random(), float(torch.rand(1)))
    _seed(3)
    O.post_transform(torch.zeros(4, 3, 8, 8), opt)
    from combat_b200.utils.dataloader import draw_params
    draw_params(4, opt)
    assert (random.random(), float(torch.rand(1))) == a


def kernel_model(x, P):
    """numpy model of csrc/augment.cu's gather (pixel-space inverse rotation, zero fill) -- the contract the CUDA kernel is
    held to on the GPU; here it is held to the kornia restatement, so the formula is checked without a GPU."""
    import math
    rows, C, H, W = x.shape
    out = np.zeros_like(x)
    cx, cy = np.float32(0.5 * (W - 1)), np.float32(0.5 * (H - 1))
    for n in range(rows):
        shx, shy, ca, sa, rot, flip = [P[n, i] for i in range(6)]
        shx, shy = int(shx), int(shy)

        def src(i, j):
            if i < 0 or i >= W or j < 0 or j >= H:
                return None
            px, py = i + shx, j + shy
            if px < 0 or px >= W or py < 0 or py >= H:
                return None
            return py, px

        for y in range(H):
            for xx in range(W):
                xf = W - 1 - xx if flip else xx
                if not rot:
                    s = src(xf, y)
                    if s:
                        out[n, :, y, xx] = x[n, :, s[0], s[1]]
                    continue
                u, v = np.float32(xf) - cx, np.float32(y) - cy
                fx, fy = np.float32(ca * u - sa * v + cx), np.float32(sa * u + ca * v + cy)
                x0, y0 = math.floor(fx), math.floor(fy)
                wx1, wy1 = fx - x0, fy - y0
                for i, j, w in ((x0, y0, (1 - wx1) * (1 - wy1)), (x0 + 1, y0, wx1 * (1 - wy1)), (x0, y0 + 1, (1 - wx1) * wy1),
                                (x0 + 1, y0 + 1, wx1 * wy1)):
                    s = src(i, j)
                    if s:
                        out[n, :, y, xx] += np.float32(w) * x[n, :, s[0], s[1]]
    return out


def test_kernel_contract_matches_kornia_restatement():
    from combat_b200.utils.dataloader import draw_params
    opt = _opt("use", "cifar10")
    seen = set()
    for seed in range(10):
        _seed(seed)
        prm = O.draw_post_transform(4, opt)
        _seed(seed)
        P = draw_params(4, opt)
        x = torch.rand(4, 3, 12, 12) * 2 - 1
        ref = O.apply_post_transform(x, prm).numpy()
        assert np.abs(ref - kernel_model(x.numpy(), P)).max() < 2e-5
        seen.add((prm["crop"], prm["rot"]))
    assert (True, True) in seen and (False, False) in seen
