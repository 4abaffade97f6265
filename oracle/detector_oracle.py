"""CPU oracle for the frequency-detector trainer (defenses/frequency_based/train.py of the reference).

TEST INFRASTRUCTURE ONLY (same rule as oracle/combat_oracle.py): nothing under ``combat_b200/`` may import this module.
It is the checker for SURVEY.md section 8(f) row 3 -- the next row of the hot-path scope table: the batched uint8 DCT of
the (clean, patched) batch and one FrequencyModel training iteration (train-mode conv -> ELU -> BatchNorm with
max-pool / dropout, cross entropy, Adadelta).  No CUDA path for this row exists yet; the oracle and its fixture come first.

What is restated: the reference's control flow, RNG consumption order (numpy global RNG for the synthetic triggers,
Python ``random`` for the batch shuffle, the torch CPU generator for dropout), layer wiring and the Adadelta update.
The arithmetic primitives are the torch / scipy CPU ops the reference itself dispatches.  Every function cites the
reference lines it follows (paths relative to defenses/frequency_based/).

Parity pin: tests/golden/detector_b8x2.npz, recorded from the UNMODIFIED reference ``train()`` by
tests/golden/make_golden_detector.py; tests/test_detector_oracle_cpu.py holds this module to it.
Not pinnable here: ``albumentations`` (GaussNoise / RandomShadow, train.py:49-62) is absent from the container.  The
fixture run and this oracle both use the stand-ins below for those two augmentations; they keep the reference's call
structure (uint8 in, uint8 out, /255) but are NOT restatements of albumentations.
"""
from __future__ import annotations

import random

import numpy as np
import scipy.fft
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# stand-ins for the absent albumentations transforms (see header)
# --------------------------------------------------------------------------
class StandInGaussNoise:
    """albumentations.GaussNoise(p=1, mean=25, var_limit=(10, 70)) stand-in: additive N(mean, var), var ~ U(var_limit),
    drawn from the numpy global RNG; uint8 in / uint8 out."""

    def __init__(self, p=1, mean=25, var_limit=(10, 70)):
        self.mean, self.var_limit = mean, var_limit

    def __call__(self, image):
        var = np.random.uniform(self.var_limit[0], self.var_limit[1])
        noise = np.random.normal(self.mean, var ** 0.5, image.shape)
        return {"image": np.clip(image.astype(np.float64) + noise, 0, 255).astype(np.uint8)}


class StandInRandomShadow:
    """albumentations.RandomShadow(p=1) stand-in: halves the brightness of a random axis-aligned rectangle in the lower
    half of the image (numpy global RNG); uint8 in / uint8 out."""

    def __init__(self, p=1):
        pass

    def __call__(self, image):
        h, w = image.shape[:2]
        x0, x1 = sorted(np.random.randint(0, w + 1, size=2))
        y0 = np.random.randint(h // 2, h)
        out = image.copy()
        out[y0:, x0:x1] = out[y0:, x0:x1] // 2
        return {"image": out}


# --------------------------------------------------------------------------
# synthetic triggers and the DCT batch (train.py:37-38, 49-62, 106-143, 188-200)
# --------------------------------------------------------------------------
def dct2(block: np.ndarray) -> np.ndarray:
    """train.py:37-38: orthonormal 2-D DCT-II of one plane (scipy float64 arithmetic on the integer input)."""
    return scipy.fft.dctn(block.astype(np.float64), norm="ortho")


def _hwc(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().transpose(1, 2, 0)  # train.py:70-72


def addnoise(img: np.ndarray, gauss_noise=StandInGaussNoise) -> np.ndarray:
    """train.py:49-53"""
    aug = gauss_noise(p=1, mean=25, var_limit=(10, 70))
    return aug(image=(img * 255).astype(np.uint8))["image"] / 255


def randshadow(img: np.ndarray, input_size=32, random_shadow=StandInRandomShadow) -> np.ndarray:
    """train.py:56-61 (cv2.resize to the image's own size is the identity; kept out of the oracle)."""
    aug = random_shadow(p=1)
    test = (img * 255).astype(np.uint8)
    assert test.shape[0] == input_size and test.shape[1] == input_size
    return aug(image=test)["image"] / 255


def patching_train(sample: torch.Tensor, train_data: torch.Tensor, n_input=3, input_size=32) -> np.ndarray:
    """train.py:106-143: one synthetic-trigger image (HWC).  numpy global RNG, in the reference's order:
    attack, patch height, patch width, [block noise | augmentation draws | blend partner], margin, corner."""
    clean = _hwc(sample)
    attack = np.random.randint(0, 5)
    px = np.random.randint(2, 8)
    py = np.random.randint(2, 8)
    output = np.copy(clean)
    if attack == 0:
        block = np.ones((px, py, n_input))
    elif attack == 1:
        block = np.random.rand(px, py, n_input)
    elif attack == 2:
        return addnoise(output)
    elif attack == 3:
        return randshadow(output, input_size)
    if attack == 4:
        randind = np.random.randint(train_data.shape[0])
        mid = output + 0.3 * _hwc(train_data[randind])
        mid[mid > 1] = 1
        return mid
    margin = np.random.randint(0, 6)
    corner = np.random.randint(0, 4)  # 0 upper left, 1 upper right, 2 lower left, 3 lower right (train.py:131-139)
    r0 = margin if corner in (0, 1) else input_size - margin - px
    c0 = margin if corner in (0, 2) else input_size - margin - py
    output[r0:r0 + px, c0:c0 + py, :] = block
    output[output > 1] = 1
    return output


def quantise(x01: np.ndarray) -> np.ndarray:
    """(plane * 255).astype(np.uint8) of train.py:197 -- truncation toward zero, float64 arithmetic."""
    return (x01 * 255).astype(np.uint8)


def make_detector_batch(x: torch.Tensor, n_input=3, input_size=32, shuffle=True):
    """train.py:188-200 (train) / :236-250 (eval, shuffle=False): x [B,C,H,W] in [0,1] ->
    (uint8 planes [2B,C,H,W], DCT coefficients float32 [2B,C,H,W], labels int64 [2B]); first B rows clean (label 0),
    last B rows patched (label 1), then one `random.shuffle` of the row order."""
    B = x.shape[0]
    poi = np.zeros((B, n_input, input_size, input_size))
    for i in range(B):
        poi[i] = np.transpose(patching_train(x[i], x, n_input, input_size), (2, 0, 1))
    planes = np.vstack((x.detach().cpu().numpy(), poi))          # float64
    labels = np.vstack((np.zeros((B, 1)), np.ones((B, 1)))).astype(np.uint8)
    q = quantise(planes)
    coef = scipy.fft.dctn(q.astype(np.float64), axes=(-2, -1), norm="ortho")  # == dct2 per (image, channel)
    idx = np.arange(2 * B)
    if shuffle:
        random.shuffle(idx)
    return q[idx], torch.tensor(coef[idx], dtype=torch.float), torch.tensor(labels[idx].flatten().astype(int).tolist())


# --------------------------------------------------------------------------
# FrequencyModel, train mode (model.py:8-52) and Adadelta (train.py:152)
# --------------------------------------------------------------------------
def frequency_model_train_forward(p: dict, b: dict, x: torch.Tensor, dropout: float = 0.2, momentum: float = 0.1,
                                  eps: float = 1e-5) -> torch.Tensor:
    """conv -> ELU -> BatchNorm(batch statistics, running stats updated in `b`) x6; after layers 2/4/6 MaxPool2d(2) then
    Dropout(0.2) from the torch default generator (module order of model.py:13-44); flatten; linear."""
    for i in range(1, 7):
        x = F.conv2d(x, p["conv%d.weight" % i], p["conv%d.bias" % i], 1, 1)
        x = F.elu(x)
        x = F.batch_norm(x, b["bn%d.running_mean" % i], b["bn%d.running_var" % i], p["bn%d.weight" % i],
                         p["bn%d.bias" % i], True, momentum, eps)
        if "bn%d.num_batches_tracked" % i in b:
            b["bn%d.num_batches_tracked" % i] += 1
        if i % 2 == 0:
            x = F.max_pool2d(x, 2)
            x = F.dropout(x, dropout, True)
    return F.linear(x.flatten(1), p["linear6.weight"], p["linear6.bias"])


def adadelta_step(params: dict, grads: dict, state: dict, lr=0.05, rho=0.9, eps=1e-6, wd=1e-4):
    """torch.optim.Adadelta (single-tensor path): g <- g + wd p; v <- rho v + (1-rho) g^2;
    delta = sqrt(u + eps) / sqrt(v + eps) * g; u <- rho u + (1-rho) delta^2; p <- p - lr delta."""
    for k, pt in params.items():
        g = grads.get(k)
        if g is None:
            continue
        if k not in state:
            state[k] = {"square_avg": torch.zeros_like(pt), "acc_delta": torch.zeros_like(pt)}
        v, u = state[k]["square_avg"], state[k]["acc_delta"]
        g = g.add(pt, alpha=wd)
        v.mul_(rho).addcmul_(g, g, value=1 - rho)
        std = v.add(eps).sqrt_()
        delta = u.add(eps).sqrt_()
        delta.div_(std).mul_(g)
        u.mul_(rho).addcmul_(delta, delta, value=1 - rho)
        pt.add_(delta, alpha=-lr)


def detector_train_step(state: dict, x: torch.Tensor, n_input=3, input_size=32) -> dict:
    """One iteration of train.py:185-209.  state: {"p": params, "b": BatchNorm buffers, "opt": Adadelta state};
    mutated in place.  Returns what the reference's loop observes (and the tensors the CUDA path will be checked on)."""
    q, coef, y = make_detector_batch(x, n_input, input_size, shuffle=True)
    p = {k: v.detach().requires_grad_(True) for k, v in state["p"].items()}
    preds = frequency_model_train_forward(p, state["b"], coef)
    loss = F.cross_entropy(preds, y)
    grads = dict(zip(p.keys(), torch.autograd.grad(loss, list(p.values()))))
    with torch.no_grad():
        adadelta_step(state["p"], grads, state["opt"])
    return {"planes_u8": q, "x_final": coef, "y_final": y, "preds": preds.detach(), "loss": float(loss.detach()),
            "correct": int((preds.argmax(1) == y).sum()), "grads": grads}


def detector_eval_batch(state: dict, x: torch.Tensor, n_input=3, input_size=32) -> dict:
    """One batch of train.py:230-253: no shuffle, eval-mode network (oracle/combat_oracle.frequency_model_forward)."""
    from .combat_oracle import frequency_model_forward

    q, coef, y = make_detector_batch(x, n_input, input_size, shuffle=False)
    with torch.no_grad():
        preds = frequency_model_forward(state["p"], state["b"], coef)
    return {"planes_u8": q, "x_final": coef, "y_final": y, "preds": preds, "correct": int((preds.argmax(1) == y).sum())}
