"""Composition check of the detector TRAINING path without a GPU: nets.FrequencyDetector.train_forward / train_backward /
adadelta_step and combat_b200.defenses.frequency_based.train.train_iteration are run with every kernel wrapper of
`combat_b200.ops` replaced by a torch-CPU model of that kernel's documented contract (include/combat_b200.h), and held to the
fixture recorded from the unmodified reference train().

What this does and does not show: the HOST logic -- layer order, saved tensors, gradient routing, OHWI gradient layout,
NHWC dropout-mask layout, flatten order of the linear layer, running-statistics and Adadelta wiring, RNG consumption -- is
the reference's.  It says nothing about the CUDA kernels themselves: those are checked on the GPU
(tests/test_detector_train_gpu.py).  The models below live in tests/ only; the product has no CPU execution mode
(NetBase refuses non-CUDA devices -- this test swaps that one check out in a private copy of NetBase.__init__)."""
import inspect
import random
import textwrap
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import combat_oracle as O
from oracle import detector_oracle as D


class FakeOps:
    """torch-CPU models of the kernel contracts the training path uses (NHWC activations, OHWI weights)."""

    def __init__(self, real, net_ref):
        self.real, self.net_ref = real, net_ref
        self.nchw_strides, self.nhwc_strides, self.F32 = real.nchw_strides, real.nhwc_strides, real.F32
        self.dt_code = real.dt_code
        self.make_wprep_table = lambda table, device: (None, 0)            # compute-layout copies are not modelled
        self._contig = real._contig

    # weights are addressed through NetBase._wptr -> (conv name, dgrad?) keys in this model
    def _w(self, key):
        name, dgrad = key
        return self.net_ref[0].store.p(name + ".weight"), dgrad          # OIHW logical view of the master weights

    def _as_nchw(self, x, geom, strides):
        N, H, W = geom
        sn, sh, sw, sc = strides
        Cc = x.numel() // (N * H * W)
        return torch.as_strided(x, (N, Cc, H, W), (sn, sc, sh, sw))

    def conv_simt(self, x, x_geom, x_strides, w, w_dt, out, out_geom, out_strides, *, Ci, Co, KH, KW, stride, pad, up=1,
                  bias=None, residual=None, act=0, post_scale=None, post_shift=None):
        W_, dgrad = self._w(w)
        xin = self._as_nchw(x, x_geom, x_strides)
        if dgrad:   # input gradient: full correlation with the flipped, (co <-> ci)-transposed filter == conv_transpose2d
            y = F.conv_transpose2d(xin, W_[:, :Co], None, 1, KH - 1 - pad)
        else:
            y = F.conv2d(xin, W_, bias, stride, pad)
        if act == 2:
            y = F.elu(y)
        out.copy_(y.permute(0, 2, 3, 1))
        return out

    def conv_wgrad_simt(self, x, x_geom, x_strides, dy, dy_geom, dy_strides, dw, *, Ci, Co, KH, KW, stride, pad, db=None):
        xin = self._as_nchw(x, x_geom, x_strides).detach()
        g = dy.permute(0, 3, 1, 2)
        w = torch.zeros(Co, Ci, KH, KW, requires_grad=True)
        F.conv2d(xin, w, None, stride, pad).backward(g)
        dw.add_(w.grad.permute(0, 2, 3, 1).reshape(-1))                    # OHWI storage order, accumulated
        if db is not None:
            db.add_(g.sum((0, 2, 3)))

    def bn_train_prepare(self, x2d, R, Cc, gamma, beta, rm, rv, momentum, eps):
        v = x2d.reshape(-1, Cc)
        mean, var = v.mean(0), v.var(0, unbiased=False)
        rm.mul_(1 - momentum).add_(momentum * mean)
        rv.mul_(1 - momentum).add_(momentum * var * R / (R - 1))
        invstd = (var + eps).rsqrt()
        return gamma * invstd, beta - mean * gamma * invstd, mean, invstd

    def affine_act(self, x, scale, shift, relu, residual=None, out=None, out_dtype=None):
        y = x * scale + shift
        return F.relu(y) if relu else y

    def maxpool2(self, x):
        return F.max_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).contiguous()

    def maxpool2_bwd(self, dy, x):
        xn = x.permute(0, 3, 1, 2).detach().clone().requires_grad_(True)
        F.max_pool2d(xn, 2).backward(dy.permute(0, 3, 1, 2))
        return xn.grad.permute(0, 2, 3, 1).contiguous()

    def mask_scale(self, x, keep, scale):
        assert keep.dtype == torch.uint8 and keep.shape == x.shape
        return x * keep * scale

    def pool_linear_fwd(self, x, P, W, b):
        assert P == 1
        feat = x.permute(0, 3, 1, 2).reshape(x.shape[0], -1)               # flatten in NCHW order
        return feat @ W.t() + b, feat

    def pool_linear_bwd(self, dlogits, pooled, W, x_shape, dtype, P, dW=None, db=None, want_dx=True):
        B, Hf, Wf, Cc = x_shape
        dW.add_(dlogits.t() @ pooled)
        db.add_(dlogits.sum(0))
        return (dlogits @ W).view(B, Cc, Hf, Wf).permute(0, 2, 3, 1).contiguous()

    def bn_bwd_train(self, dy, x, y, gamma, mean, invstd, relu, dgamma_out, dbeta_out, dadd=None, want_dres=False):
        assert not relu
        Cc = x.shape[-1]
        xh = (x - mean) * invstd
        g = dy.reshape(-1, Cc)
        dgamma_out.copy_((g * xh.reshape(-1, Cc)).sum(0))
        dbeta_out.copy_(g.sum(0))
        R = g.shape[0]
        dx = gamma * invstd * (dy - dbeta_out / R - xh * dgamma_out / R)
        return dx, None

    def elu_bwd(self, da, a):
        return da * torch.where(a > 0, torch.ones_like(a), a + 1)

    def cross_entropy(self, logits, targets, grad_scale=1.0, want_grad=True, targets2=None, loss_out=None, counts_out=None):
        lg = logits.detach().clone().requires_grad_(True)
        loss = F.cross_entropy(lg, targets)
        loss.backward()
        counts = torch.tensor([int((logits.argmax(1) == targets).sum()), 0], dtype=torch.int32)
        return loss.detach().reshape(1), lg.grad * grad_scale, counts

    def adadelta(self, p, g, square_avg, acc_delta, lr_dev, rho=0.9, eps=1e-6, wd=1e-4):
        D.adadelta_step({"w": p}, {"w": g}, {"w": {"square_avg": square_avg, "acc_delta": acc_delta}}, float(lr_dev), rho, eps, wd)


@pytest.fixture()
def cpu_detector(monkeypatch):
    from combat_b200 import nets, ops
    import combat_b200.defenses.frequency_based.train as T
    src = inspect.getsource(nets.NetBase.__init__)
    assert 'if self.device.type not in ("cuda", "meta"):' in src
    ns = dict(vars(nets))      # the copy runs with the module's own globals
    exec(compile(textwrap.dedent(src).replace('if self.device.type not in ("cuda", "meta"):', "if False:"), "<cpu-init>", "exec"), ns)
    monkeypatch.setattr(nets.NetBase, "__init__", ns["__init__"])
    holder = []
    fake = FakeOps(ops, holder)
    monkeypatch.setattr(nets, "ops", fake)
    monkeypatch.setattr(T, "ops", fake)
    monkeypatch.setattr(nets.NetBase, "prep_weights", lambda self: None)
    monkeypatch.setattr(nets.NetBase, "_wptr", lambda self, cs, dgrad=False: (cs.name, dgrad))
    monkeypatch.setattr(T, "dct_2d", lambda q: O.dct_2d(q).contiguous())    # the device launch returns a fresh contiguous tensor
    net = nets.FrequencyDetector(device="cpu", dtype=torch.float32, trainable=True)
    holder.append(net)
    return net, T


class StandInAugment:
    def addnoise(self, img):
        return D.addnoise(img)

    def randshadow(self, img, input_size=32):
        return D.randshadow(img, input_size)


def rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_training_iterations_compose_like_the_reference(cpu_detector, golden):
    net, T = cpu_detector
    fx = golden("detector_b8x2.npz")
    seed = int(fx["seed"])
    gen = torch.Generator().manual_seed(seed)
    p0, b0 = O.init_frequency_model_state(gen)
    xs = [torch.rand(8, 3, 32, 32, generator=gen) for _ in range(3)]
    net.load_state_dict({**p0, **b0})
    netC = types.SimpleNamespace(net=net)
    opt = types.SimpleNamespace(input_channel=3, input_height=32, input_width=32, device="cpu")
    lr = torch.full((1,), 0.05)
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed + 1)
    torch.set_num_threads(8)
    for i in range(2):
        loss, counts, logits = T.train_iteration(netC, xs[i], opt, lr, StandInAugment())
        assert rel(logits, fx["preds%d" % i]) < 1e-4, i
        assert abs(float(loss) - float(fx["loss%d" % i])) < 1e-5
        want = int((torch.from_numpy(fx["preds%d" % i]).argmax(1) == torch.from_numpy(fx["y_final%d" % i])).sum())
        assert int(counts[0]) == want
    sd = net.state_dict()
    for k in ("conv1.weight", "conv1.bias", "bn1.weight", "bn1.running_mean", "bn1.running_var", "bn6.running_var", "linear6.weight",
              "linear6.bias"):
        assert rel(sd[k], fx["final." + k]) < 1e-4, k
    for name, l2 in zip(fx["final_names"], fx["final_l2"]):
        assert abs(float(sd[str(name)].double().norm()) - l2) < 1e-5 * max(1.0, l2), name
    assert rel(net.store._view(net.store.mom, "linear6.weight"), fx["adadelta.linear6.weight.square_avg"]) < 1e-4
    assert rel(net.store._view(net.store.acc_delta, "linear6.weight"), fx["adadelta.linear6.weight.acc_delta"]) < 1e-4
