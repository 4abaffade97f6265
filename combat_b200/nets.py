"""Explicit forward/backward graphs of the networks on the hot path, built from the C-ABI kernels.

No autograd: every network is a hand-scheduled sequence of kernel launches with saved activations, so the
alternated step can be captured in a CUDA graph and run the "strictly needed" schedule of SURVEY.md section 3.1
(dgrad-only backward through netC / clean_model in the G-step, no generator backward in the C-step).

Reference structures restated here (paths relative to the reference root):
  PreActResNet18   classifier_models/preact_resnet.py:13-40,72-110
  ResNet18         classifier_models/resnet.py:15-37,68-106
  UnetGenerator    networks/models.py:268-341         CUnetGeneratorv1  networks/models.py:472-555
  FrequencyModel   defenses/frequency_based/model.py:8-52
Parameter names, shapes and order are the reference's state_dict keys.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import torch

from . import ops
from ._lib import check, lib


# Optional per-launch timing of the tensor-core convolutions (bench.py's roofline leg): when TC_PROFILE is a list,
# every tcgen05 launch is bracketed by CUDA events on the launching stream and (name, flops, start, end) is appended.
TC_PROFILE = None


def _tc_launch(fn, name, flops, tag=""):
    if TC_PROFILE is None:
        check(fn(), name)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    check(fn(), name)
    e.record()
    TC_PROFILE.append((name, flops, s, e, tag))


def _tag(N, H, W, cs):
    return "N%d %dx%d %d->%d k%d s%d" % (N, H, W, cs.Cin, cs.Cout, cs.k, cs.stride)


def _align(n, a=4):
    return (n + a - 1) // a * a


class ParamStore:
    """All parameters of a network in ONE flat float32 buffer (+ gradient and momentum buffers of the same
    shape) so that SGD, gradient zeroing and the data-parallel all-reduce are single launches.

    4-D (conv) weights are STORED channels-last ([Cout][KH][KW][Cin], the layout the kernels consume and the
    weight-gradient kernels produce); `p(name)` exposes them as a permuted view with the reference's OIHW shape,
    i.e. exactly a torch channels_last tensor, so state_dict round-trips keep the reference's names and shapes."""

    def __init__(self, specs, device, store_ci=None):
        """store_ci: {weight name: stored input channels >= the logical ones}.  The extra input channels of such a conv
        weight exist only in storage (zero, and they stay zero: their inputs are zero planes, so their gradient is exactly
        zero and weight decay keeps 0 at 0); `p/g/m` expose the logical OIHW shape as a slice of the stored tensor."""
        self.names = [n for n, _ in specs]
        self.shapes = {n: tuple(s) for n, s in specs}
        self.store_ci = dict(store_ci or {})
        self.offsets = {}
        off = 0
        for n, s in specs:
            self.offsets[n] = off
            off += _align(self._stored_numel(n))
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device)
        self.mom = torch.zeros(off, dtype=torch.float32, device=device)
        self.first_step = True

    def _stored_numel(self, n):
        s = self.shapes[n]
        num = 1
        for d in s:
            num *= d
        if n in self.store_ci:
            num = num // s[1] * self.store_ci[n]
        return num

    def _view(self, buf, n):
        s = self.shapes[n]
        o = self.offsets[n]
        flat = buf[o:o + self._stored_numel(n)]
        if len(s) == 4:
            co, ci, kh, kw = s
            v = flat.view(co, kh, kw, self.store_ci.get(n, ci))
            if n in self.store_ci:
                v = v[..., :ci]
            return v.permute(0, 3, 1, 2)
        return flat.view(s)

    def raw(self, buf, n):
        """contiguous storage-order view (channels-last for conv weights, stored input channels included)"""
        o = self.offsets[n]
        return buf[o:o + self._stored_numel(n)]

    def p(self, n):
        return self._view(self.flat, n)

    def g(self, n):
        return self._view(self.grad, n)

    def m(self, n):
        return self._view(self.mom, n)

    def gptr(self, n):
        return self.grad.data_ptr() + 4 * self.offsets[n]

    def load(self, sd: dict):
        for n in self.names:
            self.p(n).copy_(sd[n].to(torch.float32))

    def state(self) -> dict:
        return {n: self.p(n).contiguous().clone() for n in self.names}


@dataclass
class ConvSpec:
    name: str
    Cin: int
    Cout: int
    k: int
    stride: int
    pad: int
    bias: bool = False
    need_dgrad: bool = True
    fwd_off: int = 0
    dgrad_off: int = -1


class NetBase:
    """Shared machinery: parameter store, compute-layout weights, conv dispatch."""

    def __init__(self, device, dtype, use_tc=True):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None and torch.cuda.is_available():
            self.device = torch.device("cuda", torch.cuda.current_device())
        if self.device.type not in ("cuda", "meta"):  # "meta": shape-only introspection (names / shapes), nothing can run
            raise RuntimeError("combat_b200 networks run on CUDA devices only (no CPU fallback)")
        self.dtype = dtype
        self.dt = ops.dt_code(dtype)
        self.use_tc = bool(use_tc) and dtype == torch.bfloat16
        # conv outputs that feed a normalisation (and the residual stream) stay float32: statistics and the
        # normalisation itself then see the unrounded fp32 accumulator; activations/gradients use `dtype`
        self.pre_dtype = torch.float32
        self.fast_small = True  # direct kernels for the 3-channel boundary layers (csrc/conv_small.cu)
        # 3-input-channel convs as im2col3 + 1x1 tcgen05 conv instead of the direct kernel: measured on B200 at B=512 (A/B, 20
        # steps, twice): 12.97 ms/step with it, 12.80 without -- the 1x1 conv is bound by its 201 MB of output writes either
        # way (pure-write streams run at ~3.3 TB/s here) and the im2col pass costs more than the FMA work it removes.  Opt-in.
        self.tc_first = bool(os.environ.get("COMBAT_TC_FIRST"))
        self.tc_first_wgrad = not os.environ.get("COMBAT_NO_TC_FIRST_WGRAD")  # the weight gradient of those convs through im2col3
        # the 1x1 stride-2 shortcut's input gradient as an extra tap of the block's 3x3 stride-2 input-gradient launch
        self.fuse_shortcut_dgrad = not os.environ.get("COMBAT_NO_FUSE_SC")
        # train-mode BatchNorm backward: the reduction (sum g, sum g * xhat) in the epilogue of the input-gradient conv that
        # produces g, instead of a pass of its own over (dy, x, y)
        self.fuse_bn_bwd_reduce = not os.environ.get("COMBAT_NO_FUSE_BNB")
        # 64 -> 3 convs (generator output, input gradient of the classifiers' first conv) on the tensor pipe
        self.tc_cout3 = not os.environ.get("COMBAT_NO_TC_COUT3")
        # 3 -> 64 first convs (stride 1) on the tensor pipe, operand built in shared memory from the float32 NCHW image
        self.tc_first2 = not os.environ.get("COMBAT_NO_TC_FIRST2")
        self._w64 = {}
        self.last_stats_nblk = 0
        self.convs: dict[str, ConvSpec] = {}

    # ---- construction helpers
    def _finish_params(self, specs, conv_specs, store_ci=None):
        self.store = ParamStore(specs, self.device, store_ci)
        off = 0
        table = []
        for cs in conv_specs:
            n = cs.Cout * cs.Cin * cs.k * cs.k
            cs.fwd_off = off
            off += _align(n, 64)
            if cs.need_dgrad:
                cs.dgrad_off = off
                off += _align(n, 64)
            self.convs[cs.name] = cs
            table.append((self.store.offsets[cs.name + ".weight"], cs.fwd_off, cs.dgrad_off, cs.Cout, cs.Cin, cs.k, cs.k))
        self.wbuf = torch.zeros(off, dtype=self.dtype, device=self.device)
        if self.device.type == "meta":
            self._wtable, self._wmax, self._n_wdesc, self.esz = None, 0, len(table), 2
            return
        self._wtable, self._wmax = ops.make_wprep_table(table, self.device)
        self._n_wdesc = len(table)
        self.esz = self.wbuf.element_size()

    def prep_weights(self):
        """master OIHW float32 -> compute layouts (OHWI and flipped/transposed for dgrad) in the compute dtype."""
        ops.prep_weights(self.store.flat, self.wbuf, self._wtable, self._n_wdesc, self._wmax)
        for (name, dgrad), w64 in self._w64.items():  # [rows][27] -> [rows][27 | 0 | 27 | 0] (hi and lo halves of im2col3)
            cs = self.convs[name]
            rows = cs.Cin if dgrad else cs.Cout
            off = cs.dgrad_off if dgrad else cs.fwd_off
            src = self.wbuf[off:off + rows * 27].view(rows, 27)
            w64[:, 0:27].copy_(src)
            w64[:, 32:59].copy_(src)

    def _use_im2col(self, cs: "ConvSpec", rows, n_in, wgrad=False):
        """3 input channels, 3x3, pad 1, `rows` (a multiple of 64) output channels on the bf16 tensor-core path."""
        on = self.tc_first_wgrad if wgrad else self.tc_first
        return self.use_tc and on and n_in == 3 and cs.k == 3 and cs.pad == 1 and rows % 64 == 0

    def _w64_for(self, cs: "ConvSpec", dgrad=False):
        key = (cs.name, dgrad)
        if key not in self._w64:
            rows = cs.Cin if dgrad else cs.Cout
            self._w64[key] = torch.zeros((rows, 64), dtype=self.dtype, device=self.device)
            self.prep_weights()
        return self._w64[key]

    def cin3_tc(self, x_nchw, w64, rows, stride, out, bias=None, out2=None, bn_relu=None, stats=False, tag=""):
        """3 -> rows conv of an NCHW float32 image on the tensor cores: im2col3 + 1x1 tcgen05 conv (same epilogues)."""
        A = ops.im2col3(x_nchw, stride)
        N, Ho, Wo, _ = A.shape
        want_stats = stats and out is not None and out.dtype == torch.float32
        d = ops.conv_tc_desc(A, w64.data_ptr(), out, N, Ho, Wo, 64, Ho, Wo, rows, 1, 1, 1, 0, 1, bias=bias, out2=out2,
                             scale2=bn_relu[0] if bn_relu is not None else None, shift2=bn_relu[1] if bn_relu is not None else None,
                             stats=ops.Scratch.get(self.device) if want_stats else None)
        _tc_launch(lambda: lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc", 2.0 * N * Ho * Wo * rows * 27,
                   "N%d %dx%d 3->%d k3 s%d im2col%s" % (N, Ho * stride, Wo * stride, rows, stride, tag))
        self.last_stats_nblk = lib.combat_conv_tc_last_grid() if want_stats else 0
        return out

    def zero_grad(self):
        self.store.grad.zero_()

    def sgd_step(self, lr_dev, momentum=0.9, wd=5e-4):
        st = self.store
        ops.sgd_nesterov(st.flat, st.grad, st.mom, lr_dev, momentum, wd, st.first_step)
        st.first_step = False
        self.prep_weights()

    def _wptr(self, cs: ConvSpec, dgrad=False):
        return self.wbuf.data_ptr() + (cs.dgrad_off if dgrad else cs.fwd_off) * self.esz

    def _bias(self, cs):
        return self.store.p(cs.name + ".bias") if cs.bias else None

    # ---- conv dispatch (NHWC activations)
    def _tc_ok(self, cs: ConvSpec):
        return self.use_tc and cs.Cin % 64 == 0 and cs.Cout % 64 == 0

    def conv_fwd(self, x, cs: ConvSpec, residual=None, pre=True, bn_relu=None, want_out=True, stats=False, post=None):
        """x: NHWC [N,H,W,Cin] in the activation dtype -> NHWC out (float32 when `pre`, i.e. feeding a norm).
        bn_relu=(scale, shift): the tcgen05 epilogue ALSO writes out2 = bf16 relu(out*scale+shift) (eval BatchNorm+ReLU of
        the consumer); returns (out or None, out2) then.
        post=(scale, shift): out = (acc + bias) * scale + shift + residual (an eval-mode BatchNorm BEHIND this conv, the residual
        added after it -- the non-PreAct ResNet ordering; tcgen05 path only)."""
        N, H, W, Ct = x.shape
        Ho = (H + 2 * cs.pad - cs.k) // cs.stride + 1
        Wo = (W + 2 * cs.pad - cs.k) // cs.stride + 1
        if bn_relu is not None:
            out = torch.empty((N, Ho, Wo, cs.Cout), dtype=self.pre_dtype, device=self.device) if want_out else None
            out2 = torch.empty((N, Ho, Wo, cs.Cout), dtype=self.dtype, device=self.device)
            d = ops.conv_tc_desc(x, self._wptr(cs), out, N, H, W, cs.Cin, Ho, Wo, cs.Cout, cs.k, cs.k, cs.stride, cs.pad, 1,
                                 bias=self._bias(cs), residual=residual, out2=out2, scale2=bn_relu[0], shift2=bn_relu[1],
                                 post_scale=post[0] if post is not None else None, post_shift=post[1] if post is not None else None)
            d.res_f32 = int(residual is not None and residual.dtype == torch.float32)
            if not (self._tc_ok(cs) and Ct == cs.Cin and lib.combat_conv_tc_supported(C.byref(d))):
                raise RuntimeError("fused BatchNorm epilogue needs the tcgen05 path")
            _tc_launch(lambda: lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc", 2.0 * N * Ho * Wo * cs.Cout * cs.Cin * cs.k * cs.k,
                       _tag(N, H, W, cs) + " +bn")
            return out, out2
        out = torch.empty((N, Ho, Wo, cs.Cout), dtype=self.pre_dtype if pre else self.dtype, device=self.device)
        self.last_stats_nblk = 0  # > 0: the launch left train-mode BatchNorm partial sums of `out` in ops.Scratch
        if self._tc_ok(cs) and Ct == cs.Cin:
            want_stats = stats and pre and cs.Cout <= 512
            d = ops.conv_tc_desc(x, self._wptr(cs), out, N, H, W, cs.Cin, Ho, Wo, cs.Cout, cs.k, cs.k, cs.stride, cs.pad, 1,
                                 bias=self._bias(cs), residual=residual, stats=ops.Scratch.get(self.device) if want_stats else None,
                                 post_scale=post[0] if post is not None else None, post_shift=post[1] if post is not None else None)
            if lib.combat_conv_tc_supported(C.byref(d)):
                _tc_launch(lambda: lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc", 2.0 * N * Ho * Wo * cs.Cout * cs.Cin * cs.k * cs.k,
                           _tag(N, H, W, cs) + (" +stats" if want_stats else ""))
                if want_stats:
                    self.last_stats_nblk = lib.combat_conv_tc_last_grid()
                return out
        if post is not None:
            raise RuntimeError("fused BatchNorm epilogue needs the tcgen05 path")
        ops.conv_simt(x, (N, H, W), ops.nhwc_strides(H, W, Ct), self._wptr(cs), self.dt, out, (Ho, Wo),
                      ops.nhwc_strides(Ho, Wo, cs.Cout), Ci=cs.Cin, Co=cs.Cout, KH=cs.k, KW=cs.k, stride=cs.stride,
                      pad=cs.pad, bias=self._bias(cs), residual=residual)
        return out

    def conv_dgrad(self, dy, cs: ConvSpec, in_hw, residual=None, n_out_ch=None, mask=None, mask_scale=None, post_add=None,
                   shortcut=None, bnb=None):
        """dy: NHWC [N,Ho,Wo,Cout] -> dx NHWC [N,H,W,n_out_ch or Cin] (+ residual).
        mask/mask_scale/post_add: fused backward of the eval-mode relu(bn(.)) in front of this conv (tcgen05 path only):
        dx = (mask > 0 ? (dgrad + residual) * mask_scale[c] : 0) + post_add.
        shortcut=(dy_sc, cs_sc): the block's 1x1 stride-2 shortcut conv reads the same input; its input gradient is one more
        tap of this launch (same accumulator) instead of a launch of its own plus a residual read.
        bnb=(x, scale, shift, mean, invstd): the conv's input went through a TRAIN-mode relu(bn(x)); the epilogue masks the
        gradient with the forward's predicate and leaves the per-CTA partial sums of the BatchNorm backward reduction in
        ops.Scratch (self.last_stats_nblk blocks) -- the caller finishes with ops.bn_bwd_train_from_partials."""
        N, Ho, Wo, _ = dy.shape
        H, W = in_hw
        Cx = cs.Cin if n_out_ch is None else n_out_ch
        dx = torch.empty((N, H, W, Cx), dtype=self.dtype, device=self.device)
        padp = cs.k - 1 - cs.pad
        # the dgrad filter is stored [Cin][taps][Cout]: its first Cx rows ARE the filter of the first Cx input channels, so a
        # leading multiple-of-64 subset runs on the tensor cores as a conv with Cx output channels (stride-1 convs only)
        fuse_sc = (shortcut is not None and self.fuse_shortcut_dgrad and self._tc_ok(cs) and self._tc_ok(shortcut[1]) and Cx == cs.Cin
                   and cs.stride == 2 and cs.k == 3 and cs.pad == 1 and shortcut[1].k == 1 and shortcut[1].stride == 2
                   and shortcut[1].Cin == cs.Cin and shortcut[1].Cout == cs.Cout and residual is None)
        if bnb is not None and (not self.bnb_ok(cs, bnb[0]) or (shortcut is not None and not fuse_sc) or residual is not None
                                or mask is not None or Cx != cs.Cin):
            raise RuntimeError("fused BatchNorm backward reduction: unsupported launch (check bnb_ok first)")
        if shortcut is not None and not fuse_sc:   # unfused: the shortcut's input gradient first, added as the residual
            residual_sc = self.conv_dgrad(shortcut[0], shortcut[1], in_hw, residual=residual)
            return self.conv_dgrad(dy, cs, in_hw, residual=residual_sc, n_out_ch=n_out_ch, mask=mask, mask_scale=mask_scale,
                                   post_add=post_add)
        if self._tc_ok(cs) and (Cx == cs.Cin or (Cx % 64 == 0 and Cx < cs.Cin and cs.stride == 1)):
            extra = dict(in2=shortcut[0], w2=self._wptr(shortcut[1], True)) if fuse_sc else {}
            if bnb is not None:
                extra.update(bnb=bnb, stats=ops.Scratch.get(self.device))
            d = ops.conv_tc_desc(dy, self._wptr(cs, True), dx, N, Ho, Wo, cs.Cout, H, W, Cx, cs.k, cs.k, 1, padp,
                                 cs.stride, residual=residual, mask=mask, mask_scale=mask_scale, post_add=post_add, **extra)
            if lib.combat_conv_tc_supported(C.byref(d)):
                fl = 2.0 * N * Ho * Wo * cs.Cout * Cx * (cs.k * cs.k + (1 if fuse_sc else 0))
                _tc_launch(lambda: lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc(dgrad)", fl,
                           _tag(N, H, W, cs) + (" +sc" if fuse_sc else "") + (" +bnb" if bnb is not None else ""))
                if bnb is not None:
                    self.last_stats_nblk = lib.combat_conv_tc_last_grid()
                return dx
        if bnb is not None:
            raise RuntimeError("fused BatchNorm backward reduction needs the tcgen05 path")
        if mask is not None or post_add is not None:
            raise RuntimeError("fused BatchNorm backward epilogue needs the tcgen05 path")
        ops.conv_simt(dy, (N, Ho, Wo), ops.nhwc_strides(Ho, Wo, cs.Cout), self._wptr(cs, True), self.dt, dx, (H, W),
                      ops.nhwc_strides(H, W, Cx), Ci=cs.Cout, Co=Cx, KH=cs.k, KW=cs.k, stride=1, pad=padp, up=cs.stride,
                      residual=residual)
        return dx

    def bnb_ok(self, cs: ConvSpec, x):
        """can the input-gradient launch of `cs` carry the reduction of the train-mode BatchNorm backward in front of it?"""
        return (self.fuse_bn_bwd_reduce and self._tc_ok(cs) and cs.k == 3 and x.dtype == torch.bfloat16 and cs.Cin <= 512
                and (cs.stride == 1 or self.fuse_shortcut_dgrad))

    def conv_wgrad(self, x, dy, cs: ConvSpec, dead_bias=False):
        """accumulates dW (OIHW float32) and the bias gradient into the flat gradient buffer.
        dead_bias: the conv feeds a non-affine InstanceNorm, whose backward returns a gradient with EXACTLY zero mean per
        (sample, channel) up to rounding -- the bias gradient is that rounding noise (SURVEY trap 6; ~1e-7 of the weight
        gradients in the reference too), so it is left at zero instead of spending a column-sum launch on it; the bias still
        receives its weight-decay update."""
        N, H, W, Ct = x.shape
        _, Ho, Wo, _ = dy.shape
        dw = self.store.raw(self.store.grad, cs.name + ".weight")
        db = self.store.g(cs.name + ".bias") if (cs.bias and not dead_bias) else None
        if self._tc_ok(cs) and Ct == cs.Cin:
            d = ops.conv_tc_desc(x, None, None, N, H, W, cs.Cin, Ho, Wo, cs.Cout, cs.k, cs.k, cs.stride, cs.pad, 1)
            if lib.combat_conv_tc_supported(C.byref(d)):
                _tc_launch(lambda: lib.combat_conv_tc_wgrad(C.byref(d), ops._p(dy), ops._p(dw), ops._s()), "conv_tc_wgrad", 2.0 * N * Ho * Wo * cs.Cout * cs.Cin * cs.k * cs.k, _tag(N, H, W, cs))
                if db is not None:
                    ops.colsum(dy, cs.Cout, db)
                return
        ops.conv_wgrad_simt(x, (N, H, W), ops.nhwc_strides(H, W, Ct), dy, (Ho, Wo), ops.nhwc_strides(Ho, Wo, cs.Cout), dw,
                            Ci=cs.Cin, Co=cs.Cout, KH=cs.k, KW=cs.k, stride=cs.stride, pad=cs.pad, db=db)

    # ---- image-boundary convs (NCHW float32 on the 3-channel side), always CUDA-core kernels
    def conv_first_fwd(self, x_nchw, cs: ConvSpec, out=None, out_ctot=None, pre=True, bn_relu=None, stats=False):
        N, Cc, H, W = x_nchw.shape
        Ho = (H + 2 * cs.pad - cs.k) // cs.stride + 1
        Wo = (W + 2 * cs.pad - cs.k) // cs.stride + 1
        Ct = cs.Cout if out_ctot is None else out_ctot
        if out is None:
            out = torch.empty((N, Ho, Wo, Ct), dtype=self.pre_dtype if pre else self.dtype, device=self.device)
        self.last_stats_nblk = 0
        if (self.tc_first2 and self.use_tc and Cc == 3 and cs.k == 3 and cs.pad == 1 and cs.stride == 1 and Ct == cs.Cout == 64
                and W in (16, 32, 64, 128) and H * W >= 128):
            # tensor pipe: the epilogue warps build the [hi | lo] im2col operand in shared memory (conv_tc_first_kernel)
            out2 = torch.empty((N, Ho, Wo, Ct), dtype=self.dtype, device=self.device) if bn_relu is not None else None
            want_stats = stats and pre
            d = ops.conv_tc_desc(x_nchw, self._w64_for(cs).data_ptr(), out, N, H, W, 64, Ho, Wo, 64, 3, 3, 1, 1, 1, bias=self._bias(cs),
                                 out2=out2, scale2=bn_relu[0] if bn_relu is not None else None,
                                 shift2=bn_relu[1] if bn_relu is not None else None,
                                 stats=ops.Scratch.get(self.device) if want_stats else None, in_nchw3=True)
            _tc_launch(lambda: lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc_boundary", 2.0 * N * Ho * Wo * 64 * 27,
                       "N%d %dx%d 3->64 k3 s1 first%s" % (N, H, W, " +bn" if bn_relu is not None else (" +stats" if want_stats else "")))
            self.last_stats_nblk = lib.combat_conv_tc_last_grid() if want_stats else 0
            return (out, out2) if bn_relu is not None else out
        if Ct == cs.Cout and self._use_im2col(cs, cs.Cout, Cc):
            out2 = torch.empty((N, Ho, Wo, Ct), dtype=self.dtype, device=self.device) if bn_relu is not None else None
            self.cin3_tc(x_nchw, self._w64_for(cs), cs.Cout, cs.stride, out, bias=self._bias(cs), out2=out2, bn_relu=bn_relu,
                         stats=stats)
            return (out, out2) if bn_relu is not None else out
        if self.fast_small and Cc == 3 and cs.k == 3 and cs.pad == 1 and Ct == cs.Cout and cs.Cout % 32 == 0 and 256 % cs.Cout == 0:
            if bn_relu is not None:
                out2 = torch.empty((N, Ho, Wo, Ct), dtype=self.dtype, device=self.device)
                ops.conv_cin3(x_nchw, self._wptr(cs), self.dt, out, cs.Cout, cs.stride, bias=self._bias(cs), out2=out2,
                              scale2=bn_relu[0], shift2=bn_relu[1])
                return out, out2
            return ops.conv_cin3(x_nchw, self._wptr(cs), self.dt, out, cs.Cout, cs.stride, bias=self._bias(cs))
        if bn_relu is not None:
            raise RuntimeError("fused BatchNorm epilogue needs the direct 3-channel kernel")
        ops.conv_simt(x_nchw, (N, H, W), ops.nchw_strides(Cc, H, W), self._wptr(cs), self.dt, out, (Ho, Wo),
                      ops.nhwc_strides(Ho, Wo, Ct), Ci=cs.Cin, Co=cs.Cout, KH=cs.k, KW=cs.k, stride=cs.stride, pad=cs.pad,
                      bias=self._bias(cs))
        return out

    def conv_first_wgrad(self, x_nchw, dy, cs: ConvSpec, dy_ctot=None):
        N, Cc, H, W = x_nchw.shape
        _, Ho, Wo, Ct = dy.shape
        if Ct == cs.Cout and dy.dtype == torch.bfloat16 and self._use_im2col(cs, cs.Cout, Cc, wgrad=True):
            # dW'[co][64] over the im2col operand (recomputed: 20 us, cheaper than keeping 64 MB alive), then the hi and
            # lo halves fold into dW[co][27]
            A = ops.im2col3(x_nchw, cs.stride)
            dw64 = torch.zeros((cs.Cout, 64), dtype=torch.float32, device=self.device)
            d = ops.conv_tc_desc(A, None, None, N, Ho, Wo, 64, Ho, Wo, cs.Cout, 1, 1, 1, 0, 1)
            _tc_launch(lambda: lib.combat_conv_tc_wgrad(C.byref(d), ops._p(dy), ops._p(dw64), ops._s()), "conv_tc_boundary",
                       2.0 * N * Ho * Wo * cs.Cout * 27, "N%d %dx%d 3->%d k3 s%d im2col wgrad" % (N, H, W, cs.Cout, cs.stride))
            ops.fold_w64(dw64, self.store.raw(self.store.grad, cs.name + ".weight"), cs.Cout, 0)
            if cs.bias:
                ops.colsum(dy, cs.Cout, self.store.g(cs.name + ".bias"))
            return
        if self.fast_small and Cc == 3 and cs.k == 3 and cs.pad == 1 and Ct == cs.Cout and cs.Cout % 32 == 0 and 256 % cs.Cout == 0:
            return ops.wgrad_cin3(x_nchw, dy, self.store.raw(self.store.grad, cs.name + ".weight"),
                                  self.store.g(cs.name + ".bias") if cs.bias else None, cs.Cout, cs.stride)
        ops.conv_wgrad_simt(x_nchw, (N, H, W), ops.nchw_strides(Cc, H, W), dy, (Ho, Wo), ops.nhwc_strides(Ho, Wo, Ct),
                            self.store.raw(self.store.grad, cs.name + ".weight"), Ci=cs.Cin, Co=cs.Cout, KH=cs.k, KW=cs.k, stride=cs.stride,
                            pad=cs.pad, db=self.store.g(cs.name + ".bias") if cs.bias else None)

    def conv_first_dgrad(self, dy, cs: ConvSpec, in_hw):
        """-> dx NCHW float32 (gradient w.r.t. the input image)."""
        N, Ho, Wo, _ = dy.shape
        H, W = in_hw
        dx = torch.empty((N, cs.Cin, H, W), dtype=torch.float32, device=self.device)
        if self.fast_small and cs.Cin == 3 and cs.Cout == 64 and cs.k == 3 and cs.pad == 1 and cs.stride == 1:
            if self.tc_cout3 and self.use_tc and ops.conv_tc_cout3_ok(dy):   # tensor pipe, N = 16 accumulator columns
                return ops.conv_tc_cout3(dy, self._wptr(cs, True), dx)
            return ops.conv_cout3(dy, self._wptr(cs, True), self.dt, dx)
        ops.conv_simt(dy, (N, Ho, Wo), ops.nhwc_strides(Ho, Wo, cs.Cout), self._wptr(cs, True), self.dt, dx, (H, W),
                      ops.nchw_strides(cs.Cin, H, W), Ci=cs.Cout, Co=cs.Cin, KH=cs.k, KW=cs.k, stride=1,
                      pad=cs.k - 1 - cs.pad, up=cs.stride)
        return dx


# ===================================================================== classifiers
@dataclass
class _BN:
    name: str
    C: int


class Classifier(NetBase):
    """PreActResNet18 / ResNet18 with explicit train/eval forward and (dgrad[, wgrad]) backward."""

    def __init__(self, arch="preact_resnet18", num_classes=10, n_input=3, input_size=32, device="cuda",
                 dtype=torch.bfloat16, use_tc=True, scaler=None):
        super().__init__(device, dtype, use_tc)
        assert arch in ("preact_resnet18", "resnet18")
        self.arch, self.num_classes, self.n_input, self.input_size = arch, num_classes, n_input, input_size
        # bf16 path: the pre-normalisation tensors and the residual stream are STORED in bf16 (statistics are still taken from
        # the float32 accumulators in the conv epilogue).  With the MMA issue fixed the 64- and 128-channel conv launches that
        # read a float32 residual and write a float32 stream were HBM-bound (64->64 @32x32, batch 512: 402 MB in 80 us); bf16
        # storage: 11.13 -> 10.68 ms per step.  (In round 1, when the issuing thread paced those kernels, it made no difference:
        # 12.85 vs 12.86 ms.)  The generator keeps float32 (InstanceNorm of a bf16-rounded tensor costs 3.4e-2 vs 3e-2 allowed).
        # COMBAT_PRE_F32=1 restores float32 storage.
        if dtype == torch.bfloat16 and not os.environ.get("COMBAT_PRE_F32"):
            self.pre_dtype = torch.bfloat16
        if scaler is None:
            scaler = {32: 1, 64: 4, 224: 49}[input_size]  # reference: {32:1, 64:4}; 224 -> 49 is the natural extension
        self.scaler = scaler
        pre = arch == "preact_resnet18"
        specs, convs, self.bns, self.blocks = [], [], [], []

        def add_conv(name, ci, co, k, s, p, need_dgrad=True):
            specs.append((name + ".weight", (co, ci, k, k)))
            cs = ConvSpec(name, ci, co, k, s, p, False, need_dgrad)
            convs.append(cs)
            return cs

        def add_bn(name, c):
            specs.append((name + ".weight", (c,)))
            specs.append((name + ".bias", (c,)))
            bn = _BN(name, c)
            self.bns.append(bn)
            return bn

        self.conv1 = add_conv("conv1", n_input, 64, 3, 1, 1)
        self.bn1 = None if pre else add_bn("bn1", 64)
        in_planes = 64
        for li, (planes, stride0) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
            for bi, stride in enumerate([stride0, 1]):
                pfx = "layer%d.%d." % (li, bi)
                blk = {"stride": stride, "in": in_planes, "planes": planes}
                if pre:
                    blk["bn1"] = add_bn(pfx + "bn1", in_planes)
                    blk["conv1"] = add_conv(pfx + "conv1", in_planes, planes, 3, stride, 1)
                    blk["bn2"] = add_bn(pfx + "bn2", planes)
                    blk["conv2"] = add_conv(pfx + "conv2", planes, planes, 3, 1, 1)
                    if stride != 1 or in_planes != planes:
                        blk["sc"] = add_conv(pfx + "shortcut.0", in_planes, planes, 1, stride, 0)
                else:
                    blk["conv1"] = add_conv(pfx + "conv1", in_planes, planes, 3, stride, 1)
                    blk["bn1"] = add_bn(pfx + "bn1", planes)
                    blk["conv2"] = add_conv(pfx + "conv2", planes, planes, 3, 1, 1)
                    blk["bn2"] = add_bn(pfx + "bn2", planes)
                    if stride != 1 or in_planes != planes:
                        blk["sc"] = add_conv(pfx + "shortcut.0", in_planes, planes, 1, stride, 0)
                        blk["scbn"] = add_bn(pfx + "shortcut.1", planes)
                self.blocks.append(blk)
                in_planes = planes
        specs.append(("linear.weight", (num_classes, 512 * scaler)))
        specs.append(("linear.bias", (num_classes,)))
        self._finish_params(specs, convs)
        # buffers: running stats of every BN in one flat buffer (mean | var per layer)
        self.buf_off = {}
        off = 0
        for bn in self.bns:
            self.buf_off[bn.name] = off
            off += 2 * bn.C
        self.bufs = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.num_batches_tracked = {bn.name: 0 for bn in self.bns}
        for bn in self.bns:
            self.store.p(bn.name + ".weight").fill_(1.0)
            self.rv(bn).fill_(1.0)
        self.momentum, self.eps = 0.1, 1e-5
        # eval-mode (scale, shift) of every BatchNorm in one launch: per-channel index table into store.flat / bufs
        self.aff_off, tab, n = {}, [], 0
        for bn in self.bns:
            self.aff_off[bn.name] = n
            go, bo, ro = self.store.offsets[bn.name + ".weight"], self.store.offsets[bn.name + ".bias"], self.buf_off[bn.name]
            tab += [[go + c, bo + c, ro + c, ro + bn.C + c] for c in range(bn.C)]
            n += bn.C
        self.n_bn_ch = n
        self._aff_table = torch.tensor(tab, dtype=torch.int32).to(self.device) if self.device.type == "cuda" else None
        # the fused eval path (BatchNorm+ReLU in the tcgen05 epilogues, bf16 path): PreAct ordering (round 1) and, [r2], the plain
        # ResNet ordering conv-bn-relu-conv-bn-add-relu (post-affine residual epilogue); COMBAT_NO_FUSE_RESNET=1 for A/B
        self.fuse_eval = self.use_tc and (pre or not os.environ.get("COMBAT_NO_FUSE_RESNET"))
        self._unit = {}
        # train-mode fusions of the plain ResNet ordering (statistics in the conv epilogues, bn1's backward sums in conv2's
        # input-gradient epilogue), as the PreAct path has them: CelebA multilabel step 21.92 -> 20.94 ms (A/B in one call);
        # COMBAT_NO_FUSE_RESNET_TRAIN=1 switches them off
        self.fuse_train_plain = pre or not os.environ.get("COMBAT_NO_FUSE_RESNET_TRAIN")

    def eval_affine(self):
        out = torch.empty((2, self.n_bn_ch), dtype=torch.float32, device=self.device)
        return ops.bn_eval_affine(self.store.flat, self.bufs, self._aff_table, self.n_bn_ch, self.eps, out)

    # ---- buffers
    def rm(self, bn):
        o = self.buf_off[bn.name]
        return self.bufs[o:o + bn.C]

    def rv(self, bn):
        o = self.buf_off[bn.name] + bn.C
        return self.bufs[o:o + bn.C]

    def load_state_dict(self, sd: dict):
        self.store.load(sd)
        for bn in self.bns:
            self.rm(bn).copy_(sd[bn.name + ".running_mean"])
            self.rv(bn).copy_(sd[bn.name + ".running_var"])
            self.num_batches_tracked[bn.name] = int(sd.get(bn.name + ".num_batches_tracked", 0))
        self.prep_weights()

    def state_dict(self) -> dict:
        sd = self.store.state()
        for bn in self.bns:
            sd[bn.name + ".running_mean"] = self.rm(bn).clone()
            sd[bn.name + ".running_var"] = self.rv(bn).clone()
            sd[bn.name + ".num_batches_tracked"] = torch.tensor(self.num_batches_tracked[bn.name])
        return sd

    def bump_batches_tracked(self, n=1):
        """One train-mode forward of every BatchNorm was executed by a CUDA-graph replay (the Python of _bn_fwd did not
        run): keep the reference's num_batches_tracked buffers in step with the iterations actually executed."""
        for k in self.num_batches_tracked:
            self.num_batches_tracked[k] += n

    # ---- BN helpers
    def _bn_fwd(self, bn, x, train, relu, residual=None, stats_nblk=0):
        """stats_nblk > 0: the conv that produced x already reduced its per-CTA sums (fused epilogue statistics)."""
        Cc = bn.C
        R = x.numel() // Cc
        g, b = self.store.p(bn.name + ".weight"), self.store.p(bn.name + ".bias")
        if train:
            if stats_nblk > 0:
                scale, shift, mean, invstd = ops.bn_train_finalize(stats_nblk, R, Cc, g, b, self.rm(bn), self.rv(bn),
                                                                   self.momentum, self.eps)
            else:
                scale, shift, mean, invstd = ops.bn_train_prepare(x, R, Cc, g, b, self.rm(bn), self.rv(bn), self.momentum,
                                                                  self.eps)
            if not torch.cuda.is_current_stream_capturing():  # a capture pass executes nothing; replays are counted by
                self.num_batches_tracked[bn.name] += 1        # bump_batches_tracked() (one call per replayed train forward)
            st = (scale, mean, invstd, shift)
        else:
            scale, shift = ops.bn_eval_prepare(Cc, g, b, self.rm(bn), self.rv(bn), self.eps)
            st = (scale, None, None, shift)
        y = ops.affine_act(x, scale, shift, relu, residual=residual, out_dtype=self.dtype)
        return y, st

    def _bn_bwd(self, bn, dy, x, y, st, train, relu, need_wgrad, dadd=None, want_dres=False):
        scale, mean, invstd = st[:3]
        if train:
            if need_wgrad:
                dg, db = self.store.g(bn.name + ".weight"), self.store.g(bn.name + ".bias")
            else:
                tmp = torch.empty((2, bn.C), dtype=torch.float32, device=self.device)
                dg, db = tmp[0], tmp[1]
            return ops.bn_bwd_train(dy, x, y, self.store.p(bn.name + ".weight"), mean, invstd, relu, dg, db, dadd, want_dres)
        return ops.bn_bwd_eval(dy, y, scale, relu, dadd, want_dres)

    def _bn_bwd_tail(self, bn, g, x, st, need_wgrad, dadd=None):
        """finalize + apply of a train-mode BatchNorm backward whose reduction ran in the producing conv's epilogue"""
        if need_wgrad:
            dg, db = self.store.g(bn.name + ".weight"), self.store.g(bn.name + ".bias")
        else:
            tmp = torch.empty((2, bn.C), dtype=torch.float32, device=self.device)
            dg, db = tmp[0], tmp[1]
        return ops.bn_bwd_train_from_partials(g, x, self.store.p(bn.name + ".weight"), st[1], st[2], self.last_stats_nblk, dg, db,
                                              dadd=dadd)

    # ---- forward
    def forward(self, x_nchw, train: bool, save: bool = True, fuse: bool = True):
        """x_nchw float32 [N,C,H,W] -> (logits float32 [N,num_classes], ctx).  fuse=False keeps the pre-normalisation
        tensors of an eval-mode forward (needed only if weight gradients are wanted from it)."""
        pre = self.arch == "preact_resnet18"
        if fuse and self.fuse_eval and not train and self.fast_small and x_nchw.shape[1] == 3 and x_nchw.shape[3] % 4 == 0:
            return self._forward_eval_fused(x_nchw, save)
        ctx = {"x": x_nchw, "train": train, "blocks": []} if save else None
        h = self.conv_first_fwd(x_nchw, self.conv1, stats=train and self.fuse_train_plain)
        h0_nblk = self.last_stats_nblk
        if not pre:
            c0 = h
            h, st0 = self._bn_fwd(self.bn1, c0, train, True, stats_nblk=h0_nblk if train else 0)
            if save:
                ctx["stem"] = (c0, h, st0)
        h_nblk = h0_nblk if (train and pre) else 0  # partial-sum blocks of h left by its producer conv (train mode, tcgen05 path)
        for blk in self.blocks:
            if pre:
                o1, st1 = self._bn_fwd(blk["bn1"], h, train, True, stats_nblk=h_nblk)
                s = self.conv_fwd(o1, blk["sc"]) if "sc" in blk else h
                c1 = self.conv_fwd(o1, blk["conv1"], stats=train)
                o2, st2 = self._bn_fwd(blk["bn2"], c1, train, True, stats_nblk=self.last_stats_nblk if train else 0)
                out = self.conv_fwd(o2, blk["conv2"], residual=s, stats=train)  # statistics of the next block's bn1 input
                h_nblk = self.last_stats_nblk if train else 0
                if save:
                    ctx["blocks"].append((h, o1, c1, o2, st1, st2))
            else:
                # [r2] train mode: every conv leaves the BatchNorm statistics of its output in its epilogue (one scratch buffer:
                # each set of partial sums is finalised by its BatchNorm before the next conv overwrites it)
                nb = lambda: self.last_stats_nblk if train else 0
                ft = train and self.fuse_train_plain
                c1 = self.conv_fwd(h, blk["conv1"], stats=ft)
                o1, st1 = self._bn_fwd(blk["bn1"], c1, train, True, stats_nblk=nb())
                if "sc" in blk:
                    cs_ = self.conv_fwd(h, blk["sc"], stats=ft)
                    s, sts = self._bn_fwd(blk["scbn"], cs_, train, False, stats_nblk=nb())
                else:
                    cs_, s, sts = None, h, None
                c2 = self.conv_fwd(o1, blk["conv2"], stats=ft)
                out, st2 = self._bn_fwd(blk["bn2"], c2, train, True, residual=s, stats_nblk=nb())
                if save:
                    ctx["blocks"].append((h, c1, o1, c2, out, cs_, st1, st2, sts))
            h = out
        logits, pooled = ops.pool_linear_fwd(h, 4, self.store.p("linear.weight"), self.store.p("linear.bias"))
        if save:
            ctx["feat_shape"], ctx["pooled"] = tuple(h.shape), pooled
        return logits, ctx

    # ---- eval-mode PreAct forward/backward with BatchNorm+ReLU folded into the conv epilogues
    def _forward_eval_fused(self, x_nchw, save):
        """netC.eval() / clean_model forward (train_generator.py:214,227,228,250): per block the chain
        bn1-relu-[shortcut]-conv1-bn2-relu-conv2-add is three tcgen05 launches and no elementwise pass -- each conv's
        epilogue writes the bf16 relu(bn(.)) tensor its consumer reads (and the float32 residual stream when needed)."""
        aff = self.eval_affine()

        def sl(bn):
            o = self.aff_off[bn.name]
            return aff[0, o:o + bn.C], aff[1, o:o + bn.C]

        if self.arch != "preact_resnet18":
            return self._forward_eval_fused_resnet(x_nchw, save, sl)
        ctx = {"x": x_nchw, "train": False, "fused": True, "blocks": []} if save else None
        blocks = self.blocks
        sc1 = sl(blocks[0]["bn1"])
        h, o1 = self.conv_first_fwd(x_nchw, self.conv1, bn_relu=sc1)
        for i, blk in enumerate(blocks):
            s = self.conv_fwd(o1, blk["sc"]) if "sc" in blk else h
            sc2 = sl(blk["bn2"])
            _, o2 = self.conv_fwd(o1, blk["conv1"], bn_relu=sc2, want_out=False)
            if save:
                ctx["blocks"].append((o1, o2, sc1[0], sc2[0]))
            if i + 1 < len(blocks):
                sc1 = sl(blocks[i + 1]["bn1"])
                # the float32 residual stream is only read by an identity shortcut
                h, o1 = self.conv_fwd(o2, blk["conv2"], residual=s, bn_relu=sc1, want_out="sc" not in blocks[i + 1])
            else:
                h = self.conv_fwd(o2, blk["conv2"], residual=s)
        logits, pooled = ops.pool_linear_fwd(h, 4, self.store.p("linear.weight"), self.store.p("linear.bias"))
        if save:
            ctx["feat_shape"], ctx["pooled"] = tuple(h.shape), pooled
        return logits, ctx

    def _unit_affine(self, Cc):
        u = self._unit.get(Cc)
        if u is None:
            u = self._unit[Cc] = (torch.ones(Cc, dtype=torch.float32, device=self.device),
                                  torch.zeros(Cc, dtype=torch.float32, device=self.device))
        return u

    def _forward_eval_fused_resnet(self, x_nchw, save, sl):
        """Eval-mode forward of the plain ResNet block ordering (classifier_models/resnet.py:22-32: relu(bn1(conv1(x))),
        bn2(conv2(.)) + shortcut(x), relu) with every BatchNorm folded into a tcgen05 epilogue: conv1 writes relu(bn1(.)) directly,
        the 1x1 shortcut conv writes its BatchNorm's output (post affine), conv2 computes (acc * scale2 + shift2) + shortcut and
        writes relu(.) -- three launches per block (two without a shortcut conv), no elementwise pass."""
        ctx = {"x": x_nchw, "train": False, "fused": True, "blocks": []} if save else None
        s0 = sl(self.bn1)
        _, h = self.conv_first_fwd(x_nchw, self.conv1, bn_relu=s0)                                  # resnet.py:86
        if save:
            ctx["stem"] = (h, s0[0])
        for blk in self.blocks:
            s1, s2 = sl(blk["bn1"]), sl(blk["bn2"])
            if "sc" in blk:
                ssc = sl(blk["scbn"])
                s = self.conv_fwd(h, blk["sc"], pre=False, post=ssc)
            else:
                ssc, s = None, h
            _, o1 = self.conv_fwd(h, blk["conv1"], bn_relu=s1, want_out=False)
            _, out = self.conv_fwd(o1, blk["conv2"], residual=s, post=s2, bn_relu=self._unit_affine(blk["conv2"].Cout), want_out=False)
            if save:
                ctx["blocks"].append((o1, out, s1[0], s2[0], ssc[0] if ssc is not None else None))
            h = out
        logits, pooled = ops.pool_linear_fwd(h, 4, self.store.p("linear.weight"), self.store.p("linear.bias"))
        if save:
            ctx["feat_shape"], ctx["pooled"] = tuple(h.shape), pooled
        return logits, ctx

    def _backward_eval_fused_resnet(self, ctx, dlogits, need_dx):
        """Input gradient of the forward above: per block ONE elementwise pass (relu mask of the block output, times bn2's scale;
        the unscaled masked gradient goes to the shortcut) and two tcgen05 input-gradient launches -- conv2's carries the
        relu(bn1(.)) backward in its epilogue, conv1's adds the identity-shortcut gradient or carries the 1x1 shortcut conv as
        an extra tap."""
        st = self.store
        dh = ops.pool_linear_bwd(dlogits, ctx["pooled"], st.p("linear.weight"), ctx["feat_shape"], self.dtype, 4, dW=None, db=None)
        prev_in = [ctx["stem"][0]] + [b[1] for b in ctx["blocks"][:-1]]   # input tensor of every block (for its spatial size)
        for blk, (o1, out, scale1, scale2, scale_sc), h_in in zip(reversed(self.blocks), reversed(ctx["blocks"]), reversed(prev_in)):
            hw_in, hw_mid = h_in.shape[1:3], o1.shape[1:3]
            d_c2, d_pre = ops.bn_bwd_eval(dh, out, scale2, True, None, True)
            d_c1 = self.conv_dgrad(d_c2, blk["conv2"], hw_mid, mask=o1, mask_scale=scale1)
            if "sc" in blk:
                zero = self._unit_affine(blk["sc"].Cout)[1]
                d_cs = ops.affine_act(d_pre, scale_sc, zero, False)
                dh = self.conv_dgrad(d_c1, blk["conv1"], hw_in, shortcut=(d_cs, blk["sc"]))
            else:
                dh = self.conv_dgrad(d_c1, blk["conv1"], hw_in, residual=d_pre)
        h0, scale0 = ctx["stem"]
        d_c0, _ = ops.bn_bwd_eval(dh, h0, scale0, True, None, False)
        if need_dx:
            return self.conv_first_dgrad(d_c0, self.conv1, ctx["x"].shape[2:4])
        return None

    @staticmethod
    def slice_ctx(ctx, lo, hi):
        """The saved state of rows [lo, hi) of a fused eval-mode forward (batch-major tensors: slices are contiguous views) --
        lets one forward over [x ; x_bd] serve a metric-only half and a half that is back-propagated."""
        if not ctx.get("fused"):
            raise RuntimeError("slice_ctx needs a fused eval-mode context")
        out = dict(ctx)
        out["x"] = ctx["x"][lo:hi]
        cut = lambda t: t[lo:hi] if (torch.is_tensor(t) and t.dim() == 4) else t   # activations; per-channel scales pass through
        out["blocks"] = [tuple(cut(t) for t in blk) for blk in ctx["blocks"]]
        if "stem" in ctx:
            out["stem"] = tuple(cut(t) for t in ctx["stem"])
        out["pooled"] = ctx["pooled"][lo:hi]
        out["feat_shape"] = (hi - lo,) + tuple(ctx["feat_shape"][1:])
        return out

    def _backward_eval_fused(self, ctx, dlogits, need_dx):
        if self.arch != "preact_resnet18":
            return self._backward_eval_fused_resnet(ctx, dlogits, need_dx)
        st = self.store
        dh = ops.pool_linear_bwd(dlogits, ctx["pooled"], st.p("linear.weight"), ctx["feat_shape"], self.dtype, 4, dW=None, db=None)
        for blk, (o1, o2, scale1, scale2) in zip(reversed(self.blocks), reversed(ctx["blocks"])):
            hw_in, hw_mid = o1.shape[1:3], o2.shape[1:3]
            d_c1 = self.conv_dgrad(dh, blk["conv2"], hw_mid, mask=o2, mask_scale=scale2)
            if "sc" in blk:
                dh = self.conv_dgrad(d_c1, blk["conv1"], hw_in, mask=o1, mask_scale=scale1, shortcut=(dh, blk["sc"]))
            else:
                dh = self.conv_dgrad(d_c1, blk["conv1"], hw_in, mask=o1, mask_scale=scale1, post_add=dh)
        if need_dx:
            return self.conv_first_dgrad(dh, self.conv1, ctx["x"].shape[2:4])
        return None

    # ---- backward
    def backward(self, ctx, dlogits, need_wgrad: bool, need_dx: bool):
        """Returns dx (NCHW float32) if need_dx.  Parameter gradients are accumulated into store.grad."""
        if ctx.get("fused"):
            if need_wgrad:
                raise RuntimeError("the fused eval-mode path saves no pre-normalisation tensors: weight gradients need train mode")
            return self._backward_eval_fused(ctx, dlogits, need_dx)
        pre = self.arch == "preact_resnet18"
        train = ctx["train"]
        st = self.store
        dh = ops.pool_linear_bwd(dlogits, ctx["pooled"], st.p("linear.weight"), ctx["feat_shape"], self.dtype, 4,
                                 dW=st.g("linear.weight") if need_wgrad else None,
                                 db=st.g("linear.bias") if need_wgrad else None)
        for blk, saved in zip(reversed(self.blocks), reversed(ctx["blocks"])):
            if pre:
                h_in, o1, c1, o2, st1, st2 = saved
                hw_in, hw_mid = h_in.shape[1:3], c1.shape[1:3]
                if need_wgrad:
                    self.conv_wgrad(o2, dh, blk["conv2"])
                if train and self.bnb_ok(blk["conv2"], c1):
                    g2 = self.conv_dgrad(dh, blk["conv2"], hw_mid, bnb=(c1, st2[0], st2[3], st2[1], st2[2]))
                    d_c1 = self._bn_bwd_tail(blk["bn2"], g2, c1, st2, need_wgrad)
                else:
                    d_o2 = self.conv_dgrad(dh, blk["conv2"], hw_mid)
                    d_c1, _ = self._bn_bwd(blk["bn2"], d_o2, c1, o2, st2, train, True, need_wgrad)
                if need_wgrad:
                    self.conv_wgrad(o1, d_c1, blk["conv1"])
                fuse1 = train and self.bnb_ok(blk["conv1"], h_in)
                bnb1 = (h_in, st1[0], st1[3], st1[1], st1[2]) if fuse1 else None
                if "sc" in blk:
                    if need_wgrad:
                        self.conv_wgrad(o1, dh, blk["sc"])
                    d_o1 = self.conv_dgrad(d_c1, blk["conv1"], hw_in, shortcut=(dh, blk["sc"]), bnb=bnb1)
                    dadd = None
                else:
                    d_o1 = self.conv_dgrad(d_c1, blk["conv1"], hw_in, bnb=bnb1)
                    dadd = dh
                if fuse1:
                    dh = self._bn_bwd_tail(blk["bn1"], d_o1, h_in, st1, need_wgrad, dadd=dadd)
                else:
                    dh, _ = self._bn_bwd(blk["bn1"], d_o1, h_in, o1, st1, train, True, need_wgrad, dadd=dadd)
            else:
                h_in, c1, o1, c2, out, cs_, st1, st2, sts = saved
                hw_in, hw_mid = h_in.shape[1:3], c1.shape[1:3]
                d_c2, dres = self._bn_bwd(blk["bn2"], dh, c2, out, st2, train, True, need_wgrad, want_dres=True)
                if need_wgrad:
                    self.conv_wgrad(o1, d_c2, blk["conv2"])
                if train and self.fuse_train_plain and self.bnb_ok(blk["conv2"], c1):   # [r2] bn1's backward sums in conv2's dgrad epilogue
                    g1 = self.conv_dgrad(d_c2, blk["conv2"], hw_mid, bnb=(c1, st1[0], st1[3], st1[1], st1[2]))
                    d_c1 = self._bn_bwd_tail(blk["bn1"], g1, c1, st1, need_wgrad)
                else:
                    d_o1 = self.conv_dgrad(d_c2, blk["conv2"], hw_mid)
                    d_c1, _ = self._bn_bwd(blk["bn1"], d_o1, c1, o1, st1, train, True, need_wgrad)
                if need_wgrad:
                    self.conv_wgrad(h_in, d_c1, blk["conv1"])
                if "sc" in blk:
                    d_cs, _ = self._bn_bwd(blk["scbn"], dres, cs_, None, sts, train, False, need_wgrad)
                    if need_wgrad:
                        self.conv_wgrad(h_in, d_cs, blk["sc"])
                    dh = self.conv_dgrad(d_c1, blk["conv1"], hw_in, shortcut=(d_cs, blk["sc"]))
                else:
                    dh = self.conv_dgrad(d_c1, blk["conv1"], hw_in, residual=dres)
        if not pre:
            c0, h0, st0 = ctx["stem"]
            dh, _ = self._bn_bwd(self.bn1, dh, c0, h0, st0, train, True, need_wgrad)
        x = ctx["x"]
        if need_wgrad:
            self.conv_first_wgrad(x, dh, self.conv1)
        if need_dx:
            return self.conv_first_dgrad(dh, self.conv1, x.shape[2:4])
        return None


# ===================================================================== trigger generator
class Generator(NetBase):
    """UnetGenerator (num_classes=0) / CUnetGeneratorv1 (num_classes>0: one-hot label planes after conv0_0)."""

    LAYERS = [("conv0_0", 2), ("conv0_1", 1), ("conv1_0", 2), ("conv1_1", 1), ("conv2_0", 2), ("conv2_1", 1),
              ("conv3_0", 2), ("conv3_1", 1), ("upconv3_1", 1), ("upconv3_0", 1), ("upconv2_1", 1), ("upconv2_0", 1),
              ("upconv1_1", 1), ("upconv1_0", 1), ("upconv0_1", 1), ("upconv0_0", 1)]

    def __init__(self, in_channels=3, nf=64, num_classes=0, out_channel=None, device="cuda", dtype=torch.bfloat16,
                 use_tc=True):
        super().__init__(device, dtype, use_tc)
        if out_channel is None:
            out_channel = in_channels
        self.nf, self.cond, self.in_channels, self.out_channel = nf, num_classes, in_channels, out_channel
        ch = {"conv0_0": (in_channels, nf), "conv0_1": (nf + num_classes, nf), "conv1_0": (nf, nf * 2),
              "conv1_1": (nf * 2, nf * 2), "conv2_0": (nf * 2, nf * 4), "conv2_1": (nf * 4, nf * 4),
              "conv3_0": (nf * 4, nf * 8), "conv3_1": (nf * 8, nf * 8), "upconv3_1": (nf * 8, nf * 8),
              "upconv3_0": (nf * 8, nf * 4), "upconv2_1": (nf * 4, nf * 4), "upconv2_0": (nf * 4, nf * 2),
              "upconv1_1": (nf * 2, nf * 2), "upconv1_0": (nf * 2, nf), "upconv0_1": (nf, nf),
              "upconv0_0": (nf, out_channel)}
        # CUnetGeneratorv1.conv0_1 reads nf + num_classes channels (72 for CelebA): not a multiple of the 64-channel K chunk of
        # the tcgen05 kernels, so it ran on the generic CUDA-core kernel (1.2 ms per launch, 8.5 ms of the 34.7 ms CelebA step,
        # profiles/r01_launches_celeba_multilabel_partial.md).  That weight is therefore STORED with its input channels padded
        # to the next multiple of 64 (zeros, see ParamStore) and the concatenated activation gets the same width.
        # COMBAT_NO_PAD_COND=1 restores the unpadded layout (A/B measurements).
        self.cond_pad = (not os.environ.get("COMBAT_NO_PAD_COND")) and self.use_tc and num_classes > 0 and (nf + num_classes) % 64 != 0
        specs, convs, store_ci = [], [], {}
        for name, stride in self.LAYERS:
            ci, co = ch[name]
            specs.append((name + ".weight", (co, ci, 3, 3)))
            specs.append((name + ".bias", (co,)))
            if name == "conv0_1" and self.cond_pad:
                store_ci[name + ".weight"] = _align(ci, 64)
            convs.append(ConvSpec(name, store_ci.get(name + ".weight", ci), co, 3, stride, 1, True, need_dgrad=(name != "conv0_0")))
        self._finish_params(specs, convs, store_ci)

    def load_state_dict(self, sd):
        self.store.load(sd)
        self.prep_weights()

    def state_dict(self):
        return self.store.state()

    def forward(self, x_nchw, labels=None, save=True):
        """x float32 NCHW -> tanh output float32 NCHW (networks/models.py:318-341; :523-555 with labels)."""
        cv = self.convs
        N, _, H, W = x_nchw.shape
        nf = self.nf
        ctx = {"x": x_nchw} if save else None
        c00 = self.conv_first_fwd(x_nchw, cv["conv0_0"], pre=False)
        if self.cond:  # cat(f0, one_hot planes) then the in-place LeakyReLU (identity on the 0/1 planes)
            if self.cond_pad:  # stored width of conv0_1's input; the padding planes stay zero
                a00 = torch.zeros((N, H // 2, W // 2, cv["conv0_1"].Cin), dtype=self.dtype, device=self.device)
            else:
                a00 = torch.empty((N, H // 2, W // 2, nf + self.cond), dtype=self.dtype, device=self.device)
            ops.lrelu_into_slice(c00, a00, 0)
            ops.onehot_planes(a00, labels, nf, self.cond)
        else:
            a00 = ops.leaky_relu(c00)
        acts = {"c00": c00, "a00": a00}

        def down(name, xin, act=True):
            c = self.conv_fwd(xin, cv[name])
            y, st = ops.instnorm_fwd(c, act, out_dtype=self.dtype)
            acts[name] = (xin, c, st)
            return y

        f0 = down("conv0_1", a00)
        f1 = down("conv1_1", down("conv1_0", f0))
        f2 = down("conv2_1", down("conv2_0", f1))
        f3 = down("conv3_1", down("conv3_0", f2), act=False)

        def up_block(n1, n0, xin, skip):
            t = ops.upsample2x_act(xin)
            c1 = self.conv_fwd(t, cv[n1])
            a1, st1 = ops.instnorm_fwd(c1, True, out_dtype=self.dtype)
            acts[n1] = (t, c1, st1)
            c0 = self.conv_fwd(a1, cv[n0])
            u, st0 = ops.instnorm_fwd(c0, False, skip=skip, out_dtype=self.dtype)
            acts[n0] = (a1, c0, st0)
            return u

        u3 = up_block("upconv3_1", "upconv3_0", f3, f2)
        u2 = up_block("upconv2_1", "upconv2_0", u3, f1)
        u1 = up_block("upconv1_1", "upconv1_0", u2, f0)
        t0 = ops.upsample2x_act(u1)
        c01 = self.conv_fwd(t0, cv["upconv0_1"])
        a01, st01 = ops.instnorm_fwd(c01, True, out_dtype=self.dtype)
        acts["upconv0_1"] = (t0, c01, st01)
        cs = cv["upconv0_0"]
        out = torch.empty((N, self.out_channel, H, W), dtype=torch.float32, device=self.device)
        small = self.fast_small and nf == 64 and self.out_channel == 3
        if small and self.tc_cout3 and self.use_tc and ops.conv_tc_cout3_ok(a01):
            ops.conv_tc_cout3(a01, self._wptr(cs), out, bias=self._bias(cs), act=1)
        elif small:
            ops.conv_cout3(a01, self._wptr(cs), self.dt, out, bias=self._bias(cs), act=1)
        else:
            ops.conv_simt(a01, (N, H, W), ops.nhwc_strides(H, W, nf), self._wptr(cs), self.dt, out, (H, W),
                          ops.nchw_strides(self.out_channel, H, W), Ci=nf, Co=self.out_channel, KH=3, KW=3, stride=1, pad=1,
                          bias=self._bias(cs), act=1)
        if save:
            acts["a01"] = a01
            ctx["acts"], ctx["out"] = acts, out
        return out, ctx

    def backward(self, ctx, dout):
        """dout: gradient w.r.t. the tanh output (NCHW float32).  Accumulates all parameter gradients."""
        cv, acts = self.convs, ctx["acts"]
        x = ctx["x"]
        N, _, H, W = x.shape
        nf = self.nf
        dz = ops.tanh_bwd(dout, ctx["out"])
        cs = cv["upconv0_0"]
        a01 = acts["a01"]
        nchw_o = ops.nchw_strides(self.out_channel, H, W)
        d_a01 = torch.empty((N, H, W, nf), dtype=self.dtype, device=self.device)
        if self.fast_small and nf == 64 and self.out_channel == 3:
            if a01.dtype == torch.bfloat16 and self._use_im2col(cs, nf, 3, wgrad=True):
                # weight gradient of the 64 -> 3 conv on the tensor cores: dW'[ci][k] = sum_pix a[pix][ci] * im2col3(dz)[pix][k]
                A = ops.im2col3(dz, 1)
                dw64 = torch.zeros((nf, 64), dtype=torch.float32, device=self.device)
                d = ops.conv_tc_desc(A, None, None, N, H, W, 64, H, W, nf, 1, 1, 1, 0, 1)
                _tc_launch(lambda: lib.combat_conv_tc_wgrad(C.byref(d), ops._p(a01), ops._p(dw64), ops._s()), "conv_tc_boundary",
                           2.0 * N * H * W * nf * 27, "N%d %dx%d %d->3 k3 s1 im2col wgrad" % (N, H, W, nf))
                csum = torch.zeros(64, dtype=torch.float32, device=self.device)
                ops.colsum(A.view(-1, 64), 64, csum)
                ops.fold_w64(dw64, self.store.raw(self.store.grad, cs.name + ".weight"), nf, 1, colsum=csum,
                             db=self.store.g(cs.name + ".bias"))
            else:
                ops.wgrad_cout3(a01, dz, self.store.raw(self.store.grad, cs.name + ".weight"), self.store.g(cs.name + ".bias"))
            if self.tc_first2 and self.use_tc and nf == 64 and W in (16, 32, 64, 128) and H * W >= 128:
                # input gradient of the 64 -> 3 conv == 3 -> 64 conv of dz with the flipped filter: conv_tc_first_kernel
                d = ops.conv_tc_desc(dz, self._w64_for(cs, dgrad=True).data_ptr(), d_a01, N, H, W, 64, H, W, 64, 3, 3, 1, 1, 1,
                                     in_nchw3=True)
                _tc_launch(lambda: lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc_boundary", 2.0 * N * H * W * 64 * 27,
                           "N%d %dx%d 3->64 k3 s1 first (dgrad of 64->3)" % (N, H, W))
            elif self._use_im2col(cs, nf, 3):  # input gradient of the 64 -> 3 conv == 3 -> 64 conv with the flipped filter
                self.cin3_tc(dz, self._w64_for(cs, dgrad=True), nf, 1, d_a01, tag=" (dgrad)")
            else:
                ops.conv_cin3(dz, self._wptr(cs, True), self.dt, d_a01, nf, 1)
        else:
            ops.conv_wgrad_simt(a01, (N, H, W), ops.nhwc_strides(H, W, nf), dz, (H, W), nchw_o,
                                self.store.raw(self.store.grad, cs.name + ".weight"),
                                Ci=nf, Co=self.out_channel, KH=3, KW=3, stride=1, pad=1, db=self.store.g(cs.name + ".bias"))
            ops.conv_simt(dz, (N, H, W), nchw_o, self._wptr(cs, True), self.dt, d_a01, (H, W), ops.nhwc_strides(H, W, nf),
                          Ci=self.out_channel, Co=nf, KH=3, KW=3, stride=1, pad=1, up=1)

        def conv_in_bwd(name, dy1, dy2, act, n_out_ch=None, want_dx=True):
            """backward through  y = [lrelu](IN(conv(xin)))  given dL/dy = dy1 (+ dy2)."""
            xin, c, st = acts[name]
            d_c = ops.instnorm_bwd(dy1, dy2, c, st, act)
            self.conv_wgrad(xin, d_c, cv[name], dead_bias=True)
            if not want_dx:
                return None
            return self.conv_dgrad(d_c, cv[name], xin.shape[1:3], n_out_ch=n_out_ch)

        d_t0 = conv_in_bwd("upconv0_1", d_a01, None, True)
        d_u1 = ops.upsample2x_act_bwd(d_t0, acts["upconv0_1"][0])

        def up_block_bwd(n1, n0, d_u):
            d_a1 = conv_in_bwd(n0, d_u, None, False)
            d_t = conv_in_bwd(n1, d_a1, None, True)
            return ops.upsample2x_act_bwd(d_t, acts[n1][0])

        d_u2 = up_block_bwd("upconv1_1", "upconv1_0", d_u1)  # d_u1 is also the skip gradient of f0
        d_u3 = up_block_bwd("upconv2_1", "upconv2_0", d_u2)  # d_u2 ... of f1
        d_f3 = up_block_bwd("upconv3_1", "upconv3_0", d_u3)  # d_u3 ... of f2
        d_a30 = conv_in_bwd("conv3_1", d_f3, None, False)
        d_f2 = conv_in_bwd("conv3_0", d_a30, None, True)
        d_a20 = conv_in_bwd("conv2_1", d_f2, d_u3, True)
        d_f1 = conv_in_bwd("conv2_0", d_a20, None, True)
        d_a10 = conv_in_bwd("conv1_1", d_f1, d_u2, True)
        d_f0 = conv_in_bwd("conv1_0", d_a10, None, True)
        d_a00 = conv_in_bwd("conv0_1", d_f0, d_u1, True, n_out_ch=nf if self.cond else None)
        d_c00 = ops.leaky_relu_bwd(d_a00, acts["c00"])
        self.conv_first_wgrad(x, d_c00, cv["conv0_0"])


class GridGenerator(Generator):
    """GridGenerator of the WaNet variant (networks/models.py:344-385): the ENCODER half of UnetGenerator (same kernels: first conv
    on the 3-channel path, tcgen05 convs, InstanceNorm + LeakyReLU), global average pool + fc1 (one pool_linear launch),
    LeakyReLU, fc2, tanh -> the S x S control grid of a flow, [N, 2, S, S] float32.  Subclass of Generator for the module
    plumbing only (forward(x, labels, save) -> (y, ctx), backward(ctx, dy))."""

    LAYERS = [("conv0_0", 2), ("conv0_1", 1), ("conv1_0", 2), ("conv1_1", 1), ("conv2_0", 2), ("conv2_1", 1), ("conv3_0", 2),
              ("conv3_1", 1)]

    def __init__(self, in_channels=3, nf=64, S=2, device="cuda", dtype=torch.bfloat16, use_tc=True):
        NetBase.__init__(self, device, dtype, use_tc)
        if not 1 <= S <= 4:
            raise ValueError("GridGenerator: control grids up to 4 x 4 are supported (--s %d)" % S)
        self.nf, self.cond, self.in_channels, self.out_channel, self.S = nf, 0, in_channels, 2, S
        self.cond_pad = False
        ch = {"conv0_0": (in_channels, nf), "conv0_1": (nf, nf), "conv1_0": (nf, nf * 2), "conv1_1": (nf * 2, nf * 2),
              "conv2_0": (nf * 2, nf * 4), "conv2_1": (nf * 4, nf * 4), "conv3_0": (nf * 4, nf * 8), "conv3_1": (nf * 8, nf * 8)}
        specs, convs = [], []
        for name, stride in self.LAYERS:
            ci, co = ch[name]
            specs.append((name + ".weight", (co, ci, 3, 3)))
            specs.append((name + ".bias", (co,)))
            convs.append(ConvSpec(name, ci, co, 3, stride, 1, True, need_dgrad=(name != "conv0_0")))
        specs += [("fc1.weight", (nf, nf * 8)), ("fc1.bias", (nf,)), ("fc2.weight", (S * S * 2, nf)), ("fc2.bias", (S * S * 2,))]
        self._finish_params(specs, convs)

    def forward(self, x_nchw, labels=None, save=True):
        cv = self.convs
        N, _, H, W = x_nchw.shape
        ctx = {"x": x_nchw} if save else None
        c00 = self.conv_first_fwd(x_nchw, cv["conv0_0"], pre=False)
        a00 = ops.leaky_relu(c00)
        acts = {"c00": c00, "a00": a00}

        def down(name, xin, act=True):
            c = self.conv_fwd(xin, cv[name])
            y, st = ops.instnorm_fwd(c, act, out_dtype=self.dtype)
            acts[name] = (xin, c, st)
            return y

        f0 = down("conv0_1", a00)
        f1 = down("conv1_1", down("conv1_0", f0))
        f2 = down("conv2_1", down("conv2_0", f1))
        f3 = down("conv3_1", down("conv3_0", f2), act=False)
        if f3.shape[1] != f3.shape[2]:
            raise ValueError("GridGenerator: square images only")
        st = self.store
        h1, pooled = ops.pool_linear_fwd(f3, f3.shape[1], st.p("fc1.weight"), st.p("fc1.bias"))      # :381-382
        a1 = ops.leaky_relu(h1)
        z, _ = ops.pool_linear_fwd(a1.view(N, 1, 1, self.nf), 1, st.p("fc2.weight"), st.p("fc2.bias"))   # :383
        out = ops.tanh_fwd(z).view(N, 2, self.S, self.S)                                            # :383-384
        if save:
            ctx.update(acts=acts, out=out, f3_shape=tuple(f3.shape), pooled=pooled, h1=h1, a1=a1)
        return out, ctx

    def backward(self, ctx, dout):
        """dout: gradient w.r.t. the tanh output [N, 2, S, S] (float32).  Accumulates all parameter gradients."""
        cv, acts = self.convs, ctx["acts"]
        x = ctx["x"]
        N = x.shape[0]
        st = self.store
        dz = ops.tanh_bwd(dout.contiguous().view(N, -1), ctx["out"].view(N, -1))
        da1 = ops.pool_linear_bwd(dz, ctx["a1"], st.p("fc2.weight"), (N, 1, 1, self.nf), torch.float32, 1,
                                  dW=st.g("fc2.weight"), db=st.g("fc2.bias"))
        dh1 = ops.leaky_relu_bwd(da1.view(N, self.nf), ctx["h1"])
        d_f3 = ops.pool_linear_bwd(dh1, ctx["pooled"], st.p("fc1.weight"), ctx["f3_shape"], self.dtype, ctx["f3_shape"][1],
                                   dW=st.g("fc1.weight"), db=st.g("fc1.bias"))

        def conv_in_bwd(name, dy1, act):
            xin, c, stn = acts[name]
            d_c = ops.instnorm_bwd(dy1, None, c, stn, act)
            self.conv_wgrad(xin, d_c, cv[name], dead_bias=True)
            return self.conv_dgrad(d_c, cv[name], xin.shape[1:3])

        d = conv_in_bwd("conv3_1", d_f3, False)
        for name in ("conv3_0", "conv2_1", "conv2_0", "conv1_1", "conv1_0", "conv0_1"):
            d = conv_in_bwd(name, d, True)
        d_c00 = ops.leaky_relu_bwd(d, acts["c00"])
        self.conv_first_wgrad(x, d_c00, cv["conv0_0"])


# ===================================================================== frequency detector (forward only)
class FrequencyDetector(NetBase):
    """FrequencyModel in eval mode: conv -> ELU -> BN(eval) (x6), maxpool after 2/4/6, flatten (NCHW order), linear
    (defenses/frequency_based/model.py:8-52).  Metrics leg only (train_generator.py:245-247).

    dtype float32: CUDA-core kernels throughout.  dtype bfloat16: conv1 (3 input channels; its input are raw DCT
    coefficients of magnitude up to ~8e3) stays float32 arithmetic with float32 weights, conv2..conv6 run on the
    tcgen05 kernel with ELU + folded BatchNorm in its epilogue; channel counts below 64 are zero-padded to 64 once,
    when the (frozen) weights are loaded."""

    def __init__(self, num_classes=2, n_input=3, input_size=32, device="cuda", dtype=torch.float32, trainable=False):
        super().__init__(device, torch.float32, use_tc=False)
        self.trainable = bool(trainable)  # train_forward / train_backward / adadelta_step (float32 CUDA-core path)
        self.act_dtype = dtype
        self.tc = dtype == torch.bfloat16
        self.scaler = {32: 1, 64: 4}[input_size]
        chans = [n_input, 32, 32, 64, 64, 128, 128]
        specs, convs = [], []
        self.chans = chans
        for i in range(1, 7):
            specs.append(("conv%d.weight" % i, (chans[i], chans[i - 1], 3, 3)))
            specs.append(("conv%d.bias" % i, (chans[i],)))
            specs.append(("bn%d.weight" % i, (chans[i],)))
            specs.append(("bn%d.bias" % i, (chans[i],)))
            convs.append(ConvSpec("conv%d" % i, chans[i - 1], chans[i], 3, 1, 1, True, need_dgrad=self.trainable and i > 1))
        specs.append(("linear6.weight", (num_classes, 2048 * self.scaler)))
        specs.append(("linear6.bias", (num_classes,)))
        self._finish_params(specs, convs)
        self.rm = {i: torch.zeros(chans[i], device=self.device) for i in range(1, 7)}
        self.rv = {i: torch.ones(chans[i], device=self.device) for i in range(1, 7)}
        for i in range(1, 7):
            self.store.p("bn%d.weight" % i).fill_(1.0)
        self._affine = None
        self._padded = None

    def load_state_dict(self, sd):
        self.store.load(sd)
        for i in range(1, 7):
            self.rm[i].copy_(sd["bn%d.running_mean" % i])
            self.rv[i].copy_(sd["bn%d.running_var" % i])
        self.prep_weights()
        self._affine = None
        self._padded = None

    def state_dict(self):
        sd = self.store.state()
        for i in range(1, 7):
            sd["bn%d.running_mean" % i] = self.rm[i].clone()
            sd["bn%d.running_var" % i] = self.rv[i].clone()
            sd["bn%d.num_batches_tracked" % i] = torch.tensor(0)
        return sd

    def _prepare(self):
        """frozen network: fold BN(eval) into per-channel (scale, shift) once; build the padded bf16 weights."""
        if self._affine is None:
            self._affine = {i: ops.bn_eval_prepare(self.chans[i], self.store.p("bn%d.weight" % i), self.store.p("bn%d.bias" % i),
                                                   self.rm[i], self.rv[i], 1e-5) for i in range(1, 7)}
        if self.tc and self._padded is None:
            pad = lambda c: max(64, c)
            P = {}
            for i in range(1, 7):
                ci, co = self.chans[i - 1], self.chans[i]
                cip, cop = (ci if i == 1 else pad(ci)), pad(co)
                w = self.store.raw(self.store.flat, "conv%d.weight" % i).view(co, 9, ci)  # channels-last master
                wp = torch.zeros((cop, 9, cip), dtype=torch.float32 if i == 1 else torch.bfloat16, device=self.device)
                wp[:co, :, :ci] = w.to(wp.dtype)
                vec = torch.zeros((3, cop), dtype=torch.float32, device=self.device)  # bias, scale, shift (0 on the padding)
                vec[0, :co] = self.store.p("conv%d.bias" % i)
                vec[1, :co], vec[2, :co] = self._affine[i]
                P[i] = (wp, vec, cip, cop)
            self._padded = P

    def forward(self, x_nchw):
        """x: float32 NCHW DCT coefficients -> logits [N, num_classes]."""
        self._prepare()
        N, Cc, H, W = x_nchw.shape
        if self.tc and Cc == 3:
            return self._forward_tc(x_nchw)
        h, strides, hw = x_nchw, ops.nchw_strides(Cc, H, W), (H, W)
        for i in range(1, 7):
            cs = self.convs["conv%d" % i]
            out = torch.empty((N, hw[0], hw[1], cs.Cout), dtype=torch.float32, device=self.device)
            sc, sh = self._affine[i]
            if i == 1 and Cc == 3 and self.fast_small:
                ops.conv_cin3(h, self._wptr(cs), self.dt, out, cs.Cout, 1, bias=self._bias(cs), act=2, post_scale=sc, post_shift=sh)
            else:
                ops.conv_simt(h, (N, hw[0], hw[1]), strides, self._wptr(cs), self.dt, out, hw,
                              ops.nhwc_strides(hw[0], hw[1], cs.Cout), Ci=cs.Cin, Co=cs.Cout, KH=3, KW=3, stride=1, pad=1,
                              bias=self._bias(cs), act=2, post_scale=sc, post_shift=sh)
            h = out
            if i % 2 == 0:
                h = ops.maxpool2(h)
                hw = (hw[0] // 2, hw[1] // 2)
            strides = ops.nhwc_strides(hw[0], hw[1], cs.Cout)
        # flatten in NCHW order + linear == pool_linear with P = 1
        logits, _ = ops.pool_linear_fwd(h, 1, self.store.p("linear6.weight"), self.store.p("linear6.bias"))
        return logits

    def _forward_tc(self, x_nchw):
        N, _, H, W = x_nchw.shape
        hw = (H, W)
        wp, vec, _, cop = self._padded[1]
        h = torch.empty((N, H, W, cop), dtype=torch.bfloat16, device=self.device)
        ops.conv_cin3(x_nchw, wp.data_ptr(), ops.F32, h, cop, 1, bias=vec[0], act=2, post_scale=vec[1], post_shift=vec[2])
        for i in range(2, 7):
            wp, vec, cip, cop = self._padded[i]
            out = torch.empty((N, hw[0], hw[1], cop), dtype=torch.bfloat16, device=self.device)
            d = ops.conv_tc_desc(h, wp.data_ptr(), out, N, hw[0], hw[1], cip, hw[0], hw[1], cop, 3, 3, 1, 1, 1, bias=vec[0],
                                 act=2, post_scale=vec[1], post_shift=vec[2])
            _tc_launch(lambda: lib.combat_conv_tc(C.byref(d), ops._s()), "conv_tc", 2.0 * N * hw[0] * hw[1] * cop * cip * 9,
                       "N%d %dx%d %d->%d k3 s1 (netF)" % (N, hw[0], hw[1], cip, cop))
            h = out
            if i % 2 == 0:
                h = ops.maxpool2(h)
                hw = (hw[0] // 2, hw[1] // 2)
        logits, _ = ops.pool_linear_fwd(h, 1, self.store.p("linear6.weight"), self.store.p("linear6.bias"))
        return logits

    # ------------------------------------------------------------------ training
    # defenses/frequency_based/train.py:178-221 on the float32 CUDA-core kernels: a first CORRECT path (generic strided
    # convolutions), to be held to oracle/detector_oracle.py before any of it moves to the tensor cores.  Written at the close
    # of round 1, not yet run on a GPU; nothing in the alternated step calls it.
    DROPOUT_P = 0.2

    def draw_dropout_masks(self, N, H, W):
        """The three Dropout(0.2) keep masks of one training forward (after max-pools 1..3), drawn on the HOST from the
        torch CPU generator with the shapes and order torch's own dropout would use on the NCHW activations
        (model.py:22,33,44), returned as uint8 NHWC device tensors."""
        masks, hw = [], (H, W)
        for i in (2, 4, 6):
            hw = (hw[0] // 2, hw[1] // 2)
            keep = torch.nn.functional.dropout(torch.ones(N, self.chans[i], hw[0], hw[1]), self.DROPOUT_P, True) != 0
            masks.append(keep.permute(0, 2, 3, 1).contiguous().to(torch.uint8).to(self.device, non_blocking=True))
        return masks

    def train_forward(self, x_nchw, masks, momentum=0.1, eps=1e-5):
        """x: float32 NCHW DCT coefficients; masks: draw_dropout_masks(...).  conv -> ELU -> BatchNorm(batch statistics, running
        statistics updated) x6, MaxPool2d(2) + Dropout after layers 2 / 4 / 6, flatten (NCHW order), linear.
        Returns (logits [N, num_classes], ctx for train_backward)."""
        if not self.trainable or self.tc:
            raise RuntimeError("FrequencyDetector(trainable=True, dtype=torch.float32) is the training configuration")
        N, Cc, H, W = x_nchw.shape
        h, strides, hw = ops._contig(x_nchw), ops.nchw_strides(Cc, H, W), (H, W)   # raw-pointer kernels: dense NCHW only
        layers = []
        for i in range(1, 7):
            cs = self.convs["conv%d" % i]
            a = torch.empty((N, hw[0], hw[1], cs.Cout), dtype=torch.float32, device=self.device)
            ops.conv_simt(h, (N, hw[0], hw[1]), strides, self._wptr(cs), self.dt, a, hw, ops.nhwc_strides(hw[0], hw[1], cs.Cout),
                          Ci=cs.Cin, Co=cs.Cout, KH=3, KW=3, stride=1, pad=1, bias=self._bias(cs), act=2)   # a = elu(conv + bias)
            sc, sh, mean, invstd = ops.bn_train_prepare(a, N * hw[0] * hw[1], cs.Cout, self.store.p("bn%d.weight" % i),
                                                        self.store.p("bn%d.bias" % i), self.rm[i], self.rv[i], momentum, eps)
            y = ops.affine_act(a, sc, sh, False)
            rec = {"x": h, "x_strides": strides, "hw": hw, "a": a, "mean": mean, "invstd": invstd}
            h = y
            if i % 2 == 0:
                pooled = ops.maxpool2(y)
                rec["y"] = y
                hw = (hw[0] // 2, hw[1] // 2)
                h = ops.mask_scale(pooled, masks[i // 2 - 1], 1.0 / (1.0 - self.DROPOUT_P))
            layers.append(rec)
            strides = ops.nhwc_strides(hw[0], hw[1], cs.Cout)
        logits, feat = ops.pool_linear_fwd(h, 1, self.store.p("linear6.weight"), self.store.p("linear6.bias"))
        self._affine = None  # running statistics changed: the cached eval-mode affine is stale
        return logits, {"layers": layers, "masks": masks, "feat": feat, "h_shape": tuple(h.shape), "N": N}

    def train_backward(self, ctx, dlogits):
        """Accumulates every parameter gradient into the flat gradient buffer (call zero_grad() first)."""
        st = self.store
        d = ops.pool_linear_bwd(dlogits, ctx["feat"], st.p("linear6.weight"), ctx["h_shape"], torch.float32, 1,
                                dW=st.g("linear6.weight"), db=st.g("linear6.bias"))
        N = ctx["N"]
        for i in range(6, 0, -1):
            cs, rec = self.convs["conv%d" % i], ctx["layers"][i - 1]
            if i % 2 == 0:
                d = ops.mask_scale(d, ctx["masks"][i // 2 - 1], 1.0 / (1.0 - self.DROPOUT_P))
                d = ops.maxpool2_bwd(d, rec["y"])
            da, _ = ops.bn_bwd_train(d, rec["a"], None, st.p("bn%d.weight" % i), rec["mean"], rec["invstd"], False,
                                     st.g("bn%d.weight" % i), st.g("bn%d.bias" % i))
            dz = ops.elu_bwd(da, rec["a"])
            H, W = rec["hw"]
            ops.conv_wgrad_simt(rec["x"], (N, H, W), rec["x_strides"], dz, (H, W), ops.nhwc_strides(H, W, cs.Cout),
                                st.raw(st.grad, cs.name + ".weight"), Ci=cs.Cin, Co=cs.Cout, KH=3, KW=3, stride=1, pad=1,
                                db=st.g(cs.name + ".bias"))
            if i > 1:
                d = self.conv_dgrad(dz, cs, (H, W))

    def adadelta_step(self, lr_dev, rho=0.9, eps=1e-6, wd=1e-4):
        """torch.optim.Adadelta(lr=0.05, weight_decay=1e-4) of train.py:152 over the flat buffer (square_avg lives in the
        store's momentum buffer, acc_delta in a buffer allocated on first use), then the compute-layout weights."""
        st = self.store
        if not hasattr(st, "acc_delta"):
            st.acc_delta = torch.zeros_like(st.flat)
        ops.adadelta(st.flat, st.grad, st.mom, st.acc_delta, lr_dev, rho, eps, wd)
        self.prep_weights()
        self._padded = None

