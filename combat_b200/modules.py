"""nn.Module drop-ins with the reference's constructors, attribute paths and state_dict keys, backed by the explicit
kernel graphs of combat_b200.nets.

Each module owns a `nets.*` object; its nn.Parameters / buffers are VIEWS into that object's flat parameter store, so
`state_dict()` / `load_state_dict()` / `torch.optim.SGD` / checkpoint files keep working with the reference's key
names and shapes (conv weights are torch channels_last views).  `forward` is ONE autograd node per network whose
backward runs the hand-written backward graph -- no aten arithmetic in between.  CUDA only: there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import nets, ops


def default_dtype():
    return torch.bfloat16


class _NetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, labels, need_grad, *params):
        # need_grad is decided by the caller: grad mode is always OFF inside autograd.Function.forward
        net = mod.net
        net.prep_weights()  # parameters may have been updated in place by an external optimiser
        if isinstance(net, nets.Generator):
            y, c = net.forward(x.contiguous().float(), labels, save=need_grad)
        elif isinstance(net, nets.FrequencyDetector):
            y, c = net.forward(x.contiguous().float()), None
        else:
            # an eval-mode forward whose weight gradients are wanted (the reference computes and discards them) keeps the
            # unfused path; inference and input-gradient-only uses take the fused one
            y, c = net.forward(x.contiguous().float(), train=mod.training, save=need_grad,
                               fuse=not any(p.requires_grad for p in params))
        ctx.mod, ctx.c = mod, c
        ctx.wgrad = any(p.requires_grad for p in params)
        ctx.xgrad = x.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        mod, net = ctx.mod, ctx.mod.net
        if ctx.c is None:
            raise RuntimeError("combat_b200: backward through a forward that saved nothing")
        dy = dy.contiguous().float()
        net.zero_grad()
        if isinstance(net, nets.Generator):
            net.backward(ctx.c, dy)
            dx = None  # the reference never differentiates the generator w.r.t. its input image
        else:
            dx = net.backward(ctx.c, dy, need_wgrad=ctx.wgrad, need_dx=ctx.xgrad)
        grads = [net.store.g(n).clone() if (p.requires_grad and ctx.wgrad) else None for n, p in mod._plist]
        return (None, dx, None, None, *grads)


def _run(mod, x, labels):
    params = [p for _, p in mod._plist]
    need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
    return _NetFn.apply(mod, x, labels, need_grad, *params)


class _KernelModule(nn.Module):
    """Builds the nested attribute tree (layer1.0.bn1.weight ...) over views of the flat store."""

    def _bind(self, net, buffers=()):
        object.__setattr__(self, "net", net)
        plist = []
        for name in net.store.names:
            p = nn.Parameter(net.store.p(name), requires_grad=True)
            self._attach(name, p, is_param=True)
            plist.append((name, p))
        for name, t in buffers:
            self._attach(name, t, is_param=False)
        object.__setattr__(self, "_plist", plist)

    def _attach(self, dotted, tensor, is_param):
        parts = dotted.split(".")
        m = self
        for i, part in enumerate(parts[:-1]):
            if part not in m._modules:
                nxt = nn.Sequential() if (i + 1 < len(parts) - 1 and parts[i + 1].isdigit()) else nn.Module()
                m.add_module(part, nxt)
            m = m._modules[part]
        if is_param:
            m.register_parameter(parts[-1], tensor)
        else:
            m.register_buffer(parts[-1], tensor)

    def _apply(self, fn, recurse=True):
        # parameters are views into device buffers consumed by raw-pointer kernels: moving / casting them would
        # silently detach the module from its kernels
        probe = fn(torch.empty(0, device=self.net.device))
        if probe.device != self.net.device or probe.dtype != torch.float32:
            raise RuntimeError("combat_b200 modules live on %s in float32 master precision; .to()/.half()/.cpu() are not supported"
                               % self.net.device)
        return self

    def load_state_dict(self, state_dict, strict=True):
        out = super().load_state_dict(state_dict, strict)
        if hasattr(self.net, "num_batches_tracked"):
            for k, v in state_dict.items():
                if k.endswith("num_batches_tracked"):
                    self.net.num_batches_tracked[k[: -len(".num_batches_tracked")]] = int(v)
        self.net.prep_weights()
        if hasattr(self.net, "_affine"):
            self.net._affine, self.net._padded = None, None
        return out


class ClassifierModule(_KernelModule):
    def __init__(self, arch, num_classes, n_input, input_size, device=None, dtype=None):
        super().__init__()
        device = torch.device(device or "cuda")
        net = nets.Classifier(arch, num_classes, n_input, input_size, device=device, dtype=dtype or default_dtype())
        bufs = []
        for bn in net.bns:
            bufs += [(bn.name + ".running_mean", net.rm(bn)), (bn.name + ".running_var", net.rv(bn)),
                     (bn.name + ".num_batches_tracked", torch.zeros((), dtype=torch.long, device=device))]
        self._bind(net, bufs)
        self._init_like_torch()
        # attribute names other reference scripts rely on (defenses/fine_pruning: netC.layer4[1].conv2, .ind, .linear)
        for blk in [m for n, m in self.named_modules() if n.count(".") == 1 and n.startswith("layer")]:
            blk.ind = None

    def _init_like_torch(self):
        """nn.Conv2d / nn.Linear default initialisation, drawn from the global torch CPU generator in the reference's
        module construction order (so that `torch.manual_seed(s); PreActResNet18()` matches the reference)."""
        import math
        gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
        st = self.net.store
        for name in st.names:
            shape = st.shapes[name]
            if name.endswith(".weight") and len(shape) == 4:
                fan_in = shape[1] * shape[2] * shape[3]
                bound = math.sqrt(3.0) * gain / math.sqrt(fan_in)
                st.p(name).copy_(torch.empty(shape).uniform_(-bound, bound))
            elif name == "linear.weight":
                bound = math.sqrt(3.0) * gain / math.sqrt(shape[1])
                st.p(name).copy_(torch.empty(shape).uniform_(-bound, bound))
                st.p("linear.bias").copy_(torch.empty(shape[0]).uniform_(-1 / math.sqrt(shape[1]), 1 / math.sqrt(shape[1])))
        self.net.prep_weights()

    def forward(self, x):
        y = _run(self, x, None)
        if self.training:
            for n, b in self.named_buffers():
                if n.endswith("num_batches_tracked"):
                    b += 1
        return y


def PreActResNet18(num_classes=10, n_input=3, input_size=32, **kw):
    """classifier_models/preact_resnet.py:108 (input_size -> scaler: {32: 1, 64: 4}; KeyError otherwise)."""
    {32: 1, 64: 4}[input_size]
    return ClassifierModule("preact_resnet18", num_classes, n_input, input_size, **kw)


def ResNet18(num_classes=10, n_input=3, input_size=64, **kw):
    """classifier_models/resnet.py:104.  input_size 224 (KeyError in the reference) uses the natural scaler 49."""
    {32: 1, 64: 4, 224: 49}[input_size]
    return ClassifierModule("resnet18", num_classes, n_input, input_size, **kw)


class _GeneratorModule(_KernelModule):
    def __init__(self, opt, in_channels=3, nf=64, use_bias=True, out_channel=None, cond=0, device=None, dtype=None):
        super().__init__()
        if not use_bias:
            raise NotImplementedError("use_bias=False is not used on the hot path")
        dev = torch.device(device or getattr(opt, "device", None) or "cuda")
        if dev.type != "cuda":
            raise RuntimeError("combat_b200 modules are CUDA only")
        net = nets.Generator(in_channels, nf, cond, out_channel, device=dev, dtype=dtype or default_dtype())
        self._bind(net)
        import math
        gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
        st = net.store
        for name, _ in nets.Generator.LAYERS:  # construction order of the reference (networks/models.py:275-314)
            shape = st.shapes[name + ".weight"]
            fan_in = shape[1] * 9
            st.p(name + ".weight").copy_(torch.empty(shape).uniform_(-math.sqrt(3.0) * gain / math.sqrt(fan_in),
                                                                       math.sqrt(3.0) * gain / math.sqrt(fan_in)))
            st.p(name + ".bias").copy_(torch.empty(shape[0]).uniform_(-1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in)))
        net.prep_weights()


class UnetGenerator(_GeneratorModule):
    """networks/models.py:268-341"""

    def __init__(self, opt, in_channels=3, nf=64, use_bias=True, out_channel=None, **kw):
        super().__init__(opt, in_channels, nf, use_bias, out_channel, 0, **kw)

    def forward(self, x):
        if x.shape[0] == 0:
            return x.new_empty(x.shape)
        return _run(self, x, None)


class CUnetGeneratorv1(_GeneratorModule):
    """networks/models.py:472-555"""

    def __init__(self, opt, in_channels=3, nf=64, use_bias=True, out_channel=None, **kw):
        self.num_classes = opt.num_classes
        super().__init__(opt, in_channels, nf, use_bias, out_channel, opt.num_classes, **kw)

    def forward(self, x, y):
        return _run(self, x, y.contiguous())


class GridGenerator(_KernelModule):
    """networks/models.py:344-385 (the WaNet variant's flow generator; `opt.s` = side of the control grid)."""

    def __init__(self, opt, in_channels=3, nf=64, use_bias=True, device=None, dtype=None):
        super().__init__()
        if not use_bias:
            raise NotImplementedError("use_bias=False is not used on the hot path")
        dev = torch.device(device or getattr(opt, "device", None) or "cuda")
        if dev.type != "cuda":
            raise RuntimeError("combat_b200 modules are CUDA only")
        self.S = opt.s
        net = nets.GridGenerator(in_channels, nf, opt.s, device=dev, dtype=dtype or default_dtype())
        self._bind(net)
        import math
        gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
        st = net.store
        for name, _ in nets.GridGenerator.LAYERS:  # construction order of the reference (:350-368): convs, then fc1, fc2
            shape = st.shapes[name + ".weight"]
            fan_in = shape[1] * 9
            b = math.sqrt(3.0) * gain / math.sqrt(fan_in)
            st.p(name + ".weight").copy_(torch.empty(shape).uniform_(-b, b))
            st.p(name + ".bias").copy_(torch.empty(shape[0]).uniform_(-1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in)))
        for name in ("fc1", "fc2"):
            shape = st.shapes[name + ".weight"]
            b = math.sqrt(3.0) * gain / math.sqrt(shape[1])
            st.p(name + ".weight").copy_(torch.empty(shape).uniform_(-b, b))
            st.p(name + ".bias").copy_(torch.empty(shape[0]).uniform_(-1 / math.sqrt(shape[1]), 1 / math.sqrt(shape[1])))
        net.prep_weights()

    def forward(self, x):
        if x.shape[0] == 0:   # the reference raises here (reshape of an empty tensor, :383); an empty flow is returned instead
            return x.new_empty((0, 2, self.S, self.S))
        return _run(self, x, None)


class FrequencyModel(_KernelModule):
    """defenses/frequency_based/model.py:8-52.  `forward` is the inference path; training goes through
    combat_b200.defenses.frequency_based.train (trainable=True builds the float32 training configuration)."""

    def __init__(self, num_classes=2, n_input=3, input_size=32, device=None, dtype=None, trainable=False):
        super().__init__()
        device = torch.device(device or "cuda")
        net = nets.FrequencyDetector(num_classes, n_input, input_size, device=device, dtype=dtype or default_dtype(),
                                     trainable=trainable)
        bufs = []
        for i in range(1, 7):
            bufs += [("bn%d.running_mean" % i, net.rm[i]), ("bn%d.running_var" % i, net.rv[i]),
                     ("bn%d.num_batches_tracked" % i, torch.zeros((), dtype=torch.long, device=device))]
        self._bind(net, bufs)
        import math
        gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
        st = net.store
        for i in range(1, 7):
            shape = st.shapes["conv%d.weight" % i]
            fan_in = shape[1] * 9
            b = math.sqrt(3.0) * gain / math.sqrt(fan_in)
            st.p("conv%d.weight" % i).copy_(torch.empty(shape).uniform_(-b, b))
            st.p("conv%d.bias" % i).copy_(torch.empty(shape[0]).uniform_(-1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in)))
        shape = st.shapes["linear6.weight"]
        b = math.sqrt(3.0) * gain / math.sqrt(shape[1])
        st.p("linear6.weight").copy_(torch.empty(shape).uniform_(-b, b))
        st.p("linear6.bias").copy_(torch.empty(shape[0]).uniform_(-1 / math.sqrt(shape[1]), 1 / math.sqrt(shape[1])))
        net.prep_weights()

    def forward(self, x):
        if self.training:
            raise NotImplementedError("FrequencyModel.forward is the inference path; train through "
                                      "combat_b200.defenses.frequency_based.train.train (SURVEY.md section 8f row 3)")
        self.net._affine, self.net._padded = None, None  # weights may have been (re)loaded
        with torch.no_grad():
            return self.net.forward(x.contiguous().float())
