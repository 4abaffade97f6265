"""The alternated generator / surrogate training step (train_generator.py:170-255 of the reference) as an
explicit schedule of sm_100a kernel launches -- CUDA-graph capturable, data-parallel ready.

Schedule ("strictly needed" column; verified bit-identical to the reference on CPU):
  * the generator runs ONCE per iteration on the full batch; the C-step's poisoned rows are gathered from that
    result (same weights, same inputs as the reference's separate `netG(inputs_toChange)` call, and
    InstanceNorm is per-sample);
  * the C-step does not back-propagate into the generator (the reference zeroes those gradients at :220);
  * the G-step back-propagates through netC and clean_model with input gradients only.
Everything observable by the reference's train() -- losses, the six accuracy counters, parameter and
BatchNorm-buffer updates -- is produced.

RNG parity: one `np.random.rand(n_trg)` for the poison count, one torch CPU uniform for
the C-step blur sigma (only when num_bd > 0), one for the G-step sigma -- all drawn on the host in `plan()`
from the HOST copy of the labels, so the device never has to sync.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from types import SimpleNamespace

import numpy as np
import torch

from . import ops
from .nets import Classifier, FrequencyDetector, Generator, GridGenerator
from .utils.dataloader import PARAM_WIDTH as TF_W, draw_params as draw_tf_params


def default_opt(**kw):
    """Hot-path flags with the defaults of the reference's config.py:4-86."""
    o = SimpleNamespace(
        dataset="cifar10", input_height=32, input_width=32, input_channel=3, num_classes=10, attack_mode="all2one",
        noise_rate=0.08, target_label=0, pc=0.5, ratio=0.65, kernel_size=3, sigma=(0.1, 1.0), L2_weight=0.02,
        clean_model_weight=0.8, lr_C=1e-2, lr_G=1e-2, bs=128,
        post_transform_option="no_use", random_crop=5, random_rotation=10,   # the CLI default is "use" (config.py:75)
    )
    for k, v in kw.items():
        setattr(o, k, v)
    return o


@dataclass
class StepPlan:
    """Host-side, per-iteration decisions (integer selection + RNG draws)."""
    perm: np.ndarray          # int32 [B]: batch order of total_inputs = cat(trg_ind, ntrg_ind)
    num_bd: int
    trg_ind: np.ndarray
    ntrg_ind: np.ndarray
    bd_targets: np.ndarray    # int64 [B]
    total_targets: np.ndarray  # int64 [B]
    sigma_c: float | None
    sigma_g: float
    taps_c: tuple
    taps_g: tuple
    taps_rows: np.ndarray | None = None   # multilabel: float32 [B, 2] per-row blur taps (one sigma per class chunk)
    sigmas_g: list | None = None          # multilabel: the per-chunk sigma draws
    # PostTensorTransform parameters of the five calls of one iteration (utils/dataloader.py), float32 [5, B, 8] in STORAGE
    # order T1 (:196) | T3 (:227) | T4 (:228) | T2 (:214) | T5 (:250): [1:3] is the parameter block of netC's batched
    # [x ; x_bd] forward, [3:5] that of clean_model's.  None: --post_transform_option no_use.
    tf: np.ndarray | None = None
    # inputaware variant (train_generator_inputaware.py:238-239): blur of inputs_bd2 = x + trigger of the second loader's image
    sigma_g2: float | None = None
    taps_g2: tuple = (1.0, 0.0)


# "T6": the transform of inputs_bd2 (train_generator_inputaware.py:241), present only in that variant's parameter block
TF_SLOT = {"T1": 0, "T3": 1, "T4": 2, "T2": 3, "T5": 4, "T6": 5}
N_LOSSES = 12   # float32 scalars a step leaves on the device (see _ensure_bufs)


def _tf_on(opt) -> bool:
    return getattr(opt, "post_transform_option", "no_use") != "no_use"


def _inputaware(opt) -> bool:
    return getattr(opt, "variant", "") == "inputaware"


def _wanet(opt) -> bool:
    return getattr(opt, "variant", "") == "wanet"


def create_targets_bd_np(targets: np.ndarray, opt) -> np.ndarray:
    """train_generator.py:70-77."""
    if opt.attack_mode == "all2one":
        return np.ones_like(targets) * opt.target_label
    if opt.attack_mode == "all2all":
        return (targets + 1) % opt.num_classes
    raise Exception("{} attack mode is not implemented".format(opt.attack_mode))


def make_plan(targets_host, opt, with_metrics=True) -> StepPlan:
    """train_generator.py:173,181-183,193-194,196,214,226-228,250 -- consumes the numpy global RNG, Python's `random` and the
    torch CPU generator in the reference's order: poison count, C-step sigma, T1, [T2], G-step sigma, [T3], T4, T5 (the
    transforms draw nothing under --post_transform_option no_use; T2 / T3 belong to the metric forwards)."""
    y = np.asarray(targets_host, dtype=np.int64)
    B = y.shape[0]
    ia = _inputaware(opt)   # train_generator_inputaware.py: one more sigma (:239) and one more transform (:241) in the G-step
    wn = _wanet(opt)        # train_generator_wanet.py: warp trigger, no GaussianBlur -> no sigma draws at all
    bd = create_targets_bd_np(y, opt).astype(np.int64)
    trg = np.nonzero(y == bd)[0]
    ntrg = np.nonzero(y != bd)[0]
    num_bd = int(np.sum(np.random.rand(trg.shape[0]) < opt.pc))
    sigma_c = None
    taps_c = (1.0, 0.0)
    if num_bd > 0 and not wn:
        sigma_c = torch.empty(1).uniform_(opt.sigma[0], opt.sigma[1]).item()
        taps_c = ops.gaussian_taps(sigma_c)
    tf = None
    if _tf_on(opt):
        tf = np.zeros((6 if ia else 5, B, TF_W), dtype=np.float32)
        tf[:, :, 2] = 1.0
        tf[TF_SLOT["T1"]] = draw_tf_params(B, opt)                                    # :196
        if with_metrics:
            tf[TF_SLOT["T2"]] = draw_tf_params(B, opt)                                # :214
    sigma_g = torch.empty(1).uniform_(opt.sigma[0], opt.sigma[1]).item() if not wn else None   # :226
    sigma_g2 = torch.empty(1).uniform_(opt.sigma[0], opt.sigma[1]).item() if ia else None   # inputaware :239
    if tf is not None:
        if with_metrics:
            tf[TF_SLOT["T3"]] = draw_tf_params(B, opt)                                # :227
        if ia:
            tf[TF_SLOT["T6"]] = draw_tf_params(B, opt)                                # inputaware :241 (before inputs_bd's, :242)
        tf[TF_SLOT["T4"]] = draw_tf_params(B, opt)                                    # :228
        tf[TF_SLOT["T5"]] = draw_tf_params(B, opt)                                    # :250
    perm = np.concatenate([trg, ntrg]).astype(np.int32)
    total_y = np.concatenate([bd[trg[:num_bd]], y[trg[num_bd:]], y[ntrg]]).astype(np.int64)
    return StepPlan(perm, num_bd, trg, ntrg, bd, total_y, sigma_c, sigma_g, taps_c,
                    ops.gaussian_taps(sigma_g) if not wn else (1.0, 0.0), tf=tf, sigma_g2=sigma_g2, taps_g2=ops.gaussian_taps(sigma_g2) if ia else (1.0, 0.0))


def multilabel_chunks(bs: int, num_classes: int):
    """train_generator_multilabel.py:203-211 -- contiguous chunks of ps rows, chunk ci is pushed towards class ci."""
    ps = int((bs - 1) / num_classes) + 1
    out = []
    for ci in range(num_classes):
        si, ei = ci * ps, min(ci * ps + ps, bs)
        if si >= ei:
            break
        out.append((ci, si, ei))
    return out


def make_plan_multilabel(targets_host, opt, with_metrics=True) -> StepPlan:
    """train_generator_multilabel.py:171,74,179,190,199,203-220,225,236 -- numpy rand(bs) for the poison count (the FIRST
    num_bd rows are poisoned, labels unchanged), one torch CPU uniform for the C-step blur (only when num_bd > 0), the
    transforms T1, [T2], [T3], then one uniform per class chunk of the G-step, in chunk order, then T4, T5."""
    y = np.asarray(targets_host, dtype=np.int64)
    bs = y.shape[0]
    num_bd = int(np.sum(np.random.rand(bs) < opt.pc))
    sigma_c, taps_c = None, (1.0, 0.0)
    if num_bd > 0:
        sigma_c = torch.empty(1).uniform_(opt.sigma[0], opt.sigma[1]).item()
        taps_c = ops.gaussian_taps(sigma_c)
    tf = None
    if _tf_on(opt):
        tf = np.zeros((5, bs, TF_W), dtype=np.float32)
        tf[:, :, 2] = 1.0
        tf[TF_SLOT["T1"]] = draw_tf_params(bs, opt)                                   # :179
        if with_metrics:
            tf[TF_SLOT["T2"]] = draw_tf_params(bs, opt)                               # :190
            tf[TF_SLOT["T3"]] = draw_tf_params(bs, opt)                               # :199
    bd = np.zeros(bs, dtype=np.int64)
    taps_rows = np.zeros((bs, 2), dtype=np.float32)
    sigmas = []
    for ci, si, ei in multilabel_chunks(bs, opt.num_classes):
        sg = torch.empty(1).uniform_(opt.sigma[0], opt.sigma[1]).item()
        sigmas.append(sg)
        bd[si:ei] = ci
        taps_rows[si:ei] = ops.gaussian_taps(sg)
    if tf is not None:
        tf[TF_SLOT["T4"]] = draw_tf_params(bs, opt)                                   # :225
        tf[TF_SLOT["T5"]] = draw_tf_params(bs, opt)                                   # :236
    return StepPlan(np.arange(bs, dtype=np.int32), num_bd, np.arange(bs), np.arange(0), bd, y.copy(), sigma_c,
                    sigmas[0], taps_c, (float(taps_rows[0, 0]), float(taps_rows[0, 1])), taps_rows, sigmas, tf=tf)


def make_plan_victim(targets_host, poisoned_host, opt) -> StepPlan:
    """train_victim.py:113-130: the poisoned rows come from the DATASET's flags (utils/dataloader_cleanbd.py:124-145), not from
    an RNG draw; the C-step blur sigma is drawn only when the batch holds a poisoned row (:128-129), then T1 (:131).
    `ntrg_ind = (poisoned is False).nonzero()` at :121 raises AttributeError as shipped (`poisoned is False` is a Python bool);
    the evident intent -- the rows whose flag is False -- is what is built here.  poisoned_host=None: train_clean_classifier.py
    (:88-104), no poisoned rows, no sigma draw."""
    y = np.asarray(targets_host, dtype=np.int64)
    B = y.shape[0]
    pz = np.zeros(B, dtype=bool) if poisoned_host is None else np.asarray(poisoned_host).astype(bool)
    bd = create_targets_bd_np(y, opt).astype(np.int64)
    trg, ntrg = np.nonzero(pz)[0], np.nonzero(~pz)[0]
    num_bd = int(trg.shape[0])
    sigma_c, taps_c = None, (1.0, 0.0)
    if num_bd > 0 and not _wanet(opt):   # train_victim_wanet.py:86-97: warp trigger, no blur, nothing drawn
        sigma_c = torch.empty(1).uniform_(opt.sigma[0], opt.sigma[1]).item()
        taps_c = ops.gaussian_taps(sigma_c)
    tf = None
    if _tf_on(opt):
        tf = np.zeros((5, B, TF_W), dtype=np.float32)
        tf[:, :, 2] = 1.0
        tf[TF_SLOT["T1"]] = draw_tf_params(B, opt)
    perm = np.concatenate([trg, ntrg]).astype(np.int32)
    total_y = np.concatenate([bd[trg], y[ntrg]]).astype(np.int64)
    return StepPlan(perm, num_bd, trg, ntrg, bd, total_y, sigma_c, 1.0, taps_c, (1.0, 0.0), tf=tf)


class AlternatedStep:
    """Owns netC / clean_model / netG / netF and runs alternated iterations on one GPU.

    `step(x, y, plan)` launches everything on the current stream and returns device scalars; with
    `use_graph=True` the launches of the first call are captured and later iterations replay the graph after
    refreshing the small per-iteration parameter block (perm, targets, num_bd, blur taps, learning rates)."""

    def __init__(self, opt=None, device="cuda", dtype=torch.bfloat16, classifier="preact_resnet18", with_metrics=True,
                 use_tc=True, cond_classes=0, grad_hook=None, buf_hook=None, nets=None, multilabel=False):
        self.opt = opt or default_opt()
        o = self.opt
        self.device = torch.device(device)
        self.dtype = dtype
        self.with_metrics = with_metrics
        H = o.input_height
        mk = dict(device=self.device, dtype=dtype, use_tc=use_tc)
        if nets is not None:  # adopt networks owned by the reference-facing nn.Module wrappers
            self.netC, self.clean, self.netG, self.netF = nets
            self.dtype = self.netC.dtype
        else:
            self.netC = Classifier(classifier, o.num_classes, o.input_channel, H, **mk)
            self.clean = Classifier(classifier, o.num_classes, o.input_channel, H, **mk)
            self.netG = GridGenerator(o.input_channel, 64, o.s, **mk) if _wanet(o) else Generator(o.input_channel, 64, cond_classes, **mk)
            self.netF = FrequencyDetector(2, o.input_channel, H, device=self.device, dtype=dtype) if (with_metrics and H in (32, 64)) else None
        # multilabel=True: the step of train_generator_multilabel.py (conditional generator, class-chunked G-step)
        self.multilabel = bool(multilabel)
        if self.multilabel and not self.netG.cond:
            raise ValueError("the multilabel step needs the conditional generator (cond_classes = num_classes)")
        self.keep = int(H * o.ratio)
        # variant of the step (getattr: the base trainer's opt carries no such attribute): "imperceptible" adds the total-variation
        # term of train_generator_imperceptible.py to the G-step loss
        self.tv_weight = float(getattr(o, "tv_weight", 0.0)) if getattr(o, "variant", "") == "imperceptible" else 0.0
        # "inputaware" (train_generator_inputaware.py): a second batch x2 per iteration; the generator runs ONCE over [x ; x2], the
        # trigger of x2's rows is pasted on x (inputs_bd2), netC sees it through one more eval-mode forward + input gradient
        self.inputaware = _inputaware(o)
        self.cross_weight = float(getattr(o, "cross_weight", 0.2)) if self.inputaware else 0.0
        if self.inputaware and self.multilabel:
            raise ValueError("the inputaware variant has no multilabel form in the reference")
        # "wanet" (train_generator_wanet.py): netG is a GridGenerator (S x S flow control grid), the trigger is a WARP of the image
        # (bicubic flow, grid_sample) -- one fused kernel instead of DCT low-pass + blend + blur; loss_l2 = MSE(noise_grid, 0)
        self.wanet = _wanet(o)
        if self.wanet:
            if self.multilabel or not isinstance(self.netG, GridGenerator):
                raise ValueError("the wanet variant needs a GridGenerator (and has no multilabel form in the reference)")
            self.grid_rescale, self.S = float(o.grid_rescale), int(o.s)
        self.tf_on = _tf_on(o)      # PostTensorTransform active (five fused gather launches + two adjoints per iteration)
        self.lr_C = torch.full((1,), float(o.lr_C), dtype=torch.float32, device=self.device)
        self.lr_G = torch.full((1,), float(o.lr_G), dtype=torch.float32, device=self.device)
        self.grad_hook = grad_hook  # callable(net_name, flat_grad_tensor): data-parallel all-reduce
        self.buf_hook = buf_hook    # callable(flat_running_stats): keeps BatchNorm buffers identical across ranks
        self.launches_per_step = 0  # kernels of this library launched by one iteration (counted on the last eager/capture pass)
        self._bufs = None           # buffers of the batch size used last
        self._bufs_by_B = {}        # batch size -> buffers + captured graphs (train and eval) + ring of plan slots
        self._comm = None           # communication stream of the data-parallel exchanges
        self._comm_done = None
        self._copy_stream = None   # side stream + staging buffers of prefetch()
        self._stage = {}
        self._stage_free = None
        self._prefetched = None
        self._rd_bufs = None       # pinned double buffer of read_async()
        self._rd_slot = 0

    # ------------------------------------------------------------ state
    def load_state(self, netC=None, clean=None, netG=None, netF=None):
        dev = self.device

        def mv(sd):
            return {k: v.to(dev) for k, v in sd.items()}

        if netC is not None:
            self.netC.load_state_dict(mv(netC))
        if clean is not None:
            self.clean.load_state_dict(mv(clean))
        if netG is not None:
            self.netG.load_state_dict(mv(netG))
        if netF is not None and self.netF is not None:
            self.netF.load_state_dict(mv(netF))

    def set_lr(self, lr_C=None, lr_G=None):
        if lr_C is not None:
            self.lr_C.fill_(float(lr_C))
        if lr_G is not None:
            self.lr_G.fill_(float(lr_G))

    # ------------------------------------------------------------ per-iteration device parameter block
    # Layout of the parameter block (one contiguous device buffer, one pinned host image of it per ring slot, ONE
    # cudaMemcpyAsync per iteration): y | bd_targets | total_y (int64 [B] each) | perm (int32 [B]) | small (float32 [8]:
    # taps_c, taps_g) | num_bd (int32 [4]) | taps_rows (float32 [B, 2], multilabel only).  Every section starts 16-byte
    # aligned.  The kernels read the DEVICE copy; the host image of slot s is only rewritten after the copy that last read
    # it has executed (event per slot) -- the host may run many iterations ahead of the GPU (train() syncs every
    # --log_every iterations), and a pinned buffer that is rewritten while an earlier cudaMemcpyAsync is still queued
    # would hand that earlier iteration the labels / permutation / num_bd / blur taps of a later one.
    PLAN_SLOTS = 4

    @staticmethod
    def _plan_layout(B, multilabel, tf_on=False, n_tf=5):
        a16 = lambda n: (n + 15) // 16 * 16
        off, lay = 0, {}
        for name, nbytes in (("y", 8 * B), ("bd_targets", 8 * B), ("total_y", 8 * B), ("perm", 4 * B), ("small", 32),
                             ("num_bd", 16), ("taps_rows", 8 * B if multilabel else 0),
                             ("tf", n_tf * B * TF_W * 4 if tf_on else 0)):
            lay[name] = (off, nbytes)
            off += a16(nbytes)
        return lay, max(off, 16)

    @staticmethod
    def _plan_views(block, lay, B, n_tf=5):
        def v(name, dtype, shape=None):
            o, n = lay[name]
            t = block[o:o + n].view(dtype)
            return t.view(shape) if shape is not None else t
        return {"y": v("y", torch.int64), "bd_targets": v("bd_targets", torch.int64), "total_y": v("total_y", torch.int64),
                "perm": v("perm", torch.int32), "small": v("small", torch.float32), "num_bd": v("num_bd", torch.int32),
                "taps_rows": v("taps_rows", torch.float32, (B, 2)) if lay["taps_rows"][1] else None,
                "tf": v("tf", torch.float32, (n_tf, B, TF_W)) if lay["tf"][1] else None}

    def _ensure_bufs(self, B):
        """Buffers (and the captured graphs that reference them) are cached PER BATCH SIZE: the shorter last batch of an
        epoch, or an eval batch size different from the training one, does not throw the other size's graphs away."""
        b = self._bufs_by_B.get(B)
        if b is not None:
            self._bufs = b
            return b
        dev = self.device
        o = self.opt
        b = {"B": B, "graph": None, "gstate": {}}
        if self.inputaware:   # [x ; second loader's batch]: one generator forward / backward over both
            b["xx"] = torch.empty((2 * B, o.input_channel, o.input_height, o.input_width), dtype=torch.float32, device=dev)
            b["x"] = b["xx"][:B]
        else:
            b["x"] = torch.empty((B, o.input_channel, o.input_height, o.input_width), dtype=torch.float32, device=dev)
        b["x2"] = torch.empty((2 * B, o.input_channel, o.input_height, o.input_width), dtype=torch.float32, device=dev)  # [x ; x_bd]
        n_tf = 6 if self.inputaware else 5
        lay, nbytes = self._plan_layout(B, self.multilabel, self.tf_on, n_tf)
        b["plan_dev"] = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        dv = self._plan_views(b["plan_dev"], lay, B, n_tf)
        b["y"], b["bd_targets"], b["total_y"], b["perm"] = dv["y"], dv["bd_targets"], dv["total_y"], dv["perm"]
        b["taps_c"], b["taps_g"], b["taps_g2"] = dv["small"][0:2], dv["small"][2:4], dv["small"][4:6]
        b["num_bd"] = dv["num_bd"][0:1]
        b["taps_rows"] = dv["taps_rows"]
        b["tf"] = dv["tf"]
        b["ones"] = torch.ones(B, dtype=torch.int64, device=dev)
        if self.wanet:   # identity-grid coordinates (train_generator_wanet.py:560: torch.linspace on the host, moved to the device)
            b["ident"] = torch.linspace(-1, 1, steps=o.input_height).to(dev)
            b["noise_grid"] = torch.empty((B, o.input_height, o.input_width, 2), dtype=torch.float32, device=dev)
            b["gl_partial"] = torch.empty(B, dtype=torch.float32, device=dev)
        b["sq_partial"] = torch.empty(B * o.input_channel, dtype=torch.float32, device=dev)
        b["gl2_partial"] = torch.empty(2 * B * o.input_channel, dtype=torch.float32, device=dev)
        b["tv_partial"] = torch.empty(B * o.input_channel, dtype=torch.float32, device=dev)
        b["losses"] = torch.zeros(N_LOSSES, dtype=torch.float32, device=dev)   # loss_c, loss_ce, loss_l2, clean_model_loss, CE
        #                                                    of the metric forwards [4:7], loss_grad_l2 [7], loss_tv [8] (imperceptible)
        b["counts"] = torch.zeros(16, dtype=torch.int32, device=dev)
        # ring of pinned host images of the parameter block, each guarded by the event of the copy that last read it
        b["slots"] = []
        for _ in range(self.PLAN_SLOTS):
            h = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
            b["slots"].append({"h": h, "v": self._plan_views(h, lay, B, n_tf), "ev": None})
        b["slot"] = 0
        self._bufs_by_B[B] = b
        self._bufs = b
        return b

    def _next_slot(self, b):
        if os.environ.get("COMBAT_UNSAFE_PLAN_STAGING"):
            # round-1 behaviour (ONE unguarded pinned image), kept only so that tests/test_graph_nosync_gpu.py can be shown to
            # FAIL without the guard (profiles/r02_staging_race.md); never set in production
            return b["slots"][0]
        s = b["slots"][b["slot"]]
        b["slot"] = (b["slot"] + 1) % len(b["slots"])
        if s["ev"] is not None:
            s["ev"].synchronize()   # the H2D copy that last read this host image has executed
        return s

    def upload_plan(self, y_host, plan: StepPlan):
        """one async H2D copy of labels + the per-iteration parameter block (a few KB) from an event-guarded ring slot."""
        B = len(plan.perm)
        b = self._ensure_bufs(B)
        s = self._next_slot(b)
        v = s["v"]
        v["y"].copy_(torch.as_tensor(np.asarray(y_host, dtype=np.int64)))
        v["bd_targets"].copy_(torch.from_numpy(plan.bd_targets))
        v["total_y"].copy_(torch.from_numpy(plan.total_targets))
        v["perm"].copy_(torch.from_numpy(plan.perm))
        v["small"][0], v["small"][1] = plan.taps_c
        v["small"][2], v["small"][3] = plan.taps_g
        v["small"][4], v["small"][5] = plan.taps_g2
        v["num_bd"][0] = plan.num_bd
        if self.multilabel:
            v["taps_rows"].copy_(torch.from_numpy(plan.taps_rows))
        if self.tf_on:
            if plan.tf is None:
                raise ValueError("the engine was built with PostTensorTransform on: the plan must carry its parameters")
            if tuple(plan.tf.shape) != tuple(v["tf"].shape):
                raise ValueError("the plan's PostTensorTransform block %s does not match the engine's %s (variant mismatch)"
                                 % (tuple(plan.tf.shape), tuple(v["tf"].shape)))
            v["tf"].copy_(torch.from_numpy(plan.tf))
        b["plan_dev"].copy_(s["h"], non_blocking=True)
        s["ev"] = torch.cuda.Event()
        s["ev"].record(torch.cuda.current_stream(self.device))

    # ------------------------------------------------------------ the step
    # The iteration is five launch phases; the data-parallel exchanges (combat_b200.parallel) run on a communication stream
    # BETWEEN them, overlapped with the phase that does not need their result:
    #   A : generator forward, C-step forward + backward            -> all-reduce(netC gradients, BatchNorm buffers) starts
    #   B1: trigger batch, clean_model forward + backward             (independent of the netC update: hides that all-reduce)
    #   B2: netC SGD, netC metric/backdoor forward + backward, generator backward   -> all-reduce(netG gradients) starts
    #   C1: frequency-detector metric leg                             (hides part of it)
    #   C2: netG SGD
    # Single GPU: one CUDA graph over A+B1+B2+C1+C2 (same launches, same order).  Data parallel: one graph per phase; no
    # collective inside a captured graph (it hung under torch 2.11 / NCCL 2.28.9).
    def _phase_a(self, b, st, save_g=True, with_g=True):
        o = self.opt
        x = b["x"]
        losses, counts = b["losses"], b["counts"]
        if self.multilabel:
            # C-step trigger: conditioned on the TRUE labels (train_generator_multilabel.py:172-176); the G-step runs its own
            # forward conditioned on the chunk classes, so nothing is saved here
            noise_c_raw, _ = self.netG.forward(x, b["y"], save=False)
            noise_c = ops.plane_op(noise_c_raw, "lowfreq", keep=self.keep)
            total_x = ops.poison_blend_fwd(x, noise_c, None, 0, o.noise_rate, None, taps_dev=b["taps_c"],
                                           num_bd_dev=b["num_bd"])                            # first num_bd rows, :178
            noise_raw = ctxG = noise = None
        elif not with_g:  # train_clean_classifier.py: no generator, the batch as it is
            noise_raw = ctxG = noise = None
            total_x = x
        elif self.wanet:                                                                     # wanet :150-159
            noise_raw, ctxG = self.netG.forward(x, None, save=save_g)   # the flow control grid [B, 2, S, S], once for both steps
            noise = None
            total_x = ops.wanet_warp_fwd(x, noise_raw, b["ident"], b["perm"], 0, self.grid_rescale, self.S,
                                         num_bd_dev=b["num_bd"])
        else:
            gin = b["xx"] if (self.inputaware and save_g) else x                             # inputaware: + netG(inputs2), :238
            noise_raw, ctxG = self.netG.forward(gin, None, save=save_g)                      # :189 and :223, once
            noise = ops.plane_op(noise_raw, "lowfreq", keep=self.keep)                      # :190-191 / :224
            total_x = ops.poison_blend_fwd(x, noise[:x.shape[0]], b["perm"], 0, o.noise_rate, None, taps_dev=b["taps_c"],
                                           num_bd_dev=b["num_bd"])                            # :192-195
        total_in = ops.post_transform_fwd(total_x, b["tf"][TF_SLOT["T1"]]) if self.tf_on else total_x   # :196
        logits_c, ctxC = self.netC.forward(total_in, train=True, save=True)                  # :205
        _, dlog, _ = ops.cross_entropy(logits_c, b["total_y"], 1.0, True, loss_out=losses[0:1], counts_out=counts[0:2])
        self.netC.zero_grad()                                                                # :179
        self.netC.backward(ctxC, dlog, need_wgrad=True, need_dx=False)                       # :211
        st.update(noise_raw=noise_raw, ctxG=ctxG, noise=noise, total_x=total_x, total_in=total_in, logits_c=logits_c)

    # Phase B is split at the point where the AVERAGED netC gradient is first needed: B1 (trigger batch x_bd, MSE terms, the
    # clean_model forward / backward) does not depend on the C-step update and runs while the netC all-reduce is in flight on the
    # communication stream; B2 starts with the netC optimiser step.  Likewise C1 (frequency-detector metric leg) overlaps the
    # netG all-reduce and C2 is the netG optimiser step.  Single GPU: the same launches in the same order, one graph.
    def _trigger_batch(self, b, st, noise, out):
        """The G-step's poisoned batch inputs_bd and its distance term -> losses[2].  Additive trigger (train_generator.py:225-226,
        234): blend + clamp + blur, MSE(inputs_bd, inputs).  wanet (:196-203, 212-222): warp of the image by the generated flow,
        MSE(noise_grid, 0), and the logged finite-difference term -> losses[7]."""
        o = self.opt
        x, B, losses = b["x"], b["B"], b["losses"]
        if self.wanet:
            x_bd = ops.wanet_warp_fwd(x, st["noise_raw"], b["ident"], None, B, self.grid_rescale, self.S, out=out,
                                      noise_grid=b["noise_grid"], sq_partial=b["sq_partial"], gl_partial=b["gl_partial"])
            ops.sum_scale(b["sq_partial"][:B], 1.0 / b["noise_grid"].numel(), out=losses[2:3])
            ops.sum_scale(b["gl_partial"], 1.0 / B, out=losses[7:8])
            return x_bd
        x_bd = ops.poison_blend_fwd(x, noise, None, B, o.noise_rate, None, out=out, sq_partial=b["sq_partial"],
                                    taps_dev=b["taps_g"], taps_rows=b["taps_rows"])
        ops.sum_scale(b["sq_partial"], 1.0 / x.numel(), out=losses[2:3])
        return x_bd

    def _phase_b1(self, b, st):
        o = self.opt
        x, y, B = b["x"], b["y"], b["B"]
        losses, counts = b["losses"], b["counts"]
        noise = st["noise"]
        numel = x.numel()
        if self.inputaware:                                                                  # inputaware :238-239
            noise, noise2 = noise[:B], noise[B:]
            st["x_bd2"] = ops.poison_blend_fwd(x, noise2, None, B, o.noise_rate, None, taps_dev=b["taps_g2"])
            st.update(noise=noise, noise2=noise2)
        if self.multilabel:                                                                  # multilabel :203-221
            noise_raw, ctxG = self.netG.forward(x, b["bd_targets"], save=True)
            noise = ops.plane_op(noise_raw, "lowfreq", keep=self.keep)
            st.update(noise_raw=noise_raw, noise=noise, ctxG=ctxG)
        batched = self.with_metrics and self.netC.fuse_eval and self.clean.fuse_eval
        st["batched"] = batched
        if batched:
            # the metric forward on x (:214 / :227) and the forward on x_bd (:228 / :250) of each frozen-in-this-phase
            # classifier run as ONE eval-mode forward over [x ; x_bd]: per-sample independent, so every output is
            # bit-identical to two separate calls; half the launches, fuller waves on the deep layers.  Only the x_bd
            # half of the saved state is back-propagated.
            x2 = b["x2"]
            x2[:B].copy_(x)
            x_bd = self._trigger_batch(b, st, noise, x2[B:])                                 # :225-226, :234
            # PostTensorTransform: netC sees [T3(x) ; T4(x_bd)] (:227-228), clean_model [T2(x) ; T5(x_bd)] (:214,:250) -- one
            # gather launch per network over the 2B rows with per-row parameters
            tfK = b["tf"][TF_SLOT["T2"]:TF_SLOT["T5"] + 1].view(2 * B, TF_W) if self.tf_on else None
            x2k = ops.post_transform_fwd(x2, tfK) if self.tf_on else x2
            lgc, ctx2 = self.clean.forward(x2k, train=False, save=True)
            clean_preds, cm_preds = lgc[:B], lgc[B:]
            ops.cross_entropy(clean_preds, y, 1.0, False, loss_out=losses[4:5], counts_out=counts[2:4])
            _, dl2, _ = ops.cross_entropy(cm_preds, y, o.clean_model_weight, True, targets2=b["bd_targets"],
                                          loss_out=losses[3:4], counts_out=counts[8:10])      # :251,266-267
            g2 = self.clean.backward(self.clean.slice_ctx(ctx2, B, 2 * B), dl2, need_wgrad=False, need_dx=True)
            del ctx2
            st.update(clean_preds=clean_preds)
        else:
            tfp = (lambda name, t: ops.post_transform_fwd(t, b["tf"][TF_SLOT[name]])) if self.tf_on else (lambda name, t: t)
            if self.with_metrics:
                clean_preds, _ = self.clean.forward(tfp("T2", x), train=False, save=False)   # :214
                ops.cross_entropy(clean_preds, y, 1.0, False, loss_out=losses[4:5], counts_out=counts[2:4])
                st["clean_preds"] = clean_preds
            x_bd = self._trigger_batch(b, st, noise, None)                                   # :225-226, :234
            cm_preds, ctxK = self.clean.forward(tfp("T5", x_bd), train=False, save=True)     # :250
            _, dl2, _ = ops.cross_entropy(cm_preds, y, o.clean_model_weight, True, targets2=b["bd_targets"],
                                          loss_out=losses[3:4], counts_out=counts[8:10])      # :251,266-267
            g2 = self.clean.backward(ctxK, dl2, need_wgrad=False, need_dx=True)
            del ctxK
        if self.with_metrics and not self.multilabel and not self.inputaware and not self.wanet:
            ops.grad_l2(x, x_bd, losses[7:8], b["gl2_partial"])                              # :235-243 (logged only)
        st.update(x_bd=x_bd, clean_model_preds=cm_preds, g2=g2)

    def _phase_b2(self, b, st):
        o = self.opt
        x, y, B = b["x"], b["y"], b["B"]
        losses, counts = b["losses"], b["counts"]
        noise, x_bd, g2 = st["noise"], st["x_bd"], st["g2"]
        numel = x.numel()
        self.netC.sgd_step(self.lr_C)                                                        # :212
        if st["batched"]:
            x2 = b["x2"]
            tfC = b["tf"][TF_SLOT["T3"]:TF_SLOT["T4"] + 1].view(2 * B, TF_W) if self.tf_on else None
            x2c = ops.post_transform_fwd(x2, tfC) if self.tf_on else x2
            lg, ctx2 = self.netC.forward(x2c, train=False, save=True)
            pred_clean, pred_bd = lg[:B], lg[B:]
            ops.cross_entropy(pred_clean, y, 1.0, False, loss_out=losses[5:6], counts_out=counts[4:6])
            _, dl1, _ = ops.cross_entropy(pred_bd, b["bd_targets"], 1.0, True, loss_out=losses[1:2], counts_out=counts[6:8])
            g1 = self.netC.backward(self.netC.slice_ctx(ctx2, B, 2 * B), dl1, need_wgrad=False, need_dx=True)
            del ctx2
            st.update(pred_clean=pred_clean)
        else:
            tfp = (lambda name, t: ops.post_transform_fwd(t, b["tf"][TF_SLOT[name]])) if self.tf_on else (lambda name, t: t)
            if self.with_metrics:
                pred_clean, _ = self.netC.forward(tfp("T3", x), train=False, save=False)     # :227
                ops.cross_entropy(pred_clean, y, 1.0, False, loss_out=losses[5:6], counts_out=counts[4:6])
                st["pred_clean"] = pred_clean
            pred_bd, ctxB = self.netC.forward(tfp("T4", x_bd), train=False, save=True)       # :228
            _, dl1, _ = ops.cross_entropy(pred_bd, b["bd_targets"], 1.0, True, loss_out=losses[1:2], counts_out=counts[6:8])
            g1 = self.netC.backward(ctxB, dl1, need_wgrad=False, need_dx=True)
            del ctxB
        gx = None
        if self.inputaware:   # train_generator_inputaware.py:241,246,261: + cross_weight * CE(netC(T(inputs_bd2)), targets)
            x_bd2 = st["x_bd2"]
            xin = ops.post_transform_fwd(x_bd2, b["tf"][TF_SLOT["T6"]]) if self.tf_on else x_bd2
            pred_cross, ctxX = self.netC.forward(xin, train=False, save=True)
            _, dlx, _ = ops.cross_entropy(pred_cross, y, self.cross_weight, True, loss_out=losses[9:10], counts_out=counts[12:14])
            gx = self.netC.backward(ctxX, dlx, need_wgrad=False, need_dx=True)
            del ctxX
            if self.tf_on:
                gx = ops.post_transform_bwd(gx, b["tf"][TF_SLOT["T6"]])
            st.update(pred_cross=pred_cross)
        if self.tf_on:  # back through T4 / T5 to x_bd: both adjoints accumulate into one buffer
            gsum = ops.post_transform_bwd(g1, b["tf"][TF_SLOT["T4"]])
            ops.post_transform_bwd(g2, b["tf"][TF_SLOT["T5"]], out=gsum, accumulate=True)
            g1, g2 = gsum, None
        if self.tv_weight:   # train_generator_imperceptible.py:228,235: + tv_weight * total_variation(inputs_bd).mean()
            ops.tv_loss(x_bd, losses[8:9], grad=g1, grad_weight=self.tv_weight / B, partial=b["tv_partial"])
        if self.wanet:        # back through grid_sample / clamp / bicubic upsampling to the flow control grid (csrc/warp.cu)
            dflow = ops.wanet_warp_bwd(x, st["noise_raw"], b["ident"], g1, g2, self.grid_rescale,
                                       2.0 * o.L2_weight / b["noise_grid"].numel(), self.S)
            self.netG.zero_grad()                                                            # :195
            self.netG.backward(st.pop("ctxG"), dflow.view(B, 2, self.S, self.S))             # :236
            st.update(pred_bd=pred_bd, g1=g1, g2=g2, dnoise=dflow)
            return
        if self.inputaware:   # gradients of both trigger batches, [d noise(x) ; d noise(x2)], through ONE generator backward
            dnoise = torch.empty_like(b["xx"])
            ops.poison_blend_bwd(x, noise, x_bd, g1, g2, 2.0 * o.L2_weight / numel, o.noise_rate, None, out=dnoise[:B],
                                 taps_dev=b["taps_g"])
            ops.poison_blend_bwd(x, st["noise2"], st["x_bd2"], gx, None, 0.0, o.noise_rate, None, out=dnoise[B:],
                                 taps_dev=b["taps_g2"])
        else:
            dnoise = ops.poison_blend_bwd(x, noise, x_bd, g1, g2, 2.0 * o.L2_weight / numel, o.noise_rate, None,
                                          taps_dev=b["taps_g"], taps_rows=b["taps_rows"])
        dnoise_raw = ops.plane_op(dnoise, "lowfreq", keep=self.keep)                          # P is symmetric
        self.netG.zero_grad()                                                                # :220
        self.netG.backward(st.pop("ctxG"), dnoise_raw)                                       # :254
        st.update(pred_bd=pred_bd, g1=g1, g2=g2, dnoise=dnoise)

    def _phase_c1(self, b, st):
        losses, counts = b["losses"], b["counts"]
        if self.with_metrics and self.netF is not None:
            inputs_F = ops.plane_op(st["x_bd"], "dct", in_mode=2)                             # :245
            pred_F = self.netF.forward(inputs_F)                                              # :247
            ops.cross_entropy(pred_F, b["ones"], 1.0, False, loss_out=losses[6:7], counts_out=counts[10:12])
            st.update(inputs_F=inputs_F, pred_F=pred_F)

    def _phase_c2(self, b, st):
        self.netG.sgd_step(self.lr_G)                                                        # :255

    # ---- data-parallel exchange: started on a communication stream right after the backward that produced the gradients,
    # joined just before the optimiser step that consumes them; the flat gradient goes out in COMM_BUCKETS contiguous buckets
    COMM_BUCKETS = int(os.environ.get("COMBAT_COMM_BUCKETS", "4"))
    def _exchange_start(self, which):
        if not self._parallel:
            return
        main = torch.cuda.current_stream(self.device)
        if os.environ.get("COMBAT_DP_OVERLAP") == "0":   # A/B switch: exchange on the compute stream, nothing overlapped
            if self.grad_hook is not None:
                self.grad_hook(which, self.netC.store.grad if which == "netC" else self.netG.store.grad)
            if which == "netC" and self.buf_hook is not None:
                self.buf_hook(self.netC.bufs)
            return
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=self.device, priority=-1)   # its CTAs are scheduled ahead of the compute stream's
        ev = torch.cuda.Event()
        ev.record(main)
        self._comm.wait_event(ev)
        with torch.cuda.stream(self._comm):
            if self.grad_hook is not None:
                flat = self.netC.store.grad if which == "netC" else self.netG.store.grad
                n = flat.numel()
                step = (n + self.COMM_BUCKETS - 1) // self.COMM_BUCKETS
                for lo in range(0, n, step):
                    self.grad_hook(which, flat[lo:lo + step])
            if which == "netC" and self.buf_hook is not None:
                self.buf_hook(self.netC.bufs)
            self._comm_done = torch.cuda.Event()
            self._comm_done.record(self._comm)

    def _exchange_join(self):
        if self._parallel and self._comm_done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._comm_done)
            self._comm_done = None

    @property
    def _parallel(self):
        return self.grad_hook is not None or self.buf_hook is not None

    def _launch(self, b, keep_debug=False):
        from ._lib import launch_count
        n0 = launch_count()
        st = {}
        try:
            self._phase_a(b, st)
            self._exchange_start("netC")
            self._phase_b1(b, st)
            self._exchange_join()
            self._phase_b2(b, st)
            self._exchange_start("netG")
            self._phase_c1(b, st)
            self._exchange_join()
            self._phase_c2(b, st)
        finally:
            self.launches_per_step = launch_count() - n0
        return st if keep_debug else None

    def prefetch(self, x_host):
        """Start the host->device copy of the NEXT batch on a side stream while the current iteration runs; a following
        `step(x_host, ...)` with the same tensor waits for it and takes a device-to-device copy instead of a PCIe one."""
        if x_host.is_cuda:
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        buf = self._stage.get(tuple(x_host.shape))
        if buf is None:
            # allocated FROM THE COPY STREAM'S POOL: a block of the compute stream's pool may belong to a tensor that Python has
            # already released while kernels reading it are still queued (the host runs ahead of the GPU) -- the side-stream copy
            # below would overwrite it under them.  Seen as a wrong first-conv weight gradient (the one gradient that reads the
            # step's input images at backward time) in tests/test_wanet_gpu.py when the allocator handed such a block out.
            with torch.cuda.stream(self._copy_stream):
                buf = self._stage[tuple(x_host.shape)] = torch.empty(x_host.shape, dtype=torch.float32, device=self.device)
        # only the previous consumer of the staging buffer (last step's device-to-device copy) must have finished -- NOT the
        # iteration that was just launched, which is what this copy overlaps with
        if self._stage_free is not None:
            self._copy_stream.wait_event(self._stage_free)
        with torch.cuda.stream(self._copy_stream):
            buf.copy_(x_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._prefetched = (x_host, buf, ev)

    def _take_input(self, x, dst):
        pf = self._prefetched
        if pf is not None and pf[0] is x:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(pf[2])
            dst.copy_(pf[1], non_blocking=True)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(cur)
            self._prefetched = None
        else:
            dst.copy_(x, non_blocking=True)

    def step(self, x_dev, y_host, plan: StepPlan | None = None, use_graph=False, keep_debug=False, x2=None):
        """One alternated iteration.  x_dev: float32 NCHW tensor already on the device (or a pinned host tensor,
        which is copied asynchronously); y_host: host labels; x2: the second loader's batch (inputaware variant only).
        Returns {'losses': dev[N_LOSSES], 'counts': dev[16]}."""
        if plan is None:
            mk = make_plan_multilabel if self.multilabel else make_plan
            plan = mk(y_host, self.opt, self.with_metrics)
        B = len(plan.perm)
        b = self._ensure_bufs(B)
        self._take_input(x_dev, b["x"])
        if self.inputaware:
            if x2 is None or tuple(x2.shape) != tuple(b["x"].shape):
                raise ValueError("the inputaware step needs the second loader's batch, same shape as the first")
            b["xx"][B:].copy_(x2, non_blocking=True)
        self.upload_plan(y_host, plan)
        dbg = None
        if use_graph and not keep_debug:
            if b["graph"] is None:
                # warm-up launch outside capture (allocator pools, function attributes, first-step SGD), then capture
                self._launch(b)
                torch.cuda.current_stream().synchronize()
                pool = torch.cuda.graph_pool_handle()
                b["gstate"] = {}  # tensors handed from one captured phase to the next stay referenced here
                phases = [self._phase_a, self._phase_b1, self._phase_b2, self._phase_c1, self._phase_c2]
                if self._parallel:
                    graphs = []
                    for ph in phases:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, pool=pool):
                            ph(b, b["gstate"])
                        graphs.append(g)
                else:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool):
                        for ph in phases:
                            ph(b, b["gstate"])
                    graphs = [g]
                b["graph"] = graphs
            else:
                graphs = b["graph"]
                if len(graphs) == 1:
                    graphs[0].replay()
                else:   # A | all-reduce(netC) overlapped with B1 | B2 | all-reduce(netG) overlapped with C1 | C2
                    graphs[0].replay()
                    self._exchange_start("netC")
                    graphs[1].replay()
                    self._exchange_join()
                    graphs[2].replay()
                    self._exchange_start("netG")
                    graphs[3].replay()
                    self._exchange_join()
                    graphs[4].replay()
                self.netC.bump_batches_tracked()   # the replayed C-step forward is one train-mode pass of every BatchNorm
        else:
            dbg = self._launch(b, keep_debug)
        out = {"losses": b["losses"], "counts": b["counts"], "plan": plan}
        if dbg is not None:
            out["debug"] = dbg
        return out

    # ------------------------------------------------------------ victim / clean-classifier training
    def victim_step(self, x_dev, y_host, poisoned_host=None, plan: StepPlan | None = None, use_graph=False, keep_debug=False):
        """One iteration of train_victim.py:110-140 (poisoned_host given: the dataset's per-sample flags, the frozen generator
        builds the triggers) or of train_clean_classifier.py:88-104 (poisoned_host None and no generator): the C-step half of
        the alternated step -- batch assembly, PostTensorTransform, netC train-mode forward / backward, SGD.
        Returns {'losses': dev[8] (loss_ce at [0]), 'counts': dev[16] (correct predictions on total_targets at [0])}."""
        if plan is None:
            plan = make_plan_victim(y_host, poisoned_host, self.opt)
        B = len(plan.perm)
        b = self._ensure_bufs(B)
        self._take_input(x_dev, b["x"])
        self.upload_plan(y_host, plan)
        with_g = self.netG is not None

        def launch(st):
            self._phase_a(b, st, save_g=False, with_g=with_g)
            self._exchange_start("netC")
            self._exchange_join()
            self.netC.sgd_step(self.lr_C)

        dbg = None
        if use_graph and not keep_debug:
            if b.get("vgraph") is None:
                launch({})
                torch.cuda.current_stream().synchronize()
                b["vstate"] = {}
                if self._parallel:   # no collective inside a captured graph: forward/backward graph, exchange, optimiser graph
                    g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    pool = torch.cuda.graph_pool_handle()
                    with torch.cuda.graph(g1, pool=pool):
                        self._phase_a(b, b["vstate"], save_g=False, with_g=with_g)
                    with torch.cuda.graph(g2, pool=pool):
                        self.netC.sgd_step(self.lr_C)
                    b["vgraph"] = [g1, g2]
                else:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        launch(b["vstate"])
                    b["vgraph"] = [g]
            else:
                gs = b["vgraph"]
                gs[0].replay()
                if len(gs) > 1:
                    self._exchange_start("netC")
                    self._exchange_join()
                    gs[1].replay()
                self.netC.bump_batches_tracked()
        else:
            dbg = {}
            launch(dbg)
        out = {"losses": b["losses"], "counts": b["counts"], "plan": plan}
        if keep_debug:
            out["debug"] = dbg
        return out

    # ------------------------------------------------------------ evaluation (train_generator.py:355-391)
    def eval_step(self, x_dev, y_host, sigma=None, use_graph=False, x2=None, sigma2=None):
        """One batch of eval(): clean accuracy of netC, attack success on the NON-TARGET samples, detector and clean-model
        legs.  The trigger is built for every row of the fixed-shape batch (eval-mode networks and the blur are per-sample
        independent, so the non-target rows are bit-identical to the reference's gathered sub-batch) and the target rows are
        masked out of the counters with a negative label; one blur sigma per batch, drawn like torchvision's GaussianBlur.
        Returns device int32 counts [clean, -, bd, -, F, -, clean_model, -, bd_ba, bd_asr, -, -, cross] and the host-side n_bd.
        inputaware variant (train_generator_inputaware.py:402-413): x2 = the second loader's batch; the trigger of x2's rows
        on x, its own sigma draw (after the first), netC's accuracy on the non-target rows against their TRUE labels."""
        o = self.opt
        y = np.asarray(y_host, dtype=np.int64)
        if sigma is None and not self.wanet:   # the warp trigger has no blur: nothing is drawn (train_generator_wanet.py:374-385)
            sigma = torch.empty(1).uniform_(o.sigma[0], o.sigma[1]).item()
        if self.inputaware:
            if x2 is None:
                raise ValueError("the inputaware evaluation needs the second loader's batch")
            if sigma2 is None:
                sigma2 = torch.empty(1).uniform_(o.sigma[0], o.sigma[1]).item()
        bd = create_targets_bd_np(y, o)
        ntrg = y != o.target_label
        B = len(y)
        b = self._ensure_bufs(B)
        eb = b.setdefault("eval", {})
        if not eb:
            dev = self.device
            nbytes = 4 * B * 8 + 16
            eb["dev"] = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            eb["t"] = eb["dev"][:4 * B * 8].view(torch.int64).view(4, B)   # y | bd masked | ones masked | y masked
            eb["taps"] = eb["dev"][4 * B * 8:].view(torch.float32)[0:2]
            eb["taps2"] = eb["dev"][4 * B * 8:].view(torch.float32)[2:4]
            if self.inputaware:
                eb["x2"] = torch.empty_like(b["x"])
            eb["counts"] = torch.zeros(16, dtype=torch.int32, device=dev)
            eb["graph"] = None
            eb["slots"] = []   # event-guarded ring of pinned host images, as for the training plan (see _ensure_bufs)
            for _ in range(self.PLAN_SLOTS):
                hb = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
                eb["slots"].append({"h": hb, "t": hb[:4 * B * 8].view(torch.int64).view(4, B),
                                    "taps": hb[4 * B * 8:].view(torch.float32), "ev": None})
            eb["slot"] = 0
        s = self._next_slot(eb)
        h = s["t"]
        h[0].copy_(torch.from_numpy(y))
        h[1].copy_(torch.from_numpy(np.where(ntrg, bd, -1)))
        h[2].copy_(torch.from_numpy(np.where(ntrg, 1, -1).astype(np.int64)))
        h[3].copy_(torch.from_numpy(np.where(ntrg, y, -1)))
        s["taps"][0], s["taps"][1] = ops.gaussian_taps(sigma) if sigma is not None else (1.0, 0.0)
        if self.inputaware:
            s["taps"][2], s["taps"][3] = ops.gaussian_taps(sigma2)
            eb["x2"].copy_(x2, non_blocking=True)
        eb["dev"].copy_(s["h"], non_blocking=True)
        s["ev"] = torch.cuda.Event()
        s["ev"].record(torch.cuda.current_stream(self.device))
        b["x"].copy_(x_dev, non_blocking=True)

        def launch():
            x, t, counts = b["x"], eb["t"], eb["counts"]
            preds_clean, _ = self.netC.forward(x, train=False, save=False)                          # :360
            ops.cross_entropy(preds_clean, t[0], 1.0, False, counts_out=counts[0:2])
            noise_raw, _ = self.netG.forward(x, None, save=False)                                    # :369
            if self.wanet:
                x_bd = ops.wanet_warp_fwd(x, noise_raw, b["ident"], None, B, self.grid_rescale, self.S)
            else:
                noise = ops.plane_op(noise_raw, "lowfreq", keep=self.keep)                               # :370
                x_bd = ops.poison_blend_fwd(x, noise, None, B, o.noise_rate, None, taps_dev=eb["taps"])  # :372-373
            preds_bd, _ = self.netC.forward(x_bd, train=False, save=False)                          # :375
            ops.cross_entropy(preds_bd, t[1], 1.0, False, counts_out=counts[2:4])
            extra = {}
            if self.inputaware:                                                                      # inputaware :402-413
                noise2 = ops.plane_op(self.netG.forward(eb["x2"], None, save=False)[0], "lowfreq", keep=self.keep)
                x_bd2 = ops.poison_blend_fwd(x, noise2, None, B, o.noise_rate, None, taps_dev=eb["taps2"])
                preds_cross, _ = self.netC.forward(x_bd2, train=False, save=False)
                ops.cross_entropy(preds_cross, t[3], 1.0, False, counts_out=counts[12:14])
                extra = dict(preds_cross=preds_cross, x_bd2=x_bd2)
            if self.netF is not None:
                preds_F = self.netF.forward(ops.plane_op(x_bd, "dct", in_mode=2))                    # :381-383
                ops.cross_entropy(preds_F, t[2], 1.0, False, counts_out=counts[4:6])
            cm_clean, _ = self.clean.forward(x, train=False, save=False)                             # :387
            ops.cross_entropy(cm_clean, t[0], 1.0, False, counts_out=counts[6:8])
            cm_bd, _ = self.clean.forward(x_bd, train=False, save=False)                             # :389
            ops.cross_entropy(cm_bd, t[3], 1.0, False, targets2=t[1], counts_out=counts[8:10])
            return dict(preds_clean=preds_clean, preds_bd=preds_bd, x_bd=x_bd, cm_clean=cm_clean, cm_bd=cm_bd, **extra)

        dbg = None
        if use_graph:
            if eb["graph"] is None:
                launch()
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    eb["keep"] = launch()
                eb["graph"] = g
            else:
                eb["graph"].replay()
        else:
            dbg = launch()
        return {"counts": eb["counts"], "n_bd": int(ntrg.sum()), "n": B, "sigma": sigma, "sigma2": sigma2, "debug": dbg}

    class _Pending:
        """Handle of an asynchronous device->host read of a step's scalars (stream-ordered right after that step)."""

        def __init__(self, l, c, ev):
            self._l, self._c, self._ev = l, c, ev

        def get(self) -> dict:
            self._ev.synchronize()
            return AlternatedStep._to_dict(self._l.numpy().copy(), self._c.numpy().copy())

    def read_async(self, out):
        """Enqueue the device->host copy of this step's losses / counters into pinned memory and return a handle; calling
        `.get()` one iteration later keeps the host from draining the GPU queue after every step."""
        slot = self._rd_slot
        self._rd_slot = slot ^ 1
        if self._rd_bufs is None:
            self._rd_bufs = [(torch.empty(N_LOSSES, dtype=torch.float32).pin_memory(), torch.empty(16, dtype=torch.int32).pin_memory())
                             for _ in range(2)]
        l, c = self._rd_bufs[slot]
        l.copy_(out["losses"], non_blocking=True)
        c.copy_(out["counts"], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return AlternatedStep._Pending(l, c, ev)

    @staticmethod
    def _to_dict(l, c) -> dict:
        return dict(loss_c=float(l[0]), loss_ce=float(l[1]), loss_l2=float(l[2]), clean_model_loss=float(l[3]),
                    loss_grad_l2=float(l[7]), loss_tv=float(l[8]), loss_cross=float(l[9]), n_cross_correct=int(c[12]),
                    n_total_correct=int(c[0]), n_clean_model_correct=int(c[2]), n_clean_correct=int(c[4]),
                    n_bd_correct=int(c[6]), n_clean_model_bd_ba=int(c[8]), n_clean_model_bd_asr=int(c[9]),
                    n_F_correct=int(c[10]))

    @staticmethod
    def unpack(out) -> dict:
        """One D2H read of the step's scalars (call sparingly)."""
        return AlternatedStep._to_dict(out["losses"].cpu().numpy(), out["counts"].cpu().numpy())
