#!/usr/bin/env python
"""bench.py -- alternated-step images/sec at CIFAR-10 shape (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours      : one alternated generator/surrogate iteration (train_generator.py:170-255 semantics, all six metric
            forwards included) per step on N B200s, bf16 tensor-core path, batch 512 per GPU (BASELINE configs[1]; weak
            scaling for N > 1 with NCCL gradient all-reduce).  `value` = device-resident inputs, CUDA-graph replay;
            `e2e` = the same metric through the public API with HOST inputs (pinned H2D of the batch and D2H of the
            step's losses/counters inside the timed region).
reference : the UNMODIFIED reference `train_generator.train()` on the box's host cores, from baseline/_ref/ (a git-ignored
            verbatim install of the reference's .py files made by build(), oracle/install_reference.py; it ships with the
            snapshot), BASELINE configs[0] (batch 128), same metric; the oracle port only if baseline/_ref is absent.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_C, F_G, F_F = 1.1107e9, 0.5234e9, 0.0772e9          # forward FLOPs per image
FLOPS_PER_IMG_UPDATES = 7 * F_C + 3 * F_G              # 9.35 GFLOP: what the parameter updates need
FLOPS_PER_IMG_FULL = 9 * F_C + 3 * F_G + F_F           # 11.65 GFLOP: + the three metric forwards and netF


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs: NVML every 10 ms (nvidia_ml_py), falling back to
    `nvidia-smi` polling when NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.mx, self.reasons = index, False, [], [], set()

    def _run_nvml(self):
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        self.mx.append(float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)))
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        while not self.stop_flag:
            self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            self.reasons |= {k for k, b in bits.items() if r & b}
            time.sleep(0.01)

    def _run_smi(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [c.strip() for c in out.split(",")]
                if len(r) >= 6:
                    self.sm.append(float(r[0]))
                    self.mx.append(float(r[1]))
                    self.reasons |= {self.NAMES[i] for i in range(4) if r[2 + i].lower().startswith("active")}
            except Exception:
                pass
            time.sleep(0.05)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def synthetic_state(seed):
    from oracle import combat_oracle as O
    gen = torch.Generator().manual_seed(seed)
    netC_p, netC_b = O.init_preact_resnet18_state(gen)
    clean_p, clean_b = O.init_preact_resnet18_state(gen)
    netG_p = O.init_unet_state(gen)
    netF_p, netF_b = O.init_frequency_model_state(gen)
    return dict(netC_p=netC_p, netC_b=netC_b, clean_p=clean_p, clean_b=clean_b, netG_p=netG_p, netF_p=netF_p, netF_b=netF_b,
                momC={}, momG={})


def cpu_reference_rate(batch, steps, warmup, seed=0):
    """images/sec of the reference algorithm (oracle restatement) on the host cores."""
    from oracle import combat_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    state = synthetic_state(seed)
    opt = O.default_opt()
    g = torch.Generator().manual_seed(seed + 1)
    times = []
    for i in range(warmup + steps):
        x = torch.rand(batch, 3, 32, 32, generator=g) * 2 - 1
        y = torch.randint(0, 10, (batch,), generator=g)
        t0 = time.perf_counter()
        O.alternated_step(state, x, y, opt)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), sum(times) / len(times)


def reference_installed():
    """baseline/_ref/ holds the UNMODIFIED reference (oracle/install_reference.py, run by build() in the build container)."""
    return os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "train_generator.py"))


def reference_train_rate(batch, steps, warmup, seed=0, anomaly=True):
    """images/sec of the UNMODIFIED reference `train_generator.train()` (baseline/_ref, `--device cpu --post_transform_option
    no_use`) on the host cores.  Nothing of the reference is patched: the per-step times are observed from
    OUTSIDE, by a loader object that notes the time each time train() asks for the next batch.  anomaly=False measures the
    same code with `torch.autograd.set_detect_anomaly` (which train() switches on at :147) made inert for the call."""
    import contextlib
    import random
    import tempfile

    from oracle.ref_loader import NullWriter, load_reference
    tg, config = load_reference(os.path.join(ROOT, "baseline", "_ref"))
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    opt = config.get_arguments().parse_args(["--device", "cpu", "--post_transform_option", "no_use", "--bs", str(batch)])
    opt.input_height = opt.input_width = 32
    opt.input_channel = 3
    netC, optC, schC, netG, optG, schG, netF, clean = tg.get_model(opt)
    netF.eval()
    clean.eval()

    class TimedBatches:
        def __init__(self, n):
            self.n, self.t = n, []

        def __len__(self):
            return self.n

        def __iter__(self):
            for _ in range(self.n):
                xy = (torch.rand(batch, 3, 32, 32) * 2 - 1, torch.randint(0, 10, (batch,)))
                self.t.append(time.perf_counter())
                yield xy
            self.t.append(time.perf_counter())

    dl = TimedBatches(warmup + steps)
    real_setter = torch.autograd.set_detect_anomaly
    cwd = os.getcwd()
    try:
        os.chdir(tempfile.mkdtemp())
        if not anomaly:
            torch.autograd.set_detect_anomaly = lambda *a, **k: None
        with open(os.devnull, "w") as null, contextlib.redirect_stdout(null):   # the reference's progress bar writes to stdout
            tg.train(netC, optC, schC, netG, optG, schG, netF, clean, dl, NullWriter(), 1, opt)
    finally:
        torch.autograd.set_detect_anomaly = real_setter
        real_setter(False)
        os.chdir(cwd)
    dt = np.diff(np.array(dl.t))[warmup:]
    return batch * len(dt) / float(dt.sum()), float(np.median(dt)), float(dt.mean())


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _timed_kernel(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def sub_hbm_kernels(dev, peak_bw):
    """BASELINE configs[4] + the north_star's HBM targets, timed alone with CUDA events on tensors far larger than the 126 MB
    L2: batched 2-D DCT / IDCT / low_freq over 65,536 CIFAR-shape and 16,384 CelebA-shape images (float32 and uint8 input),
    the fused poison-blend forward / backward over 65,536 rows, PostTensorTransform, the flat Nesterov-SGD update.
    bytes = ALGORITHMIC bytes: one read + one write of every plane, etc."""
    from combat_b200 import ops
    from combat_b200.utils.dataloader import draw_params
    out = []

    def rec(name, nbytes, sec, units=None):
        gbs = nbytes / sec / 1e9
        r = {"kernel": name, "bytes": nbytes, "us": round(sec * 1e6, 2), "GB/s": round(gbs, 1), "frac": round(gbs / peak_bw, 4)}
        if units:
            r["images_per_s"] = round(units / sec, 1)
        out.append(r)

    for N, NI, keep in ((32, 65536, 20), (64, 16384, 41)):
        x = torch.rand(NI, 3, N, N, device=dev) * 2 - 1
        o = torch.empty_like(x)
        for kind in ("dct", "idct", "lowfreq"):
            t = _timed_kernel(lambda: ops.plane_op(x, kind, keep=keep, out=o))
            rec("dct%d %s fp32 %dx3x%dx%d" % (N, kind, NI, N, N), 2 * x.numel() * 4, t, NI)
        xu = (torch.rand(NI, 3, N, N, device=dev) * 255).to(torch.uint8)
        t = _timed_kernel(lambda: ops.plane_op(xu, "dct", in_mode=1, out=o))
        rec("dct%d dct uint8-in %dx3x%dx%d" % (N, NI, N, N), x.numel() * 5, t, NI)
        # full-size correctness through size-independent properties (round trip, Parseval)
        X = ops.plane_op(x, "dct", out=o)
        e = float((x.double() ** 2).sum())
        pars = abs(float((X.double() ** 2).sum()) - e) / e
        rt = float((ops.plane_op(X, "idct") - x).abs().max())
        assert rt < 5e-6 and pars < 1e-6, (N, rt, pars)
        out[-1]["properties"] = {"idct(dct(x)) max abs err": rt, "Parseval rel err": pars}
        del x, o, xu, X
    B = 65536
    x = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
    noise = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
    o = torch.empty_like(x)
    taps = ops.gaussian_taps(0.6)
    sq = torch.empty(B * 3, device=dev)
    t = _timed_kernel(lambda: ops.poison_blend_fwd(x, noise, None, B, 0.08, taps, out=o, sq_partial=sq))
    rec("poison_blend_fwd (blend+clamp+blur+MSE partials) 65536 rows", 3 * x.numel() * 4, t, B)
    g1 = torch.randn(B, 3, 32, 32, device=dev)
    dn = torch.empty_like(x)
    t = _timed_kernel(lambda: ops.poison_blend_bwd(x, noise, o, g1, None, 1e-6, 0.08, taps, out=dn))
    rec("poison_blend_bwd 65536 rows", 5 * x.numel() * 4, t, B)
    opt_tf = argparse.Namespace(post_transform_option="use", random_crop=5, random_rotation=10, dataset="cifar10")
    import random as _r
    _r.seed(1)   # seed 1: crop, rotation and flip gates all on
    P = None
    while P is None or not (P[0, 4] == 1 and abs(P[:, 0]).sum() > 0):
        P = draw_params(B, opt_tf)
    Pd = torch.from_numpy(P).to(dev)
    t = _timed_kernel(lambda: ops.post_transform_fwd(x, Pd, out=o))
    rec("post_transform_fwd (crop+rotate+flip) 65536 rows", 2 * x.numel() * 4, t, B)
    # WaNet warp trigger (train_generator_wanet.py:196-203): flow -> bicubic table -> clamp -> bilinear gather, and its backward
    S_ = 2
    flow = torch.tanh(torch.randn(B, 2, S_, S_, device=dev))
    ident = torch.linspace(-1, 1, steps=32).to(dev)
    t = _timed_kernel(lambda: ops.wanet_warp_fwd(x, flow, ident, None, B, 0.15, S_, out=o, sq_partial=sq))
    rec("wanet_warp_fwd (bicubic flow + clamp + grid_sample) 65536 rows", 2 * x.numel() * 4, t, B)
    t = _timed_kernel(lambda: ops.wanet_warp_bwd(x, flow, ident, g1, None, 0.15, 1e-6, S_))
    rec("wanet_warp_bwd 65536 rows", 2 * x.numel() * 4, t, B)
    del x, noise, o, g1, dn, sq, flow
    n = 20541389
    p_, g_, m_ = (torch.randn(n, device=dev) for _ in range(3))
    lr = torch.full((1,), 1e-2, device=dev)
    t = _timed_kernel(lambda: ops.sgd_nesterov(p_, g_, m_, lr, 0.9, 5e-4, False))
    rec("sgd_nesterov 20.5M params", 20 * n, t)
    del p_, g_, m_
    torch.cuda.empty_cache()
    return out


def sub_step_config(name, dataset, S, ncls, B, multilabel, flops, lr, dev, rank, world, sync, barrier, peak_tf, steps=5, warmup=3,
                    post_transform="no_use"):
    """One of the OTHER step shapes BASELINE.json names, same engine, device-resident inputs, CUDA-graph replay, weak DP."""
    import torch.distributed as dist
    from combat_b200 import config
    from combat_b200 import train_generator as tg
    from combat_b200 import train_generator_multilabel as tgm
    from combat_b200.engine import AlternatedStep
    mod = tgm if multilabel else tg
    opt = config.get_arguments().parse_args(["--device", str(dev), "--post_transform_option", post_transform, "--dataset", dataset])
    opt.input_height = opt.input_width = S
    opt.input_channel, opt.num_classes = 3, ncls
    opt.lr_C = opt.lr_G = lr
    torch.manual_seed(0)
    netC, _, _, netG, _, _, netF, clean_model = mod.get_model(opt)
    eng = AlternatedStep(opt, device=dev, with_metrics=True, multilabel=multilabel,
                         nets=(netC.net, clean_model.net, netG.net, netF.net if netF is not None else None),
                         grad_hook=sync.grad_hook if world > 1 else None, buf_hook=sync.buf_hook if world > 1 else None)
    g = torch.Generator().manual_seed(7 + rank)
    xs = [(torch.rand(B, 3, S, S, generator=g) * 2 - 1).to(dev) for _ in range(2)]
    ys = [torch.randint(0, ncls, (B,), generator=g).numpy() for _ in range(2)]
    for i in range(warmup):
        out = eng.step(xs[i % 2], ys[i % 2], use_graph=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = eng.step(xs[i % 2], ys[i % 2], use_graph=True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    ms /= steps
    losses = [float(v) for v in out["losses"].cpu()]
    rate = world * B / (ms * 1e-3)
    r = {"case": name, "images_per_s": rate, "ms_per_step": ms, "batch_per_gpu": B, "n_gpus": world, "steps": steps, "warmup": warmup,
         "dtype": "bf16", "lr": lr, "post_transform_option": post_transform, "update_flops_per_image": flops,
         "step_frac_of_tensor_peak": rate / world * flops / 1e12 / peak_tf, "finite": bool(np.all(np.isfinite(losses))),
         "max_memory_GB": torch.cuda.max_memory_allocated() / 1e9, "launches_per_step": eng.launches_per_step}
    del eng, netC, netG, netF, clean_model, xs
    torch.cuda.empty_cache()
    return r


def gpu_eager_baseline(dev, batch=512, steps=5, warmup=5):
    """'PyTorch eager on the same B200' comparator: the reference's algorithm (oracle restatement:
    aten / cuDNN / cuFFT kernels, autograd, fp32) with every tensor on the GPU.  A reported baseline next to cpu_baseline;
    never the product path."""
    from oracle import combat_oracle as O
    res = {}
    for tf32 in (False, True):
        torch.backends.cudnn.benchmark = True
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = tf32
        gen = torch.Generator().manual_seed(0)
        st = synthetic_state(0)
        to = lambda d: {k: v.to(dev) for k, v in d.items()}
        state = {k: (to(v) if isinstance(v, dict) and k not in ("momC", "momG") else v) for k, v in st.items()}
        opt = O.default_opt()
        xs = [(torch.rand(batch, 3, 32, 32, generator=gen) * 2 - 1).to(dev) for _ in range(2)]
        ys = [torch.randint(0, 10, (batch,), generator=gen).to(dev) for _ in range(2)]
        times = []
        for i in range(warmup + steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            O.alternated_step(state, xs[i % 2], ys[i % 2], opt)
            torch.cuda.synchronize()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        ms = 1e3 * sum(times) / len(times)
        res["tf32" if tf32 else "fp32"] = {"value": batch / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms}
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    res.update(batch=batch, steps=steps, warmup=warmup,
               impl="oracle/combat_oracle.py with all tensors on cuda:0 (torch %s eager: cuDNN convs, cuFFT DCT, autograd)" % torch.__version__)
    torch.cuda.empty_cache()
    return res


WORKLOAD = ("CIFAR-10 shape alternated step (train_generator.py:170-255), PreActResNet18 + UnetGenerator, pc=0.5 noise_rate=0.08, "
            "bf16 tcgen05 convs, all metric forwards included (BASELINE configs[1])")


def workload_config(B, world, use_graph):
    return {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
            "cuda_graph": use_graph,
            "l2_policy": "working set per step (several GB of activations) is far larger than the 126 MB L2",
            "flops_per_image": FLOPS_PER_IMG_FULL}


def cpu_baseline_record(batch, steps, warmup, with_anomaly_off=True):
    """The reference arm / cpu_baseline measurement.  kind "reference": the UNMODIFIED reference train() from baseline/_ref
    (anomaly mode as shipped = ON is the value; the same code with anomaly mode off is reported beside it);
    kind "port": the oracle restatement (only when baseline/_ref did not travel)."""
    if reference_installed():
        rate, tmed, tmean = reference_train_rate(batch, steps, warmup, anomaly=True)
        rec = {"value": rate, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "reference", "cpu_model": cpu_model(),
               "sample": "%d iterations of the unmodified reference train_generator.train() at batch %d (BASELINE configs[0]) after %d "
                         "warm-up, --device cpu --post_transform_option no_use, anomaly mode as shipped (on), all host threads"
                         % (steps, batch, warmup),
               "median_s_per_step": tmed}
        if with_anomaly_off:
            r2, tm2, _ = reference_train_rate(batch, max(1, min(steps, 5)), min(warmup, 1), anomaly=False)
            rec["value_anomaly_off"] = r2
            rec["median_s_per_step_anomaly_off"] = tm2
        return rec, tmean
    rate, tstep = cpu_reference_rate(batch, steps, warmup)
    return {"value": rate, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "cpu_model": cpu_model(),
            "sample": "%d alternated steps of batch %d on the host cores (oracle/combat_oracle.py; baseline/_ref absent)"
                      % (steps, batch)}, tstep


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # BASELINE configs[0]: batch 128, every step a bounded sample of the CUDA arm's workload (same synthetic distribution, fp32,
    # all host threads); ~1.5 s per step on 16 cores, so K + W steps end within a few minutes for any sensible K
    batch = 128
    steps = min(args.steps, 40)
    cpu, tstep = cpu_baseline_record(batch, steps, min(args.warmup, 5))
    rate = cpu["value"]
    line = {
        "impl": "reference", "metric": "alternated-step images/sec at CIFAR-10 shape", "value": rate, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 5), "ms_per_step": tstep * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.batch, max(1, args.gpus), not args.no_graph), reference_sample_batch=batch),
        "cpu_baseline": cpu,
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch.distributed as dist

    from combat_b200 import config, nets, parallel
    from combat_b200 import train_generator as tg
    from combat_b200.engine import AlternatedStep
    rank, world, local = parallel.env_rank()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    parallel.init(device=dev)
    B = args.batch
    # the reference's own construction path (get_model, train_generator.py:80-128): random init in the reference's
    # order from seed 0 on every rank -> identical replicas; no checkpoint, no dataset (synthetic)
    opt = config.get_arguments().parse_args(["--device", str(dev), "--post_transform_option", "no_use"])
    opt.input_height = opt.input_width = 32
    opt.input_channel = 3
    torch.manual_seed(0)
    netC, _, _, netG, _, _, netF, clean_model = tg.get_model(opt)
    sync = parallel.GradSync()
    eng = AlternatedStep(opt, device=dev, with_metrics=True, nets=(netC.net, clean_model.net, netG.net, netF.net),
                         grad_hook=sync.grad_hook if world > 1 else None, buf_hook=sync.buf_hook if world > 1 else None)
    parallel.seed_rank(1234, rank)
    g = torch.Generator().manual_seed(99 + rank)
    n_host = 4
    xs_host = [(torch.rand(B, 3, 32, 32, generator=g) * 2 - 1).pin_memory() for _ in range(n_host)]
    ys_host = [torch.randint(0, 10, (B,), generator=g).numpy() for _ in range(n_host)]
    xs_dev = [x.to(dev) for x in xs_host]
    use_graph = not args.no_graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # ---- device-resident leg
    def step_resident(i):
        eng.step(xs_dev[i % n_host], ys_host[i % n_host], use_graph=use_graph)

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps)
    sampler.stop_flag = True
    value = world * B * args.steps / (ms / 1e3)
    launches = eng.launches_per_step * args.steps

    # ---- end-to-end leg: host batch in, host scalars out, every step
    sink = []
    pend = [None]

    def step_e2e(i):
        # the public loop (combat_b200.train_generator.train): this batch's pinned host tensor in, the NEXT batch's copy
        # started on the side stream while the graph runs, every step's scalars copied to the host (stream-ordered after
        # the step) and consumed one iteration later so that the host never drains the GPU queue
        out = eng.step(xs_host[i % n_host], ys_host[i % n_host], use_graph=use_graph)
        eng.prefetch(xs_host[(i + 1) % n_host])
        p, pend[0] = pend[0], eng.read_async(out)
        if p is not None:
            sink.append(p.get()["loss_c"])

    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    sink.append(pend[0].get()["loss_c"])  # the last step's scalars (their copy was enqueued inside the timed region)
    assert all(np.isfinite(v) for v in sink)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    h2d = xs_host[0].numel() * 4 + B * 8 * 3 + B * 4 + 20
    from combat_b200.engine import N_LOSSES
    d2h = N_LOSSES * 4 + 16 * 4

    # ---- roofline leg: per-launch CUDA-event timing of the tcgen05 convolutions in one eager iteration
    roof = None
    if rank == 0:
        # park the GPU behind a ~150 ms spin so that the host (eager launches through ctypes) runs ahead of it: each
        # event pair then brackets device time only, not the host-side launch latency of the kernel between them.
        # Two passes, per-launch minimum: a host hiccup (allocator growth, page fault) in one pass does not pollute it.
        passes = []
        eng.grad_hook = eng.buf_hook = None  # rank-0-only eager passes: no collective may be issued here (measurement is over)
        for _ in range(2):
            nets.TC_PROFILE = []
            torch.cuda._sleep(int(0.15 * 1.9e9))
            eng.step(xs_dev[0], ys_host[0], use_graph=False)
            torch.cuda.synchronize()
            passes.append([(n, f, s.elapsed_time(e), t) for n, f, s, e, t in nets.TC_PROFILE])
        nets.TC_PROFILE = None
        if len(passes[0]) == len(passes[1]):
            prof = [(a[0], a[1], min(a[2], b[2]), a[3]) for a, b in zip(*passes)]
        else:
            prof = passes[-1]
        # the K = 27 image-boundary launches (first conv, its weight gradient through im2col) are memory-bound data movement on the
        # tensor pipe: listed in by_kind, excluded from the family's roofline ("exclude K=27 first conv and Cout=3 last conv")
        tot_ms = sum(ms_ for n_, _, ms_, _ in prof if n_ != "conv_tc_boundary")
        tot_fl = sum(f for n_, f, _, _ in prof if n_ != "conv_tc_boundary")
        peak_tf, peak_bw, which = peaks()
        by = {}
        layers = {}
        for name, f, ms_, tag in prof:
            for d, k in ((by, name), (layers, name + " | " + tag)):
                a = d.setdefault(k, [0.0, 0.0, 0])
                a[0] += f
                a[1] += ms_
                a[2] += 1
        if args.dump_layers:
            with open(args.dump_layers, "w") as fh:
                for k, v in sorted(layers.items(), key=lambda kv: -kv[1][1]):
                    fh.write("%-60s launches %3d  ms %8.3f  TFLOP/s %7.1f\n" % (k, v[2], v[1], v[0] / (v[1] * 1e-3) / 1e12))
        ach = tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
        # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/r02_ncu_traffic.json):
        # per launch of conv_tc_kernel<128,0> at its most frequent shape, next to its algorithmic bytes (the bf16 output of
        # that launch, 33.5 MB, is still in the 126 MB L2 when the kernel ends: traffic < algorithmic, no wasted re-reads)
        traffic, traffic_of = None, None
        tp = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
        if os.path.exists(tp):
            kname, kd = next(iter(json.load(open(tp))["kernels"].items()))
            traffic, traffic_of = kd["dram_bytes_read"] + kd["dram_bytes_write"], {"kernel": kname, "algorithmic_bytes": kd["algorithmic_bytes"]}
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel / conv_tc64_kernel / conv_tc_wgrad*_kernel (tcgen05 implicit GEMM)", "achieved": ach,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic, "traffic_of": traffic_of,
                "peak_source": which + " (sustained bf16)",
                "launches": sum(1 for n_, _, _, _ in prof if n_ != "conv_tc_boundary"), "conv_ms_per_step": tot_ms,
                "boundary_ms_per_step": sum(ms_ for n_, _, ms_, _ in prof if n_ == "conv_tc_boundary"),
                "by_kind": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0, "ms": v[1], "launches": v[2]} for k, v in by.items()},
                "step_frac_of_tensor_peak": (value / world) * FLOPS_PER_IMG_FULL / 1e12 / peak_tf}

    # ---- CPU baseline (rank 0, N == 1 only): the reference algorithm on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_baseline_record(128, 3, 1, with_anomaly_off=False)
    # ---- sub-records: the other targets BASELINE.json / north_star name, measured by the same process (not the headline)
    sub = {}
    if not args.no_sub:
        peak_tf, peak_bw, _ = peaks()
        torch.cuda.empty_cache()
        if rank == 0 and world == 1:
            sub["hbm_kernels"] = sub_hbm_kernels(dev, peak_bw)
        steps_cfg = []
        cases = [("configs[4] CelebA 64x64 multilabel step (ResNet18(8) + CUnetGeneratorv1), batch 256 per GPU", "celeba", 64, 8, 256, True, 37.4e9, 1e-2)]
        if world in (1, 8):
            # 224x224 with the scaler-49 head diverges within a few iterations at the reference's lr 1e-2 on random data (the
            # reference itself raises KeyError for this size); throughput does not depend on lr: timed at 1e-4, stated here
            cases.append(("configs[3] ImageNet-10 shape 224x224 (ResNet18 scaler-49 + UnetGenerator), batch 256 per GPU", "imagenet10", 224, 10, 256, False, 458e9, 1e-4))
        if world == 1:
            cases.append(("configs[1] with the reference's DEFAULT --post_transform_option use, batch 512", "cifar10", 32, 10, 512, False, 9.35e9, 1e-2))
        if world in (2, 4):
            cases.append(("configs[2] CIFAR-10 STRONG scaling point: global batch 4096 (%d per GPU)" % (4096 // world), "cifar10", 32, 10, 4096 // world, False, 9.35e9, 1e-2))
        for c in cases:
            try:
                torch.cuda.reset_peak_memory_stats()
                ptf = "use" if "DEFAULT" in c[0] else "no_use"
                steps_cfg.append(sub_step_config(*c, dev, rank, world, sync, barrier, peak_tf, post_transform=ptf))
            except Exception as e:  # a failing extra shape must not cost the headline line
                steps_cfg.append({"case": c[0], "error": "%s: %s" % (type(e).__name__, str(e)[:300])})
        sub["step_configs"] = steps_cfg
    eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            eager = gpu_eager_baseline(dev)
        except Exception as e:
            eager = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    if rank == 0:
        line = {
            "metric": "alternated-step images/sec at CIFAR-10 shape", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(B, world, use_graph),
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu,
            "gpu_eager_baseline": eager, "sub": sub,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="batch per GPU")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (HBM kernels, other step shapes)")
    ap.add_argument("--dump-layers", default=None, help="write the per-layer tcgen05 conv timing table to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
