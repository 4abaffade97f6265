#!/usr/bin/env python
"""bench.py -- alternated-step images/sec at CIFAR-10 shape (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours      : one alternated generator/surrogate iteration (train_generator.py:170-255 semantics, all six metric
            forwards included) per step on N B200s, bf16 tensor-core path, batch 512 per GPU (BASELINE configs[1]; weak
            scaling for N > 1 with NCCL gradient all-reduce).  `value` = device-resident inputs, CUDA-graph replay;
            `e2e` = the same metric through the public API with HOST inputs (pinned H2D of the batch and D2H of the
            step's losses/counters inside the timed region).
reference : the reference's algorithm on the box's host cores (oracle/combat_oracle.py -- the CPU restatement pinned
            to the unmodified reference; /root/reference does not exist on the GPU box), same metric.
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


WORKLOAD = ("CIFAR-10 shape alternated step (train_generator.py:170-255), PreActResNet18 + UnetGenerator, pc=0.5 noise_rate=0.08, "
            "bf16 tcgen05 convs, all metric forwards included (BASELINE configs[1])")


def workload_config(B, world, use_graph):
    return {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
            "cuda_graph": use_graph,
            "l2_policy": "working set per step (several GB of activations) is far larger than the 126 MB L2",
            "flops_per_image": FLOPS_PER_IMG_FULL}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # size the per-step sample so that the whole run ends within a few minutes
    r32, t32 = cpu_reference_rate(32, 1, 0)
    budget = 150.0
    batch = 128
    while batch > 16 and (args.steps + args.warmup) * t32 * batch / 32 > budget:
        batch //= 2
    rate, tstep = cpu_reference_rate(batch, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "alternated-step images/sec at CIFAR-10 shape", "value": rate, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tstep * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same workload as the CUDA arm; every timed step is a bounded SAMPLE of it (one alternated step over `batch` images of
        # the same synthetic distribution, fp32, all host threads) -- images/s is batch-size independent on the CPU
        "config": dict(workload_config(args.batch, max(1, args.gpus), not args.no_graph), reference_sample_batch=batch),
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d alternated steps of batch %d on the host cores (oracle/combat_oracle.py)" % (args.steps, batch)},
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
    d2h = 8 * 4 + 16 * 4

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
        tot_ms = sum(ms_ for _, _, ms_, _ in prof)
        tot_fl = sum(f for _, f, _, _ in prof)
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
        # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/r01_ncu_traffic.json):
        # per launch of the most frequent shape of conv_tc_kernel<128,0>, next to its algorithmic bytes
        traffic, traffic_of = None, None
        tp = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
        if os.path.exists(tp):
            kname, kd = next(iter(json.load(open(tp))["kernels"].items()))
            traffic, traffic_of = kd["dram_bytes_read"] + kd["dram_bytes_write"], {"kernel": kname, "algorithmic_bytes": kd["algorithmic_bytes"]}
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel / conv_tc64_kernel / conv_tc_wgrad*_kernel (tcgen05 implicit GEMM)", "achieved": ach,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic, "traffic_of": traffic_of,
                "peak_source": which + " (sustained bf16)",
                "launches": len(prof), "conv_ms_per_step": tot_ms,
                "by_kind": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0, "ms": v[1], "launches": v[2]} for k, v in by.items()},
                "step_frac_of_tensor_peak": (value / world) * FLOPS_PER_IMG_FULL / 1e12 / peak_tf}

    # ---- CPU baseline (rank 0, N == 1 only): the reference algorithm on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, tstep = cpu_reference_rate(128, 2, 1)
        cpu = {"value": rate, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "2 alternated steps of batch 128 (BASELINE configs[0]) after 1 warm-up, oracle/combat_oracle.py on the host cores"}
    if rank == 0:
        line = {
            "metric": "alternated-step images/sec at CIFAR-10 shape", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(B, world, use_graph),
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu,
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
    ap.add_argument("--dump-layers", default=None, help="write the per-layer tcgen05 conv timing table to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
