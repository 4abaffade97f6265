#!/bin/bash
# programmatic dependent launch on every kernel: correctness (all GPU tests), then A/B of the step with / without it
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py -x -q -m gpu > gpurun_out/t_pdl_quick.log 2>&1; rc=$?; echo "quick tests rc=$rc"; tail -3 gpurun_out/t_pdl_quick.log
[ $rc -ne 0 ] && exit 1
timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; echo "bench pdl rc=$?"
COMBAT_NO_PDL=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "bench nopdl rc=$?"
python - <<'PY'
import json
for n in ("pdl", "nopdl"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "conv frac %.4f" % d["roofline"]["frac"], "launches/step", d["gpu_launches"] / d["steps"])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"; tail -4 gpurun_out/t_all.log
