#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_nets_gpu.py tests/test_step_gpu.py -x -q -m gpu > gpurun_out/t_bnfused.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -4 gpurun_out/t_bnfused.log
[ $rc -ne 0 ] && exit 1
timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/bench_bnf.json 2> gpurun_out/bench_bnf.err; echo "bench rc=$?"
COMBAT_NO_FUSED_BN_BWD=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/bench_nobnf.json 2> gpurun_out/bench_nobnf.err
python - <<'PY'
import json
for n in ("bnf", "nobnf"):
    d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "conv frac %.4f" % d["roofline"]["frac"], "launches/step", d["gpu_launches"] / d["steps"])
PY
