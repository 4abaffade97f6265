#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "conv_tc" > gpurun_out/t_kernels_tc.log 2>&1; rc=$?; echo "tc kernel tests rc=$rc"; tail -3 gpurun_out/t_kernels_tc.log
[ $rc -ne 0 ] && exit 1
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"; tail -6 gpurun_out/t_all.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline --dump-layers gpurun_out/conv_layers_sc.txt > gpurun_out/bench_sc.json 2> gpurun_out/bench_sc.err; echo "bench rc=$?"
COMBAT_NO_FUSE_SC=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/bench_nosc.json 2> gpurun_out/bench_nosc.err
python - <<'PY'
import json
for n in ("sc", "nosc"):
    d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "conv frac %.4f" % d["roofline"]["frac"], "conv ms %.3f" % d["roofline"]["conv_ms_per_step"], "launches/step", d["gpu_launches"] / d["steps"])
PY
