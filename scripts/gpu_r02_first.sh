#!/bin/bash
# tensor-pipe first conv (3 -> 64): parity, then step A/B
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "conv_tc_first" > gpurun_out/t_first_kernels.log 2>&1; rc=$?; echo "first-conv kernel tests rc=$rc"; tail -3 gpurun_out/t_first_kernels.log
[ $rc -ne 0 ] && { grep -E "^E |^FAILED|Error" gpurun_out/t_first_kernels.log | head -8; exit 1; }
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_first_all.log 2>&1; echo "all gpu tests rc=$?"; tail -4 gpurun_out/t_first_all.log; grep -E "^E " gpurun_out/t_first_all.log | head -5
timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench rc=$?"
COMBAT_NO_TC_FIRST2=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/bench_nofirst.json 2> gpurun_out/bench_nofirst.err
python - <<'PY'
import json
for n in ("first", "nofirst"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "conv frac %.4f" % d["roofline"]["frac"], "conv ms %.3f" % d["roofline"]["conv_ms_per_step"], d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
