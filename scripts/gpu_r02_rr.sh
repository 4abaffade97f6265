#!/bin/bash
# row-reuse two-tile kernel for 128 output channels + bf16 pre-normalisation storage: parity, micro-benchmark, step
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "conv_tc or tc" > gpurun_out/t_rr_kernels.log 2>&1; rc=$?; echo "tc kernel tests rc=$rc"; tail -3 gpurun_out/t_rr_kernels.log
[ $rc -ne 0 ] && { grep -E "^E |^FAILED|Error" gpurun_out/t_rr_kernels.log | head -8; exit 1; }
timeout 300 python scripts/bench_conv.py --only 128x16 > gpurun_out/conv_rr_128x16.txt 2>&1; cat gpurun_out/conv_rr_128x16.txt
COMBAT_TC_DBG=1 timeout 120 python scripts/bench_conv.py --only 128x16 --variant "fwd bn only" 2>&1 | tail -1
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/t_rr_all.log 2>&1; echo "all gpu tests rc=$?"; tail -4 gpurun_out/t_rr_all.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline --dump-layers gpurun_out/conv_layers_rr.txt > gpurun_out/bench_rr.json 2> gpurun_out/bench_rr.err; echo "bench rc=$?"
COMBAT_NO_RR=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/bench_norr.json 2> gpurun_out/bench_norr.err
python - <<'PY'
import json
for n in ("rr", "norr"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "conv frac %.4f" % d["roofline"]["frac"], "conv ms %.3f" % d["roofline"]["conv_ms_per_step"], d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
