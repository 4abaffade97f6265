#!/bin/bash
# converged-warp issue (elect.sync) of the TMA / MMA warps: parity, then micro-benchmark and step
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "conv_tc or tc" > gpurun_out/t_issue_kernels.log 2>&1; rc=$?; echo "tc kernel tests rc=$rc"; tail -3 gpurun_out/t_issue_kernels.log
[ $rc -ne 0 ] && { grep -E "^E " gpurun_out/t_issue_kernels.log | head -5; exit 1; }
for sh in 64x32 128x16 256x8 512x4; do
  timeout 300 python scripts/bench_conv.py --only $sh > gpurun_out/conv_issue_$sh.txt 2>&1
  cat gpurun_out/conv_issue_$sh.txt
done
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_step_gpu.py tests/test_bf16_parity_gpu.py -x -q -m gpu > gpurun_out/t_issue_nets.log 2>&1; echo "nets/step tests rc=$?"; tail -3 gpurun_out/t_issue_nets.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline --dump-layers gpurun_out/conv_layers_issue.txt > gpurun_out/bench_issue.json 2> gpurun_out/bench_issue.err; echo "bench rc=$?"
COMBAT_PAIR128=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline --dump-layers gpurun_out/conv_layers_issue_pair128.txt > gpurun_out/bench_issue_pair128.json 2> gpurun_out/bench_issue_pair128.err
python - <<'PY'
import json
for n in ("issue", "issue_pair128"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "conv frac %.4f" % d["roofline"]["frac"], "conv ms %.3f" % d["roofline"]["conv_ms_per_step"], d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
