#!/bin/bash
# Round 2, third GPU call: first run of the CTA-pair (cta_group::2) convolution kernel -- parity first (under timeout: a protocol
# slip hangs), then the micro-benchmark A/B against the single-CTA kernel, then the reworked bf16 parity tests and the step.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "conv_tc" > gpurun_out/t_pair_kernels.log 2>&1; rc=$?; echo "pair kernel tests rc=$rc"
tail -5 gpurun_out/t_pair_kernels.log
if [ $rc -eq 0 ]; then
  for sh in 128x16 256x8 512x4; do
    timeout 300 python scripts/bench_conv.py --only $sh > gpurun_out/conv_pair_$sh.txt 2>&1
    COMBAT_NO_PAIR=1 timeout 300 python scripts/bench_conv.py --only $sh > gpurun_out/conv_nopair_$sh.txt 2>&1
  done
  paste -d'|' gpurun_out/conv_pair_256x8.txt gpurun_out/conv_nopair_256x8.txt
  timeout 900 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-sub --dump-layers gpurun_out/conv_layers_pair.txt > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; echo "bench rc=$?"
  COMBAT_NO_PAIR=1 timeout 900 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/bench_nopair.json 2> gpurun_out/bench_nopair.err
  python - <<'PY'
import json
for n in ("pair", "nopair"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "ms/step %.3f" % d["ms_per_step"], "conv frac %.3f" % d["roofline"]["frac"], "conv ms %.3f" % d["roofline"]["conv_ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
fi
COMBAT_PARITY_DUMP=gpurun_out/parity.jsonl timeout 1500 python -m pytest tests/test_bf16_parity_gpu.py -q -m gpu -s > gpurun_out/t_bf16_parity.log 2>&1; echo "bf16 parity rc=$?"
tail -c 2500 gpurun_out/t_bf16_parity.log
