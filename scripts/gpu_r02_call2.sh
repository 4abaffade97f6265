#!/bin/bash
# Round 2, second GPU call: fixed tests, the per-tensor bf16 parity measurement (dumped to gpurun_out/parity.jsonl), the
# rewritten transform kernel's bandwidth, a longer run of the eager-torch comparator.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
for t in tests/test_graph_nosync_gpu.py tests/test_post_transform_gpu.py tests/test_api_gpu.py; do
  timeout 900 python -m pytest $t -x -q -m gpu > gpurun_out/t_$(basename $t .py).log 2>&1; echo "$t rc=$?"
done
COMBAT_PARITY_DUMP=gpurun_out/parity.jsonl COMBAT_GRAD_TOL=1.0 timeout 1200 python -m pytest tests/test_bf16_parity_gpu.py -q -m gpu -s > gpurun_out/t_bf16_parity.log 2>&1; echo "bf16 parity rc=$?"
timeout 600 python scripts/bench_eager_torch.py --steps 10 --warmup 5 > gpurun_out/eager_b512.json 2> gpurun_out/eager.err; echo "eager rc=$?"
timeout 600 python scripts/bench_eager_torch.py --steps 10 --warmup 5 --batch 128 > gpurun_out/eager_b128.json 2>> gpurun_out/eager.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/t_bf16_parity.log
