#!/bin/bash
# Round 2, first GPU call: everything that was written but never ran (detector trainer, padded conditional generator), the new
# staging-ring / PostTensorTransform / resume tests, the bench line with its sub-records and the eager-torch comparator.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
# new / previously unrun tests first, each in its own process so that one crash does not hide the others
for t in tests/test_graph_nosync_gpu.py tests/test_post_transform_gpu.py tests/test_detector_train_gpu.py tests/test_mstep_gpu.py tests/test_api_gpu.py; do
  timeout 900 python -m pytest $t -x -q -m gpu > gpurun_out/t_$(basename $t .py).log 2>&1; echo "$t rc=$?"
done
# the race, demonstrated: the same no-sync test on round 1's single unguarded staging buffer must FAIL
COMBAT_UNSAFE_PLAN_STAGING=1 timeout 600 python -m pytest tests/test_graph_nosync_gpu.py -q -m gpu -k without_host_sync > gpurun_out/t_race_unsafe.log 2>&1; echo "unsafe staging rc=$? (expected non-zero)"
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --dump-layers gpurun_out/conv_layers.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/t_all.log
