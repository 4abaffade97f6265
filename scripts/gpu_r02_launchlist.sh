#!/bin/bash
# ncu launch list (per-launch device time, cold-cache, serialised) of one eager alternated step at batch 512
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-sub"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_r02b.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu.log; wc -l gpurun_out/launches_r02b.csv
