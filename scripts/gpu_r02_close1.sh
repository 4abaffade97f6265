#!/bin/bash
# round-2 closing evidence, call 1: all GPU tests, smoke, the bench line (with sub-records and the per-layer conv table),
# the reference arm, the ncu launch list of one eager step, and the ncu --set full capture of the dominant conv kernels
set -u
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --dump-layers gpurun_out/conv_layers.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/summary.txt
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-sub"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_r02_close.csv $CMD > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?" >> gpurun_out/summary.txt
CMD2='python scripts/bench_conv_ncu.py'
rm -f gpurun_out/prof_r02_conv_close.ncu-rep
timeout 120 $CMD2 > gpurun_out/plain_ncu_conv.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv_tc' -s 6 -c 3 -f -o gpurun_out/prof_r02_conv_close $CMD2 > gpurun_out/ncu_conv.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_all.log; cut -c1-600 gpurun_out/bench.json; cut -c1-300 gpurun_out/bench_ref.json
