#!/bin/bash
# round-2 closing evidence, call 3: ncu launch list of the CelebA multilabel step (BASELINE configs[4]) and the ncu --set full
# capture of the WaNet warp kernels (one forward, one backward launch of scripts/bench_warp.py)
set -u
mkdir -p gpurun_out
timeout 90 python scripts/bench_configs.py "CelebA" --eager > gpurun_out/eager_celeba.json 2> gpurun_out/eager_celeba.err && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_celeba_r02.csv python scripts/bench_configs.py "CelebA" --eager > gpurun_out/ncu_celeba.log 2>&1
echo "celeba launch list rc=$?"; cut -c1-200 gpurun_out/eager_celeba.json; wc -l gpurun_out/launches_celeba_r02.csv
timeout 120 python scripts/bench_warp.py > gpurun_out/bench_warp.txt 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'wanet_warp' -s 12 -c 2 -f -o gpurun_out/prof_r02_warp python scripts/bench_warp.py > gpurun_out/ncu_warp.log 2>&1
echo "ncu warp rc=$?"; cat gpurun_out/bench_warp.txt; tail -2 gpurun_out/ncu_warp.log
