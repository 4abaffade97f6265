#!/bin/bash
# r02 ncu --set full captures of the three dominant convolution kernels on the micro-benchmark shapes of the step
# (one ncu invocation per gpurun call is the pool rule: all three kernels are captured by ONE ncu run over one command).
set -u
mkdir -p gpurun_out; rm -f gpurun_out/prof_r02_conv.ncu-rep
CMD='python scripts/bench_conv_ncu.py'
$CMD > gpurun_out/plain_ncu_conv.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv_tc' -s 6 -c 3 -f -o gpurun_out/prof_r02_conv $CMD > gpurun_out/ncu_conv.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_conv.log; ls -la gpurun_out/*.ncu-rep
