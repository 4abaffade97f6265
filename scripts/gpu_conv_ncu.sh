mkdir -p gpurun_out
python scripts/bench_conv.py > gpurun_out/bench_conv.txt 2>&1; cat gpurun_out/bench_conv.txt
CMD="python scripts/bench_conv.py --only 64x32 --iters 4"
$CMD > gpurun_out/plain_conv.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc64 -s 6 -c 3 -f -o gpurun_out/prof_tc64 $CMD > gpurun_out/ncu_conv.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_conv.log
