mkdir -p gpurun_out
python scripts/bench_conv.py --only 64x32 --iters 4 --variant "bwd mask+add" > gpurun_out/plain_conv.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc64 -s 5 -c 1 -f -o gpurun_out/prof_tc64_m1 python scripts/bench_conv.py --only 64x32 --iters 4 --variant "bwd mask+add" > gpurun_out/ncu_conv.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_conv.log
