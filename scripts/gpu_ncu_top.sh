mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
CMD='python scripts/bench_conv.py --only 128x16 --iters 4 --variant'
$CMD "fwd res+out+bn" > gpurun_out/plain_top.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 5 -c 1 -f -o gpurun_out/prof_tc128 $CMD "fwd res+out+bn" > gpurun_out/ncu_top.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_top.log; ls -la gpurun_out/*.ncu-rep
