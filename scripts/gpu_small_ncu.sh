mkdir -p gpurun_out
python scripts/bench_small.py > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:"conv_cin3_k|conv_cout3_k|wgrad3_k" -s 3 -c 1 -f -o gpurun_out/prof_cin3 python scripts/bench_small.py "conv_cin3 f32 out" > gpurun_out/ncu_small.log 2>&1
ncu --set full --clock-control none -k regex:"conv_cin3_k|conv_cout3_k|wgrad3_k" -s 3 -c 1 -f -o gpurun_out/prof_cout3 python scripts/bench_small.py "conv_cout3" >> gpurun_out/ncu_small.log 2>&1
ncu --set full --clock-control none -k regex:"conv_cin3_k|conv_cout3_k|wgrad3_k" -s 3 -c 1 -f -o gpurun_out/prof_wgrad3 python scripts/bench_small.py "wgrad_cin3" >> gpurun_out/ncu_small.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_small.log; ls -la gpurun_out/*.ncu-rep
