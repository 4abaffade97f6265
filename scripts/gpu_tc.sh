mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_tc" > gpurun_out/t_tc.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|Error|^E " gpurun_out/t_tc.log | head -10
python scripts/bench_conv.py --only 64x32 > gpurun_out/bench_conv_64.txt 2>&1; grep -E "wgrad|plain f32" gpurun_out/bench_conv_64.txt
