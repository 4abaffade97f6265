mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "small" > gpurun_out/t_small.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_small.log
python scripts/bench_small.py > gpurun_out/bench_small.txt 2>&1; cat gpurun_out/bench_small.txt
