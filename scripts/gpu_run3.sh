mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python scripts/calib.py > gpurun_out/calib.log 2>&1; echo "calib rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
