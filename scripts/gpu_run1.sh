mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not conv_tc" --timeout 300 -x > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_tc" --timeout 120 > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_nets_gpu.py -q -m gpu -k "fp32 or golden or shipped" --timeout 300 > gpurun_out/t_nets_fp32.log 2>&1; echo "nets fp32 rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_step_gpu.py -q -m gpu -k "fp32 or known" --timeout 400 > gpurun_out/t_step_fp32.log 2>&1; echo "step fp32 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
