mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_nets_gpu.py -q -m gpu --timeout 300 -s > gpurun_out/t_nets.log 2>&1; echo "nets rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_step_gpu.py -q -m gpu --timeout 500 -s > gpurun_out/t_step.log 2>&1; echo "step rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
