mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dp_gpu.py -q -m gpu --timeout 250 > gpurun_out/t_dp.log 2>&1; echo "dp tests rc=$?"; tail -3 gpurun_out/t_dp.log
