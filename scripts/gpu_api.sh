mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_api_gpu.py -q -m gpu -x > gpurun_out/t_api.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|Error|^E " gpurun_out/t_api.log | head -20
