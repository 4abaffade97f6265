mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_mstep_gpu.py -q -m gpu -x -s > gpurun_out/t_mstep.log 2>&1; echo "tests rc=$?"; grep -E "two-iteration|passed|failed|Error|assert" gpurun_out/t_mstep.log | head -20
