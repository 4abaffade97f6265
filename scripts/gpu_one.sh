mkdir -p gpurun_out
timeout 600 python -m pytest "$@" -q -m gpu -x > gpurun_out/t_one.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|Error|^E " gpurun_out/t_one.log | head -20
