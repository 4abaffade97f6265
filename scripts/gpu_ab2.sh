mkdir -p gpurun_out
COMBAT_PRE_BF16=1 timeout 600 python -m pytest tests/test_nets_gpu.py tests/test_step_gpu.py tests/test_mstep_gpu.py -q -m gpu -s 2>&1 | grep -E "two-iteration|passed|failed|^E  |FAILED" | head -30
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pre fp32 ms', d['ms_per_step'])"
COMBAT_PRE_BF16=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pre bf16 ms', d['ms_per_step'])"
