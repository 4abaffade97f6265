mkdir -p gpurun_out
for i in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('A (new default) ms', d['ms_per_step'])"
COMBAT_NO_TC_FIRST_WGRAD=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B (old) ms', d['ms_per_step'])"
done
