mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_api_gpu.py tests/test_step_gpu.py -q -m gpu -x -k "known_answer or graph or multilabel_train" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e img/s', d['e2e']['value'])"
