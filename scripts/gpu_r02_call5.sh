#!/bin/bash
# all GPU tests after the phase split (overlapped exchanges) + bench; then (2 GPUs) the DP parity test and the 2-GPU bench
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"; tail -4 gpurun_out/t_all.log
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 900 python -m pytest tests/test_dp_gpu.py -q -m gpu > gpurun_out/t_dp.log 2>&1; echo "dp tests rc=$?"; tail -3 gpurun_out/t_dp.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
  python -c "
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value']); print(d['sub'])"
fi
timeout 900 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
