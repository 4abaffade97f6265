#!/bin/bash
set -u
mkdir -p gpurun_out
run() { env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 30 --warmup 5 --no-sub > gpurun_out/bench_n2_$1.json 2> gpurun_out/bench_n2_$1.err; echo "$1 rc=$?"; }
run reserve0 "COMBAT_DP_RESERVE_SMS=0" 29603
run nooverlap "COMBAT_DP_OVERLAP=0" 29604
timeout 600 python -m pytest tests/test_dp_gpu.py -q -m gpu > gpurun_out/t_dp.log 2>&1; echo "dp tests rc=$?"
python - <<'PY'
import json
    try:
        d = json.loads(open("gpurun_out/bench_n2_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"])
    except Exception as e:
        print(n, "failed", e)
PY
