#!/bin/bash
# A/B at N = 2: number of buckets of the flat-gradient all-reduces (the persistent conv kernels leave the collectives no SM to overlap on)
set -u
mkdir -p gpurun_out
for nb in 1 2 4; do
  COMBAT_COMM_BUCKETS=$nb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$nb bench.py --gpus 2 --steps 30 --warmup 5 --no-sub > gpurun_out/bench_n2_b$nb.json 2> gpurun_out/bench_n2_b$nb.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n2_b$nb.json") if l.startswith("{")][-1])
print("buckets $nb: ms/step %.3f value %.0f e2e %.0f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
PY
done
