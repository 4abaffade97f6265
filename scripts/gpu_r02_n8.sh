#!/bin/bash
# 8-GPU run: the bench line with its sub-records (CelebA multilabel, ImageNet-10 shape B=256 per GPU: BASELINE configs[3]),
# then the A/B of the overlapped vs in-stream gradient exchange, then N=4.
set -u
mkdir -p gpurun_out
run() { # name, nproc, extra env, extra args
  env $3 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 295$2$2 bench.py --gpus $2 --steps 20 --warmup 5 $4 > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; echo "$1 rc=$?"
}
run n8 8 "COMBAT_X=1" ""
run n8_nooverlap 8 "COMBAT_DP_OVERLAP=0" "--no-sub"
run n4 4 "COMBAT_X=1" "--no-sub"
run n4_nooverlap 4 "COMBAT_DP_OVERLAP=0" "--no-sub"
timeout 600 python bench.py --steps 20 --warmup 5 --no-sub --no-cpu-baseline > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err
python - <<'PY'
import json
for n in ("n1b", "n4", "n4_nooverlap", "n8", "n8_nooverlap"):
    try:
        d = json.loads(open("gpurun_out/bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"])
        for s in (d.get("sub") or {}).get("step_configs", []):
            print("   ", s.get("case", "")[:60], s.get("images_per_s"), s.get("ms_per_step"), s.get("error"))
    except Exception as e:
        print(n, "failed", e)
PY
