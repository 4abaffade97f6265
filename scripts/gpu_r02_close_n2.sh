#!/bin/bash
# round-2 closing evidence on 2 GPUs (gpurun --gpus 2): the data-parallel tests that the 1-GPU runs skip, the N = 2 bench line
# (weak scaling, 512 images per GPU; sub-record: the strong-scaling point of BASELINE configs[2]), the reference arm under torchrun
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -q -m gpu --timeout 400 > gpurun_out/t_dp.log 2>&1; echo "dp tests rc=$?"; tail -3 gpurun_out/t_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
cut -c1-300 gpurun_out/bench_n2.json
COMBAT_DP_OVERLAP=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-sub > gpurun_out/bench_n2_instream.json 2> gpurun_out/bench_n2_instream.err; echo "bench n2 in-stream rc=$?"
cut -c1-300 gpurun_out/bench_n2_instream.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2>/dev/null; echo "ref arm n2 rc=$?"; cut -c1-200 gpurun_out/bench_ref_n2.json
