mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n4 rc=$?"
cut -c1-330 gpurun_out/bench_n8.json; tail -2 gpurun_out/bench_n8.err
