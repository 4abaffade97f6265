mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 240 python -m pytest tests/test_dp_gpu.py -q -m gpu --timeout 200 -x > gpurun_out/t_dp.log 2>&1; echo "dp tests rc=$?"
tail -5 gpurun_out/t_dp.log
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
