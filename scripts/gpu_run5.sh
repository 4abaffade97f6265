mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; echo "bench nograph rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
