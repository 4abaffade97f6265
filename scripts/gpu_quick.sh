mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 400 -x > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --dump-layers gpurun_out/conv_layers.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/t_all.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
