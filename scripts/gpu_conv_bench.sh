mkdir -p gpurun_out
python scripts/bench_conv.py "$@" > gpurun_out/bench_conv.txt 2>&1; cat gpurun_out/bench_conv.txt
