# DCT kernels: parity tests, then the bench at BASELINE configs[4] sizes; optional extra commands as arguments.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/t_kernels.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_kernels.log
timeout 300 python scripts/bench_dct.py > gpurun_out/dct_kernels.jsonl 2>&1; echo "bench rc=$?"; cut -c1-175 gpurun_out/dct_kernels.jsonl
for extra in "$@"; do eval "$extra"; done
