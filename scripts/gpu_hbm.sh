mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "dct or blend" > gpurun_out/t_dct.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_dct.log
python scripts/bench_hbm_kernels.py > gpurun_out/hbm_kernels.jsonl 2>&1; cut -c1-200 gpurun_out/hbm_kernels.jsonl
