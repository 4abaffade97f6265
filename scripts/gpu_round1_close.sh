# Closing validation of round 1 after the DCT rewrite: all GPU tests, smoke, the bench line, the DCT kernels at
# BASELINE configs[4] size, the alternated step at the other BASELINE shapes, one ncu --set full capture of the DCT kernels.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 200 python scripts/bench_dct.py > gpurun_out/dct_kernels.jsonl 2>&1; echo "dct rc=$?" >> gpurun_out/summary.txt
timeout 300 python scripts/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?" >> gpurun_out/summary.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dct -c 3 -f -o gpurun_out/prof_dct python scripts/ncu_dct.py > gpurun_out/ncu_dct.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -2 gpurun_out/t_all.log; cut -c1-250 gpurun_out/bench.json; cut -c1-200 gpurun_out/dct_kernels.jsonl; cut -c1-400 gpurun_out/configs.jsonl; tail -2 gpurun_out/configs.err
