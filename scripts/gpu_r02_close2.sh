#!/bin/bash
# round-2 closing evidence, call 2 (after the input-aware / WaNet / multilabel-eval additions and the prefetch fix):
# all GPU tests, smoke, the bench line with sub-records, the reference arm (unmodified reference from baseline/_ref)
set -u
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --dump-layers gpurun_out/conv_layers.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_all.log; cut -c1-400 gpurun_out/bench.json; cut -c1-300 gpurun_out/bench_ref.json
