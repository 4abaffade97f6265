#!/bin/bash
# the input-aware variant's GPU tests (engine step eager / graph / bf16 / transforms, eval, public API vs the reference fixture,
# victim evaluator) + the reference arm with the unmodified reference from baseline/_ref
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "inputaware" --timeout 400 > gpurun_out/t_inputaware.log 2>&1; echo "inputaware tests rc=$?"
tail -25 gpurun_out/t_inputaware.log
ls baseline/_ref | head -3
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err; echo "bench ref rc=$?"
cut -c1-1500 gpurun_out/bench_ref2.json; tail -3 gpurun_out/bench_ref2.err
