mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/summary.txt
timeout 300 python -m combat_b200.train_generator --debug --bs 128 --post_transform_option no_use > gpurun_out/cli.log 2>&1; echo "cli rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_all.log; cut -c1-300 gpurun_out/bench.json; cut -c1-300 gpurun_out/bench_ref.json; tail -3 gpurun_out/cli.log
