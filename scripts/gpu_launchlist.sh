mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
