# ncu launch lists (gpu__time_duration.sum) of one eager alternated iteration at the CelebA multilabel and ImageNet-10 shapes;
# each command first runs without ncu.  Summarise with scripts/summarize_launches.py.
mkdir -p gpurun_out
for c in "CelebA" "batch 64"; do
  tag=$(echo $c | tr -d ' ' | tr 'A-Z' 'a-z')
  timeout 60 python scripts/bench_configs.py "$c" --eager > gpurun_out/eager_$tag.json 2> gpurun_out/eager_$tag.err && \
  timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$tag.csv python scripts/bench_configs.py "$c" --eager > gpurun_out/ncu_$tag.log 2>&1
  echo "$c rc=$?"; cut -c1-200 gpurun_out/eager_$tag.json; wc -l gpurun_out/launches_$tag.csv
done
