for lv in 99 1 0; do
  for w in synthetic teapot bunny; do
    RT_SORT_MIN_LEVEL=$lv python bench.py --workload $w --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ab_sort${lv}_$w.json 2>/dev/null
  done
done
