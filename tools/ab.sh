for v in $(cd cs184-raytracer_b200 && ls -d lib libv* 2>/dev/null); do
  for w in synthetic teapot refraction3; do
  RT_B200_LIB_DIR=$PWD/cs184-raytracer_b200/$v python bench.py --workload $w --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ab_${v}_$w.json 2>/dev/null
  done
done
