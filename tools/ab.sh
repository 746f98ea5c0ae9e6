for v in $(cd cs184-raytracer_b200 && ls -d lib libv*); do
  RT_B200_LIB_DIR=$PWD/cs184-raytracer_b200/$v python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ab_$v.json 2>/dev/null
done
