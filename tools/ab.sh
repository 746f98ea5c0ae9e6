#!/bin/bash
# A/B of kernel build variants on one GPU box: tools/ab.sh <workload> <libdir> [<libdir> ...]
# Each variant is a directory under cs184-raytracer_b200/ built with `make LIBDIR=<dir> EXTRA=-D...`.
w=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  RT_B200_LIB_DIR=$PWD/cs184-raytracer_b200/$v python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --configs none \
      > gpurun_out/ab_${w}_$v.json 2> gpurun_out/ab_${w}_$v.err
  python - "$v" gpurun_out/ab_${w}_$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = d["roofline"].get("ms_kernel_serial_frame") or d["roofline"]["ms_kernel_per_step"]
    print(f"{sys.argv[1]:10s} {d['value']:8.1f} Mrays/s  {d['ms_per_step']:8.2f} ms  " + "  ".join(f"{a} {b:.1f}" for a, b in k.items()))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
