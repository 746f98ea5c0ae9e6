#!/bin/bash
# A/B of one environment variable on one GPU box: tools/ab_env.sh <workload> <VAR> <value> [<value> ...]
w=$1; var=$2; shift; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $var=$v python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --configs none \
      > gpurun_out/abenv_${w}_${var}_$v.json 2> gpurun_out/abenv_${w}_${var}_$v.err
  python - "$var=$v" gpurun_out/abenv_${w}_${var}_$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = d["roofline"].get("ms_kernel_serial_frame") or d["roofline"]["ms_kernel_per_step"]
    print(f"{sys.argv[1]:24s} {d['value']:8.1f} Mrays/s  {d['ms_per_step']:8.2f} ms  launches {d['gpu_launches']}  " + "  ".join(f"{a} {b:.1f}" for a, b in k.items()))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
