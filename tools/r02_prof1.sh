#!/bin/bash
# round-2 call 1: sanity + per-line captures of the bounce-level traversal kernels (existing build)
set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
python -m pytest tests -x -q -m gpu > gpurun_out/r02_c1_pytest.log 2>&1
$B > gpurun_out/r02_c1_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 7 -c 1 -o gpurun_out/r02_trace_l1 $B > gpurun_out/r02_c1_ncu1.log 2>&1
ncu -i gpurun_out/r02_trace_l1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_trace_l1_src.csv 2>/dev/null
ncu -i gpurun_out/r02_trace_l1.ncu-rep --page raw --csv > gpurun_out/r02_trace_l1_raw.csv 2>/dev/null
rm -f gpurun_out/r02_trace_l1.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:k_shadow -s 7 -c 1 -o gpurun_out/r02_shadow_l1 $B > gpurun_out/r02_c1_ncu2.log 2>&1
ncu -i gpurun_out/r02_shadow_l1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02_shadow_l1_src.csv 2>/dev/null
ncu -i gpurun_out/r02_shadow_l1.ncu-rep --page raw --csv > gpurun_out/r02_shadow_l1_raw.csv 2>/dev/null
rm -f gpurun_out/r02_shadow_l1.ncu-rep
python bench.py --workload synthetic4k --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_c1_synth4k.json 2>&1
ls -la gpurun_out | head -50
