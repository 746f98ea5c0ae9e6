#!/bin/bash
# round-2 call 5 (1 GPU): why does a rank's 1/8 share run 20 % slower per ray?  ncu raw metrics of levels 0..1 of the emulated rank
set -x
export RT_BENCH_EMULATE_RANK=0/8
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --configs none"
$B > gpurun_out/r02_c5_plain.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:"k_trace|k_shadow" -s 12 -c 4 -o gpurun_out/r02_emu8 $B > gpurun_out/r02_c5_ncu.log 2>&1
ncu -i gpurun_out/r02_emu8.ncu-rep --page raw --csv > gpurun_out/r02_emu8_raw.csv 2>/dev/null
rm -f gpurun_out/r02_emu8.ncu-rep
unset RT_BENCH_EMULATE_RANK
ncu --set full --clock-control none -k regex:"k_trace|k_shadow" -s 12 -c 4 -o gpurun_out/r02_full $B > gpurun_out/r02_c5_ncu2.log 2>&1
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_full.ncu-rep
# the new default library (split queue regions) on the headline + refraction3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --configs none > gpurun_out/r02_c5_synth.json 2>/dev/null
python bench.py --workload refraction3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --configs none > gpurun_out/r02_c5_refr3.json 2>/dev/null
python tools/walltime.py --skip-reference --inputs 01,05,08 > gpurun_out/r02_c5_walltime.json 2> gpurun_out/r02_c5_walltime.err
