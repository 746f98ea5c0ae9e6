#!/bin/bash
# round-2 call 6 (1 GPU): cheap knobs on a rank's 1/8 share (emulated rank 0/8)
set -x
export RT_BENCH_EMULATE_RANK=0/8
bash tools/ab_env.sh synthetic RT_SHADOW_MODE 1 0 > gpurun_out/r02_c6_knobs.txt 2>&1
bash tools/ab_env.sh synthetic RT_HIT_SORT_BITS 32 24 16 >> gpurun_out/r02_c6_knobs.txt 2>&1
bash tools/ab_env.sh synthetic RT_HIT_SORT_MIN_RAYS 262144 65536 1000000 >> gpurun_out/r02_c6_knobs.txt 2>&1
unset RT_BENCH_EMULATE_RANK
bash tools/ab_env.sh synthetic RT_SHADOW_MODE 1 0 >> gpurun_out/r02_c6_knobs.txt 2>&1
grep -v "^+" gpurun_out/r02_c6_knobs.txt
