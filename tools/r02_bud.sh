#!/bin/bash
set -x
export RT_B200_LIB_DIR=$PWD/cs184-raytracer_b200/libvbud
RT_TRACE_BUDGET=24 python -m pytest tests -x -q -m gpu -k "not cli and not pathb" > gpurun_out/r02_bud_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_bud_pytest.log
bash tools/ab_env.sh synthetic RT_TRACE_BUDGET 0 16 24 32 48 > gpurun_out/r02_bud_ab.txt 2>&1
RT_TRACE_BUDGET2=200 bash tools/ab_env.sh synthetic RT_TRACE_BUDGET 24 >> gpurun_out/r02_bud_ab.txt 2>&1
RT_TRACE_BUDGET_FIRST_LEVEL=0 bash tools/ab_env.sh synthetic RT_TRACE_BUDGET 32 >> gpurun_out/r02_bud_ab.txt 2>&1
for w in teapot bunny refraction3; do bash tools/ab_env.sh $w RT_TRACE_BUDGET 0 24; done >> gpurun_out/r02_bud_ab.txt 2>&1
RT_BENCH_EMULATE_RANK=0/8 bash tools/ab_env.sh synthetic RT_TRACE_BUDGET 0 24 >> gpurun_out/r02_bud_ab.txt 2>&1
grep -v "^+" gpurun_out/r02_bud_ab.txt; tail -3 gpurun_out/r02_bud_pytest.log
