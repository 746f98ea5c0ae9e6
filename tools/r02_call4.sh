#!/bin/bash
# round-2 call 4 (1 GPU): round-2 tests, TEX-pipe A/B, concurrent sub-contexts per GPU (tail overlap)
set -x
python -m pytest tests/test_gpu_round2.py -x -q -m gpu > gpurun_out/r02_c4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c4_pytest.log
bash tools/ab.sh synthetic lib libvtex1 libvtex2 > gpurun_out/r02_c4_ab_tex.txt 2>&1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --configs none"
for sub in 1 2 3 4; do
  RT_BENCH_EMULATE_RANK=0/8 RT_BENCH_SUB=$sub $B > gpurun_out/r02_c4_emu8_sub$sub.json 2> gpurun_out/r02_c4_emu8_sub$sub.err
done
RT_BENCH_EMULATE_RANK=3/8 RT_BENCH_SUB=1 $B > gpurun_out/r02_c4_emu8r3_sub1.json 2> /dev/null
RT_BENCH_EMULATE_RANK=3/8 RT_BENCH_SUB=2 $B > gpurun_out/r02_c4_emu8r3_sub2.json 2> /dev/null
RT_BENCH_EMULATE_RANK=0/1 RT_BENCH_SUB=2 $B > gpurun_out/r02_c4_emu1_sub2.json 2> gpurun_out/r02_c4_emu1_sub2.err
cat gpurun_out/r02_c4_ab_tex.txt
tail -c 300 gpurun_out/r02_c4_pytest.log
