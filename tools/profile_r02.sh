#!/bin/bash
# Round-2 evidence run on one B200: GPU suite, smoke, default bench, ncu launch list, ncu --set full of every traversal
# launch of one frame (converted to CSV on the box), whole-process wall-time table.  Outputs under gpurun_out/.
set -x
python -m pytest tests -q -m gpu > gpurun_out/r02_final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1
python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --configs none"
$B > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_prof_ncu_launches.log 2>&1
B1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --configs none"
$B1 > gpurun_out/r02_prof_plain1.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_trace|k_shadow" -s 12 -c 12 -o gpurun_out/r02_trav $B1 > gpurun_out/r02_prof_ncu_full.log 2>&1
ncu -i gpurun_out/r02_trav.ncu-rep --page raw --csv > gpurun_out/r02_trav_raw.csv 2>/dev/null
rm -f gpurun_out/r02_trav.ncu-rep
python tools/walltime.py > gpurun_out/r02_walltime.json 2> gpurun_out/r02_walltime.err
tail -3 gpurun_out/r02_final_pytest.log
