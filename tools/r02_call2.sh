#!/bin/bash
# round-2 call 2: GPU suite on the refactored library + default bench (all configs)
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02_c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c2_smoke.log 2>&1
( time python bench.py ) > gpurun_out/r02_c2_bench.json 2> gpurun_out/r02_c2_bench.err
tail -c 600 gpurun_out/r02_c2_pytest.log
