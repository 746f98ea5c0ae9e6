#!/bin/bash
set -x
RT_B200_LIB_DIR=$PWD/cs184-raytracer_b200/libvbb python -m pytest tests -x -q -m gpu -k "not cli and not pathb" > gpurun_out/r02_c7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c7_pytest.log
for w in synthetic teapot bunny; do bash tools/ab.sh $w lib libvbb; done > gpurun_out/r02_c7_ab.txt 2>&1
grep -v "^+" gpurun_out/r02_c7_ab.txt; tail -3 gpurun_out/r02_c7_pytest.log
