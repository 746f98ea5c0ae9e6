#!/bin/bash
set -x
RT_B200_LIB_DIR=$PWD/cs184-raytracer_b200/libvskip python -m pytest tests -x -q -m gpu -k "not cli and not pathb" > gpurun_out/r02_c8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c8_pytest.log
for w in synthetic teapot bunny refraction3; do bash tools/ab.sh $w libvbb libvskip; done > gpurun_out/r02_c8_ab.txt 2>&1
grep -v "^+" gpurun_out/r02_c8_ab.txt; tail -3 gpurun_out/r02_c8_pytest.log
