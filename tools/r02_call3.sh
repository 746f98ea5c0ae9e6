#!/bin/bash
# round-2 call 3 (2 GPUs): new tests on two real devices, 2-rank bench (peer / nccl A-B), wall-time table
set -x
python -m pytest tests/test_gpu_round2.py -x -q -m gpu > gpurun_out/r02_c3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c3_pytest.log
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$R bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_c3_n2_peer.json 2> gpurun_out/r02_c3_n2_peer.err
$R bench.py --gpus 2 --steps 10 --warmup 3 --assemble nccl --configs none > gpurun_out/r02_c3_n2_nccl.json 2> gpurun_out/r02_c3_n2_nccl.err
python tools/walltime.py > gpurun_out/r02_c3_walltime.json 2> gpurun_out/r02_c3_walltime.err
tail -c 400 gpurun_out/r02_c3_pytest.log
