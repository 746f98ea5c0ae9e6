# 1/2/4/8-GPU scaling of the five BASELINE configs on one box (run under gpurun --gpus 8): bench.py at every N,
# launched the way the driver launches it; tools/collect_scaling.py turns the lines into profiles/r02_scaling.json
T=${1:-r02}
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${T}_scale_n$n.json 2> gpurun_out/${T}_scale_n$n.err
done
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/${T}_scale_n1.json 2> gpurun_out/${T}_scale_n1.err
