#!/bin/bash
# 8-GPU weak-scaling run of the headline bench (the driver's SCALE run uses the same command)
mkdir -p gpurun_out
nvidia-smi -L | wc -l | tee gpurun_out/gpus8.txt
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/bench_n$n.err | tail -1 | tee gpurun_out/bench_n$n.json | cut -c1-260
done
