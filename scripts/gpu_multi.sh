#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/ddp_check.py 2>&1 | tail -5 | tee gpurun_out/ddp_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench_n2.err | tail -1 | tee gpurun_out/bench_n2.json
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 2> gpurun_out/bench_n1.err | tail -1 | tee gpurun_out/bench_n1.json
tail -3 gpurun_out/bench_n2.err
