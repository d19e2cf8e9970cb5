#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/ddp_check.py 2>&1 | tail -5 | tee gpurun_out/ddp_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench_n2.err | tail -1 | tee gpurun_out/bench_n2.json
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_multi.log
tail -3 gpurun_out/bench_n2.err
