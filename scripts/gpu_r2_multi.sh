#!/bin/bash
# multi-GPU: bench.py --gpus N (ddp equivalence, frame split, sharded config-4 training step) + the multi-GPU tests
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r02_gpus_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
cat gpurun_out/r02_bench_n$N.json; tail -5 gpurun_out/r02_bench_n$N.err
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_epoch.py -m gpu -q --tb=short -k "multi or sharded or device" 2>&1 | tail -5 | tee gpurun_out/r02_pytest_multi_n$N.log
