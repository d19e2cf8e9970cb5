#!/bin/bash
# round-2 final evidence pass (1 GPU): whole GPU suite, both bench arms, smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu_final.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-300 gpurun_out/r02_bench_reference.json
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; cut -c1-400 gpurun_out/r02_bench_final.json; tail -2 gpurun_out/r02_bench_final.err
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
