#!/bin/bash
# round 2, call A: suite + both bench arms at N=1 (train_step parity, psnr on all maps, reference = staged reference modules)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 | tee gpurun_out/r02_pytest_a.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cat gpurun_out/r02_bench_reference.json
timeout 900 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; cat gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
