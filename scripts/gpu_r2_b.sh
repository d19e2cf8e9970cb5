#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 | tee gpurun_out/r02_pytest_b.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cat gpurun_out/r02_bench_reference.json
CLIFT_BENCH_CPU_PORT=1 timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_reference_port.json 2>/dev/null; cat gpurun_out/r02_bench_reference_port.json
