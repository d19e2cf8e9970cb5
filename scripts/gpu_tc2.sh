#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -k "inference_golden or instance_and_segment or tc_gemm" 2>&1 | tail -40 | tee gpurun_out/pytest_tc2.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -k "full_size" 2>&1 | tail -40 | tee -a gpurun_out/pytest_tc2.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; cat gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
