#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -k "inference_golden or tc_gemm or instance_and_segment or full_size" 2>&1 | tail -5 | tee gpurun_out/pytest_tc3.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['achieved'])"; tail -5 gpurun_out/bench_tc.err
timeout 300 python scripts/tc_trace.py 2>&1 | head -22 | tee gpurun_out/tc_trace.txt
