#!/bin/bash
# bring-up of the fp16-split tensor-core heads: GEMM core vs fp64, parity suite, A/B bench against the 3xTF32 path, trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q --tb=short -x -s 2>&1 | tail -70 > gpurun_out/pytest_tc16.log; tail -15 gpurun_out/pytest_tc16.log
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 | tee gpurun_out/pytest_all.log
timeout 600 python bench.py --heads tensor16 --steps 5 --no-cpu > gpurun_out/bench_tc16.json 2> gpurun_out/bench_tc16.err; cat gpurun_out/bench_tc16.json; tail -3 gpurun_out/bench_tc16.err
timeout 600 python bench.py --heads tensor --steps 5 --no-cpu > gpurun_out/bench_tc32.json 2> gpurun_out/bench_tc32.err; cat gpurun_out/bench_tc32.json
timeout 300 python scripts/tc_trace.py 2>&1 | head -26 > gpurun_out/tc16_trace.txt; cat gpurun_out/tc16_trace.txt
