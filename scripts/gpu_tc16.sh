#!/bin/bash
# fp16-split tensor-core heads: GEMM core vs fp64, parity suite, A/B benches of the development switches, trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q --tb=short -x 2>&1 | tail -5 | tee gpurun_out/pytest_tc16.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 | tee gpurun_out/pytest_all.log
short() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value'],3), 'Mrays/s', d['stage_ms'], d['clocks'])" $1; }
timeout 600 python bench.py --heads tensor16 --steps 5 --no-cpu > gpurun_out/bench_tc16.json 2> gpurun_out/bench_tc16.err; short gpurun_out/bench_tc16.json; tail -3 gpurun_out/bench_tc16.err
CLIFT_TC16_PARK=0 timeout 600 python bench.py --heads tensor16 --steps 5 --no-cpu > gpurun_out/bench_tc16_nopark.json 2>/dev/null; short gpurun_out/bench_tc16_nopark.json
CLIFT_TC16_SETS=0 timeout 600 python bench.py --heads tensor16 --steps 5 --no-cpu > gpurun_out/bench_tc16_nosets.json 2>/dev/null; short gpurun_out/bench_tc16_nosets.json
CLIFT_TC16_STREAM=1 timeout 600 python bench.py --heads tensor16 --steps 5 --no-cpu > gpurun_out/bench_tc16_stream.json 2>/dev/null; short gpurun_out/bench_tc16_stream.json
timeout 300 python scripts/tc_trace.py 2>&1 | head -26 > gpurun_out/tc16_trace.txt; cat gpurun_out/tc16_trace.txt
