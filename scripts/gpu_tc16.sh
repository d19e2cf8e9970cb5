#!/bin/bash
# fp16-split tensor-core heads: GEMM core vs fp64, parity suite, A/B benches of the development switches, trace
mkdir -p gpurun_out
short() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value'],3), 'Mrays/s', d['stage_ms'], d['clocks'])" $1; }
if [ "$1" == "pair" ]; then
  CLIFT_TC16_PAIR=1 timeout 180 python -m pytest tests/test_gpu_parity.py -q --tb=short -x -k "tcgen05_f16" 2>&1 | tail -15 | tee gpurun_out/pytest_pair.log
  CLIFT_TC16_PAIR=1 timeout 180 python bench.py --heads tensor16 --steps 5 --no-cpu > gpurun_out/bench_tc16_pair.json 2> gpurun_out/bench_tc16_pair.err; short gpurun_out/bench_tc16_pair.json; tail -3 gpurun_out/bench_tc16_pair.err
  CLIFT_TC16_PAIR=1 CLIFT_TC16_STREAM=1 timeout 180 python -m pytest tests/test_gpu_parity.py -q --tb=short -x -k "tcgen05_f16" 2>&1 | tail -5 | tee gpurun_out/pytest_pair_stream.log
  CLIFT_TC16_PAIR=1 CLIFT_TC16_STREAM=1 timeout 180 python bench.py --heads tensor16 --steps 5 --no-cpu > gpurun_out/bench_tc16_pair_stream.json 2>/dev/null; short gpurun_out/bench_tc16_pair_stream.json
  CLIFT_TC16_PAIR=1 CLIFT_TC16_STREAM=1 timeout 120 python scripts/tc_trace.py 2>&1 | head -26 > gpurun_out/tc16_trace_pair_stream.txt; cat gpurun_out/tc16_trace_pair_stream.txt
  CLIFT_TC16_PAIR=1 timeout 120 python scripts/tc_trace.py 2>&1 | head -26 > gpurun_out/tc16_trace_pair.txt; cat gpurun_out/tc16_trace_pair.txt
fi
timeout 600 python -m pytest tests/test_gpu_tc.py -q --tb=short -x 2>&1 | tail -5 | tee gpurun_out/pytest_tc16.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 | tee gpurun_out/pytest_all.log
timeout 600 python bench.py --heads tensor16 --steps 5 --no-cpu > gpurun_out/bench_tc16.json 2> gpurun_out/bench_tc16.err; short gpurun_out/bench_tc16.json; tail -3 gpurun_out/bench_tc16.err
