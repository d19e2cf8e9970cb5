#!/bin/bash
# pipelined xyz-stack kernel: inference parity first (short timeouts: a stuck mbarrier must not hold the box), then timings
mkdir -p gpurun_out
for mode in 1 0; do
  echo "== CLIFT_X16_PAIR=$mode"
  CLIFT_X16_PAIR=$mode timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -q --tb=short -x -k "inference or golden or tc or instance or segment" 2>&1 | tail -4 | tee gpurun_out/r02_pytest_x16_pair$mode.log
  if grep -q "passed" gpurun_out/r02_pytest_x16_pair$mode.log && ! grep -q "failed\|error" gpurun_out/r02_pytest_x16_pair$mode.log; then
    CLIFT_X16_PAIR=$mode timeout 200 python scripts/heads_split_time.py 2>&1 | tail -1
    CLIFT_X16_PAIR=$mode timeout 200 python scripts/x16_trace.py 2>&1 | tail -27 > gpurun_out/r02_x16_trace_pair$mode.txt
  fi
done
