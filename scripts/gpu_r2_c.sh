#!/bin/bash
# tensor-core dgrad engine: gradient tests in both modes (under timeouts: a stuck mbarrier must not hold the box), training step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "grad or training or instance" 2>&1 | tail -15 | tee gpurun_out/r02_pytest_dgrad_tc.log
CLIFT_DGRAD_FMA=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "grad or training or instance" 2>&1 | tail -5 | tee gpurun_out/r02_pytest_dgrad_fma.log
timeout 300 python scripts/train_step_bench.py --no-cpu --parity > gpurun_out/r02_train_step_tc.json 2>gpurun_out/r02_train_step_tc.err; cat gpurun_out/r02_train_step_tc.json; tail -3 gpurun_out/r02_train_step_tc.err
CLIFT_DGRAD_FMA=1 timeout 300 python scripts/train_step_bench.py --no-cpu > gpurun_out/r02_train_step_fma.json 2>/dev/null; cat gpurun_out/r02_train_step_fma.json
