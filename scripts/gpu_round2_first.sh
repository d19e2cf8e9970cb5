#!/bin/bash
# First GPU call of the next round: what was written after round 1's GPU budget ran out, before anything else.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh'
mkdir -p gpurun_out
# 1. the regular suite (includes the CPU-resident upsample and the ray-range split tests added at the end of round 1)
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 | tee gpurun_out/pytest_all.log
# 2. the opt-in tests that never ran on a B200 (RGB-only heads = BASELINE config 1), under their own short timeout:
#    an mbarrier wait that never completes would otherwise hold the box
CLIFT_RUN_UNVERIFIED=1 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k config1 2>&1 | tail -15 | tee gpurun_out/pytest_unverified.log
# 3. bench with the psnr_match key and the corrected (in-place ReLU) CPU port; both arms
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
