#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 3 -c 1 -o gpurun_out/prof_wgrad_tc -f \
    python scripts/train_step_bench.py --profile-steps 2 > gpurun_out/ncu_wgrad_tc.log 2>&1
tail -3 gpurun_out/ncu_wgrad_tc.log
python scripts/ncu_summary.py gpurun_out/prof_wgrad_tc.ncu-rep gpurun_out/ncu_wgrad_tc.md "wgrad_tc_kernel" | tail -30
