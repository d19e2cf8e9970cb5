#!/bin/bash
# ncu evidence for the round: launch list of the default bench command + full captures of the two main kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heads_tc_forward -s 3 -c 1 -o gpurun_out/prof_heads_tc -f \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_heads_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 3 -c 1 -o gpurun_out/prof_march -f \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_march.log 2>&1
ls -la gpurun_out/*.ncu-rep
