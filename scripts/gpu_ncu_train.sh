#!/bin/bash
# ncu --set full captures of the two other big kernels of the training step (main-pass chunk)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heads_backward_kernel -s 3 -c 1 -o gpurun_out/prof_heads_backward -f \
    python scripts/train_step_bench.py --profile-steps 2 > gpurun_out/ncu_heads_backward.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heads_tc16_forward -s 3 -c 1 -o gpurun_out/prof_heads_tc16_train -f \
    python scripts/train_step_bench.py --profile-steps 2 > gpurun_out/ncu_heads_tc16_train.log 2>&1
ls -la gpurun_out/prof_heads_backward.ncu-rep gpurun_out/prof_heads_tc16_train.ncu-rep
timeout 300 python bench.py --no-cpu --steps 3 2>&1 | cut -c1-160
