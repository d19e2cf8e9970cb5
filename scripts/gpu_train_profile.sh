#!/bin/bash
# training-step bench + its ncu launch list (per-kernel share of the step)
mkdir -p gpurun_out
timeout 600 python scripts/train_step_bench.py 2>&1 | tail -1 | tee gpurun_out/train_step.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv \
    python scripts/train_step_bench.py --profile-steps 2 > gpurun_out/ncu_launch_train.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_train.csv gpurun_out/launches_train.md "training step launch list" | head -60
