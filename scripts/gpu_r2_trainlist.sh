#!/bin/bash
# training step (BASELINE config 3) launch lists at 2 and 6 steps: the difference is 4 steady-state steps (setup launches cancel)
mkdir -p gpurun_out
for n in 2 6; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_$n.csv \
    python scripts/train_step_bench.py --profile-steps $n > gpurun_out/ncu_launch_train_$n.log 2>&1
done
python scripts/launch_summary.py gpurun_out/launches_train_6.csv gpurun_out/launches_train.md "training step launch list (4 steady-state steps)" --minus gpurun_out/launches_train_2.csv | head -70
