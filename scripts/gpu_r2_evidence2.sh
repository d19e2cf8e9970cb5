#!/bin/bash
# second evidence pass of round 2: ncu of the backward kernel after the fused mask pass, of the training forward on the pipelined
# kernel, of the grid-mode head kernel (summaries made on the box: the reports together exceed what gpurun copies back; only
# the backward kernel's report is kept for the source page); inference launch list
mkdir -p gpurun_out/ev2
O=gpurun_out/ev2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heads_backward_kernel -s 3 -c 1 -o $O/r02_prof_heads_backward -f \
    python scripts/train_step_bench.py --profile-steps 2 > $O/ncu_heads_backward.log 2>&1
python scripts/ncu_summary.py $O/r02_prof_heads_backward.ncu-rep $O/r02_ncu_heads_backward.md "Round 2 - heads_backward_kernel: tcgen05 dgrad engine + ReLU mask folded into the emit pass (main-pass chunk, 2048 rays)" > /dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heads_x16_kernel -s 3 -c 1 -o $O/x16_train -f \
    python scripts/train_step_bench.py --profile-steps 2 > $O/ncu_x16_train.log 2>&1
python scripts/ncu_summary.py $O/x16_train.ncu-rep $O/r02_ncu_heads_x16_train.md "Round 2 - heads_x16_kernel<pair, kStash> (training forward of the xyz stacks, main-pass chunk, 2048 rays)" > /dev/null
rm -f $O/x16_train.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heads_tc16_forward -s 1 -c 1 -o $O/grid -f \
    python scripts/grid_mode_time.py > $O/ncu_grid.log 2>&1
python scripts/ncu_summary.py $O/grid.ncu-rep $O/r02_ncu_heads_grid.md "Round 2 - heads_tc16_forward_kernel with grid-mode semantic + instance heads (allgrid, 800x800 x 512, inference)" > /dev/null
rm -f $O/grid.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_infer.csv python bench.py --no-cpu --no-train --steps 2 --warmup 3 > $O/ncu_launch_infer.log 2>&1
python scripts/launch_summary.py $O/r02_launches_infer.csv $O/r02_launches_infer.md "inference frame launch list (bench.py --steps 2 --warmup 3, 800x800 x 512, all heads)" > /dev/null
ls -la $O | head -20; du -sh gpurun_out
