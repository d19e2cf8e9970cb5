#!/bin/bash
# evidence pass: ncu of the fused march, launch list of the inference frame, bench line
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -o gpurun_out/r02_prof_march -f python bench.py --no-cpu --no-train --steps 1 --warmup 3 > gpurun_out/ncu_march.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_infer.csv python bench.py --no-cpu --no-train --steps 2 --warmup 3 > gpurun_out/ncu_launch_infer.log 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; cut -c1-1500 gpurun_out/r02_bench_g.json
