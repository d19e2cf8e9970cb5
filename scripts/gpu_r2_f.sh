#!/bin/bash
# full suite + bench + ncu captures of the new head kernels and the march
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 | tee gpurun_out/r02_pytest_f.log
timeout 900 python bench.py > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; cut -c1-3000 gpurun_out/r02_bench_f.json; tail -3 gpurun_out/r02_bench_f.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heads_x16_kernel -s 2 -c 1 -o gpurun_out/r02_prof_heads_x16 -f python bench.py --no-cpu --no-train --steps 1 --warmup 3 > gpurun_out/ncu_x16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:heads_tc16_forward -s 2 -c 1 -o gpurun_out/r02_prof_heads_rgb -f python bench.py --no-cpu --no-train --steps 1 --warmup 3 > gpurun_out/ncu_rgb.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_infer.csv python bench.py --no-cpu --no-train --steps 2 --warmup 3 > gpurun_out/ncu_launch_infer.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
