#!/bin/bash
# first GPU pass: forward parity tests, short bench, launch list + two full ncu captures
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -k "not gradients" 2>&1 | tail -40 > gpurun_out/pytest_fwd.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --frame 400 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -o gpurun_out/prof_march -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --frame 400 > gpurun_out/ncu_march.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heads_forward -s 2 -c 1 -o gpurun_out/prof_heads -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --frame 400 > gpurun_out/ncu_heads.log 2>&1
cat gpurun_out/pytest_fwd.log | tail -30
cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
cat gpurun_out/bench_ref.json
