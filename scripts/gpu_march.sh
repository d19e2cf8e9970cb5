#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -k "density or weights or inference_golden or sampling or full_size or gradients" 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_m.json')); print(d['value'], d['ms_per_step'], d['stage_ms'])"; tail -3 gpurun_out/bench_m.err
