#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 | tee gpurun_out/pytest_all.log
timeout 900 python scripts/train_step_bench.py 2>&1 | tail -3 | tee gpurun_out/train_step.json
