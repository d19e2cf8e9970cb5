#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -s --tb=short 2>&1 | tail -40 | tee gpurun_out/pytest_tc.log
