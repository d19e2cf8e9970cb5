#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/sweep.py 2>&1 | tee gpurun_out/sweep.md | tail -20
