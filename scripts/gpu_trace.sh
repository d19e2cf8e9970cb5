#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tc_trace.py 2>&1 | tee gpurun_out/tc_trace.txt
