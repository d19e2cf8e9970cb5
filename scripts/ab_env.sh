#!/bin/bash
# A/B of one environment switch on the training step: scripts/ab_env.sh VAR  -> ms per step with VAR=0 and VAR=1, alternating
# three times in fresh processes on the same box (box-to-box and power-state noise is larger than most kernel-level effects)
VAR=$1
for rep in 1 2 3; do
  for v in 0 1; do
    ms=$(env $VAR=$v timeout 300 python scripts/train_step_bench.py 2>/dev/null | tail -1 | python -c "import json,sys; print('%.3f' % json.loads(sys.stdin.read())['ms_per_step'])")
    echo "$VAR=$v rep $rep: $ms ms/step"
  done
done
