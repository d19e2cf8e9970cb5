#!/bin/bash
# end-of-iteration check: all GPU tests, smoke, default bench (both arms), training-step bench + launch list, sweep, ncu evidence
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 | tee gpurun_out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python scripts/train_step_bench.py 2>&1 | tail -1 | tee gpurun_out/train_step.json
CLIFT_TRAIN_FWD_FMA=1 CLIFT_WGRAD_FMA=1 timeout 600 python scripts/train_step_bench.py 2>&1 | tail -1 | tee gpurun_out/train_step_fma.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv \
    python scripts/train_step_bench.py --profile-steps 2 > gpurun_out/ncu_launch_train.log 2>&1
timeout 300 python scripts/tc_trace.py 2>&1 | head -26 > gpurun_out/tc_trace.txt
timeout 300 python scripts/wgrad_trace.py 2>&1 | tail -150 > gpurun_out/wgrad_trace.txt
timeout 900 python scripts/sweep.py > gpurun_out/sweep.md 2> gpurun_out/sweep.err; tail -20 gpurun_out/sweep.md
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc16.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-train > gpurun_out/ncu_launch_tc16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heads_tc16_forward -s 3 -c 1 -o gpurun_out/prof_heads_tc16 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-train > gpurun_out/ncu_heads_tc16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 3 -c 1 -o gpurun_out/prof_march -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-train > gpurun_out/ncu_march.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 3 -c 1 -o gpurun_out/prof_wgrad_tc -f \
    python scripts/train_step_bench.py --profile-steps 2 > gpurun_out/ncu_wgrad_tc.log 2>&1
ls -la gpurun_out/*.ncu-rep
