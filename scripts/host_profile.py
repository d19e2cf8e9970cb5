"""Host-side profile of the training step: cProfile over the timed steps of scripts/train_step_bench.measure() on a batch small
enough that the GPU is never the limiter (the N = 8 shard of BASELINE config 4: 1024 rays), so tottime ranks what the
Python / ctypes / torch-dispatch side of a step costs.  Writes the table to gpurun_out/host_profile.txt."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import train_step_bench as tsb

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prof = cProfile.Profile()
real = tsb.gpu_step
calls = [0]
wall = []


def wrapped(*a, **k):
    calls[0] += 1
    on = calls[0] > 3
    t0 = time.perf_counter()
    if on:
        prof.enable()
    try:
        return real(*a, **k)
    finally:
        if on:
            prof.disable()
            wall.append(time.perf_counter() - t0)


tsb.gpu_step = wrapped
out = tsb.measure(steps=20, warmup=3, rays=rays, classes=2)
s = io.StringIO()
st = pstats.Stats(prof, stream=s)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumulative").print_stats(60)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "host_profile.txt"), "w") as f:
    f.write("ms_per_step (CUDA events) %.3f; host wall per step (profiled) %.3f ms; launches/step %.0f\n"
            % (out["ms_per_step"], 1e3 * sum(wall) / max(len(wall), 1), out["clift_launches_per_step"]))
    f.write(s.getvalue())
print(out["ms_per_step"], 1e3 * sum(wall) / max(len(wall), 1))
