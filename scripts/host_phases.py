"""Host wall time of the phases of one training step (perf_counter around the public entry points, no profiler), next to the
CUDA-event step time: shows which part of a step the host, not the GPU, is paying for.  Usage: host_phases.py [rays] [classes]"""
import collections
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import contrastive_lift_b200 as cl
from contrastive_lift_b200 import field, loss, optim, renderer
import train_step_bench as tsb

acc = collections.defaultdict(float)
cnt = collections.defaultdict(int)
live = [False]


def timed(owner, name, label):
    fn = getattr(owner, name)
    raw = fn.__func__ if isinstance(fn, staticmethod) else fn

    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return raw(*a, **k)
        finally:
            if live[0]:
                acc[label] += time.perf_counter() - t0
                cnt[label] += 1
    setattr(owner, name, staticmethod(w) if isinstance(owner.__dict__.get(name), staticmethod) else w)


timed(field.PackedField, "refresh", "  packed.refresh")
timed(field.PackedField, "matches", "  packed.matches")
timed(renderer._Render, "_forward", " render forward (incl. refresh, stats sync)")
timed(renderer._Render, "_backward", " render backward")
timed(renderer.TensoRFRenderer, "last_stats", "  last_stats (D2H sync)")
timed(renderer.TensoRFRenderer, "_run", "render _run (forward + autograd apply)")
timed(optim.FusedAdam, "step", "FusedAdam.step")
timed(torch.Tensor, "backward", "loss.backward()")
timed(loss, "total_tv", "total_tv")
timed(tsb, "main_loss", "main_loss (2 renders + loss math)")
timed(cl, "slow_fast_loss", "slow_fast_loss")
timed(cl, "ema_update_slownet", "ema_update_slownet")

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
classes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
real = tsb.gpu_step
n = [0]
wall = []


def step(*a, **k):
    n[0] += 1
    live[0] = n[0] > 3
    t0 = time.perf_counter()
    r = real(*a, **k)
    if live[0]:
        wall.append(time.perf_counter() - t0)
    return r


segs = []


def step_counted(*a, **k):
    r = step(*a, **k)
    s = torch.cuda.memory_stats()
    segs.append((s["segment.all.allocated"], s["segment.all.freed"], s["num_alloc_retries"]))
    return r


# inside render backward: the scratch allocation, the library call, the gradient unpack
lib = cl.lib.load() if hasattr(cl, "lib") else None
from contrastive_lift_b200 import lib as LL
_lib = LL.load()
for nm in ("clift_render_backward", "clift_render_forward", "clift_pack_batch"):
    fn = getattr(_lib, nm)

    def mk(fn, nm):
        def w(*a):
            t0 = time.perf_counter()
            try:
                return fn(*a)
            finally:
                if live[0]:
                    acc["   C call " + nm] += time.perf_counter() - t0
                    cnt["   C call " + nm] += 1
        return w
    setattr(_lib, nm, mk(fn, nm))
_empty = torch.empty


def empty(*a, **k):
    t0 = time.perf_counter()
    try:
        return _empty(*a, **k)
    finally:
        if live[0]:
            acc["   torch.empty (all)"] += time.perf_counter() - t0
            cnt["   torch.empty (all)"] += 1


torch.empty = empty
tsb.gpu_step = step_counted
out = tsb.measure(steps=20, warmup=3, rays=rays, classes=classes)
print("cudaMalloc segments (allocated, freed, retries) first/last timed step:", segs[3], segs[-1])
steps = len(wall)
lines = ["rays %d classes %d: CUDA-event %.3f ms/step, host wall %.3f ms/step, clift launches/step %.0f"
         % (rays, classes, out["ms_per_step"], 1e3 * sum(wall) / steps, out["clift_launches_per_step"])]
for k in sorted(acc, key=lambda k: -acc[k]):
    lines.append("%8.3f ms/step  %5.1f calls/step  %s" % (1e3 * acc[k] / steps, cnt[k] / steps, k))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "host_phases_%d.txt" % rays), "w") as f:
    f.write("\n".join(lines) + "\n")
print("\n".join(lines))
