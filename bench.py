#!/usr/bin/env python
"""bench.py - the hot path's headline metric: rendered Mrays/s at 512 samples/ray (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frame 800] [--samples 512]

One *step* = one pass of the render hot path over one synthetic frame: ray generation (R1-R4), ray march with
TensoRF density lookup + transmittance scan (S1-C1), the RGB / semantic / instance heads on the active samples
(F2-H3) and compositing + per-ray epilogue (C3), all heads on (C=21 classes, 3+3 slow-fast embedding dims).
N>1: one process per GPU (torchrun), every rank renders its own frame with replicated parameters - the camera of rank r
is rolled by 90 degrees x r about its optical axis, a different ray set of exactly the same cost (the cube and the ball are
symmetric under that roll), so per-GPU work is fixed: rays shard with no data-path collective, scaling is "weak" and
`value` = all ranks' rays / max time.

`value`   : device-resident inputs (camera pose -> rays generated on the GPU), CUDA-event time per step, L2
            flushed between steps outside the timed spans.
`e2e`     : the same frame through the public TensoRFRenderer.forward call with the rays in pinned HOST memory
            (as the reference's dataset produces them): H2D of the rays, render, D2H of the four output maps.
`roofline`: the dominant kernel - the pipelined tensor-core kernel of the semantic / instance stacks (92 % of the head
            FLOPs; `heads_all` gives the figure over both head kernels) - from CUDA events recorded inside libclift_b200.so
            around its launch; `roofline_march`: the march (L1/TEX-bound gather; labelled as not-a-roofline, with its
            binding unit and DRAM bytes over compulsory bytes).
`train_step`: (N=1) BASELINE config 3 through the same public classes - ms per training step and `parity` (step-0 losses
            and every parameter gradient against the CPU oracle on the same RNG draws; scripts/train_step_bench.py);
            `train_step_cfg4`: BASELINE config 4's 8192-ray batch on this many GPUs.
N > 1 adds  `ddp_check` (N-GPU == 1-GPU equivalence, run before anything is timed), `frame_split` (ONE frame sharded over the
            ranks + all-gather of the maps, strong scaling) and `train_step_cfg4` (batch sharded over the ranks, the
            gradient all-reduce - clift_allreduce_grads - inside the timed step).
`cpu_baseline` / `--impl reference`: the reference's own CPU implementation on the host cores: the unmodified reference
            modules staged under oracle/_ref by oracle/vendor_reference.py (kind "reference"; the reference-pinned port in
            oracle/ - kind "port" - only when they are absent), replaying render_panopli.py's chunk loop (chunk=2048) on a
            bounded ray sample of the same frame.  `psnr_match`: the CUDA render of those rays against the CPU maps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/sec (512 samples/ray)"
GRID = (128, 128, 128)
N_CLS, N_INS = 21, 3
MAX_ACTIVE_PER_RAY = 160


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sust=1400.0, src="fallback")


def ncu_traffic(name: str, frame: int, samples: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` summary
    (profiles/<name>.json, written by scripts/ncu_summary.py for the default 800x800 x 512 workload), else None."""
    p = os.path.join(ROOT, "profiles", name + ".json")
    if frame != 800 or samples != 512 or not os.path.exists(p):
        return None
    d = json.load(open(p))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        return sum(d[k] * scale[d[k + "__unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except KeyError:
        return None


def ncu_metric(name: str, key: str, frame: int, samples: int):
    """One metric of the committed ncu summary profiles/<name>.json (default workload only), else None."""
    p = os.path.join(ROOT, "profiles", name + ".json")
    if frame != 800 or samples != 512 or not os.path.exists(p):
        return None
    return json.load(open(p)).get(key)


def head_flops_per_sample(params, part: str = "all") -> float:
    """2*MACs of every Linear the heads evaluate per active sample (~1.016 MFLOP at C=21, d=3).
    part "xyz": the semantic + instance stacks (the pipelined kernel), "rgb": basis + rgb MLP, "all": both."""
    macs = 0
    for k, v in params.items():
        if k.endswith(".weight") and ("mlp" in k or "basis" in k):
            is_rgb = "appearance" in k
            if part == "all" or (part == "rgb") == is_rgb:
                macs += v.shape[0] * v.shape[1]
    return 2.0 * macs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_renderer(frame: int, samples: int, rank_yaw: float = 0.0):
    """The reference's CPU implementation of the path for the bench scene -> (render(rays)->(rgb, sem, ins, depth), rays, kind).
    kind "reference": the UNMODIFIED reference modules (TensoRFRenderer.forward on TensorVMSplit) from /root/reference or
    their staged copy oracle/_ref (oracle/vendor_reference.py); kind "port": the reference-pinned restatement in
    oracle/clift_oracle.py when neither is present.  Both replay render_panopli.py:114-120 (no_grad, chunk = 2048)."""
    from contrastive_lift_b200 import synthetic as syn
    from oracle import clift_oracle as orc
    from oracle import refload
    torch.set_num_threads(os.cpu_count() or 1)
    params = syn.make_field_params(0, GRID, N_CLS, N_INS)
    aabb = syn.default_aabb()
    ratio = orc.ratio_for_samples(aabb, GRID, samples)
    k, c2w = syn.camera(frame, frame, yaw_deg=rank_yaw)
    if refload.available() and os.environ.get("CLIFT_BENCH_CPU_PORT") != "1":
        ref = refload.load()
        model = refload.build_model(params, GRID, N_CLS, N_INS, slow_fast=True, semantic_softmax=True)
        rend = refload.build_renderer(aabb, GRID, semantic_softmax=True)
        rend.update_step_ratio(ratio)
        assert rend.n_samples == samples, (rend.n_samples, samples)
        dirs = ref.ray.get_ray_directions_with_intrinsics(frame, frame, k.numpy())
        o, d = ref.ray.get_rays(dirs, c2w)
        far = ref.ray.rays_intersect_sphere(o, d, r=1)
        rays = torch.cat([o, d, 0.01 * torch.ones_like(o[:, :1]), far[:, None]], 1).float().contiguous()   # dataset/base.py:216-219

        def render(r):
            outs = []
            with torch.no_grad():
                for i in range(0, r.shape[0], 2048):
                    outs.append(rend(model, r[i:i + 2048], 1.0, False, False)[:4])
            return tuple(torch.cat([o[j] for o in outs]) for j in range(4))
        return render, rays, "reference"
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=GRID, step_ratio=ratio).refresh()
    assert cfg.n_samples == samples
    rays = orc.make_rays(frame, frame, k, c2w)
    return (lambda r: tuple(orc.render_chunked(params, cfg, r, chunk=2048)[:4])), rays, "port"


def cpu_reference_rate(frame: int, samples: int, budget_s: float, max_rays: int, rank_yaw: float = 0.0, keep: bool = False):
    """Reference CPU path on the host cores: render_panopli.py:114-120 chunk loop, chunk=2048, on a strided ray sample.
    -> (Mrays/s, rays done, seconds, kind[, rays, rgb, semantic, instance, depth])"""
    render, rays, kind = cpu_renderer(frame, samples, rank_yaw)
    stride = max(1, rays.shape[0] // max_rays)
    sub = rays[::stride][:max_rays].contiguous()            # strided: same in-box/active statistics as the frame
    render(sub[:256])                                       # warm-up (thread pools, allocator)
    done, t0, kept = 0, time.perf_counter(), []
    while done < sub.shape[0]:
        out = render(sub[done:done + 2048])
        if keep:
            kept.append(out)
        done += min(2048, sub.shape[0] - done)
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    if keep:
        return (done / dt / 1e6, done, dt, kind, sub[:done]) + tuple(torch.cat([k[j] for k in kept]) for j in range(4))
    return done / dt / 1e6, done, dt, kind


def psnr_match(rend, model, cpu_rays, cpu_maps, dev):
    """CUDA render of ``cpu_rays`` against the CPU reference's maps of the same rays: PSNR of the rgb map
    (util/metrics.py:25-26) + worst deviations of every map (semantic compared in probability space)."""
    import math
    cpu_rgb, cpu_sem, cpu_ins, cpu_depth = cpu_maps
    with torch.no_grad():
        out = rend(model, cpu_rays.to(dev).contiguous(), 1.0, False, False)
    rgb, depth = out[0].float().cpu(), out[3].float().cpu()
    mse = float(((rgb - cpu_rgb) ** 2).mean())
    res = {"psnr_db": min(200.0, -10.0 * math.log10(mse)) if mse > 0.0 else 200.0, "cap_db": 200.0,
           "rgb_max_abs_err": float((rgb - cpu_rgb).abs().max()),
           "depth_max_rel_err": float((depth - cpu_depth).abs().max() / cpu_depth.abs().max().clamp_min(1e-12)),
           "rays": int(cpu_rays.shape[0]), "against": "CPU reference, same rays (cpu_baseline sample)"}
    if out[1] is not None and cpu_sem is not None:
        res["semantic_prob_max_abs_err"] = float((out[1].float().cpu().exp() - cpu_sem.exp()).abs().max())
    over = at_thres = 0
    if out[2] is not None and cpu_ins is not None:
        ins = out[2].float().cpu()
        per_ray = (ins - cpu_ins).abs().amax(-1) / cpu_ins.abs().max().clamp_min(1e-12)
        res["instance_max_rel_err"] = float(per_ray.max())
        over = int((per_ray >= 1e-4).sum())
        res["instance_rays_over_1e-4"] = over
        if 0 < over <= 4096:
            # are those the rays with a sample at the activity threshold?  (CPU oracle weights of just those rays; the same
            # 2e-3 relative margin the parity tests use, tests/gpu_util.py: flip_risk)
            from contrastive_lift_b200 import synthetic as syn
            from oracle import clift_oracle as orc
            samples = int(rend.n_samples)
            cfg = orc.RenderConfig(aabb=syn.default_aabb(), grid_dim=GRID, step_ratio=orc.ratio_for_samples(syn.default_aabb(), GRID, samples)).refresh()
            with torch.no_grad():
                w = orc._march(syn.make_field_params(0, GRID, N_CLS, N_INS), cfg, cpu_rays[per_ray >= 1e-4], None)[6]
            at_thres = int((((w - cfg.weight_thres).abs() < 2e-3 * cfg.weight_thres).sum(-1) > 0).sum())
            res["instance_rays_over_1e-4_with_a_sample_at_the_threshold"] = at_thres
    res["tolerance"] = ("1e-4 scale-relative (max |diff| / max |reference|) per map; semantic in probability space.  A sample "
                        "whose weight straddles raymarch_weight_thres = 1e-4 within fp32 rounding may be active in one "
                        "implementation and not the other (renderer:103); that moves its ray's un-normalised instance "
                        "embedding by thres x |embedding| ~ 1e-4 of the scale, so rays_over counts such rays")
    res["ok_strict"] = bool(res["rgb_max_abs_err"] < 1e-4 and res["depth_max_rel_err"] < 1e-4 and
                            res.get("semantic_prob_max_abs_err", 0.0) < 1e-4 and res.get("instance_max_rel_err", 0.0) < 1e-4)
    # ok: every map inside 1e-4 except rays that demonstrably carry a sample at the threshold (one flip moves the
    # un-normalised embedding by at most thres x |embedding|, i.e. ~1e-4 of the scale: 2e-4 bounds two flips)
    res["ok"] = bool(res["rgb_max_abs_err"] < 1e-4 and res["depth_max_rel_err"] < 1e-4 and
                     res.get("semantic_prob_max_abs_err", 0.0) < 1e-4 and res.get("instance_max_rel_err", 0.0) < 2e-4 and
                     (over == 0 or at_thres == over))
    return res


def host_training_pass(frame_rays):
    """One training-style forward + backward of the path on the host (CPU oracle, 4096 rays in two chunks + the instance
    pass).  Measured on this pool's GPU boxes (oracle/cpu_regime_probe.py): after such a pass the
    reference's no_grad chunk loop runs ~1.85x FASTER in the same process (0.22 s instead of 0.41 s per 2048-ray chunk;
    smaller passes, plain large allocations or thread-pool resets do not trigger it - host memory state, not arithmetic).
    The reference arm runs it as part of its warm-up so that the CPU baseline is the reference's best regime, and reports
    the cold rate next to it."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import train_step_bench as tsb
    from contrastive_lift_b200 import synthetic as syn
    from oracle import clift_oracle as orc
    params = syn.make_field_params(0, GRID, N_CLS, N_INS)
    cfg = orc.RenderConfig(aabb=syn.default_aabb(), grid_dim=GRID).refresh()
    g = torch.Generator().manual_seed(3)
    b = 4096
    batch = (frame_rays[::150][:b].contiguous(), torch.rand(b, 3, generator=g), torch.softmax(torch.randn(b, N_CLS, generator=g), -1),
             torch.rand(b, generator=g), frame_rays[7::600][:1024].contiguous(), torch.randint(1, 8, (1024,), generator=g),
             torch.rand(1024, generator=g))
    tsb.cpu_losses(params, cfg, batch, tsb.replay_draws(b, 1024, 0))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    per_step = 2048          # one reference chunk (config.chunk, render_panopli.py:114) per step: ~0.3 s of host work
    # model and renderer are built once; every step renders a DIFFERENT strided 2048-ray subset of the frame (steady state of
    # the chunk loop, thread pools warm - the same regime the cpu_baseline leg of the GPU arm measures over ~15 s)
    render, frame_rays, kind = cpu_renderer(args.frame, args.samples)
    stride = max(1, frame_rays.shape[0] // per_step)
    subset = lambda i: frame_rays[(i % stride)::stride][:per_step].contiguous()
    render(subset(0)[:256])
    t0 = time.perf_counter()
    render(subset(1))
    cold = per_step / (time.perf_counter() - t0) / 1e6
    if args.frame * args.frame >= 4096 * 150:
        host_training_pass(frame_rays)
    vals = []
    for i in range(warm + steps):
        sub = subset(i)
        t0 = time.perf_counter()
        render(sub)
        dt = time.perf_counter() - t0
        if i >= warm:
            vals.append((sub.shape[0] / dt / 1e6, sub.shape[0], dt))
    rays = sum(n for _, n, _ in vals)
    secs = sum(dt for _, _, dt in vals)
    value = rays / secs / 1e6
    cores = os.cpu_count() or 1
    sample = (f"{per_step} rays/step strided from the {args.frame}x{args.frame} frame, S={args.samples}, chunk=2048, all heads; "
              + ("the unmodified reference modules (TensoRFRenderer.forward)" if kind == "reference" else "oracle port"))
    cfg = workload_config(args, 1)
    cfg["rays_per_step_per_gpu"] = 0
    cfg["rays_per_step_on_host"] = per_step      # what this arm times per step (a bounded sample of the frame; per-ray rate)
    cfg["parallelism"] = f"host cores only ({cores} threads), no GPU"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample,
                             "cold_value": cold,
                             "regime": "value: after one training-style forward+backward pass on the host in this process "
                                       "(the reference's faster regime on these boxes, see host_training_pass); cold_value: "
                                       "first full chunk of a fresh process"},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"synthetic {args.frame}x{args.frame} frame (solid-ball TensoRF scene, G=128^3), {args.samples} samples/ray, "
                        f"all heads (RGB + semantic C={N_CLS} + slow-fast instance d={N_INS}+{N_INS}), inference",
            "frame": args.frame, "samples_per_ray": args.samples, "grid": list(GRID), "rays_per_step_per_gpu": args.frame ** 2,
            "parallelism": f"rays sharded, parameters replicated, {world} GPU(s), no data-path collective; rank r renders the "
                           "frame of the camera rolled by 90 deg x r (same cost by symmetry: fixed per-GPU work)",
            "l2": "256 MB scratch write between timed steps (outside the timed spans); per-step working set ~1.5 GB >> 126 MB L2"}


def run_ours(args):
    import ctypes as C
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - libclift_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    import contrastive_lift_b200 as cl
    from contrastive_lift_b200 import lib as L
    from contrastive_lift_b200 import synthetic as syn

    lib = L.load()
    params = syn.make_field_params(0, GRID, N_CLS, N_INS)
    aabb = syn.default_aabb()
    ratio = syn.ratio_for_samples(aabb, GRID, args.samples)
    model = cl.TensorVMSplit(list(GRID), num_semantics_comps=(32, 32, 32), num_instance_comps=(32, 32, 32),
                             num_semantic_classes=N_CLS, dim_feature_instance=2 * N_INS, use_semantic_mlp=True,
                             use_instance_mlp=True, slow_fast_mode=True)
    model.load_state_dict(params)
    rend = cl.TensoRFRenderer(aabb, list(GRID), semantic_weight_mode="softmax")
    rend.update_step_ratio(ratio)
    assert rend.n_samples == args.samples, (rend.n_samples, args.samples)
    model, rend = model.to(dev), rend.to(dev)
    rend.max_active_per_ray = MAX_ACTIVE_PER_RAY
    rend.check_overflow = False       # checked once after warm-up below; keeps the timed call free of host syncs
    rend.head_path = {"auto": L.HEADS_AUTO, "fma": L.HEADS_FMA, "tensor": L.HEADS_TENSOR, "tensor16": L.HEADS_TENSOR16}[args.heads]
    tensor_heads = args.heads != "fma"
    f16_heads = args.heads in ("auto", "tensor16")
    H = W = args.frame
    n_rays = H * W
    k, c2w = syn.camera(H, W)
    roll = [(1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0)][rank % 4]            # (cos, sin) of 90 degrees x rank
    c2w[:3, :3] = c2w[:3, :3] @ torch.tensor([[roll[0], -roll[1], 0.0], [roll[1], roll[0], 0.0], [0.0, 0.0, 1.0]])
    k_np, c2w_np = k.numpy(), c2w.numpy()
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)

    def step_device():
        rays, _bad = cl.get_rays(H, W, k_np, c2w_np, device=dev)
        return rend(model, rays, 1.0, False, False)

    host_rays = cl.get_rays_checked(H, W, k_np, c2w_np, device=dev).cpu().pin_memory()
    d_rays = torch.empty_like(host_rays, device=dev)
    host_out = [torch.empty((n_rays, c), pin_memory=True) for c in (3, N_CLS, 2 * N_INS)] + [torch.empty((n_rays,), pin_memory=True)]

    def step_e2e():
        d_rays.copy_(host_rays, non_blocking=True)
        out = rend(model, d_rays, 1.0, False, False)
        for h, o in zip(host_out, out[:4]):
            h.copy_(o, non_blocking=True)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """Per-step CUDA-event spans on the launching (current) stream; L2 flushed between spans."""
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize(dev)
        return [a.elapsed_time(b) for a, b in evs]

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            step_device()
        n_act, n_in, overflow, n_tiles = rend.last_stats(dev)
        if overflow:
            raise SystemExit("bench.py: active-sample list overflowed; raise MAX_ACTIVE_PER_RAY")
        sampler = ClockSampler(local)
        barrier()
        if rank == 0:
            sampler.start()
        l0 = L.launch_count()
        ms = timed(step_device, args.steps)
        launches = L.launch_count() - l0
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        # stage timing (separate pass so the event records do not sit inside the headline spans)
        L.check(lib.clift_profile_enable(1))
        stage = [0.0, 0.0, 0.0, 0.0]
        split = [0.0, 0.0]
        reps = min(args.steps, 5)
        for _ in range(reps):
            flush.fill_(1)
            step_device()
            buf = (C.c_float * 4)()
            L.check(lib.clift_profile_stage_ms(buf))
            stage = [s + float(b) for s, b in zip(stage, buf)]
            buf2 = (C.c_float * 2)()
            L.check(lib.clift_profile_heads_split_ms(buf2))
            split = [s + float(b) for s, b in zip(split, buf2)]
        stage = [s / reps for s in stage]
        split = [s / reps for s in split]
        L.check(lib.clift_profile_enable(0))
        for _ in range(2):
            step_e2e()
        barrier()
        ms_e2e = timed(step_e2e, args.steps)
        barrier()

    total_ms = torch.tensor([sum(ms), sum(ms_e2e)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = (float(v) for v in total_ms.cpu())
    multi = {}
    if dist is not None and not args.no_train:
        # The exchange steps of the path, on every rank: (0) N-GPU == 1-GPU equivalence before anything is timed,
        # (1) ONE frame split over the ranks + all-gather of the output maps (strong scaling of inference),
        # (2) BASELINE config 4: one 8192-ray batch sharded over the ranks, gradient all-reduce inside the timed step
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import ddp_check
        import train_step_bench
        from contrastive_lift_b200 import parallel as par
        try:
            multi["ddp_check"] = ddp_check.check(dev, verbose=False)
        except Exception as e:
            multi["ddp_check"] = {"ok": False, "error": f"{type(e).__name__}: {e}"}
        c2w0 = syn.camera(H, W)[1].numpy()

        def step_split():
            rays, _bad = cl.get_rays(H, W, k_np, c2w0, device=dev)          # every rank: the same frame, its own rows
            out = rend(model, par.shard_rows_cyclic(rays, H, W, rank, world).contiguous(), 1.0, False, False)
            maps = torch.cat([out[0], out[1], out[2], out[3][:, None]], 1)
            return par.gather_rows_cyclic(maps, H, W)

        with torch.no_grad():
            for _ in range(3):
                full = step_split()
            barrier()
            ms_split = timed(step_split, args.steps)
            barrier()
            single = None
            if rank == 0:       # the gathered frame against the same frame rendered whole on this GPU
                rays, _bad = cl.get_rays(H, W, k_np, c2w0, device=dev)
                o = rend(model, rays, 1.0, False, False)
                single = torch.cat([o[0], o[1], o[2], o[3][:, None]], 1)
                # per-ray sums of a ray whose samples straddle two head tiles are combined with atomics: last-bit differences
                split_equal = bool(torch.allclose(full, single, rtol=1e-5, atol=2e-6))
                split_err = float((full - single).abs().max())
            del full, single
        t = torch.tensor([sum(ms_split)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        multi["frame_split"] = {"workload": f"ONE {H}x{W} frame, image rows dealt round-robin to {world} GPUs (equal cost per rank), "
                                            f"maps all-gathered ({(3 + N_CLS + 2 * N_INS + 1) * 4} B/ray) to every rank in ray order, "
                                            "inside the timed region",
                                "ms_per_frame": float(t) / args.steps, "Mrays_per_s": n_rays * args.steps / (float(t) * 1e-3) / 1e6,
                                "scaling": "strong"}
        if rank == 0:
            multi["frame_split"].update({"matches_single_gpu_frame": split_equal, "max_abs_diff": split_err,
                                         "tolerance": "allclose(rtol 1e-5, atol 2e-6)"})
        flush = d_rays = None
        torch.cuda.empty_cache()
        try:
            multi["train_step_cfg4"] = train_step_bench.measure(steps=10, warmup=3, device_index=local, rays=8192, classes=2,
                                                                distributed=True)
        except Exception as e:
            multi["train_step_cfg4"] = {"error": f"{type(e).__name__}: {e}"}
    if dist is not None:
        from contrastive_lift_b200 import parallel as _par
        torch.cuda.synchronize(dev)
        _par.destroy_nccl_communicators()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = world * n_rays * args.steps / (t_dev * 1e-3) / 1e6
    e2e = world * n_rays * args.steps / (t_e2e * 1e-3) / 1e6
    flops_all = head_flops_per_sample(params) * n_act
    two_kernels = split[0] > 0.0          # inference default: pipelined xyz-stack kernel + rgb-stack kernel
    # dominant kernel: the xyz stacks (92 % of the head FLOPs) when they run on their own kernel, else the one head kernel
    flops = head_flops_per_sample(params, "xyz") * n_act if two_kernels else flops_all
    dom_ms = split[0] if two_kernels else stage[2]
    heads_tflops = flops / (dom_ms * 1e-3) / 1e12
    march_bytes = 32.0 * n_rays + 1152.0 * n_in + 4.0 * n_rays * args.samples + 16.0 * n_rays
    march_gbs = march_bytes / (stage[0] * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e, "unit": "Mrays/s", "h2d_bytes_per_step": host_rays.numel() * 4,
                "d2h_bytes_per_step": sum(h.numel() for h in host_out) * 4, "ms_per_step": t_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "scene": {"n_inbox_per_ray": n_in / n_rays, "n_active_per_ray": n_act / n_rays, "head_tiles": n_tiles},
        "stage_ms": {"march": stage[0], "compact": stage[1], "heads": stage[2], "epilogue": stage[3],
                     "heads_xyz_kernel": split[0], "heads_rgb_kernel": split[1]},
        "heads_all": {"tflops": flops_all / (stage[2] * 1e-3) / 1e12, "frac_of_bf16_sustained": flops_all / (stage[2] * 1e-3) / 1e12 / pk["bf16_sust"],
                      "note": "all head FLOPs over the whole heads stage (both kernels)"},
        "roofline": {"kernel": ("heads_x16_kernel (xyz stacks: tcgen05 kind::f16 cta_group::2, 3-product fp16 split, pipelined)"
                                if two_kernels else
                                ("heads_tc16_forward_kernel (tcgen05 kind::f16, 3-product fp16 split)" if f16_heads else
                                 "heads_tc_forward_kernel (tcgen05 3xTF32)") if tensor_heads else "heads_forward_kernel (FP32 FMA)"),
                     "bound": "tensor", "achieved": heads_tflops, "peak": pk["bf16_sust"],
                     "unit": "TFLOP/s", "frac": heads_tflops / pk["bf16_sust"], "launch_ms": dom_ms,
                     "traffic": ncu_traffic("r02_ncu_heads_x16" if two_kernels else
                                            (("r01_ncu_heads_tc16" if f16_heads else "r01_ncu_heads_tc") if tensor_heads else "r01_ncu_heads_fma"),
                                            args.frame, args.samples),
                     "traffic_source": "dram__bytes of one launch from the COMMITTED `ncu --set full` capture of this workload "
                                       "under profiles/ (not measured in this run)",
                     "algorithmic_flops_per_launch": flops,
                     "note": f"algorithmic 2*MAC FLOPs of the head Linears x active samples; peak = {pk['src']} sustained bf16 "
                             + ("(fp32-faithful heads issue 3 kind::f16 MMAs per product = 3x the bf16 cost, so the reachable "
                                "fraction of this peak is 1/3 by construction)" if f16_heads else
                                "(fp32-faithful heads issue 3 tf32 MMAs per product = 6x the bf16 cost, so the reachable "
                                "fraction of this peak is 1/6 by construction)")},
        "roofline_march": {"kernel": "march_kernel", "bound": "hbm", "achieved": march_gbs, "peak": pk["hbm"], "unit": "GB/s",
                           "frac": march_gbs / pk["hbm"], "traffic": ncu_traffic("r02_ncu_march", args.frame, args.samples),
                           "traffic_source": "committed ncu capture profiles/r02_ncu_march.json (march with fused compaction), not measured in this run",
                           "algorithmic_bytes_per_launch": march_bytes,
                           "not_a_roofline": "frac > 1: the 12.6 MB of factors are L1/L2-resident, so the ALGORITHMIC gather rate "
                                             "(1152 B per in-box sample) exceeds the HBM peak; HBM does not bound this kernel",
                           "binding_unit": {"unit": "L1/TEX", "pct_of_peak": ncu_metric("r02_ncu_march", "l1tex__throughput.avg.pct_of_peak_sustained_active", args.frame, args.samples),
                                            "l1_hit_pct": ncu_metric("r02_ncu_march", "l1tex__t_sector_hit_rate.pct", args.frame, args.samples),
                                            "source": "profiles/r02_ncu_march.json"},
                           "dram_bytes_over_compulsory": (lambda t, c: None if t is None else t / c)(
                               ncu_traffic("r02_ncu_march", args.frame, args.samples),
                               32.0 * n_rays + 12.6e6 + 16.0 * n_rays + 24.0 * n_act),
                           "compulsory_note": "compulsory = rays in (32 B/ray) + one read of the factors (12.6 MB) + per-ray outputs "
                                              "(16 B/ray) + the active-sample records the march now emits itself (24 B each); round 1's "
                                              "two-pass form moved 30x its compulsory bytes (dense [B,S] weights written and re-read)",
                           "note": "algorithmic gather bytes (1152 B per in-box sample) + ray/weight streams; factors are "
                                   "L2-resident so DRAM traffic is far below this by design"},
    }
    if world == 1 and not args.no_train:
        # BASELINE config 3 next to the headline: one training step (4096-ray main pass + 1024-ray instance pass with the
        # slow-fast loss, both backwards, two fused Adam steps) through the same public classes; reported, not the metric
        try:
            flush = d_rays = None        # release the render bench's buffers first
            torch.cuda.empty_cache()
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import train_step_bench
            line["train_step"] = train_step_bench.measure(steps=10, warmup=3, device_index=local, with_cpu=False, parity=True)
            # BASELINE config 4 on ONE GPU: the N = 1 point of the strong-scaling training figure the N > 1 lines carry
            line["train_step_cfg4"] = train_step_bench.measure(steps=10, warmup=3, device_index=local, rays=8192, classes=2)
        except Exception as e:      # never lose the bench line over the extra figure
            line["train_step"] = {"error": f"{type(e).__name__}: {e}"}
    line.update(multi)
    if world == 1 and not args.no_cpu:
        v, n, dt, kind, cpu_rays, *cpu_maps = cpu_reference_rate(args.frame, args.samples, args.cpu_seconds, 131072, keep=True)
        line["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": kind,
                                "sample": f"{n} rays strided from the same {H}x{W} frame, S={args.samples}, chunk=2048, all heads, {dt:.1f} s"
                                          + ("; the unmodified reference modules" if kind == "reference" else "; reference-pinned port")}
        # "PSNR match" of BASELINE.json's metric: the rays the CPU reference just rendered, rendered again by the CUDA path,
        # PSNR = -10 log10(mse) of the two rgb maps (util/metrics.py:25-26); identical images give +inf, reported capped
        try:
            line["psnr_match"] = psnr_match(rend, model, cpu_rays, cpu_maps, dev)
        except Exception as e:      # never lose the bench line over the extra figure
            line["psnr_match"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frame", type=int, default=800)
    ap.add_argument("--samples", type=int, default=512)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the extra training-step figure (N=1 only)")
    ap.add_argument("--heads", default="auto", choices=["auto", "fma", "tensor", "tensor16"], help="MLP-head kernel family")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
