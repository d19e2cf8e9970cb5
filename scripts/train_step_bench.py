"""BASELINE configs 3 and 4: one training step - ray batch, slow-fast contrastive loss, forward + backward - on 1..N GPUs.

Replays the hot-path calls of TensoRFTrainer.training_step (trainer/train_panopli_tensorf.py:148-228):
  (A) main pass: chunked render (chunk 2048) -> MSE + TV + dist-reg + semantic CE -> backward -> Adam step
  (B) instance pass: forward_instance_feature on 1024 rays of one image -> EMA -> slow-fast loss -> backward -> Adam step
N > 1 (torchrun, one process per GPU; what the reference gets from Lightning's DDPStrategy, trainer/__init__.py:95-108): the
main-pass batch is sharded over the ranks (parallel.shard_range), every rank renders one instance image of its own (the
slow-fast loss couples the rays of an image, so images stay whole - DDP's sampler does the same), and each backward is
followed by ONE sum-all-reduce of that optimizer's gradient arena (parallel.GradientArena) INSIDE the timed region; the 1/N of
the mean rides in FusedAdam(grad_scale=1/N).

`measure()` returns one dict: ms per step (CUDA events, max over ranks), the all-reduces' own device time and bytes, and - with
`parity=True` (1 GPU) - step-0 losses and every parameter gradient against the CPU oracle fed the SAME jitter / background coin.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contrastive_lift_b200 as cl  # noqa: E402
from contrastive_lift_b200 import lib as L, parallel as par, synthetic as syn  # noqa: E402

GRID, D = (128, 128, 128), 3
CHUNK, N_INS = 2048, 1024


class Cfg:
    lambda_tv_density, lambda_tv_appearance, lambda_tv_semantics, lambda_tv_instances = 0.1, 0.01, 0.0, 0.0
    late_semantic_optimization, instance_optimization_epoch = 0, 0


def main_loss(rend, model, rays, rgbs, probs, confs):
    outs = [rend(model, rays[i:i + CHUNK], 1.0, False, True) for i in range(0, rays.shape[0], CHUNK)]
    rgb = torch.cat([o[0] for o in outs])
    sem = torch.cat([o[1] for o in outs])
    dist = torch.stack([o[5] for o in outs]).mean()
    return ((rgb - rgbs) ** 2).mean() + model.total_tv_loss(None, Cfg, 5) + 0.005 * dist \
        + 0.1 * (-(probs * torch.log_softmax(sem, -1)).sum(-1) * confs).mean()


def gpu_step(model, rend, opt_main, opt_ins, batch, arenas=None, timed=False):
    rays, rgbs, probs, confs, ins_rays, labels, ins_conf = batch
    loss = main_loss(rend, model, rays, rgbs, probs, confs)
    opt_main.zero_grad(set_to_none=True)
    loss.backward()
    if arenas is not None:
        arenas[0].reduce(timed=timed)
    opt_main.step()
    feats, _ = rend.forward_instance_feature(model, ins_rays, 1.0, True)
    cl.ema_update_slownet(model.render_instance_mlp.slow_mlp, model.render_instance_mlp.mlp)
    l_ins = cl.slow_fast_loss(feats, labels, ins_conf)
    opt_ins.zero_grad(set_to_none=True)
    l_ins.backward()
    if arenas is not None:
        arenas[1].reduce(timed=timed)
    opt_ins.step()
    return loss.detach(), l_ins.detach()


def cpu_losses(params, cfg, batch, draws):
    """The same two passes on the CPU oracle (test infrastructure; the only user of oracle/ in this file), with the jitter
    and background-coin values the GPU pass consumed.  -> (loss_main, loss_ins, {name: grad})"""
    from oracle import clift_oracle as orc
    rays, rgbs, probs, confs, ins_rays, labels, ins_conf = (t.cpu() for t in batch)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    outs = []
    for i, (u, coin) in zip(range(0, rays.shape[0], CHUNK), draws["main"]):
        outs.append(orc.render_forward(p, cfg, rays[i:i + CHUNK], u, coin))
    rgb = torch.cat([o[0] for o in outs])
    sem = torch.cat([o[1] for o in outs])
    dist = torch.stack([o[5] for o in outs]).mean()
    loss = ((rgb - rgbs) ** 2).mean() + orc.total_tv_loss(p) + 0.005 * dist \
        + 0.1 * (-(probs * torch.log_softmax(sem, -1)).sum(-1) * confs).mean()
    loss.backward()
    g_main = {k: v.grad.clone() for k, v in p.items() if v.grad is not None}
    for v in p.values():
        v.grad = None
    feats, _ = orc.render_instance_feature(p, cfg, ins_rays, draws["ins"])
    l_ins = orc.slow_fast_loss(feats, labels, ins_conf)
    l_ins.backward()
    g_ins = {k: v.grad.clone() for k, v in p.items() if v.grad is not None}
    return float(loss), float(l_ins), g_main, g_ins


def replay_draws(n_main, n_ins, seed):
    """The CPU-generator draws TensoRFRenderer makes for one step after torch.manual_seed(seed), in its order
    (renderer:807-810 jitter then :164 coin per chunk call; jitter alone for forward_instance_feature)."""
    torch.manual_seed(seed)
    main = []
    for i in range(0, n_main, CHUNK):
        u = torch.rand((min(CHUNK, n_main - i), 1))
        main.append((u, bool(torch.rand((1,)) < 0.5)))
    return {"main": main, "ins": torch.rand((n_ins, 1))}


def parity_check(model, rend, params, aabb, batch, seed=777):
    """Step-0 losses and every parameter gradient of both passes, CUDA path vs CPU oracle on identical inputs and RNG."""
    from oracle import clift_oracle as orc
    rays, rgbs, probs, confs, ins_rays, labels, ins_conf = batch
    torch.manual_seed(seed)
    model.zero_grad(set_to_none=True)
    loss = main_loss(rend, model, rays, rgbs, probs, confs)
    loss.backward()
    g_main = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    feats, _ = rend.forward_instance_feature(model, ins_rays, 1.0, True)
    l_ins = cl.slow_fast_loss(feats, labels, ins_conf)
    l_ins.backward()
    g_ins = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=GRID).refresh()
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    ref_loss, ref_ins, r_main, r_ins = cpu_losses(params, cfg, batch, replay_draws(rays.shape[0], ins_rays.shape[0], seed))
    cpu_s = time.perf_counter() - t0
    worst, worst_name, missing = 0.0, "", []
    for got, ref in ((g_main, r_main), (g_ins, r_ins)):
        for k, r in ref.items():
            if float(r.abs().max()) == 0.0 and k not in got:
                continue
            if k not in got:
                missing.append(k)
                continue
            err = float((got[k] - r).norm() / r.norm().clamp_min(1e-30))
            if err > worst:
                worst, worst_name = err, k
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-12)
    out = {"loss_main": float(loss), "loss_main_cpu": ref_loss, "loss_main_rel_err": rel(float(loss), ref_loss),
           "loss_slow_fast": float(l_ins), "loss_slow_fast_cpu": ref_ins, "loss_slow_fast_rel_err": rel(float(l_ins), ref_ins),
           "grad_tensors_compared": len(r_main) + len(r_ins), "grad_worst_rel_l2": worst, "grad_worst_tensor": worst_name,
           "grad_missing": missing, "tolerance": {"loss_rel": 1e-3, "grad_rel_l2": 2e-3},
           "cpu_oracle_s": cpu_s, "cpu_cores": os.cpu_count(),
           "against": "CPU oracle (reference-pinned restatement), same rays, same jitter and background coin (seed %d)" % seed}
    out["ok"] = bool(out["loss_main_rel_err"] < 1e-3 and out["loss_slow_fast_rel_err"] < 1e-3 and worst < 2e-3 and not missing)
    return out


def measure(steps=10, warmup=3, device_index=0, with_cpu=False, torch_adam=False, profile_steps=0, rays=None, classes=None,
            parity=False, distributed=False):
    """One training-step measurement -> dict (bench.py reports it next to the render metric).  ``rays`` / ``classes``
    override the (global) main-pass batch and the class count: BASELINE config 3 = 4096 rays, C = 21 on one GPU;
    config 4 = 8192 rays, C = 2, sharded over the ranks of the initialised process group (``distributed``)."""
    import torch.distributed as dist
    B = int(rays) if rays else 4096
    C = int(classes) if classes else 21
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    dev = torch.device("cuda", device_index)
    params = syn.make_field_params(0, GRID, C, D)
    aabb = syn.default_aabb()
    model = cl.TensorVMSplit(list(GRID), num_semantic_classes=C, dim_feature_instance=2 * D, use_semantic_mlp=True,
                             use_instance_mlp=True, slow_fast_mode=True)
    model.load_state_dict(params)
    rend = cl.TensoRFRenderer(aabb, list(GRID), semantic_weight_mode="softmax")          # step_ratio 0.5 -> S = 440
    model, rend = model.to(dev), rend.to(dev)
    k, c2w = syn.camera(400, 400)
    frame = cl.get_rays_checked(400, 400, k.numpy(), c2w.numpy(), device=dev)
    g = torch.Generator().manual_seed(3)
    pick = torch.randperm(frame.shape[0], generator=g)
    b0, b1 = par.shard_range(B, rank, world)
    # global batch drawn identically on every rank, then sharded; the instance image is this rank's own
    g_rgb, g_prob, g_conf = torch.rand(B, 3, generator=g), torch.softmax(torch.randn(B, C, generator=g), -1), torch.rand(B, generator=g)
    rays_l = frame[pick[:B][b0:b1].to(dev)].contiguous()
    ins_pick = pick[B + rank * N_INS:B + (rank + 1) * N_INS]
    ins_rays = frame[ins_pick.to(dev)].contiguous()
    gi = torch.Generator().manual_seed(100 + rank)
    batch = (rays_l, g_rgb[b0:b1].to(dev), g_prob[b0:b1].to(dev), g_conf[b0:b1].to(dev), ins_rays,
             torch.randint(1, 8, (N_INS,), generator=gi).to(dev), torch.rand(N_INS, generator=gi).to(dev))
    adam = torch.optim.Adam if torch_adam else cl.FusedAdam      # one launch per param group
    kw = {} if torch_adam else {"grad_scale": 1.0 / world}
    opt_main = adam(model.get_optimizable_parameters(0.02, 0.001, 1e-8), betas=(0.9, 0.99), **kw)
    opt_ins = adam(model.get_optimizable_instance_parameters(0.02, 0.001, using_DINO=True), betas=(0.9, 0.999), **kw)
    arenas = None
    if distributed and world > 1:
        par.broadcast_parameters(model)
        arenas = [par.GradientArena([p for gr in o.param_groups for p in gr["params"]]) for o in (opt_main, opt_ins)]
    out = {}
    if parity:
        out["parity"] = parity_check(model, rend, params, aabb, batch)
    torch.manual_seed(123)           # seed_everything: every rank draws the same jitter / coin sequence (on different rays)
    if profile_steps:       # under ncu: a few bare steps, no timing, no CPU leg
        for _ in range(profile_steps):
            gpu_step(model, rend, opt_main, opt_ins, batch, arenas)
        torch.cuda.synchronize(dev)
        return None
    for _ in range(warmup):
        gpu_step(model, rend, opt_main, opt_ins, batch, arenas)
    torch.cuda.synchronize(dev)
    if distributed:
        dist.barrier()
        torch.cuda.synchronize(dev)
    torch.cuda.reset_peak_memory_stats(dev)
    l0 = L.launch_count()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        losses = gpu_step(model, rend, opt_main, opt_ins, batch, arenas, timed=True)
    t1.record()
    torch.cuda.synchronize(dev)
    rend.synchronize_overflow_checks()      # every render of the timed steps fitted its capacity (raises otherwise)
    ar_ms = [ar.drain_ms() for ar in arenas] if arenas is not None else [0.0, 0.0]
    ms = t0.elapsed_time(t1) / steps
    if distributed:
        t = torch.tensor([ms, ar_ms[0] / steps, ar_ms[1] / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ar0, ar1 = (float(v) for v in t.cpu())
    else:
        ar0 = ar1 = 0.0
    launches = (L.launch_count() - l0) / steps
    out.update({"workload": "training step: %d-ray main pass over %d GPU(s) (%d rays, %d chunk(s) per rank; MSE+TV+dist+CE, Adam) + "
                            "one 1024-ray instance image per rank (slow-fast loss, EMA, Adam), S=%d, G=128^3, C=%d, d=3+3"
                            % (B, world, b1 - b0, (b1 - b0 + CHUNK - 1) // CHUNK, rend.n_samples, C),
                "n_gpus": world, "ms_per_step": ms, "train_Mrays_per_s": (B + N_INS * world) / ms / 1e3,
                "clift_launches_per_step": launches, "optimizer": adam.__name__, "loss_main": float(losses[0]),
                "loss_slow_fast": float(losses[1])})
    if arenas is not None:
        out["allreduce"] = {"collective": "clift_allreduce_grads (NCCL sum-all-reduce through the C ABI) of one persistent fp32 arena per optimizer, inside the timed step",
                            "main_pass_ms": ar0, "instance_pass_ms": ar1,
                            "main_pass_bytes": arenas[0].nbytes, "instance_pass_bytes": arenas[1].nbytes,
                            "note": "device time between CUDA events around gather-copy + all_reduce + scatter-copy, max over "
                                    "ranks; includes waiting for the slowest rank's backward"}
    if with_cpu:
        from oracle import clift_oracle as orc
        cfg = orc.RenderConfig(aabb=aabb, grid_dim=GRID).refresh()
        torch.set_num_threads(os.cpu_count() or 1)
        c0 = time.perf_counter()
        cpu_losses(params, cfg, batch, replay_draws(b1 - b0, N_INS, 0))
        cpu_s = time.perf_counter() - c0
        out.update({"cpu_oracle_s_per_step": cpu_s, "cpu_cores": os.cpu_count(), "speedup_vs_cpu": cpu_s * 1e3 / ms})
    out["gpu_mem_peak_GB"] = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    return out


def main():
    prof = int(sys.argv[sys.argv.index("--profile-steps") + 1]) if "--profile-steps" in sys.argv else 0
    arg = lambda k: int(sys.argv[sys.argv.index(k) + 1]) if k in sys.argv else None
    distributed = "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if distributed:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = measure(with_cpu="--no-cpu" not in sys.argv and not distributed, torch_adam="--torch-adam" in sys.argv, profile_steps=prof,
                  rays=arg("--rays"), classes=arg("--classes"), parity="--parity" in sys.argv, device_index=local,
                  distributed=distributed)
    if out is not None and int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out))
    if distributed:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
