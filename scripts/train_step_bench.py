"""BASELINE config 3: one training step on 1 GPU - 4096-ray batch, slow-fast contrastive loss, forward + backward.

Replays the hot-path calls of TensoRFTrainer.training_step (trainer/train_panopli_tensorf.py:148-228):
  (A) main pass: chunked render (chunk 2048) -> MSE + TV + dist-reg + semantic CE -> backward -> Adam step
  (B) instance pass: forward_instance_feature on 1024 rays of one image -> EMA -> slow-fast loss -> backward -> Adam step
and prints one JSON line with the step time (CUDA events) next to the CPU oracle's time for the same step.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contrastive_lift_b200 as cl  # noqa: E402
from contrastive_lift_b200 import lib as L, synthetic as syn  # noqa: E402

GRID, C, D = (128, 128, 128), 21, 3
B, CHUNK, N_INS = 4096, 2048, 1024


class Cfg:
    lambda_tv_density, lambda_tv_appearance, lambda_tv_semantics, lambda_tv_instances = 0.1, 0.01, 0.0, 0.0
    late_semantic_optimization, instance_optimization_epoch = 0, 0


def gpu_step(model, rend, opt_main, opt_ins, batch):
    rays, rgbs, probs, confs, ins_rays, labels, ins_conf = batch
    outs = [rend(model, rays[i:i + CHUNK], 1.0, False, True) for i in range(0, B, CHUNK)]
    rgb = torch.cat([o[0] for o in outs])
    sem = torch.cat([o[1] for o in outs])
    dist = torch.stack([o[5] for o in outs]).mean()
    loss = ((rgb - rgbs) ** 2).mean() + model.total_tv_loss(None, Cfg, 5) + 0.005 * dist \
        + 0.1 * (-(probs * torch.log_softmax(sem, -1)).sum(-1) * confs).mean()
    opt_main.zero_grad(set_to_none=True)
    loss.backward()
    opt_main.step()
    feats, _ = rend.forward_instance_feature(model, ins_rays, 1.0, True)
    cl.ema_update_slownet(model.render_instance_mlp.slow_mlp, model.render_instance_mlp.mlp)
    l_ins = cl.slow_fast_loss(feats, labels, ins_conf)
    opt_ins.zero_grad(set_to_none=True)
    l_ins.backward()
    opt_ins.step()
    return loss.detach(), l_ins.detach()


def cpu_step(params, cfg, batch):
    from oracle import clift_oracle as orc      # the CPU leg is the only user of the oracle
    rays, rgbs, probs, confs, ins_rays, labels, ins_conf = (t.cpu() for t in batch)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    g = torch.Generator().manual_seed(0)
    outs = [orc.render_forward(p, cfg, rays[i:i + CHUNK], torch.rand(min(CHUNK, B - i), 1, generator=g), False) for i in range(0, B, CHUNK)]
    rgb = torch.cat([o[0] for o in outs])
    sem = torch.cat([o[1] for o in outs])
    dist = torch.stack([o[5] for o in outs]).mean()
    loss = ((rgb - rgbs) ** 2).mean() + orc.total_tv_loss(p) + 0.005 * dist + 0.1 * (-(probs * torch.log_softmax(sem, -1)).sum(-1) * confs).mean()
    loss.backward()
    feats, _ = orc.render_instance_feature(p, cfg, ins_rays, torch.rand(N_INS, 1, generator=g))
    orc.slow_fast_loss(feats, labels, ins_conf).backward()


def measure(steps=10, warmup=3, device_index=0, with_cpu=False, torch_adam=False, profile_steps=0, rays=None, classes=None):
    """One training-step measurement -> dict (bench.py reports it next to the render metric).  ``rays`` / ``classes``
    override the main-pass batch and the class count (BASELINE config 4 per-GPU share: 1024 rays, C = 2)."""
    global B, C
    if rays:
        B = int(rays)
    if classes:
        C = int(classes)
    dev = torch.device("cuda", device_index)
    params = syn.make_field_params(0, GRID, C, D)
    aabb = syn.default_aabb()
    model = cl.TensorVMSplit(list(GRID), num_semantic_classes=C, dim_feature_instance=2 * D, use_semantic_mlp=True,
                             use_instance_mlp=True, slow_fast_mode=True)
    model.load_state_dict(params)
    rend = cl.TensoRFRenderer(aabb, list(GRID), semantic_weight_mode="softmax")          # step_ratio 0.5 -> S = 440
    model, rend = model.to(dev), rend.to(dev)
    k, c2w = syn.camera(400, 400)
    frame = cl.get_rays_checked(400, 400, k.numpy(), c2w.numpy())
    g = torch.Generator().manual_seed(3)
    pick = torch.randperm(frame.shape[0], generator=g)
    rays = frame[pick[:B].to(dev)].contiguous()
    ins_rays = frame[pick[B:B + N_INS].to(dev)].contiguous()
    batch = (rays, torch.rand(B, 3, generator=g).to(dev), torch.softmax(torch.randn(B, C, generator=g), -1).to(dev),
             torch.rand(B, generator=g).to(dev), ins_rays, torch.randint(1, 8, (N_INS,), generator=g).to(dev),
             torch.rand(N_INS, generator=g).to(dev))
    adam = torch.optim.Adam if torch_adam else cl.FusedAdam      # one launch per param group
    opt_main = adam(model.get_optimizable_parameters(0.02, 0.001, 1e-8), betas=(0.9, 0.99))
    opt_ins = adam(model.get_optimizable_instance_parameters(0.02, 0.001, using_DINO=True), betas=(0.9, 0.999))
    torch.manual_seed(123)
    if profile_steps:       # under ncu: a few bare steps, no timing, no CPU leg
        for _ in range(profile_steps):
            gpu_step(model, rend, opt_main, opt_ins, batch)
        torch.cuda.synchronize(dev)
        return None
    for _ in range(warmup):
        gpu_step(model, rend, opt_main, opt_ins, batch)
    torch.cuda.synchronize(dev)
    l0 = L.launch_count()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        losses = gpu_step(model, rend, opt_main, opt_ins, batch)
    t1.record()
    torch.cuda.synchronize(dev)
    ms = t0.elapsed_time(t1) / steps
    launches = (L.launch_count() - l0) / steps
    out = {"workload": "training step: %d-ray main pass (%d chunk(s), MSE+TV+dist+CE, Adam) + 1024-ray instance pass "
                       "(slow-fast loss, EMA, Adam), S=%d, G=128^3, C=%d, d=3+3" % (B, (B + CHUNK - 1) // CHUNK, rend.n_samples, C),
           "ms_per_step": ms, "train_Mrays_per_s": (B + N_INS) / ms / 1e3, "clift_launches_per_step": launches,
           "optimizer": adam.__name__, "loss_main": float(losses[0]), "loss_slow_fast": float(losses[1])}
    if with_cpu:
        from oracle import clift_oracle as orc
        cfg = orc.RenderConfig(aabb=aabb, grid_dim=GRID).refresh()
        torch.set_num_threads(os.cpu_count() or 1)
        c0 = time.perf_counter()
        cpu_step(params, cfg, batch)
        cpu_s = time.perf_counter() - c0
        out.update({"cpu_oracle_s_per_step": cpu_s, "cpu_cores": os.cpu_count(), "speedup_vs_cpu": cpu_s * 1e3 / ms})
    out["gpu_mem_peak_GB"] = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    return out


def main():
    prof = int(sys.argv[sys.argv.index("--profile-steps") + 1]) if "--profile-steps" in sys.argv else 0
    arg = lambda k: int(sys.argv[sys.argv.index(k) + 1]) if k in sys.argv else None
    out = measure(with_cpu="--no-cpu" not in sys.argv, torch_adam="--torch-adam" in sys.argv, profile_steps=prof,
                  rays=arg("--rays"), classes=arg("--classes"))
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
