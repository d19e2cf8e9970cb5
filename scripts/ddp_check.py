"""torchrun script: data-parallel equivalence of the render path on N GPUs.

Every rank renders its shard of one ray batch with replicated parameters, gradients are summed through
parallel.allreduce_gradients (one flat NCCL all-reduce), and rank 0 checks them against the gradient of the whole
batch rendered on one GPU.  Also checks that sharded inference gathers back to the single-GPU frame.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/ddp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contrastive_lift_b200 as cl  # noqa: E402
from contrastive_lift_b200 import parallel as par, synthetic as syn  # noqa: E402


def build(dev):
    grid = (32, 32, 32)
    params = syn.make_field_params(5, grid, 7, 3)
    model = cl.TensorVMSplit(list(grid), num_semantic_classes=7, dim_feature_instance=6, use_semantic_mlp=True,
                             use_instance_mlp=True, slow_fast_mode=True)
    model.load_state_dict(params)
    rend = cl.TensoRFRenderer(syn.default_aabb(), list(grid), semantic_weight_mode="softmax")
    return model.to(dev), rend.to(dev)


def loss_of(out, tgt, n_total):
    rgb, sem, ins, depth, _, dist_reg = out
    # sums over local rays / global count: the all-reduced gradient is then the gradient of the global mean
    return ((rgb - tgt) ** 2).sum() / (3 * n_total) + 0.1 * (-sem[:, 1]).sum() / n_total + 0.05 * ins.sum() / n_total


def check(dev, verbose=True):
    """Runs on every rank of an initialised NCCL process group -> dict (identical on all ranks): sharded training gradients
    after the arena all-reduce vs the whole batch on rank 0, and sharded inference gathered back vs the single-GPU frame."""
    rank, world = dist.get_rank(), dist.get_world_size()
    model, rend = build(dev)
    par.broadcast_parameters(model)
    rays = syn.random_rays(9, 1024).to(dev)
    tgt = torch.rand(1024, 3, generator=torch.Generator().manual_seed(1)).to(dev)
    b, e = par.shard_range(rays.shape[0], rank, world)
    out = rend(model, rays[b:e].contiguous(), 0.0, True, True)       # perturb 0 / white bg: no RNG in the comparison
    loss_of(out, tgt[b:e], rays.shape[0]).backward()
    nbytes = par.allreduce_gradients(model.parameters(), average=False)
    grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    with torch.no_grad():
        inf = rend(model, rays[b:e].contiguous(), 0.0, True, False)
        full_rgb = par.gather_rays_output(inf[0], rays.shape[0])
    res = torch.zeros(3, device=dev, dtype=torch.float64)            # ok, worst gradient error, inference equal
    if rank == 0:
        ok = True
        model.zero_grad(set_to_none=True)
        out = rend(model, rays, 0.0, True, True)
        loss_of(out, tgt, rays.shape[0]).backward()
        worst = 0.0
        for k, p in model.named_parameters():
            if p.grad is None:
                continue
            ref = p.grad
            err = float((grads[k] - ref).abs().max() / ref.abs().max().clamp_min(1e-20))
            worst = max(worst, err)
            if err > 2e-3:
                ok = False
                if verbose:
                    print(f"MISMATCH {k}: {err:.3e}")
        with torch.no_grad():
            single = rend(model, rays, 0.0, True, False)[0]
        same = bool(torch.allclose(single, full_rgb, rtol=1e-5, atol=1e-6))
        ok = ok and same
        res = torch.tensor([1.0 if ok else 0.0, worst, 1.0 if same else 0.0], device=dev, dtype=torch.float64)
    dist.broadcast(res, 0)
    ok, worst, same = (float(v) for v in res.cpu())
    return {"ok": bool(ok), "world": world, "allreduced_bytes": int(nbytes), "worst_rel_grad_err": worst,
            "sharded_inference_equals_single_gpu": bool(same), "tolerance": 2e-3}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    r = check(dev)
    if rank == 0:
        print(f"ddp_check world={world}: all-reduced {r['allreduced_bytes']} B, worst relative gradient error "
              f"{r['worst_rel_grad_err']:.2e}, sharded inference == single-GPU frame: {r['sharded_inference_equals_single_gpu']} "
              f"-> {'OK' if r['ok'] else 'FAIL'}")
    dist.destroy_process_group()
    sys.exit(0 if r["ok"] else 1)


if __name__ == "__main__":
    main()
