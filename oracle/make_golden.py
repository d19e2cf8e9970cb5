"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container:  python -m oracle.make_golden
Every vector stored here is an output of the reference's own code on seeded inputs; the script
also asserts that oracle/clift_oracle.py reproduces each of them (that is what "pins" the oracle).
Fixtures keep seeds + reference outputs only (inputs are regenerated from the seeds by
contrastive_lift_b200.synthetic, guarded by a parameter checksum).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from contrastive_lift_b200 import synthetic as syn   # noqa: E402
from oracle import clift_oracle as orc                # noqa: E402
from oracle import refload                            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, grid_dim, classes, max_instances, n_samples forced, softmax, slow_fast)
RENDER_CASES = [
    ("render_a", (20, 24, 16), 21, 3, 48, True, True),
    ("render_b", (16, 16, 16), 2, 3, 37, True, True),      # MOS-like: C=2, S not a multiple of 32
    ("render_c", (24, 18, 20), 5, 4, 64, False, False),    # semantic_weight_mode none, no slow net
]
# grid-mode heads (allgrid.yaml / instGRIDsemMLP.yaml / onlyRGBsegGRID.yaml; tensoRF.py:72-85, 142-156):
# name -> (semantic grid comps, instance grid comps); None = that head stays an MLP on xyz
GRID_CASES = {
    "render_d": (32, 32),      # both heads on VM grids, slow-fast instance nets sharing one feature
    "render_e": (16, None),    # semantic grid (16 comps), instance MLP, no slow net, semantic_weight_mode none
    "render_f": (None, 32),    # instGRIDsemMLP: semantic MLP, instance grid
}
RENDER_CASES += [
    ("render_d", (18, 20, 16), 5, 3, 40, True, True),
    ("render_e", (16, 18, 20), 4, 4, 48, False, False),
    ("render_f", (20, 16, 18), 3, 3, 33, True, False),
]


def case_rays(seed: int) -> torch.Tensor:
    """A 12x10 camera frame plus shuffled rays (axis-aligned directions, box misses)."""
    k, c2w = syn.camera(10, 12)
    frame = orc.make_rays(10, 12, k, c2w)
    rnd = syn.random_rays(seed, 40)
    miss = rnd[:4].clone()
    miss[:, 0:3] = torch.tensor([0.0, 0.95, 0.0])
    miss[:, 3:6] = torch.tensor([1.0, 0.0, 0.0])            # skims outside the +y face? stays inside sphere
    miss[:, 7] = orc.sphere_far(miss[:, 0:3], miss[:, 3:6])
    return torch.cat([frame, rnd, miss], 0)


def t2n(x):
    return x.detach().cpu().numpy()


def grad_digest(g: torch.Tensor) -> np.ndarray:
    """[sum, abs-sum, sq-sum] + a strided sample: keeps big MLP-weight gradients small in the fixture."""
    f = g.detach().double().flatten()
    head = torch.stack([f.sum(), f.abs().sum(), (f * f).sum()])
    return np.concatenate([t2n(head), t2n(f[::97][:512])])


def render_case(name, grid, n_cls, n_ins, n_samples, softmax, slow_fast, seed):
    ref = refload.load()
    sem_grid, ins_grid = GRID_CASES.get(name, (None, None))
    params = syn.make_field_params(seed, grid, n_cls, n_ins, slow_fast=slow_fast, sem_grid_comps=sem_grid,
                                   ins_grid_comps=ins_grid)
    aabb = syn.default_aabb()
    if name == "render_c":
        aabb = torch.tensor([[-0.9, -0.8, -1.0], [1.0, 0.7, 0.85]])
    ratio = orc.ratio_for_samples(aabb, grid, n_samples)
    model = refload.build_model(params, grid, n_cls, n_ins, slow_fast, softmax, sem_grid_comps=sem_grid,
                                ins_grid_comps=ins_grid)
    rend = refload.build_renderer(aabb, grid, softmax, True, 0.5)
    rend.update_step_ratio(ratio)
    assert rend.n_samples == n_samples, (rend.n_samples, n_samples)
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=grid, step_ratio=ratio, semantic_softmax=softmax,
                           slow_fast=slow_fast).refresh()
    assert cfg.n_samples == n_samples and torch.equal(cfg.step_size, rend.step_size)
    rays = case_rays(seed + 1)
    fx = dict(seed=seed, grid=np.array(grid), n_cls=n_cls, n_ins=n_ins, n_samples=n_samples,
              softmax=int(softmax), slow_fast=int(slow_fast), sem_grid=int(sem_grid or 0), ins_grid=int(ins_grid or 0),
              aabb=t2n(aabb), step_ratio=ratio,
              step_size=t2n(rend.step_size), inv_extent=t2n(rend.inv_box_extent), units=t2n(rend.units),
              params_checksum=syn.params_checksum(params), rays=t2n(rays))

    # ---- inference (is_train False) --------------------------------------------------
    with torch.no_grad():
        r_ref = rend(model, rays, 1.0, False, False)
        pts, z, inbox = ref.renderer.sample_points_in_box(rays, rend.bbox_aabb, rend.n_samples, rend.step_size, 1.0, False)
        xyz = rend.normalize_coordinates(pts)
        sigma = torch.zeros(z.shape)
        sigma[inbox] = model.compute_density(xyz[inbox])
        dists = torch.cat((z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])), dim=-1)
        _, weight, _ = rend.raw_to_alpha(sigma, dists * rend.distance_scale)
        r_orc, det = orc.render_forward(params, cfg, rays, None, False, detail=True)
    for a, b in zip(r_ref, r_orc):
        assert torch.equal(a, b), (name, (a - b).abs().max())
    assert torch.equal(z.expand(rays.shape[0], -1), det["z"].expand(rays.shape[0], -1))
    assert torch.equal(inbox, det["inbox"]) and torch.equal(sigma, det["sigma"]) and torch.equal(weight, det["weight"])
    # explicit-index bilinear restatement agrees with the library sampler
    s_exp = orc.density(params, xyz[inbox], explicit=True)
    assert torch.allclose(s_exp, sigma[inbox], rtol=2e-5, atol=1e-6), (s_exp - sigma[inbox]).abs().max()
    fx.update(inf_rgb=t2n(r_ref[0]), inf_sem=t2n(r_ref[1]), inf_ins=t2n(r_ref[2]), inf_depth=t2n(r_ref[3]),
              inf_dist=t2n(r_ref[5]), inf_z=t2n(z.expand(rays.shape[0], -1)), inf_inbox=t2n(inbox),
              inf_xyz=t2n(xyz), inf_sigma=t2n(sigma), inf_weight=t2n(weight),
              inf_active=t2n(weight > rend.raymarch_weight_thres))

    # ---- training forward + backward (jitter, random background, all gradients) -----------
    for bg_seed, tag in ((7, "trn"), (8, "trn2")):
        torch.manual_seed(bg_seed)
        model.zero_grad(set_to_none=True)
        out = rend(model, rays, 1.0, False, True)
        torch.manual_seed(bg_seed)
        jitter = 1.0 * torch.rand(rays.shape[0], 1)
        coin = bool(torch.rand((1,)) < 0.5)
        p_req = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        o2 = orc.render_forward(p_req, cfg, rays, jitter, coin)
        for a, b in zip(out, o2):
            assert torch.equal(a, b), (name, tag)
        gen = torch.Generator().manual_seed(100 + bg_seed)
        tgt_rgb = torch.rand(rays.shape[0], 3, generator=gen)
        probs = torch.softmax(torch.randn(rays.shape[0], n_cls, generator=gen), -1)
        w_ins = torch.randn(rays.shape[0], out[2].shape[1], generator=gen)

        def loss_fn(o):
            ce = -(probs * torch.log_softmax(o[1], -1)).sum(-1).mean()
            return ((o[0] - tgt_rgb) ** 2).mean() + 0.37 * o[5] + 0.1 * ce + 0.05 * (o[2] * w_ins).sum(-1).mean()

        l_ref = loss_fn(out)
        l_ref.backward()
        l_orc = loss_fn(o2)
        l_orc.backward()
        fx[f"{tag}_jitter"] = t2n(jitter)
        fx[f"{tag}_coin"] = int(coin)
        fx[f"{tag}_tgt_rgb"] = t2n(tgt_rgb)
        fx[f"{tag}_probs"] = t2n(probs)
        fx[f"{tag}_w_ins"] = t2n(w_ins)
        fx[f"{tag}_loss"] = t2n(l_ref)
        for i, key in enumerate(("rgb", "sem", "ins", "depth")):
            fx[f"{tag}_{key}"] = t2n(out[i])
        fx[f"{tag}_dist"] = t2n(out[5])
        for k, prm in model.named_parameters():
            g_ref = prm.grad if prm.grad is not None else torch.zeros_like(prm)
            g_orc = p_req[k].grad if p_req[k].grad is not None else torch.zeros_like(prm)
            assert torch.allclose(g_ref, g_orc, rtol=1e-5, atol=1e-9), (name, tag, k)
            if g_ref.numel() <= 20000:
                fx[f"{tag}_grad/{k}"] = t2n(g_ref)
            fx[f"{tag}_gdig/{k}"] = grad_digest(g_ref)

    # ---- instance / segment passes ------------------------------------------------------------
    torch.manual_seed(11)
    ins_map, pts_xyz = rend.forward_instance_feature(model, rays, 1.0, True)
    torch.manual_seed(11)
    jit = 1.0 * torch.rand(rays.shape[0], 1)
    oi, op = orc.render_instance_feature(params, cfg, rays, jit)
    assert torch.equal(ins_map, oi) and torch.equal(pts_xyz, op)
    torch.manual_seed(12)
    seg_map = rend.forward_segment_feature(model, rays, 1.0, True)
    torch.manual_seed(12)
    jit2 = 1.0 * torch.rand(rays.shape[0], 1)
    os_ = orc.render_segment_feature(params, cfg, rays, jit2)
    assert torch.equal(seg_map, os_)
    fx.update(insf_jitter=t2n(jit), insf_map=t2n(ins_map), insf_pts=t2n(pts_xyz),
              segf_jitter=t2n(jit2), segf_map=t2n(seg_map))
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **fx)
    print(f"{name}: rays={rays.shape[0]} S={n_samples} in-box={float(inbox.float().mean()):.3f} "
          f"active={float((weight > 1e-4).float().mean()):.3f} opacity={float(weight.sum(-1).mean()):.3f}")


def ray_cases():
    ref = refload.load()
    fx = {}
    cams = [(10, 12, 0.8, (0.0, 0.0, -0.6), 0.0), (7, 5, 1.1, (0.2, -0.1, -0.5), 25.0), (33, 64, 0.8, (0.0, 0.3, 0.4), 160.0)]
    for i, (h, w, f, pos, yaw) in enumerate(cams):
        k, c2w = syn.camera(h, w, f, pos, yaw)
        dirs = ref.ray.get_ray_directions_with_intrinsics(h, w, k.numpy())
        o, d = ref.ray.get_rays(dirs, c2w)
        far = ref.ray.rays_intersect_sphere(o, d, r=1)
        rays = torch.cat([o, d, 0.01 * torch.ones_like(far[:, None]), far[:, None]], 1)
        mine = orc.make_rays(h, w, k, c2w)
        assert torch.equal(rays, mine), (rays - mine).abs().max()
        fx[f"cam{i}_hw"] = np.array([h, w])
        fx[f"cam{i}_K"] = t2n(k)
        fx[f"cam{i}_c2w"] = t2n(c2w)
        fx[f"cam{i}_rays"] = t2n(rays)
    np.savez_compressed(os.path.join(OUT, "rays.npz"), **fx)
    print("rays: ok")


def loss_cases():
    ref = refload.load()
    fx = {}
    gen = torch.Generator().manual_seed(5)
    params = syn.make_field_params(3, (8, 8, 8), 4, 3)
    model = refload.build_model(params, (8, 8, 8), 4, 3)
    # (N, d, label range, scale)
    cases = [(1024, 3, 7, 1.0), (257, 3, 3, 0.5), (64, 6, 40, 2.0), (2, 3, 1, 1.0), (1, 3, 1, 1.0), (40, 3, 2, 1.0)]
    for ci, (n, d, nl, scale) in enumerate(cases):
        feats = (scale * torch.randn(n, 2 * d, generator=gen)).requires_grad_(True)
        labels = torch.randint(1, nl + 1, (n,), generator=gen)
        if ci == 5:     # disjoint label sets in the two halves: no positives, empty intersection
            labels[: n // 2] = 1
            labels[n // 2:] = 2
        conf = torch.rand(n, generator=gen)
        model.dim_feature_instance = 2 * d
        l_ref = refload.slow_fast_loss(model, labels, feats, conf, use_ema=False)
        f2 = feats.detach().clone().requires_grad_(True)
        l_orc = orc.slow_fast_loss(f2, labels, conf)
        assert torch.equal(l_ref.detach(), l_orc.detach()) or (torch.isnan(l_ref) and torch.isnan(l_orc)), (ci, l_ref, l_orc)
        grad = torch.zeros_like(feats)
        if l_ref.requires_grad and not torch.isnan(l_ref):
            l_ref.backward()
            l_orc.backward()
            grad = feats.grad
            assert torch.allclose(grad, f2.grad, rtol=1e-6, atol=1e-10)
        fx.update({f"sf{ci}_feats": t2n(feats), f"sf{ci}_labels": t2n(labels), f"sf{ci}_conf": t2n(conf),
                   f"sf{ci}_loss": t2n(l_ref), f"sf{ci}_grad": t2n(grad)})
        print(f"slow_fast case {ci}: N={n} d={d} loss={float(l_ref):.6f}")
    # EMA (trainer:325-329) on the instance MLP pair
    model = refload.build_model(params, (8, 8, 8), 4, 3)
    holder = type("H", (), {})()
    ref.trainer.TensoRFTrainer.ema_update_slownet(holder, model.render_instance_mlp.slow_mlp, model.render_instance_mlp.mlp, 0.9)
    slow = [params[f"render_instance_mlp.slow_mlp.{k}.{t}"].clone() for k in (0, 2, 4, 6) for t in ("weight", "bias")]
    fast = [params[f"render_instance_mlp.mlp.{k}.{t}"] for k in (0, 2, 4, 6) for t in ("weight", "bias")]
    orc.ema_update(slow, fast, 0.9)
    for a, b in zip(slow, model.render_instance_mlp.slow_mlp.parameters()):
        assert torch.equal(a, b)
    fx["ema_seed"] = 3
    fx["ema_last_weight"] = t2n(slow[-2])
    fx["ema_first_bias"] = t2n(slow[1])
    # vanilla contrastive (loss.py:62-82)
    for ci, (n, d, nl, temp) in enumerate([(512, 3, 6, 100.0), (33, 3, 3, 1.0), (128, 8, 128, 10.0)]):
        feats = torch.randn(n, d, generator=gen).requires_grad_(True)
        labels = torch.randint(1, nl + 1, (n,), generator=gen)
        l_ref = ref.loss.contrastive_loss(feats, labels, temp)
        l_ref.backward()
        f2 = feats.detach().clone().requires_grad_(True)
        l_orc = orc.contrastive_loss(f2, labels, temp)
        l_orc.backward()
        assert torch.equal(l_ref.detach(), l_orc.detach()) and torch.allclose(feats.grad, f2.grad, rtol=1e-6, atol=1e-10)
        fx.update({f"ct{ci}_feats": t2n(feats), f"ct{ci}_labels": t2n(labels), f"ct{ci}_temp": temp,
                   f"ct{ci}_loss": t2n(l_ref), f"ct{ci}_grad": t2n(feats.grad)})
    # TV (loss.py:9-26) + total_tv_loss (tensoRF.py:281-290) on a small non-square field
    p2 = syn.make_field_params(9, (10, 14, 12), 3, 3)
    m2 = refload.build_model(p2, (10, 14, 12), 3, 3)
    tv = ref.loss.TVLoss()
    cfgns = type("C", (), dict(late_semantic_optimization=1, instance_optimization_epoch=4, lambda_tv_density=0.1,
                               lambda_tv_appearance=0.01, lambda_tv_semantics=0.02, lambda_tv_instances=0.02))()
    tot = m2.total_tv_loss(tv, cfgns, 5)
    tot.backward()
    pq = {k: v.clone().requires_grad_(True) for k, v in p2.items()}
    tot_o = orc.total_tv_loss(pq)
    tot_o.backward()
    assert torch.allclose(tot.detach(), tot_o.detach(), rtol=1e-6)
    fx["tv_seed"] = 9
    fx["tv_grid"] = np.array((10, 14, 12))
    fx["tv_total"] = t2n(tot)
    fx["tv_plane0"] = t2n(tv(p2["density_plane.0"]))
    fx["tv_grad_density_plane.1"] = t2n(m2.density_plane[1].grad)
    fx["tv_grad_appearance_plane.2"] = t2n(m2.appearance_plane[2].grad[:, :8])
    assert torch.allclose(m2.density_plane[1].grad, pq["density_plane.1"].grad, rtol=1e-5, atol=1e-10)
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **fx)
    print("losses: ok")


def distloss_case():
    """The restated eff_distloss vs the O(S^2) definition + fp64 gradcheck (parity unpinned vs the package)."""
    gen = torch.Generator().manual_seed(21)
    w = torch.rand(6, 17, generator=gen, dtype=torch.float64) * 0.1
    z = torch.cumsum(torch.rand(6, 18, generator=gen, dtype=torch.float64) * 0.1, -1)
    m = (z[:, 1:] + z[:, :-1]) / 2
    iv = z[:, 1:] - z[:, :-1]
    a = orc.distortion_loss(w, m, iv)
    b = orc.distortion_loss_bruteforce(w, m, iv)
    assert torch.allclose(a, b, rtol=1e-10), (a, b)
    wg = w.clone().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda t: orc.distortion_loss(t, m, iv), (wg,))
    np.savez_compressed(os.path.join(OUT, "distloss.npz"), w=t2n(w), m=t2n(m), iv=t2n(iv), value=t2n(b))
    print("distloss: ok")


def epoch_cases():
    """the reference's update_bbox_aabb_and_shrink / upsample_volume_grid / get_target_resolution
    on a ball scene vs the oracle restatement; rank 2: restated Adam vs torch.optim.Adam."""
    import contextlib
    import io
    out = {}
    for tag, grid, seed, lenience in (("a", (20, 24, 18), 61, 1.0), ("b", (26, 22, 30), 62, 1.2)):
        params = syn.make_field_params(seed, grid, 4, 3, ball=0.3, ball_gain=3.5)
        aabb = syn.default_aabb()
        model = refload.build_model(params, grid, 4, 3)
        rend = refload.build_renderer(aabb, grid)
        cfg = orc.RenderConfig(aabb=aabb.clone(), grid_dim=grid).refresh()
        alpha_ref, xyz_ref = rend.get_dense_alpha(model)
        alpha, xyz = orc.dense_alpha(params, cfg)
        assert torch.equal(xyz, xyz_ref) and torch.allclose(alpha, alpha_ref, rtol=1e-6, atol=1e-9), tag
        lo, hi, n_valid = orc.alpha_bbox(alpha_ref, xyz_ref, rend.alpha_mask_threshold)
        new_aabb, t_l, b_r = orc.shrink_plan(cfg, lo, hi, lenience)
        with contextlib.redirect_stdout(io.StringIO()):
            rend.update_bbox_aabb_and_shrink(model, lenience)
        assert torch.equal(rend.bbox_aabb, new_aabb), (tag, rend.bbox_aabb, new_aabb)
        assert rend.grid_dim.tolist() == (b_r - t_l).tolist(), tag
        shr = orc.shrink_params(params, t_l, b_r)
        sd = model.state_dict()
        for k in shr:
            assert torch.equal(sd[k], shr[k]), (tag, k)
        n_vox = int(grid[0] * grid[1] * grid[2] * 2.5)
        res = rend.get_target_resolution(n_vox)
        assert tuple(res) == orc.target_resolution(new_aabb, n_vox), tag
        model.upsample_volume_grid(res)
        ups = orc.upsample_params(shr, res)
        sd = model.state_dict()
        for k in ups:
            assert torch.equal(sd[k], ups[k]), (tag, k)
        out.update({f"{tag}_grid": np.array(grid), f"{tag}_seed": np.array(seed), f"{tag}_lenience": np.array(lenience),
                    f"{tag}_alpha": t2n(alpha_ref), f"{tag}_bbox_lo": t2n(lo), f"{tag}_bbox_hi": t2n(hi),
                    f"{tag}_n_valid": np.array(n_valid), f"{tag}_new_aabb": t2n(new_aabb), f"{tag}_t_l": t2n(t_l),
                    f"{tag}_b_r": t2n(b_r), f"{tag}_res": np.array(res),
                    f"{tag}_up_digest": np.array([float(ups[k].double().sum()) for k in sorted(ups) if "plane" in k or "line" in k]),
                    f"{tag}_up_plane0": t2n(ups["density_plane.0"])})
    # Adam: restatement vs torch.optim.Adam, 5 steps, two groups (weight decay on / off)
    gen = torch.Generator().manual_seed(77)
    ps = [torch.randn(37, 5, generator=gen), torch.randn(130, generator=gen)]
    ref_p = [torch.nn.Parameter(p.clone()) for p in ps]
    opt = torch.optim.Adam([{"params": [ref_p[0]], "lr": 0.02, "weight_decay": 1e-2}, {"params": [ref_p[1]], "lr": 0.001}],
                           betas=(0.9, 0.99))
    mine = [p.clone() for p in ps]
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    grads = []
    for step in range(1, 6):
        gs = [torch.randn(p.shape, generator=gen) * (0.5 ** step) for p in ps]
        grads.append(gs)
        for rp, g in zip(ref_p, gs):
            rp.grad = g.clone()
        opt.step()
        orc.adam_step(mine[0], gs[0], ms[0], vs[0], step, 0.02, (0.9, 0.99), 1e-8, 1e-2)
        orc.adam_step(mine[1], gs[1], ms[1], vs[1], step, 0.001, (0.9, 0.99), 1e-8, 0.0)
    for a, b in zip(mine, ref_p):
        assert torch.allclose(a, b.data, rtol=1e-6, atol=1e-8), float((a - b.data).abs().max())
    out.update({"adam_p0": t2n(ps[0]), "adam_p1": t2n(ps[1]), "adam_out0": t2n(ref_p[0].data), "adam_out1": t2n(ref_p[1].data),
                "adam_g0": np.stack([t2n(g[0]) for g in grads]), "adam_g1": np.stack([t2n(g[1]) for g in grads])})
    np.savez_compressed(os.path.join(OUT, "epoch.npz"), **out)
    print("epoch (bbox / shrink / upsample / adam): ok")


def grid_epoch_case():
    """Grid-mode heads through the parameter-only paths: total_tv_loss with semantic / instance planes AND lines
    (tensoRF.py:260-290), shrink and upsample_volume_grid of all four factor sets (tensoRF.py:158-197)."""
    ref = refload.load()
    grid, seed = (12, 16, 14), 71
    params = syn.make_field_params(seed, grid, 4, 3, sem_grid_comps=32, ins_grid_comps=32)
    model = refload.build_model(params, grid, 4, 3, sem_grid_comps=32, ins_grid_comps=32)
    tv = ref.loss.TVLoss()
    cfgns = type("C", (), dict(late_semantic_optimization=1, instance_optimization_epoch=4, lambda_tv_density=0.1,
                               lambda_tv_appearance=0.01, lambda_tv_semantics=0.02, lambda_tv_instances=0.02))()
    tot = model.total_tv_loss(tv, cfgns, 5)
    tot.backward()
    pq = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    tot_o = orc.total_tv_loss(pq)
    tot_o.backward()
    assert torch.allclose(tot.detach(), tot_o.detach(), rtol=1e-6)
    fx = dict(seed=seed, grid=np.array(grid), tv_total=t2n(tot), tv_early=t2n(model.total_tv_loss(tv, cfgns, 2)))
    for k in ("semantic_plane.1", "semantic_line.0", "instance_plane.2", "instance_line.1"):
        g = dict(model.named_parameters())[k].grad
        assert torch.allclose(g, pq[k].grad, rtol=1e-5, atol=1e-10), k
        fx[f"tv_grad/{k}"] = t2n(g)
    t_l, b_r = torch.tensor([1, 2, 0]), torch.tensor([11, 15, 13])
    with torch.no_grad():
        model.shrink(t_l, b_r)
    shr = orc.shrink_params(params, t_l, b_r)
    sd = model.state_dict()
    for k in shr:
        assert torch.equal(sd[k], shr[k]), k
    res = (17, 21, 19)
    model.upsample_volume_grid(res)
    ups = orc.upsample_params(shr, res)
    sd = model.state_dict()
    for k in ups:
        assert torch.equal(sd[k], ups[k]), k
    fx.update(t_l=t2n(t_l), b_r=t2n(b_r), res=np.array(res),
              up_digest=np.array([float(ups[k].double().sum()) for k in sorted(ups) if "plane" in k or "line" in k]),
              up_semantic_plane0=t2n(ups["semantic_plane.0"][:, :4]), up_instance_line2=t2n(ups["instance_line.2"]))
    np.savez_compressed(os.path.join(OUT, "grid_epoch.npz"), **fx)
    print("grid_epoch (tv / shrink / upsample with semantic + instance grids): ok")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    only = set(sys.argv[1:])          # e.g. `python -m oracle.make_golden render_d grid_epoch`: regenerate just these
    want = lambda n: not only or n in only
    if want("rays"):
        ray_cases()
    if want("distloss"):
        distloss_case()
    if want("losses"):
        loss_cases()
    if want("epoch"):
        epoch_cases()
    if want("grid_epoch"):
        grid_epoch_case()
    for i, (name, grid, c, d, s, sm, sf) in enumerate(RENDER_CASES):
        if want(name):
            render_case(name, grid, c, d, s, sm, sf, seed=40 + i)
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"fixtures: {tot / 1e6:.2f} MB in {OUT}")


if __name__ == "__main__":
    main()
