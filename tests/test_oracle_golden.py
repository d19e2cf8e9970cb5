"""The CPU oracle against the reference's own outputs (tests/golden, written by oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from contrastive_lift_b200 import synthetic as syn
from oracle import clift_oracle as orc
from oracle import refload
import golden_util as gu


def tn(x):
    return torch.from_numpy(np.asarray(x))


def test_rays_golden():
    fx = gu.load("rays")
    for i in range(3):
        h, w = (int(v) for v in fx[f"cam{i}_hw"])
        rays = orc.make_rays(h, w, tn(fx[f"cam{i}_K"]), tn(fx[f"cam{i}_c2w"]))
        assert torch.equal(rays, tn(fx[f"cam{i}_rays"]))


def test_distloss_matches_definition_and_gradcheck():
    fx = gu.load("distloss")
    w, m, iv = tn(fx["w"]), tn(fx["m"]), tn(fx["iv"])
    assert torch.allclose(orc.distortion_loss(w, m, iv), tn(fx["value"]), rtol=1e-10)
    assert torch.allclose(orc.distortion_loss_bruteforce(w, m, iv), tn(fx["value"]), rtol=1e-12)
    wg = w.clone().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda t: orc.distortion_loss(t, m, iv), (wg,))


@pytest.mark.parametrize("name", gu.RENDER_CASES)
def test_render_inference_golden(name):
    fx = gu.load(name)
    params, cfg, rays = gu.render_inputs(fx)
    with torch.no_grad():
        out, det = orc.render_forward(params, cfg, rays, None, False, detail=True)
    # indices / positions / masks: bit-exact
    assert torch.equal(det["z"].expand(rays.shape[0], -1), tn(fx["inf_z"]))
    assert torch.equal(det["inbox"], tn(fx["inf_inbox"]))
    assert torch.equal(det["xyz"], tn(fx["inf_xyz"]))
    assert torch.equal(det["active"], tn(fx["inf_active"]))
    assert torch.equal(det["sigma"], tn(fx["inf_sigma"]))
    assert torch.equal(det["weight"], tn(fx["inf_weight"]))
    for got, key in zip(out[:4], ("inf_rgb", "inf_sem", "inf_ins", "inf_depth")):
        assert torch.equal(got, tn(fx[key])), key
    assert torch.equal(out[5], tn(fx["inf_dist"]))


@pytest.mark.parametrize("name", gu.RENDER_CASES)
@pytest.mark.parametrize("tag", ["trn", "trn2"])
def test_render_training_golden(name, tag):
    fx = gu.load(name)
    params, cfg, rays = gu.render_inputs(fx)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out = orc.render_forward(p, cfg, rays, tn(fx[f"{tag}_jitter"]), bool(fx[f"{tag}_coin"]))
    for i, key in enumerate(("rgb", "sem", "ins", "depth")):
        assert torch.equal(out[i], tn(fx[f"{tag}_{key}"])), key
    loss = gu.train_loss(out, fx, tag)
    assert torch.allclose(loss.detach(), tn(fx[f"{tag}_loss"]), rtol=1e-6)
    loss.backward()
    for k, v in p.items():
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        dig = gu.grad_digest(g)
        assert np.allclose(dig, fx[f"{tag}_gdig/{k}"], rtol=1e-4, atol=1e-9), k
        if f"{tag}_grad/{k}" in fx.files:
            assert torch.allclose(g, tn(fx[f"{tag}_grad/{k}"]), rtol=1e-5, atol=1e-9), k
    # the instance head never receives gradient through a loss on rgb/sem only; here w_ins gives it one:
    assert p["render_instance_mlp.mlp.0.weight"].grad is not None
    # density factors get no gradient from the semantic / instance maps (stop_semantic_grad)


@pytest.mark.parametrize("name", gu.RENDER_CASES)
def test_instance_and_segment_golden(name):
    fx = gu.load(name)
    params, cfg, rays = gu.render_inputs(fx)
    ins, pts = orc.render_instance_feature(params, cfg, rays, tn(fx["insf_jitter"]))
    assert torch.equal(ins, tn(fx["insf_map"])) and torch.equal(pts, tn(fx["insf_pts"]))
    seg = orc.render_segment_feature(params, cfg, rays, tn(fx["segf_jitter"]))
    assert torch.equal(seg, tn(fx["segf_map"]))


def test_losses_golden():
    fx = gu.load("losses")
    ci = 0
    while f"sf{ci}_feats" in fx.files:
        feats = tn(fx[f"sf{ci}_feats"]).requires_grad_(True)
        loss = orc.slow_fast_loss(feats, tn(fx[f"sf{ci}_labels"]), tn(fx[f"sf{ci}_conf"]))
        ref = tn(fx[f"sf{ci}_loss"])
        if torch.isnan(ref):
            assert torch.isnan(loss)
        else:
            assert torch.allclose(loss.detach(), ref, rtol=1e-6, atol=0)
            if loss.requires_grad:
                loss.backward()
                assert torch.allclose(feats.grad, tn(fx[f"sf{ci}_grad"]), rtol=1e-5, atol=1e-9)
        ci += 1
    assert ci == 6
    ci = 0
    while f"ct{ci}_feats" in fx.files:
        feats = tn(fx[f"ct{ci}_feats"]).requires_grad_(True)
        loss = orc.contrastive_loss(feats, tn(fx[f"ct{ci}_labels"]), float(fx[f"ct{ci}_temp"]))
        assert torch.allclose(loss.detach(), tn(fx[f"ct{ci}_loss"]), rtol=1e-6)
        loss.backward()
        assert torch.allclose(feats.grad, tn(fx[f"ct{ci}_grad"]), rtol=1e-5, atol=1e-9)
        ci += 1
    assert ci == 3
    # EMA
    p = syn.make_field_params(int(fx["ema_seed"]), (8, 8, 8), 4, 3)
    slow = [p[f"render_instance_mlp.slow_mlp.{k}.{t}"].clone() for k in (0, 2, 4, 6) for t in ("weight", "bias")]
    fast = [p[f"render_instance_mlp.mlp.{k}.{t}"] for k in (0, 2, 4, 6) for t in ("weight", "bias")]
    orc.ema_update(slow, fast, 0.9)
    assert torch.equal(slow[-2], tn(fx["ema_last_weight"])) and torch.equal(slow[1], tn(fx["ema_first_bias"]))
    # TV
    p2 = syn.make_field_params(int(fx["tv_seed"]), tuple(int(v) for v in fx["tv_grid"]), 3, 3)
    pq = {k: v.clone().requires_grad_(True) for k, v in p2.items()}
    tot = orc.total_tv_loss(pq)
    assert torch.allclose(tot.detach(), tn(fx["tv_total"]), rtol=1e-6)
    assert torch.allclose(orc.tv_loss(p2["density_plane.0"]), tn(fx["tv_plane0"]), rtol=1e-6)
    tot.backward()
    assert torch.allclose(pq["density_plane.1"].grad, tn(fx["tv_grad_density_plane.1"]), rtol=1e-5, atol=1e-10)


def test_explicit_bilinear_matches_library_sampler():
    params = syn.make_field_params(1, (9, 13, 11), 3, 3)
    g = torch.Generator().manual_seed(0)
    xyz = torch.rand(500, 3, generator=g) * 2.2 - 1.1          # includes out-of-range taps (zero padding)
    a = orc.density(params, xyz)
    b = orc.density(params, xyz, explicit=True)
    assert torch.allclose(a, b, rtol=2e-5, atol=1e-6)
    fa = orc.vm_feature(params, "appearance", xyz)
    fb = orc.vm_feature(params, "appearance", xyz, explicit=True)
    assert torch.allclose(fa, fb, rtol=1e-4, atol=1e-6)


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_oracle_against_live_reference_random_config():
    """Fresh, un-fixtured comparison with the imported reference (build container only)."""
    grid = (12, 10, 14)
    params = syn.make_field_params(77, grid, 6, 2)
    aabb = syn.default_aabb()
    model = refload.build_model(params, grid, 6, 2)
    rend = refload.build_renderer(aabb, grid)
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=grid).refresh()
    assert cfg.n_samples == rend.n_samples
    rays = syn.random_rays(5, 64)
    with torch.no_grad():
        ref = rend(model, rays, 1.0, False, False)
        got = orc.render_forward(params, cfg, rays)
    for a, b in zip(ref, got):
        assert torch.equal(a, b)


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("sem_grid,ins_grid,slow_fast", [(32, 32, False), (48, None, True), (None, 16, True)])
def test_oracle_against_live_reference_grid_heads(sem_grid, ins_grid, slow_fast):
    """Grid-mode semantic / instance heads (allgrid.yaml family): forward, instance pass and gradients against the
    imported reference on a fresh configuration (build container only)."""
    grid = (10, 12, 9)
    params = syn.make_field_params(81, grid, 5, 3, slow_fast=slow_fast, sem_grid_comps=sem_grid, ins_grid_comps=ins_grid)
    aabb = syn.default_aabb()
    model = refload.build_model(params, grid, 5, 3, slow_fast, True, sem_grid_comps=sem_grid, ins_grid_comps=ins_grid)
    rend = refload.build_renderer(aabb, grid)
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=grid, slow_fast=slow_fast).refresh()
    rays = syn.random_rays(6, 48)
    torch.manual_seed(3)
    ref = rend(model, rays, 1.0, False, True)
    torch.manual_seed(3)
    jitter = 1.0 * torch.rand(rays.shape[0], 1)
    coin = bool(torch.rand((1,)) < 0.5)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    got = orc.render_forward(p, cfg, rays, jitter, coin)
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
    (ref[0].sum() + ref[1].exp().sum() + (ref[2] ** 2).sum()).backward()
    (got[0].sum() + got[1].exp().sum() + (got[2] ** 2).sum()).backward()
    for k, prm in model.named_parameters():
        g_ref = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        g = p[k].grad if p[k].grad is not None else torch.zeros_like(prm)
        assert torch.allclose(g_ref, g, rtol=1e-5, atol=1e-9), k
    torch.manual_seed(4)
    ins_ref, pts_ref = rend.forward_instance_feature(model, rays, 1.0, True)
    torch.manual_seed(4)
    ins, pts = orc.render_instance_feature(params, cfg, rays, 1.0 * torch.rand(rays.shape[0], 1))
    assert torch.equal(ins_ref, ins) and torch.equal(pts_ref, pts)


# ---- epoch-boundary volume operations, Adam, nearest centroid -------------------------------
EPOCH_CASES = {"a": 4, "b": 4}


def epoch_inputs(fx, tag):
    grid = tuple(int(v) for v in fx[f"{tag}_grid"])
    params = syn.make_field_params(int(fx[f"{tag}_seed"]), grid, 4, 3, ball=0.3, ball_gain=3.5)
    cfg = orc.RenderConfig(aabb=syn.default_aabb(), grid_dim=grid).refresh()
    return params, cfg


@pytest.mark.parametrize("tag", sorted(EPOCH_CASES))
def test_bbox_shrink_upsample_golden(tag):
    fx = gu.load("epoch")
    params, cfg = epoch_inputs(fx, tag)
    alpha, xyz = orc.dense_alpha(params, cfg)
    assert torch.allclose(alpha, tn(fx[f"{tag}_alpha"]), rtol=1e-6, atol=1e-9)
    lo, hi, n_valid = orc.alpha_bbox(tn(fx[f"{tag}_alpha"]), xyz)
    assert torch.equal(lo, tn(fx[f"{tag}_bbox_lo"])) and torch.equal(hi, tn(fx[f"{tag}_bbox_hi"]))
    assert n_valid == int(fx[f"{tag}_n_valid"])
    new_aabb, t_l, b_r = orc.shrink_plan(cfg, lo, hi, float(fx[f"{tag}_lenience"]))
    assert torch.equal(new_aabb, tn(fx[f"{tag}_new_aabb"]))
    assert torch.equal(t_l, tn(fx[f"{tag}_t_l"])) and torch.equal(b_r, tn(fx[f"{tag}_b_r"]))
    res = orc.target_resolution(new_aabb, int(cfg.grid_dim[0] * cfg.grid_dim[1] * cfg.grid_dim[2] * 2.5))
    assert list(res) == [int(v) for v in fx[f"{tag}_res"]]
    ups = orc.upsample_params(orc.shrink_params(params, t_l, b_r), res)
    assert torch.equal(ups["density_plane.0"], tn(fx[f"{tag}_up_plane0"]))
    digest = np.array([float(ups[k].double().sum()) for k in sorted(ups) if "plane" in k or "line" in k])
    assert np.allclose(digest, fx[f"{tag}_up_digest"], rtol=1e-12)


def test_adam_restatement_golden():
    fx = gu.load("epoch")
    for i, (lr, wd) in enumerate(((0.02, 1e-2), (0.001, 0.0))):
        p = tn(fx[f"adam_p{i}"]).clone()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for step in range(1, 6):
            orc.adam_step(p, tn(fx[f"adam_g{i}"][step - 1]), m, v, step, lr, (0.9, 0.99), 1e-8, wd)
        assert torch.allclose(p, tn(fx[f"adam_out{i}"]), rtol=1e-6, atol=1e-8)


def test_fused_adam_state_layout_matches_torch_adam():
    """The drop-in keeps torch.optim.Adam's param_group keys and per-parameter state names (checkpoint format)."""
    import contrastive_lift_b200 as cl
    p = [torch.nn.Parameter(torch.zeros(3))]
    a = torch.optim.Adam(p, lr=0.1, betas=(0.9, 0.99), weight_decay=0.5)
    b = cl.FusedAdam(p, lr=0.1, betas=(0.9, 0.99), weight_decay=0.5)
    assert set(a.param_groups[0]) <= set(b.param_groups[0]) | {"decoupled_weight_decay"}
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad", "maximize"):
        assert a.param_groups[0][k] == b.param_groups[0][k], k
    p[0].grad = torch.ones(3)
    with pytest.raises(cl.lib.CliftError):      # no CPU path
        b.step()


@pytest.mark.skipif(not refload.available(), reason="/root/reference not present")
def test_bbox_shrink_live_reference():
    import contextlib
    import io
    grid = (18, 20, 22)
    params = syn.make_field_params(91, grid, 3, 2, ball=0.4, ball_gain=3.0)
    model = refload.build_model(params, grid, 3, 2)
    rend = refload.build_renderer(syn.default_aabb(), grid)
    cfg = orc.RenderConfig(aabb=syn.default_aabb(), grid_dim=grid).refresh()
    alpha, xyz = orc.dense_alpha(params, cfg)
    lo, hi, _ = orc.alpha_bbox(alpha, xyz, rend.alpha_mask_threshold)
    new_aabb, t_l, b_r = orc.shrink_plan(cfg, lo, hi, 1.0)
    with contextlib.redirect_stdout(io.StringIO()):
        rend.update_bbox_aabb_and_shrink(model)
    assert torch.equal(rend.bbox_aabb, new_aabb)
    assert rend.grid_dim.tolist() == (b_r - t_l).tolist()


# ---- grid-mode heads (allgrid.yaml family): parameter-only paths ------------------------------------------
def test_grid_heads_tv_shrink_upsample_golden():
    fx = gu.load("grid_epoch")
    grid = tuple(int(v) for v in fx["grid"])
    params = syn.make_field_params(int(fx["seed"]), grid, 4, 3, sem_grid_comps=32, ins_grid_comps=32)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    tot = orc.total_tv_loss(p)
    assert torch.allclose(tot.detach(), tn(fx["tv_total"]), rtol=1e-6)
    # epoch 2 of the fixture: past late_semantic_optimization (1), before instance_optimization_epoch (4)
    early = orc.total_tv_loss(params, lambda_instances=0.0)
    assert torch.allclose(early, tn(fx["tv_early"]), rtol=1e-6)
    tot.backward()
    for k in ("semantic_plane.1", "semantic_line.0", "instance_plane.2", "instance_line.1"):
        assert torch.allclose(p[k].grad, tn(fx[f"tv_grad/{k}"]), rtol=1e-5, atol=1e-10), k
    shr = orc.shrink_params(params, tn(fx["t_l"]), tn(fx["b_r"]))
    ups = orc.upsample_params(shr, [int(v) for v in fx["res"]])
    digest = np.array([float(ups[k].double().sum()) for k in sorted(ups) if "plane" in k or "line" in k])
    assert np.allclose(digest, fx["up_digest"], rtol=1e-12)
    assert torch.equal(ups["semantic_plane.0"][:, :4], tn(fx["up_semantic_plane0"]))
    assert torch.equal(ups["instance_line.2"], tn(fx["up_instance_line2"]))


def test_grid_head_model_mirrors_reference_layout():
    """Host mirror in grid mode: same state_dict keys / shapes as the parameter dict the reference loads (no GPU needed)."""
    import contrastive_lift_b200 as cl
    grid = (8, 10, 12)
    for sem_grid, ins_grid, slow_fast in ((32, 32, True), (16, None, False), (None, 32, False)):
        params = syn.make_field_params(5, grid, 4, 3, slow_fast=slow_fast, sem_grid_comps=sem_grid, ins_grid_comps=ins_grid)
        model = cl.TensorVMSplit(list(grid), num_semantics_comps=(sem_grid or 32,) * 3, num_instance_comps=(ins_grid or 32,) * 3,
                                 num_semantic_classes=4, dim_feature_instance=6 if slow_fast else 3,
                                 use_semantic_mlp=not sem_grid, use_instance_mlp=not ins_grid, slow_fast_mode=slow_fast)
        res = model.load_state_dict(params, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        names = {n for g in model.get_optimizable_parameters(1e-2, 1e-3) for p in g["params"]
                 for n, q in model.named_parameters() if q is p}
        assert ("semantic_plane.0" in names) == bool(sem_grid) and "render_semantic_mlp.mlp.0.weight" in names
        ins_names = {n for g in model.get_optimizable_instance_parameters(1e-2, 1e-3) for p in g["params"]
                     for n, q in model.named_parameters() if q is p}
        assert ("instance_basis_mat.weight" in ins_names) == bool(ins_grid)


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_config1_density_rgb_slice_against_live_reference_pieces():
    """BASELINE config 1 (64x64 frame, 64 samples/ray, density + RGB heads only).  TensoRFRenderer.forward cannot skip the
    semantic / instance heads (renderer:119-131), so the slice is the reference's own pieces called one by one:
    sample_points_in_box -> normalize_coordinates -> compute_density -> raw_to_alpha -> compute_appearance_feature ->
    render_appearance_mlp -> compositing."""
    ref = refload.load()
    grid = (128, 128, 128)
    params = syn.make_field_params(0, grid, 21, 3)
    aabb = syn.default_aabb()
    ratio = orc.ratio_for_samples(aabb, grid, 64)
    model = refload.build_model(params, grid, 21, 3)
    rend = refload.build_renderer(aabb, grid)
    rend.update_step_ratio(ratio)
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=grid, step_ratio=ratio).refresh()
    assert rend.n_samples == cfg.n_samples == 64
    k, c2w = syn.camera(64, 64)
    rays = orc.make_rays(64, 64, k, c2w)
    with torch.no_grad():
        pts, z, inbox = ref.renderer.sample_points_in_box(rays, rend.bbox_aabb, rend.n_samples, rend.step_size, 1.0, False)
        dists = torch.cat((z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])), dim=-1)
        xyz = rend.normalize_coordinates(pts)
        sigma = torch.zeros(xyz.shape[:-1])
        sigma[inbox] = model.compute_density(xyz[inbox])
        _, weight, _ = rend.raw_to_alpha(sigma, dists * rend.distance_scale)
        active = weight > rend.raymarch_weight_thres
        assert int(active.sum()) > 1000                                     # the ball scene is not degenerate at S = 64
        rgb = torch.zeros((*xyz.shape[:2], 3))
        viewdirs = rays[:, 3:6].view(-1, 1, 3).expand(xyz.shape)
        rgb[active] = model.render_appearance_mlp(viewdirs[active], model.compute_appearance_feature(xyz[active]))
        rgb_ref = torch.sum(weight[..., None] * rgb, -2).clamp(0, 1)
        depth_ref = torch.sum(weight * z.expand_as(weight), -1)
    rgb_got, depth_got = orc.density_rgb_only(params, cfg, rays)
    assert torch.equal(rgb_got, rgb_ref) and torch.equal(depth_got, depth_ref)
    # and the slice is what the full forward returns in its rgb / depth slots
    full = orc.render_forward(params, cfg, rays)
    assert torch.equal(full[0], rgb_got) and torch.equal(full[3], depth_got)
