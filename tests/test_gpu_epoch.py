"""-m gpu parity tests of the SURVEY 8(f) rows: fused Adam, dense-alpha / bounding box / shrink, bilinear factor
upsampling, nearest-centroid assignment - through the C ABI, against the reference-pinned fixtures and the oracle."""
import numpy as np
import pytest
import torch

import contrastive_lift_b200 as cl
from contrastive_lift_b200 import lib as L
from contrastive_lift_b200 import synthetic as syn
from oracle import clift_oracle as orc
import golden_util as gu
import gpu_util as gpu

pytestmark = pytest.mark.gpu


def tn(x):
    return torch.from_numpy(np.asarray(x))


def epoch_case(fx, tag):
    grid = tuple(int(v) for v in fx[f"{tag}_grid"])
    params = syn.make_field_params(int(fx[f"{tag}_seed"]), grid, 4, 3, ball=0.3, ball_gain=3.5)
    model, rend = gpu.build(params, grid, 4, 3, True, True, syn.default_aabb(), 0.5)
    return params, grid, model, rend


@pytest.mark.parametrize("tag", ["a", "b"])
def test_dense_alpha_bbox_shrink_upsample_golden(tag):
    fx = gu.load("epoch")
    params, grid, model, rend = epoch_case(fx, tag)
    alpha, xyz = rend.get_dense_alpha(model)
    ref_alpha = tn(fx[f"{tag}_alpha"])
    assert gpu.rel_err(alpha, ref_alpha) < 1e-5
    cfg = orc.RenderConfig(aabb=syn.default_aabb(), grid_dim=grid).refresh()
    assert torch.equal(xyz.cpu(), orc.dense_alpha(params, cfg)[1])            # lattice positions: bit-exact
    lo, hi, n_valid = rend.alpha_bbox(model)
    assert torch.equal(lo.cpu(), tn(fx[f"{tag}_bbox_lo"])) and torch.equal(hi.cpu(), tn(fx[f"{tag}_bbox_hi"]))
    assert n_valid == int(fx[f"{tag}_n_valid"])
    rend.update_bbox_aabb_and_shrink(model, float(fx[f"{tag}_lenience"]))
    assert torch.equal(rend.bbox_aabb.cpu(), tn(fx[f"{tag}_new_aabb"]))      # bit-exact new box
    t_l, b_r = tn(fx[f"{tag}_t_l"]), tn(fx[f"{tag}_b_r"])
    assert rend.grid_dim.tolist() == (b_r - t_l).tolist() and list(model.grid_dim()) == (b_r - t_l).tolist()
    shr = orc.shrink_params(params, t_l, b_r)
    sd = model.state_dict()
    for k in ("density_plane.0", "density_line.2", "appearance_plane.1", "appearance_line.0"):
        assert torch.equal(sd[k].cpu(), shr[k]), k
    res = rend.get_target_resolution(int(grid[0] * grid[1] * grid[2] * 2.5))
    assert list(res) == [int(v) for v in fx[f"{tag}_res"]]
    model.upsample_volume_grid(res)
    rend.update_step_size(res)
    ups = orc.upsample_params(shr, res)
    sd = model.state_dict()
    worst = max(gpu.rel_err(sd[k], ups[k]) for k in ups if "plane" in k or "line" in k)
    assert worst < 1e-6, worst
    assert gpu.rel_err(sd["density_plane.0"], tn(fx[f"{tag}_up_plane0"])) < 1e-6
    # the resized model renders (packed view rebuilt, renderer geometry follows)
    rays = syn.random_rays(3, 64).cuda()
    with torch.no_grad():
        out = rend(model, rays, 1.0, False, False)
    params2 = {k: v.cpu() for k, v in model.state_dict().items()}
    cfg2 = orc.RenderConfig(aabb=rend.bbox_aabb.cpu(), grid_dim=tuple(res), step_ratio=rend.step_ratio).refresh()
    ref = orc.render_forward(params2, cfg2, rays.cpu())
    assert gpu.rel_err(out[0], ref[0]) < 1e-4 and gpu.rel_err(out[3], ref[3]) < 1e-4


def test_empty_volume_keeps_the_box():
    grid = (12, 12, 12)
    params = syn.make_field_params(5, grid, 4, 3, ball=None)         # sigma ~ softplus(-10): nothing above threshold
    model, rend = gpu.build(params, grid, 4, 3, True, True, syn.default_aabb(), 0.5)
    before = rend.bbox_aabb.clone()
    lo, hi, n_valid = rend.alpha_bbox(model)
    assert n_valid == 0
    rend.update_bbox_aabb_and_shrink(model)
    assert torch.equal(rend.bbox_aabb, before) and list(model.grid_dim()) == list(grid)


def test_full_size_alpha_sweep_properties():
    """128^3 lattice (2.1 M voxels): the ball scene is mirror-symmetric, so is its bounding box; the count matches a
    dense threshold of the max-pooled alpha done with torch ops on the kernel's own alpha."""
    grid = (128, 128, 128)
    params = syn.make_field_params(0, grid, 21, 3)
    model, rend = gpu.build(params, grid, 21, 3, True, True, syn.default_aabb(), 0.5)
    alpha, xyz = rend.get_dense_alpha(model)
    lo, hi, n_valid = rend.alpha_bbox(model)
    pooled = torch.nn.functional.max_pool3d(alpha.clamp(0, 1)[None, None], 3, 1, 1)[0, 0]
    valid = pooled >= rend.alpha_mask_threshold
    assert n_valid == int(valid.sum())
    pts = xyz[valid]
    assert torch.equal(lo, pts.amin(0)) and torch.equal(hi, pts.amax(0))
    assert 0 < n_valid < 128 ** 3 and float((lo + hi).abs().max()) < 0.1


def _adam_run(opt_cls, fx, device):
    ps = [torch.nn.Parameter(tn(fx["adam_p0"]).clone().to(device)), torch.nn.Parameter(tn(fx["adam_p1"]).clone().to(device))]
    opt = opt_cls([{"params": [ps[0]], "lr": 0.02, "weight_decay": 1e-2}, {"params": [ps[1]], "lr": 0.001}], betas=(0.9, 0.99))
    for step in range(5):
        ps[0].grad = tn(fx["adam_g0"][step]).to(device)
        ps[1].grad = tn(fx["adam_g1"][step]).to(device)
        opt.step()
    return ps, opt


def test_fused_adam_golden_and_state_interchange():
    fx = gu.load("epoch")
    launches = L.load().clift_launch_count()
    ps, opt = _adam_run(cl.FusedAdam, fx, "cuda")
    assert L.load().clift_launch_count() - launches == 5                # one launch per step for both groups
    for i in range(2):
        assert gpu.rel_err(ps[i].data, tn(fx[f"adam_out{i}"])) < 2e-6
    # state_dict interchange with the stock optimizer: continue the run with torch.optim.Adam and vice versa
    ref_ps, ref_opt = _adam_run(torch.optim.Adam, fx, "cuda")
    stock = torch.optim.Adam([{"params": [ps[0]], "lr": 0.02, "weight_decay": 1e-2}, {"params": [ps[1]], "lr": 0.001}], betas=(0.9, 0.99))
    stock.load_state_dict(opt.state_dict())
    mine = cl.FusedAdam([{"params": [ref_ps[0]], "lr": 0.02, "weight_decay": 1e-2}, {"params": [ref_ps[1]], "lr": 0.001}], betas=(0.9, 0.99))
    mine.load_state_dict(ref_opt.state_dict())
    g = [torch.full_like(p, 0.01) for p in ps]
    for p, q, gg in zip(ps, ref_ps, g):
        p.grad, q.grad = gg.clone(), gg.clone()
    stock.step()
    mine.step()
    for p, q in zip(ps, ref_ps):
        assert gpu.rel_err(p.data, q.data) < 2e-6


def test_fused_adam_large_ragged_group():
    """One launch over tensors of very different sizes (incl. an unaligned view and an empty-grad parameter)."""
    gen = torch.Generator().manual_seed(3)
    shapes = [(1, 48, 192, 192), (1, 16, 192, 1), (27, 144), (3,), (1027,)]
    base = [torch.randn(s, generator=gen) for s in shapes]
    a = [torch.nn.Parameter(b.clone().cuda()) for b in base]
    b = [torch.nn.Parameter(b.clone().cuda()) for b in base]
    skipped_a, skipped_b = torch.nn.Parameter(torch.ones(5).cuda()), torch.nn.Parameter(torch.ones(5).cuda())
    oa = cl.FusedAdam(a + [skipped_a], lr=0.02, betas=(0.9, 0.99), weight_decay=1e-8)
    ob = torch.optim.Adam(b + [skipped_b], lr=0.02, betas=(0.9, 0.99), weight_decay=1e-8)
    for step in range(3):
        for x, y in zip(a, b):
            g = torch.randn(x.shape, generator=gen).cuda()
            x.grad, y.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    for x, y in zip(a, b):
        assert gpu.rel_err(x.data, y.data) < 2e-6
    assert torch.equal(skipped_a.data, skipped_b.data)


def test_fused_adam_many_groups_and_mixed_step_histories():
    """More (group, step-history) segments than one launch holds (8): 11 param groups with their own learning rates and decay,
    one of them with a parameter that joins two steps late (its own step count), against torch.optim.Adam."""
    gen = torch.Generator().manual_seed(11)
    base = [torch.randn(3 + 7 * i, generator=gen) for i in range(12)]
    a = [torch.nn.Parameter(b.clone().cuda()) for b in base]
    b = [torch.nn.Parameter(b.clone().cuda()) for b in base]
    groups = lambda ps: [{"params": [ps[i]] if i else [ps[0], ps[11]], "lr": 0.01 * (i + 1), "weight_decay": 1e-3 * (i % 3)}
                         for i in range(11)]
    oa = cl.FusedAdam(groups(a), betas=(0.9, 0.99))
    ob = torch.optim.Adam(groups(b), betas=(0.9, 0.99))
    launches = L.load().clift_launch_count()
    for step in range(4):
        for i, (x, y) in enumerate(zip(a, b)):
            if i == 11 and step < 2:
                continue            # no gradient yet: its step count starts later than its group-mate's
            g = torch.randn(x.shape, generator=gen).cuda()
            x.grad, y.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    # steps 0-1: 11 segments -> 2 launches; steps 2-3: 12 segments (group 0 splits by step history) -> 2 launches
    assert L.load().clift_launch_count() - launches == 8
    for x, y in zip(a, b):
        assert gpu.rel_err(x.data, y.data) < 2e-6
    assert int(oa.state[a[11]]["step"]) == 2 and int(oa.state[a[0]]["step"]) == 4


@pytest.mark.parametrize("n,k,d", [(1000, 7, 3), (1, 1, 3), (4099, 40, 3), (50000, 300, 6), (0, 3, 3)])
def test_nearest_centroid_matches_cdist_argmin(n, k, d):
    gen = torch.Generator().manual_seed(n + k)
    feats = torch.randn(n, d, generator=gen)
    cents = torch.randn(k, d, generator=gen)
    labels, dist = cl.nearest_centroid(feats.cuda(), cents, return_distance=True)
    assert labels.shape == (n,)
    if n == 0:
        return
    dmat = torch.cdist(feats.double(), cents.double())
    ref = dmat.argmin(-1)
    got = labels.cpu().long()
    chosen = dmat.gather(1, got[:, None])[:, 0]
    assert torch.all(chosen <= dmat.min(-1).values + 1e-6)             # same minimum (ties / fp32 near-ties allowed)
    assert float((got != ref).float().mean()) < 1e-3
    assert torch.allclose(dist.cpu().double(), chosen, rtol=1e-5, atol=1e-6)


def test_assign_clusters_matches_reference_bookkeeping():
    """render_panopli.py:371-419 replayed with numpy/torch on the CPU vs cl.assign_clusters."""
    gen = torch.Generator().manual_seed(11)
    n_img, n_pix, d, C = 2, 300, 3, 5
    N = n_img * n_pix
    sem = [torch.randn(n_pix, C, generator=gen) for _ in range(n_img)]
    feats = torch.randn(N, d, generator=gen).numpy()
    pad = np.full((N, d + 1), -np.inf, dtype=np.float32)
    pad[:, 1:] = feats
    sem_arg = torch.cat(sem).argmax(-1).numpy()
    stuff = ~np.isin(sem_arg, [2, 3, 4])
    pad[stuff, 0] = np.inf
    cents = {c: torch.randn(4 + c, d, generator=gen).numpy() for c in (2, 3, 4)}
    got = cl.assign_clusters(pad, sem, cents, "cuda", num_images=n_img).cpu()
    # reference bookkeeping with the oracle's argmin
    labels = np.zeros(N, dtype=np.int64)
    thing = ~stuff
    max_label = 0
    tl = np.zeros(int(thing.sum()), dtype=np.int64)
    ts = sem_arg[thing]
    for c in np.unique(ts):
        m = ts == c
        lab = orc.nearest_centroid(torch.from_numpy(feats[thing][m]), torch.from_numpy(cents[c])).numpy() + max_label
        max_label = lab.max() + 1
        tl[m] = lab
    labels[thing] = tl
    labels[~thing] = -1
    labels += 1
    assert got.shape == (n_img, n_pix, labels.max() + 1)
    assert float((got.argmax(-1).view(-1).numpy() != labels).mean()) < 5e-3


def test_grid_heads_tv_shrink_upsample_golden():
    """Grid-mode heads (allgrid.yaml family): TV over semantic / instance planes AND lines, shrink and upsample of all
    four factor sets, against the reference-generated fixture."""
    fx = gu.load("grid_epoch")
    grid = tuple(int(v) for v in fx["grid"])
    params = syn.make_field_params(int(fx["seed"]), grid, 4, 3, sem_grid_comps=32, ins_grid_comps=32)
    model, rend = gpu.build(params, grid, 4, 3, True, True, syn.default_aabb(), 0.5, sem_grid=32, ins_grid=32)

    class Cfg:
        late_semantic_optimization, instance_optimization_epoch = 1, 4
        lambda_tv_density, lambda_tv_appearance, lambda_tv_semantics, lambda_tv_instances = 0.1, 0.01, 0.02, 0.02

    early = model.total_tv_loss(None, Cfg, 2)
    assert abs(float(early.detach()) - float(fx["tv_early"])) < 1e-5 * abs(float(fx["tv_early"]))
    tot = model.total_tv_loss(None, Cfg, 5)
    assert abs(float(tot) - float(fx["tv_total"])) < 1e-5 * abs(float(fx["tv_total"]))
    tot.backward()
    named = dict(model.named_parameters())
    for k in ("semantic_plane.1", "semantic_line.0", "instance_plane.2", "instance_line.1"):
        assert gpu.rel_err(named[k].grad, tn(fx[f"tv_grad/{k}"])) < 1e-5, k
    t_l, b_r, res = tn(fx["t_l"]), tn(fx["b_r"]), [int(v) for v in fx["res"]]
    model.shrink(t_l, b_r)
    shr = orc.shrink_params(params, t_l, b_r)
    sd = model.state_dict()
    for k in ("semantic_plane.2", "semantic_line.1", "instance_plane.0", "instance_line.2"):
        assert torch.equal(sd[k].cpu(), shr[k]), k
    model.upsample_volume_grid(res)
    ups = orc.upsample_params(shr, res)
    sd = model.state_dict()
    worst = max(gpu.rel_err(sd[k], ups[k]) for k in ups if "plane" in k or "line" in k)
    assert worst < 1e-6, worst
    # the resized grid-mode model renders
    rend.update_step_size(tuple(res))
    with torch.no_grad():
        out = rend(model, syn.random_rays(3, 64).cuda(), 1.0, False, False)
    assert all(torch.isfinite(o).all() for o in out[:4])


def test_upsample_of_a_cpu_resident_model_is_staged_through_the_gpu():
    """Checkpoint resume: on_load_checkpoint (trainer:461-469) calls upsample_volume_grid with the CPU LongTensor
    renderer.grid_dim while Lightning still holds the module on the CPU.  Same kernel, parameters stay on the CPU."""
    grid = (8, 8, 8)
    params = syn.make_field_params(0, grid, 4, 3)
    on_gpu, _ = gpu.build(params, grid, 4, 3, True, True, syn.default_aabb(), 0.5)
    on_cpu, _ = gpu.build(params, grid, 4, 3, True, True, syn.default_aabb(), 0.5, device="cpu")
    on_gpu.upsample_volume_grid((12, 10, 14))
    on_cpu.upsample_volume_grid(torch.tensor([12, 10, 14]))
    assert list(on_cpu.grid_dim()) == [12, 10, 14]
    a, b = on_gpu.state_dict(), on_cpu.state_dict()
    for k in a:
        assert not b[k].is_cuda and a[k].shape == b[k].shape, k
        assert torch.equal(a[k].cpu(), b[k]), k


def test_assign_clusters_missing_centroids_raise_keyerror_and_empty_input():
    """A thing class with points but no centroid table is the reference's KeyError (render_panopli.py:389); classes without
    points consume no labels; zero points give an empty one-hot with only the stuff column."""
    d, C = 3, 4
    sem = [torch.eye(C)[[1, 1, 2, 3]]]
    pad = np.full((4, d + 1), -np.inf, dtype=np.float32)
    pad[:, 1:] = np.arange(12, dtype=np.float32).reshape(4, 3)
    pad[3, 0] = np.inf                                            # last point is stuff
    cents = {1: np.array([[0, 1, 2], [3, 4, 5]], np.float32), 2: np.array([[100, 100, 100], [6, 7, 8]], np.float32),
             0: np.zeros((5, 3), np.float32)}                     # class 0 has centroids but no points: no labels consumed
    got = cl.assign_clusters(pad, sem, cents, "cuda", num_images=1).cpu()
    # class 1: points 0,1 -> centroids 0,1 -> labels 1,2; class 2: point 2 -> centroid 1 -> label 2 + 1 + 1 = 4 (range length 2)
    assert got.dtype == torch.float64 and got.shape == (1, 4, 5)
    assert got[0].argmax(-1).tolist() == [1, 2, 4, 0]
    with pytest.raises(KeyError):
        cl.assign_clusters(pad, sem, {1: cents[1]}, "cuda", num_images=1)
    empty = cl.assign_clusters(np.zeros((0, d + 1), np.float32), [torch.zeros(0, C)], cents, "cuda", num_images=None)
    assert empty.shape == (0, 1)


def test_fused_adam_step_invalidates_cached_packed_parameters():
    """render(no_grad) -> FusedAdam.step -> render(no_grad): the optimizer rewrites parameters through raw pointers (no
    autograd version bump), so it must bump the packed-parameter epoch or the second render reuses stale packed weights."""
    grid = (16, 16, 16)
    params = syn.make_field_params(2, grid, 4, 3)
    aabb = syn.default_aabb()
    model, rend = gpu.build(params, grid, 4, 3, True, True, aabb, 0.5)
    rays = syn.random_rays(4, 64).cuda()
    with torch.no_grad():
        before = rend(model, rays, 0.0, False, False)[0].clone()
    opt = cl.FusedAdam(model.get_optimizable_parameters(0.05, 0.05), betas=(0.9, 0.99))
    for p in model.parameters():
        p.grad = torch.ones_like(p)
    e0 = L.param_epoch()
    opt.step()
    assert L.param_epoch() > e0
    with torch.no_grad():
        after = rend(model, rays, 0.0, False, False)[0]
        model.invalidate_packed()
        fresh = rend(model, rays, 0.0, False, False)[0]
    assert not torch.equal(before, after)
    assert torch.equal(after, fresh)


def test_calls_target_the_tensors_device_not_the_current_one():
    """A model on cuda:1 while cuda:0 is current must launch on cuda:1 (every call site runs under lib.on(device))."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    grid = (16, 16, 16)
    params = syn.make_field_params(2, grid, 4, 3)
    aabb = syn.default_aabb()
    m0, r0 = gpu.build(params, grid, 4, 3, True, True, aabb, 0.5, device="cuda:0")
    m1, r1 = gpu.build(params, grid, 4, 3, True, True, aabb, 0.5, device="cuda:1")
    rays = syn.random_rays(4, 64)
    torch.cuda.set_device(0)
    with torch.no_grad():
        a = r0(m0, rays.to("cuda:0"), 0.0, False, False)
        b = r1(m1, rays.to("cuda:1"), 0.0, False, False)
    for x, y in zip(a[:4], b[:4]):
        assert y.device == torch.device("cuda:1") and torch.equal(x.cpu(), y.cpu())
    f = torch.randn(64, 6, device="cuda:1", requires_grad=True)
    loss = cl.slow_fast_loss(f, torch.randint(1, 4, (64,), device="cuda:1"), torch.rand(64, device="cuda:1"))
    loss.backward()
    assert f.grad.device == torch.device("cuda:1") and torch.isfinite(loss)
