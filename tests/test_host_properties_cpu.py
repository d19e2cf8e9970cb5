"""Property tests (hypothesis) of the host-side geometry the kernels are configured with: TensoRFRenderer's step bookkeeping
(renderer:59-78) against the oracle restatement, and the sample-count forcing used by bench.py / the sweep."""
import torch
from hypothesis import given, settings, strategies as st

import contrastive_lift_b200 as cl
from contrastive_lift_b200 import synthetic as syn
from oracle import clift_oracle as orc

grids = st.tuples(st.integers(8, 200), st.integers(8, 200), st.integers(8, 200))
boxes = st.tuples(st.floats(0.4, 1.5), st.floats(0.4, 1.5), st.floats(0.4, 1.5), st.floats(0.4, 1.5), st.floats(0.4, 1.5),
                  st.floats(0.4, 1.5))


def _aabb(b):
    return torch.tensor([[-b[0], -b[1], -b[2]], [b[3], b[4], b[5]]], dtype=torch.float32)


@settings(max_examples=60, deadline=None)
@given(grid=grids, box=boxes, ratio=st.floats(0.15, 0.95))
def test_step_geometry_matches_reference_restatement(grid, box, ratio):
    aabb = _aabb(box)
    rend = cl.TensoRFRenderer(aabb, list(grid), step_ratio=0.5)
    rend.update_step_ratio(ratio)
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=grid, step_ratio=ratio).refresh()
    assert rend.n_samples == cfg.n_samples
    assert torch.equal(rend.step_size, cfg.step_size) and torch.equal(rend.units, cfg.units)
    assert torch.equal(rend.inv_box_extent, cfg.inv_extent)
    # update_step_size after a resize (trainer:452-456) follows the same arithmetic
    new_grid = tuple(max(4, g // 2 + 3) for g in grid)
    rend.update_step_size(new_grid)
    cfg2 = orc.RenderConfig(aabb=aabb, grid_dim=new_grid, step_ratio=ratio).refresh()
    assert rend.n_samples == cfg2.n_samples and torch.equal(rend.step_size, cfg2.step_size)


@settings(max_examples=60, deadline=None)
@given(grid=grids, box=boxes, n_samples=st.integers(16, 1400))
def test_ratio_for_samples_forces_the_requested_count(grid, box, n_samples):
    aabb = _aabb(box)
    ratio = syn.ratio_for_samples(aabb, grid, n_samples)
    rend = cl.TensoRFRenderer(aabb, list(grid))
    rend.update_step_ratio(ratio)
    assert rend.n_samples == n_samples
    assert ratio == orc.ratio_for_samples(aabb, grid, n_samples)


@settings(max_examples=40, deadline=None)
@given(box=boxes, n_voxels=st.integers(1000, 8_000_000))
def test_target_resolution_matches_reference_restatement(box, n_voxels):
    aabb = _aabb(box)
    rend = cl.TensoRFRenderer(aabb, [16, 16, 16])
    assert tuple(rend.get_target_resolution(n_voxels)) == orc.target_resolution(aabb, n_voxels)


from oracle import refload  # noqa: E402
import pytest  # noqa: E402


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
@settings(max_examples=25, deadline=None)
@given(grid=grids, box=boxes, ratio=st.floats(0.15, 0.95))
def test_step_geometry_matches_the_live_reference(grid, box, ratio):
    """The same bookkeeping against the imported reference class itself (build container only)."""
    aabb = _aabb(box)
    ref = refload.build_renderer(aabb, grid, step_ratio=0.5)
    ref.update_step_ratio(ratio)
    rend = cl.TensoRFRenderer(aabb, list(grid), step_ratio=0.5)
    rend.update_step_ratio(ratio)
    assert rend.n_samples == ref.n_samples and torch.equal(rend.step_size, ref.step_size)
    assert torch.equal(rend.units, ref.units) and torch.equal(rend.inv_box_extent, ref.inv_box_extent)
    assert tuple(rend.get_target_resolution(262144)) == tuple(ref.get_target_resolution(262144))


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
@settings(max_examples=12, deadline=None)
@given(seed=st.integers(0, 10_000), n_samples=st.integers(9, 70), n_cls=st.integers(1, 7), slow_fast=st.booleans(),
       softmax=st.booleans(), empty_scene=st.booleans())
def test_oracle_forward_is_bit_equal_to_the_live_reference(seed, n_samples, n_cls, slow_fast, softmax, empty_scene):
    """Random small scenes and ray sets - axis-aligned directions, rays that miss the box, sample counts that are not
    multiples of 32, scenes without a single active sample - through the restated forward and the reference's own."""
    grid = (9 + seed % 5, 8 + seed % 7, 10 + seed % 3)
    params = syn.make_field_params(seed, grid, n_cls, 2, slow_fast=slow_fast, ball=None if empty_scene else 0.4)
    aabb = syn.default_aabb()
    ratio = syn.ratio_for_samples(aabb, grid, n_samples)
    model = refload.build_model(params, grid, n_cls, 2, slow_fast, softmax)
    rend = refload.build_renderer(aabb, grid, softmax)
    rend.update_step_ratio(ratio)
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=grid, step_ratio=ratio, semantic_softmax=softmax, slow_fast=slow_fast).refresh()
    assert rend.n_samples == n_samples == cfg.n_samples
    rays = syn.random_rays(seed + 1, 24)
    rays[:3, 0:3] = torch.tensor([0.0, 0.97, 0.0])      # skim outside the +y face: every sample out of the box
    rays[:3, 3:6] = torch.tensor([1.0, 0.0, 0.0])
    rays[:3, 7] = orc.sphere_far(rays[:3, 0:3], rays[:3, 3:6])
    with torch.no_grad():
        ref = rend(model, rays, 1.0, False, False)
        got, det = orc.render_forward(params, cfg, rays, None, False, detail=True)
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
    if empty_scene and n_samples >= 56:      # sigma = softplus(-10 +- 0.1): alpha = sigma * delta * 25 stays below the 1e-4 threshold
        assert int(det["active"].sum()) == 0


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 90), d=st.integers(1, 4), n_labels=st.integers(1, 9), seed=st.integers(0, 10_000))
def test_slow_fast_loss_oracle_matches_the_live_reference(n, d, n_labels, seed):
    """trainer:256-310 on random batches: odd N, N = 1, a single label, labels missing from one half."""
    gen = torch.Generator().manual_seed(seed)
    params = syn.make_field_params(3, (8, 8, 8), 4, 3)
    model = refload.build_model(params, (8, 8, 8), 4, 3)
    model.dim_feature_instance = 2 * d
    feats = torch.randn(n, 2 * d, generator=gen).requires_grad_(True)
    labels = torch.randint(1, n_labels + 1, (n,), generator=gen)
    conf = torch.rand(n, generator=gen)
    l_ref = refload.slow_fast_loss(model, labels, feats, conf, use_ema=False)
    f2 = feats.detach().clone().requires_grad_(True)
    l_orc = orc.slow_fast_loss(f2, labels, conf)
    assert torch.equal(l_ref.detach(), l_orc.detach()) or (torch.isnan(l_ref) and torch.isnan(l_orc))
    if l_ref.requires_grad and not torch.isnan(l_ref):
        l_ref.backward()
        l_orc.backward()
        assert torch.allclose(feats.grad, f2.grad, rtol=1e-6, atol=1e-10)


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
@settings(max_examples=25, deadline=None)
@given(n=st.integers(2, 70), d=st.integers(1, 6), n_labels=st.integers(1, 8), temp=st.sampled_from([1.0, 10.0, 100.0]),
       seed=st.integers(0, 10_000))
def test_contrastive_loss_oracle_matches_the_live_reference(n, d, n_labels, temp, seed):
    """model/loss/loss.py:62-82 on random batches."""
    ref = refload.load()
    gen = torch.Generator().manual_seed(seed)
    feats = torch.randn(n, d, generator=gen).requires_grad_(True)
    labels = torch.randint(1, n_labels + 1, (n,), generator=gen)
    l_ref = ref.loss.contrastive_loss(feats, labels, temp)
    f2 = feats.detach().clone().requires_grad_(True)
    l_orc = orc.contrastive_loss(f2, labels, temp)
    assert torch.equal(l_ref.detach(), l_orc.detach()) or (torch.isnan(l_ref) and torch.isnan(l_orc))
    if not torch.isnan(l_ref):
        l_ref.backward()
        l_orc.backward()
        assert torch.allclose(feats.grad, f2.grad, rtol=1e-6, atol=1e-10)


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
@settings(max_examples=8, deadline=None)
@given(seed=st.integers(0, 10_000), n_samples=st.integers(12, 60), n_cls=st.integers(2, 6), slow_fast=st.booleans(),
       softmax=st.booleans(), white_bg=st.booleans())
def test_oracle_training_pass_matches_the_live_reference(seed, n_samples, n_cls, slow_fast, softmax, white_bg):
    """Training-mode forward (the reference draws jitter and the background coin from the CPU generator, renderer:807-810,164)
    and the gradients of a loss over every output, on random small scenes: the restatement used to check the CUDA
    backward must agree with the reference's autograd on every parameter."""
    grid = (8 + seed % 4, 9 + seed % 5, 8 + seed % 3)
    params = syn.make_field_params(seed, grid, n_cls, 2, slow_fast=slow_fast, ball=0.45)
    aabb = syn.default_aabb()
    ratio = syn.ratio_for_samples(aabb, grid, n_samples)
    model = refload.build_model(params, grid, n_cls, 2, slow_fast, softmax)
    rend = refload.build_renderer(aabb, grid, softmax)
    rend.update_step_ratio(ratio)
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=grid, step_ratio=ratio, semantic_softmax=softmax, slow_fast=slow_fast).refresh()
    rays = syn.random_rays(seed + 3, 20)
    gen = torch.Generator().manual_seed(seed)
    w_rgb, w_sem = torch.rand(20, 3, generator=gen), torch.rand(20, n_cls, generator=gen)
    w_ins = torch.rand(20, 4 if slow_fast else 2, generator=gen)

    def loss_of(out):
        return (out[0] * w_rgb).sum() + 0.3 * (out[1] * w_sem).mean() + 0.2 * (out[2] * w_ins).sum() + 0.37 * out[5]

    # the reference consumes the generator itself; replay the same draws for the restatement
    torch.manual_seed(seed + 7)
    ref = rend(model, rays, 1.0, white_bg, True)
    torch.manual_seed(seed + 7)
    jitter = 1.0 * torch.rand((rays.shape[0], 1))
    add_bg = bool(white_bg or bool(torch.rand((1,)) < 0.5))
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    got = orc.render_forward(p, cfg, rays, jitter, add_bg)
    for a, b in zip(ref[:4], got[:4]):
        assert torch.equal(a, b)
    assert torch.equal(ref[5], got[5])
    loss_of(ref).backward()
    loss_of(got).backward()
    for k, v in model.named_parameters():
        g_ref = v.grad if v.grad is not None else torch.zeros_like(v)
        g_got = p[k].grad if p[k].grad is not None else torch.zeros_like(v)
        scale = float(g_ref.abs().max())
        assert torch.allclose(g_got, g_ref, rtol=1e-4, atol=1e-6 * max(scale, 1e-12) + 1e-12), k


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
@settings(max_examples=8, deadline=None)
@given(seed=st.integers(0, 10_000), n_samples=st.integers(12, 60), n_cls=st.integers(2, 6), slow_fast=st.booleans(),
       softmax=st.booleans(), train=st.booleans())
def test_oracle_instance_and_segment_passes_match_the_live_reference(seed, n_samples, n_cls, slow_fast, softmax, train):
    """forward_instance_feature / forward_segment_feature (renderer:178-217, 259-300): maps bit-equal, gradients reach only
    the instance (resp. semantic) head - density is evaluated under no_grad there."""
    grid = (8 + seed % 4, 9 + seed % 5, 8 + seed % 3)
    params = syn.make_field_params(seed, grid, n_cls, 2, slow_fast=slow_fast, ball=0.45)
    aabb = syn.default_aabb()
    ratio = syn.ratio_for_samples(aabb, grid, n_samples)
    model = refload.build_model(params, grid, n_cls, 2, slow_fast, softmax)
    rend = refload.build_renderer(aabb, grid, softmax)
    rend.update_step_ratio(ratio)
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=grid, step_ratio=ratio, semantic_softmax=softmax, slow_fast=slow_fast).refresh()
    rays = syn.random_rays(seed + 5, 18)
    torch.manual_seed(seed + 11)
    ref_ins, ref_pts = rend.forward_instance_feature(model, rays, 1.0, train)
    ref_seg = rend.forward_segment_feature(model, rays, 1.0, train)
    torch.manual_seed(seed + 11)
    j1 = torch.rand((rays.shape[0], 1)) if train else None
    j2 = torch.rand((rays.shape[0], 1)) if train else None
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ins, pts = orc.render_instance_feature(p, cfg, rays, j1)
    seg = orc.render_segment_feature(p, cfg, rays, j2)
    assert torch.equal(ins, ref_ins) and torch.equal(pts, ref_pts) and torch.equal(seg, ref_seg)
    (ref_ins.sum() + 0.5 * ref_seg.mean()).backward()
    (ins.sum() + 0.5 * seg.mean()).backward()
    for k, v in model.named_parameters():
        head = k.startswith(("render_instance_mlp", "render_semantic_mlp"))
        if not head:
            assert v.grad is None and p[k].grad is None, k
            continue
        g_ref = v.grad if v.grad is not None else torch.zeros_like(v)
        g_got = p[k].grad if p[k].grad is not None else torch.zeros_like(v)
        assert torch.allclose(g_got, g_ref, rtol=1e-4, atol=1e-6 * max(float(g_ref.abs().max()), 1e-12) + 1e-12), k
