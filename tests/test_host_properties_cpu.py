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
