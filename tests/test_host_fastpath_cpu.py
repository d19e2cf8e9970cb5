"""Host-side fast paths of the training step (no device work): the parameter walk the render entries use, the recorded /
replayed packing plan, and the deferred active-sample-capacity check."""
import collections

import pytest
import torch

import contrastive_lift_b200 as cl
from contrastive_lift_b200 import field as F
from contrastive_lift_b200 import lib as L


@pytest.mark.parametrize("kw", [dict(use_semantic_mlp=True, use_instance_mlp=True, slow_fast_mode=True),
                                dict(num_semantics_comps=(32, 32, 32), num_instance_comps=(32, 32, 32), use_semantic_mlp=False,
                                     use_instance_mlp=False, slow_fast_mode=False)], ids=["mlp_heads", "grid_heads"])
def test_param_list_is_named_parameters_order(kw):
    model = cl.TensorVMSplit([8, 8, 8], num_semantic_classes=5, dim_feature_instance=6, **kw)
    ref = [p for _, p in model.named_parameters()]
    got = F.param_list(model)
    assert len(got) == len(ref) and all(a is b for a, b in zip(got, ref))
    # a module registered twice and a parameter shared by two modules are listed once, where named_parameters lists them
    lin = torch.nn.Linear(2, 2)
    twin = torch.nn.Linear(2, 2)
    twin.weight = lin.weight
    net = torch.nn.Sequential(lin, torch.nn.ReLU(), lin, twin)
    ref = [p for _, p in net.named_parameters()]
    got = F.param_list(net)
    assert len(got) == len(ref) == 3 and all(a is b for a, b in zip(got, ref))


def test_pack_plan_records_stream_ordered_calls_and_replays_them_on_the_current_stream():
    log = []

    class FakeLib:
        def clift_pack_batch(self, table, n, tiles, stream):
            log.append(("pack_batch", table, n, tiles, stream))
            return 0

        def clift_tc16_weight_bytes(self, n_out, n_in, bias):       # a size query: passes through, never replayed
            log.append(("bytes",))
            return 64

        def clift_pack_linear_tc16(self, w, b, dst, n_out, n_in, bound, floor, stream):
            log.append(("tc16", w, dst, stream))
            return 0

    plan = F._PackPlan(FakeLib(), key=(1, 2))
    assert plan.clift_pack_batch(100, 3, 7, 0xAA) == 0 and plan.clift_tc16_weight_bytes(4, 4, 1) == 64
    assert plan.clift_pack_linear_tc16(5, None, 6, 4, 4, None, 1.0, 0xAA) == 0
    assert len(plan.calls) == 2
    del log[:]
    plan.replay(0xBB)
    assert log == [("pack_batch", 100, 3, 7, 0xBB), ("tc16", 5, 6, 0xBB)]


class _Event:
    def __init__(self, done):
        self.done, self.waited = done, False

    def query(self):
        return self.done

    def synchronize(self):
        self.waited, self.done = True, True


def _renderer():
    return cl.TensoRFRenderer([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], [8, 8, 8], semantic_weight_mode="softmax")


def test_capacity_follows_a_decayed_maximum_per_head_set():
    r = _renderer()
    static = r.max_active(100, True, 3)
    assert static == 100 * min(192, r.n_samples)
    r._note_active(3, 1000, 100)                      # 10 active samples per ray with head set 3
    assert r.max_active(100, True, 3) == 100 * (int(10 * 1.25) + 8)
    r._note_active(3, 200, 100)                       # a sparser batch: the history decays by 10 % per call, not at once
    assert r._active_hist[3] == pytest.approx(9.0)
    r._note_active(4, 3000, 100)                      # another head set (the instance pass) keeps its own history
    assert r._active_hist[4] == pytest.approx(30.0) and r._active_hist[3] == pytest.approx(9.0)
    assert r.max_active(100, False, 3) == static      # inference renders use the static bound


def test_deferred_overflow_is_reported_by_a_later_call():
    r = _renderer()
    pend = lambda done, n_act, overflow, rays=100, cap=1500, heads=3: \
        (_Event(done), torch.tensor([n_act, 5000, overflow, 12], dtype=torch.int64), rays, cap, heads)
    r._pending = collections.deque([pend(True, 1000, 0), pend(False, 1200, 0)])
    r._poll_overflow()                                # reads only what has landed, never waits
    assert len(r._pending) == 1 and r._active_hist[3] == pytest.approx(10.0) and len(r._pinned) == 1
    r.synchronize_overflow_checks()                   # waits for the rest
    assert not r._pending and r._active_hist[3] == pytest.approx(12.0)
    r._pending = collections.deque([pend(True, 4000, 1), pend(True, 900, 0)])
    with pytest.raises(L.CliftError, match="4000 active samples but room for 1500"):
        r._poll_overflow()
    assert r._active_hist[3] == pytest.approx(40.0) and not r._pending      # the next render is sized for it
    r.overflow_policy = "warn"
    r._pending = collections.deque([pend(True, 5000, 1)])
    with pytest.warns(RuntimeWarning, match="5000 active samples"):
        r._poll_overflow()
    assert r._active_hist[3] == pytest.approx(50.0)


def test_a_copied_renderer_starts_without_outstanding_checks():
    import copy
    import pickle
    r = _renderer()
    r._note_active(3, 1000, 100)
    r._pending = collections.deque([(_Event(False), torch.zeros(4, dtype=torch.int64), 100, 1500, 3)])
    r._pinned = [torch.zeros(4, dtype=torch.int64)]
    for twin in (copy.deepcopy(r), pickle.loads(pickle.dumps(r))):
        assert not twin._pending and not twin._pinned and twin._active_hist == {3: 10.0}
        assert torch.equal(twin.bbox_aabb, r.bbox_aabb) and twin.n_samples == r.n_samples
    assert len(r._pending) == 1          # the original keeps its own


def test_unpack_plan_lays_gradients_out_in_one_flat_buffer():
    """_UnpackPlan turns the planning run's jobs (absolute destination pointers of ~45 separate tensors) into offsets of one
    flat buffer: 256-byte aligned slots in gradient order, every job mapped to the slot of the tensor it wrote."""
    from contrastive_lift_b200 import renderer as R
    grads = {"a.weight": torch.zeros(5, 7), "a.bias": torch.zeros(5), "plane.0": torch.zeros(1, 16, 9, 9)}

    class Batch:
        jobs, tiles = [], 11
    for g in (grads["plane.0"], grads["a.weight"], grads["a.bias"]):       # job order differs from gradient order
        j = L.PackJob()
        j.src, j.dst, j.d_rows, j.d_cols = 0x1000, g.data_ptr(), 3, 4
        Batch.jobs.append(j)
    plan = R._UnpackPlan(Batch, grads)
    assert [(n, s, k, o) for n, s, k, o in plan.items] == [("a.weight", (5, 7), 35, 0), ("a.bias", (5,), 5, 64),
                                                            ("plane.0", (1, 16, 9, 9), 1296, 128)]
    assert plan.total == 128 + 1344 and plan.tiles == 11                  # 1296 floats round up to 21 slots of 64
    assert [off for _, off in plan.jobs] == [128 * 4, 0, 64 * 4]
    assert all(j.src == 0x1000 and j.d_rows == 3 for j, _ in plan.jobs)
