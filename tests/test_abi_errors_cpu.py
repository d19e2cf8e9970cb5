"""C-ABI error behaviour (include/clift_b200.h): bad arguments are refused with a status + clift_last_error() BEFORE any
CUDA call, so these run without a GPU.  The reference only asserts (SURVEY 8b "Errors"); the C boundary never throws."""
import ctypes as C

import pytest

from contrastive_lift_b200 import lib as L

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4
FAKE = 0x1000          # non-null placeholder address: validation never dereferences device pointers


def cfg(n_samples=64, step=0.01, heads=L.HEAD_ALL):
    c = L.RenderCfg()
    L.fill3(c.aabb_min, [-1.0] * 3)
    L.fill3(c.aabb_max, [1.0] * 3)
    L.fill3(c.inv_extent, [1.0] * 3)
    c.step_size, c.n_samples, c.distance_scale, c.weight_thres = step, n_samples, 25.0, 1e-4
    c.semantic_softmax, c.heads, c.head_path = 1, heads, L.HEADS_AUTO
    return c


def mlp(m, dims):
    m.n_layers = len(dims) - 1
    for i, d in enumerate(dims):
        m.dims[i] = d
    for i in range(len(dims) - 1):
        m.wt[i] = m.bias[i] = FAKE


def field(grid=(16, 16, 16), n_cls=21, d_ins=3, density_comps=16):
    f = L.Field()
    L.fill3(f.grid, grid)
    f.density_comps, f.appearance_comps, f.dim_appearance = density_comps, 48, 27
    f.pe_view, f.pe_feat, f.pe_sem, f.pe_ins = 2, 2, 0, 0
    f.num_classes, f.dim_instance, f.slow_fast, f.density_shift = n_cls, d_ins, 1, -10.0
    for i in range(3):
        f.density_plane[i] = f.density_line[i] = f.appearance_plane[i] = f.appearance_line[i] = FAKE
    f.basis = FAKE
    mlp(f.rgb, [150, 128, 128, 3])
    mlp(f.semantic, [3, 256, 256, 256, 256, n_cls])
    mlp(f.instance_fast, [3, 256, 256, 256, d_ins])
    mlp(f.instance_slow, [3, 256, 256, 256, d_ins])
    return f


def out():
    o = L.RenderOut()
    for name, _ in L.RenderOut._fields_:
        if name != "save_for_backward":
            setattr(o, name, FAKE)
    return o


def forward(lib, c, f, n_rays=8, rays=FAKE, ws=FAKE, ws_bytes=1 << 40, o=None):
    o = o if o is not None else out()
    return lib.clift_render_forward(C.byref(c) if c is not None else None, C.byref(f) if f is not None else None, rays, None, n_rays,
                                    0, ws, ws_bytes, 0, C.byref(o), None)


def err(lib):
    return lib.clift_last_error().decode()


def test_null_and_malformed_descriptors_are_refused():
    lib = L.load()
    assert forward(lib, None, field()) == ERR_ARG and "null cfg" in err(lib)
    assert forward(lib, cfg(n_samples=1), field()) == ERR_ARG and "n_samples" in err(lib)
    assert forward(lib, cfg(step=0.0), field()) == ERR_ARG and "step_size" in err(lib)
    assert forward(lib, cfg(), None) == ERR_ARG and "null field" in err(lib)
    assert forward(lib, cfg(), field(grid=(16, 1, 16))) == ERR_ARG and "grid" in err(lib)
    assert forward(lib, cfg(), field(), n_rays=-1) == ERR_ARG
    f = field()
    f.density_line[1] = None
    assert forward(lib, cfg(), f) == ERR_ARG and "density" in err(lib)
    f = field()
    f.semantic.wt[2] = None
    assert forward(lib, cfg(), f) == ERR_ARG and "semantic mlp" in err(lib) and "layer 2" in err(lib)
    f = field()
    f.semantic.dims[5] = 7                         # last layer width != num_classes
    assert forward(lib, cfg(), f) == ERR_ARG and "output width" in err(lib)
    f = field()
    f.instance_fast.dims[0] = 9                    # MLP-mode instance head reads xyz (+PE): 3 + 6*pe_ins inputs
    assert forward(lib, cfg(), f) == ERR_ARG and "instance mlp input width" in err(lib)
    f = field()
    f.rgb.dims[0] = 149                            # 27*(1+2*2) + 3*(1+2*2) = 150
    assert forward(lib, cfg(), f) == ERR_ARG and "rgb mlp input width" in err(lib)
    # a head that is not requested is not validated: the instance pass needs no rgb / semantic description
    f = field()
    f.rgb.n_layers = 0
    f.semantic.n_layers = 0
    assert forward(lib, cfg(heads=L.HEAD_INSTANCE), f, rays=None) == ERR_ARG and "null pointer" in err(lib)


def test_configurations_outside_the_compiled_envelope_are_unsupported_not_crashes():
    lib = L.load()
    assert forward(lib, cfg(), field(density_comps=8)) == ERR_UNSUPPORTED and "density_comps" in err(lib)
    assert forward(lib, cfg(), field(n_cls=65)) == ERR_UNSUPPORTED and "num_classes" in err(lib)
    assert forward(lib, cfg(), field(d_ins=0)) == ERR_UNSUPPORTED and "dim_instance" in err(lib)
    f = field()
    f.pe_view = f.pe_feat = 0
    assert forward(lib, cfg(), f) == ERR_UNSUPPORTED and "view-independent" in err(lib)
    f = field()
    f.semantic.n_layers = 9
    assert forward(lib, cfg(), f) == ERR_UNSUPPORTED and "n_layers" in err(lib)
    f = field()
    f.semantic.dims[2] = 4096
    assert forward(lib, cfg(), f) == ERR_UNSUPPORTED and "width" in err(lib)


def test_call_size_limit_workspace_and_required_outputs():
    lib = L.load()
    c, f = cfg(n_samples=1024), field()
    # one call handles n_rays * n_samples < 2^31 (the Python host splits larger frames into ray ranges)
    assert forward(lib, c, f, n_rays=(1 << 21)) == ERR_ARG and "2^31" in err(lib)
    # workspace: the size query is pure host arithmetic; a smaller buffer is refused with the required size in the message
    need = lib.clift_render_workspace_bytes(C.byref(c), C.byref(f), 4096, 4096 * 160, 0)
    need_train = lib.clift_render_workspace_bytes(C.byref(c), C.byref(f), 4096, 4096 * 160, 1)
    assert 0 < need < need_train                   # the training stash (A + Z, ~25 KB per active-sample slot) is extra
    assert lib.clift_render_workspace_bytes(C.byref(c), C.byref(f), 8192, 8192 * 160, 0) > need
    assert lib.clift_render_workspace_bytes(None, C.byref(f), 8, 0, 0) == ERR_ARG
    assert lib.clift_render_workspace_bytes(C.byref(c), C.byref(f), -1, 0, 0) == ERR_ARG
    o = out()
    assert lib.clift_render_forward(C.byref(c), C.byref(f), FAKE, None, 4096, 0, FAKE, need - 1, 4096 * 160, C.byref(o),
                                    None) == ERR_WORKSPACE
    assert str(need) in err(lib)
    for missing, word in (("depth", "depth"), ("rgb_raw", "rgb"), ("semantic", "semantic"), ("instance", "instance"),
                          ("dist_ray", "dist_ray")):
        o = out()
        setattr(o, missing, None)
        assert forward(lib, c, f, n_rays=64, o=o) == ERR_ARG and word in err(lib), missing


def test_python_host_turns_status_codes_into_exceptions():
    lib = L.load()
    assert forward(lib, None, field()) == ERR_ARG
    with pytest.raises(L.CliftError, match="null cfg"):
        L.check(ERR_ARG)
    L.check(OK)


def test_loss_ray_and_epoch_entries_validate_before_touching_the_device():
    lib = L.load()
    # slow-fast loss (trainer:256-310): null inputs, non-positive width, and batches beyond one cluster's shared memory
    assert lib.clift_slowfast_loss(None, FAKE, FAKE, 64, 6, FAKE, None, None) == ERR_ARG
    assert lib.clift_slowfast_loss(FAKE, FAKE, FAKE, 64, 0, FAKE, None, None) == ERR_ARG
    assert lib.clift_slowfast_loss(FAKE, FAKE, FAKE, -1, 6, FAKE, None, None) == ERR_ARG
    assert lib.clift_slowfast_loss(FAKE, FAKE, FAKE, 1 << 20, 6, FAKE, None, None) == ERR_UNSUPPORTED and "slow-fast" in err(lib)
    # vanilla contrastive loss (loss.py:62-82)
    assert lib.clift_contrastive_loss(FAKE, None, 64, 3, 100.0, FAKE, None, None) == ERR_ARG
    assert lib.clift_contrastive_loss(FAKE, FAKE, 0, 3, 100.0, FAKE, None, None) == ERR_ARG
    assert lib.clift_contrastive_loss(FAKE, FAKE, 1 << 20, 3, 100.0, FAKE, None, None) == ERR_UNSUPPORTED
    # EMA, TV
    assert lib.clift_ema_update(None, FAKE, 16, 0.9, None) == ERR_ARG
    assert lib.clift_ema_update(FAKE, FAKE, -1, 0.9, None) == ERR_ARG
    assert lib.clift_tv_loss(None, 16, 8, 8, FAKE, None, 1.0, None) == ERR_ARG
    assert lib.clift_tv_loss(FAKE, 16, 0, 8, FAKE, None, 1.0, None) == ERR_ARG
    assert lib.clift_ema_update_batch(None, 3, 16, 0.9, None) == ERR_ARG
    assert lib.clift_ema_update_batch(FAKE, -1, 16, 0.9, None) == ERR_ARG
    assert lib.clift_ema_update_batch(None, 0, 0, 0.9, None) == 0                      # nothing to do
    assert lib.clift_tv_loss_batch(None, 2, 64, None) == ERR_ARG
    assert lib.clift_tv_loss_batch(FAKE, 5000, 64, None) == ERR_UNSUPPORTED
    # grouped Adam: the groups must tile the table in order, with step >= 1, at most CLIFT_MAX_ADAM_GROUPS of them
    grp = (L.AdamGroup * 2)()
    grp[0].lr, grp[0].beta1, grp[0].beta2, grp[0].eps, grp[0].first, grp[0].count, grp[0].step = 0.01, 0.9, 0.99, 1e-8, 0, 2, 1
    grp[1].lr, grp[1].beta1, grp[1].beta2, grp[1].eps, grp[1].first, grp[1].count, grp[1].step = 0.01, 0.9, 0.99, 1e-8, 2, 1, 1
    assert lib.clift_adam_step_groups(None, 3, 16, grp, 2, 1.0, None) == ERR_ARG
    assert lib.clift_adam_step_groups(FAKE, 4, 16, grp, 2, 1.0, None) == ERR_ARG and "cover" in err(lib)
    grp[1].first = 1
    assert lib.clift_adam_step_groups(FAKE, 3, 16, grp, 2, 1.0, None) == ERR_ARG and "tile" in err(lib)
    grp[1].first, grp[1].step = 2, 0
    assert lib.clift_adam_step_groups(FAKE, 3, 16, grp, 2, 1.0, None) == ERR_ARG
    assert lib.clift_adam_step_groups(FAKE, 3, 16, grp, L.MAX_ADAM_GROUPS + 1, 1.0, None) == ERR_UNSUPPORTED
    # ray generation (util/ray.py): null camera, empty frame, misaligned output (one ray = two 16-byte stores)
    k = (C.c_float * 9)(*[1.0] * 9)
    pose = (C.c_float * 16)(*[0.0] * 16)
    assert lib.clift_gen_rays(None, pose, 8, 8, 0.01, 1.0, FAKE, FAKE, None) == ERR_ARG
    assert lib.clift_gen_rays(k, pose, 0, 8, 0.01, 1.0, FAKE, FAKE, None) == ERR_ARG
    assert lib.clift_gen_rays(k, pose, 8, 8, 0.01, 1.0, FAKE + 4, FAKE, None) == ERR_ARG and "aligned" in err(lib)
    # nearest-centroid assignment (render_panopli.py:389-396)
    assert lib.clift_assign_centroids(FAKE, 10, 3, 2, FAKE, 4, FAKE, None, None) == ERR_ARG          # stride < dim
    assert lib.clift_assign_centroids(FAKE, 10, 3, 3, None, 4, FAKE, None, None) == ERR_ARG
    assert lib.clift_assign_centroids(FAKE, 10, 32, 32, FAKE, 4, FAKE, None, None) == ERR_UNSUPPORTED
    assert lib.clift_assign_centroids(None, 0, 3, 3, None, 4, None, None, None) == OK                 # empty input: nothing to do
    # plane upsample, layout packing
    assert lib.clift_upsample_bilinear(FAKE, FAKE, 16, 8, 8, 0, 12, None) == ERR_ARG
    assert lib.clift_pack_plane(None, FAKE, 16, 8, 8, None) == ERR_ARG
    assert lib.clift_pack_linear_dgrad(FAKE, FAKE, 4096, 16, None) == ERR_UNSUPPORTED


def test_fused_adam_step_bumps_the_packed_parameter_epoch():
    """FusedAdam rewrites parameters through raw pointers; cached packed copies are keyed on lib.param_epoch()."""
    import torch
    import contrastive_lift_b200 as cl
    p = torch.nn.Parameter(torch.zeros(3))          # no .grad: nothing to launch, the epoch must still move
    opt = cl.FusedAdam([p], lr=0.1)
    e0 = L.param_epoch()
    opt.step()
    assert L.param_epoch() == e0 + 1


def test_allreduce_entry_refuses_a_null_communicator():
    """clift_allreduce_grads (the path's one collective) validates its arguments before touching NCCL."""
    lib = L.load()
    assert lib.clift_allreduce_grads(None, None, 16, None) == -1
    assert b"communicator" in lib.clift_last_error()
    assert lib.clift_allreduce_grads(C.c_void_p(1), None, 0, None) == 0          # empty arena: nothing to do
