"""CPU-side checks of the boundary: the C-ABI library loads here (no GPU) and exports every symbol the header
declares; the Python host mirrors the reference's constructor/state_dict surface; compute calls refuse CPU tensors."""
import os
import re

import pytest
import torch

import contrastive_lift_b200 as cl
from contrastive_lift_b200 import lib as L
from contrastive_lift_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "clift_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clift_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(L.SIGNATURES) == names, "lib.py SIGNATURES must list exactly the header's entry points"
    assert lib.clift_abi_version() == L.ABI_VERSION


def test_struct_sizes_match_header_layout():
    import ctypes as C
    assert C.sizeof(L.Mlp) == 4 + 4 * 9 + 8 * 8 * 7
    assert C.sizeof(L.RenderCfg) == 4 * 9 + 4 * 7
    assert C.sizeof(L.RenderOut) == 8 * 11 + 8 + 8


def test_no_cpu_fallback():
    with pytest.raises(L.CliftError):
        L.ptr(torch.zeros(4))
    params = syn.make_field_params(0, (8, 8, 8), 4, 3)
    model = cl.TensorVMSplit([8, 8, 8], num_semantic_classes=4, dim_feature_instance=6, use_semantic_mlp=True,
                             use_instance_mlp=True, slow_fast_mode=True)
    model.load_state_dict(params)
    rend = cl.TensoRFRenderer(syn.default_aabb(), [8, 8, 8], semantic_weight_mode="softmax")
    with pytest.raises(L.CliftError):
        rend(model, syn.random_rays(0, 8), 1.0, False, False)
    with pytest.raises(L.CliftError):
        cl.slow_fast_loss(torch.zeros(8, 6), torch.zeros(8, dtype=torch.long), torch.zeros(8))


def test_state_dict_keys_and_shapes_match_reference_checkpoint_layout():
    grid = (12, 10, 14)
    params = syn.make_field_params(1, grid, 21, 3)
    model = cl.TensorVMSplit(list(grid), num_semantics_comps=(32, 32, 32), num_instance_comps=(32, 32, 32),
                             num_semantic_classes=21, dim_feature_instance=6, use_semantic_mlp=True,
                             use_instance_mlp=True, slow_fast_mode=True)
    sd = model.state_dict()
    assert sorted(sd) == sorted(params)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(params[k].shape), k
    rend = cl.TensoRFRenderer(syn.default_aabb(), list(grid), semantic_weight_mode="softmax")
    assert sorted(rend.state_dict()) == ["bbox_aabb", "grid_dim", "inv_box_extent", "units"]
    assert rend.n_samples == int((12 ** 0.5) / float(rend.step_size)) + 1


def test_renderer_geometry_matches_oracle():
    from oracle import clift_oracle as orc
    for grid, ratio in (((128, 128, 128), 0.5), ((20, 24, 16), 0.37), ((192, 150, 171), 0.25)):
        aabb = torch.tensor([[-1.0, -0.9, -1.0], [1.0, 0.8, 0.7]])
        rend = cl.TensoRFRenderer(aabb, list(grid), step_ratio=0.5)
        rend.update_step_ratio(ratio)
        cfg = orc.RenderConfig(aabb=aabb, grid_dim=grid, step_ratio=ratio).refresh()
        assert rend.n_samples == cfg.n_samples and torch.equal(rend.step_size, cfg.step_size)
        assert torch.equal(rend.inv_box_extent, cfg.inv_extent) and torch.equal(rend.units, cfg.units)


def test_optimizer_groups_follow_reference():
    model = cl.TensorVMSplit([8, 8, 8], num_semantic_classes=4, dim_feature_instance=6, use_semantic_mlp=True,
                             use_instance_mlp=True, slow_fast_mode=True)
    g = model.get_optimizable_parameters(0.02, 0.001, weight_decay=1e-8)
    assert [x["lr"] for x in g] == [0.02, 0.02, 0.02, 0.02, 0.001, 0.001, 0.001]
    assert g[0]["weight_decay"] == 1e-8 and "weight_decay" not in g[1]
    gi = model.get_optimizable_instance_parameters(0.02, 0.001, using_DINO=True)
    assert len(gi) == 1 and len(list(gi[0]["params"])) == 8
    assert len(model.get_optimizable_instance_parameters(0.02, 0.001, using_DINO=False)) == 2


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(L.CliftError):
        cl.TensorVMSplit([8, 8, 8], num_semantic_classes=4, dim_feature_instance=6, use_semantic_mlp=False)
    with pytest.raises(L.CliftError):
        cl.TensoRFRenderer(syn.default_aabb(), [8, 8, 8], stop_semantic_grad=False)


def test_ctypes_structs_match_the_compiled_header(tmp_path):
    """Every struct of include/clift_b200.h, compiled as plain C by gcc, has the size of its ctypes mirror in lib.py."""
    import ctypes as C
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    pairs = [("clift_mlp", L.Mlp), ("clift_mlp_grad", L.MlpGrad), ("clift_grid_head", L.GridHead),
             ("clift_grid_head_grad", L.GridHeadGrad), ("clift_field", L.Field), ("clift_field_grad", L.FieldGrad),
             ("clift_render_cfg", L.RenderCfg), ("clift_render_out", L.RenderOut), ("clift_adam_tensor", L.AdamTensor),
             ("clift_pack_job", L.PackJob), ("clift_tc16_job", L.Tc16Job)]
    body = "".join(f'printf("%zu\\n", sizeof({c}));' for c, _ in pairs)
    src = tmp_path / "sizes.c"
    src.write_text(f'#include <stdio.h>\n#include "clift_b200.h"\nint main(void) {{ {body} return 0; }}\n')
    exe = tmp_path / "sizes"
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    for (cname, ctype), size in zip(pairs, sizes):
        assert C.sizeof(ctype) == size, (cname, C.sizeof(ctype), size)


def test_batched_pack_job_tables(monkeypatch):
    """Host side of clift_pack_batch / clift_pack_linear_tc16_batch: tile and block prefixes, job shapes (no GPU needed)."""
    monkeypatch.setattr(L, "ptr", lambda t: None if t is None else 0x1000)
    b = L.PackBatch()
    x = torch.zeros(1)
    b.plane(x, x, 48, 20, 24)              # (1,48,20,24) -> [480][48]
    b.linear(x, x, x, x, 21, 256)          # W [21][256] -> W^T [256][64] + bias [64]
    b.dgrad(x, x, 21, 256)                 # -> [32][256]
    b.unlinear(x, x, x, None, 3, 150)      # packed grad [160][64] -> [3][150], no bias
    j = b.jobs
    assert [(q.d_rows, q.d_cols, q.kind) for q in j] == [(480, 48, 0), (256, 64, 0), (1, 64, 1), (32, 256, 1), (3, 150, 0)]
    tiles = [15 * 2, 8 * 2, 1 * 2, 1 * 8, 1 * 5]
    assert [q.first_tile for q in j] == [sum(tiles[:i]) for i in range(5)] and b.tiles == sum(tiles)
    assert (j[1].s_pitch, j[1].s_rows, j[1].s_cols) == (256, 21, 256) and (j[4].s_pitch, j[4].s_rows, j[4].s_cols) == (64, 150, 3)
    t = L.Tc16Batch()
    t.add(x, None, x, 27, 144, 0x2000, 0.0, 0)       # basis: no bias, 9 slabs x 16 x 32
    t.add(x, x, x, 128, 150, 0x3000, 1.0, 0)         # rgb layer 0: 10 + 1 slabs x 16 x 128
    t.add(x, x, x, 256, 3, None, 1.0, 1)             # xyz layer 0 of another chain: 1 + 1 slabs x 16 x 256
    assert [q.first_block for q in t.jobs] == [0, 18, 18 + 88] and t.blocks == 18 + 88 + 32 and t.chains == 2


def test_semantic_weight_mode_and_output_activation_must_agree():
    """One C-ABI flag switches the per-sample Softmax of the semantic MLP (tensoRF.py:594) and the renderer's
    log-normalisation (renderer:160-162); the reference's callers always build them together (trainer:54,67).  A pair that
    disagrees, or the arg-max compositing mode (renderer:142-143), is refused instead of being computed differently."""
    import gpu_util as gpu
    grid = (8, 8, 8)
    for softmax in (True, False):
        for sem_grid, ins_grid in ((None, None), (32, 32)):
            params = syn.make_field_params(0, grid, 4, 3, sem_grid_comps=sem_grid, ins_grid_comps=ins_grid) \
                if sem_grid else syn.make_field_params(0, grid, 4, 3)
            model, rend = gpu.build(params, grid, 4, 3, True, softmax, syn.default_aabb(), 0.5, device="cpu",
                                    sem_grid=sem_grid, ins_grid=ins_grid)
            for heads in (L.HEAD_ALL, L.HEAD_INSTANCE, L.HEAD_SEMANTIC, 0):      # what the GPU tests / bench / smoke build
                assert rend._cfg(model, heads).semantic_softmax == int(softmax)
    model, rend = gpu.build(syn.make_field_params(0, grid, 4, 3), grid, 4, 3, True, True, syn.default_aabb(), 0.5, device="cpu")
    rend.semantic_weight_mode = "none"
    with pytest.raises(L.CliftError, match="output_mlp_semantics"):
        rend._cfg(model, L.HEAD_ALL)
    rend._cfg(model, L.HEAD_INSTANCE)                   # the instance pass never evaluates the semantic head
    rend.semantic_weight_mode = "softmax"
    model.render_semantic_mlp.output_activation = torch.nn.Sigmoid()
    with pytest.raises(L.CliftError, match="Softmax"):
        rend._cfg(model, L.HEAD_SEMANTIC)
    with pytest.raises(L.CliftError, match="argmax"):
        cl.TensoRFRenderer(syn.default_aabb(), list(grid), semantic_weight_mode="argmax")
    rend.semantic_weight_mode = "argmax"                # changed after construction: caught at the next call
    with pytest.raises(L.CliftError, match="argmax"):
        rend._cfg(model, 0)


def test_renderer_descriptor_follows_buffers_replaced_behind_its_back():
    """on_load_checkpoint assigns renderer.bbox_aabb directly (trainer:466); in-place edits and .data swaps happen in
    user code too.  The cached host copy of the geometry must follow all of them."""
    import gpu_util as gpu
    grid = (8, 8, 8)
    model, rend = gpu.build(syn.make_field_params(0, grid, 4, 3), grid, 4, 3, True, True, syn.default_aabb(), 0.5, device="cpu")
    assert list(rend._cfg(model, 0).aabb_max) == [1.0, 1.0, 1.0]
    rend.bbox_aabb = torch.tensor([[-1.0, -1.0, -1.0], [0.5, 0.75, 1.0]])          # buffer replaced
    assert list(rend._cfg(model, 0).aabb_max) == [0.5, 0.75, 1.0]
    rend.bbox_aabb[1, 0] = 0.25                                                     # written in place
    assert list(rend._cfg(model, 0).aabb_max) == [0.25, 0.75, 1.0]
    rend.bbox_aabb.data = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 0.5]])       # storage swapped
    assert list(rend._cfg(model, 0).aabb_max) == [1.0, 1.0, 0.5]
    rend.update_step_size(rend.grid_dim)
    c = rend._cfg(model, 0)
    assert abs(c.inv_extent[2] - 2.0 / 1.5) < 1e-7 and c.step_size == float(rend.step_size) and c.n_samples == rend.n_samples


def test_upsample_of_a_cpu_model_without_a_gpu_fails_loudly():
    """on_load_checkpoint (trainer:461-469) resizes the factors while Lightning still holds the module on the CPU: the
    mirror stages them through the CUDA device; with no device there is no silent CPU resize."""
    if torch.cuda.is_available():
        pytest.skip("covered by the -m gpu test of the staged path")
    import gpu_util as gpu
    grid = (8, 8, 8)
    model, _ = gpu.build(syn.make_field_params(0, grid, 4, 3), grid, 4, 3, True, True, syn.default_aabb(), 0.5, device="cpu")
    with pytest.raises(L.CliftError, match="CUDA device"):
        model.upsample_volume_grid((12, 10, 14))
    assert list(model.grid_dim()) == [8, 8, 8]                  # nothing was replaced


def test_missing_or_stale_library_fails_loudly(monkeypatch, tmp_path):
    """No extension, no product: a missing libclift_b200.so (or one built from another header revision) raises at the first
    compute call - nothing falls back to PyTorch or to the oracle."""
    monkeypatch.setattr(L, "_LIB", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "libclift_b200.so"))
    with pytest.raises(L.CliftError, match="is missing"):
        L.load()
    with pytest.raises(L.CliftError, match="is missing"):
        cl.contrastive_loss(torch.zeros(4, 3), torch.zeros(4, dtype=torch.long), 100.0)
    # a library of another ABI revision is refused too
    monkeypatch.undo()
    real = L.load()
    monkeypatch.setattr(L, "_LIB", None)
    monkeypatch.setattr(L, "ABI_VERSION", real.clift_abi_version() + 1)
    with pytest.raises(L.CliftError, match="rebuild"):
        L.load()
    monkeypatch.undo()
    assert L.load().clift_abi_version() == L.ABI_VERSION


def test_integration_doc_binding_matches_the_library():
    """INTEGRATION.md shows the ctypes stub a maintainer would write; its struct fields, ABI number and the entry points it
    calls must be the library's."""
    import re
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class RenderCfg\(C\.Structure\):.*?_fields_ = \[(.*?)\]\s*#", doc, re.S)
    assert m, "RenderCfg stub not found in INTEGRATION.md"
    names = re.findall(r'\("(\w+)"', m.group(1))
    assert names == [n for n, _ in L.RenderCfg._fields_]
    assert f"clift_abi_version() == {L.ABI_VERSION}" in doc
    for fn in set(re.findall(r"lib\.(clift_\w+)", doc)):
        assert fn in L.SIGNATURES, fn


def test_models_survive_deepcopy_and_pickle_with_a_live_packed_view():
    """ddp_spawn, EMA copies and torch.save(model) copy the module; the packed parameter view (ctypes structs full of raw
    device pointers) is a cache and must not travel."""
    import copy
    import io
    import pickle

    class FakePacked:                     # stands in for the PackedField a CUDA render leaves behind
        def __init__(self):
            self.field = L.Field()
            self.field.basis = 0x1000

    model = cl.TensorVMSplit([8, 8, 8], num_semantic_classes=4, dim_feature_instance=6, use_semantic_mlp=True,
                             use_instance_mlp=True, slow_fast_mode=True)
    model._packed = FakePacked()
    with pytest.raises(ValueError):
        pickle.dumps(model._packed.field)     # the hazard: ctypes objects with pointers do not pickle
    for clone in (copy.deepcopy(model), pickle.loads(pickle.dumps(model))):
        assert clone._packed is None and model._packed is not None
        for (ka, a), (kb, b) in zip(model.state_dict().items(), clone.state_dict().items()):
            assert ka == kb and torch.equal(a, b) and a.data_ptr() != b.data_ptr()
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    assert torch.load(buf, weights_only=False)._packed is None
    rend = cl.TensoRFRenderer(syn.default_aabb(), [8, 8, 8], semantic_weight_mode="softmax")
    rend._cfg(model, 0)                   # fills the host geometry cache
    twin = copy.deepcopy(rend)
    twin.bbox_aabb[1, 2] = 0.5            # the copy owns its buffers and its cache follows them
    assert list(twin._cfg(model, 0).aabb_max) == [1.0, 1.0, 0.5] and list(rend._cfg(model, 0).aabb_max) == [1.0, 1.0, 1.0]


def test_packed_view_cache_validation_handles_inference_tensors():
    """PackedField.refresh trusts its packed copy only while every parameter's version counter AND storage address and the
    library's parameter epoch are unchanged, separately for training packings (data-gradient operands) and inference packings.
    Parameters created under torch.inference_mode() keep no counter: the cache is then never trusted (repack on every call)
    instead of raising 'Inference tensors do not track version counter'."""
    from contrastive_lift_b200 import field as F
    normal = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.ones(2))]
    with torch.inference_mode():
        frozen = [torch.nn.Parameter(torch.zeros(3))]
    assert F._param_version(normal[0]) == 0 and F._param_version(frozen[0]) is None

    def key(params):
        return tuple(F._param_version(p) for p in params) + tuple(p.data_ptr() for p in params) + (L.param_epoch(),)

    def stub(params, train=False, infer=True):
        pk = object.__new__(F.PackedField)          # no device work: only the cache check at the top of refresh() runs
        pk.model_params, pk.tc_stale = params, False
        pk.has_train, pk.has_infer = train, infer
        pk.versions = key(params)
        return pk

    assert stub(normal).refresh(False) is None      # valid cache: returns before touching the library
    assert stub(normal, train=True).refresh(True) is None      # the chunk renders of one training step share a packing
    with pytest.raises(AttributeError):             # an inference packing does not carry the training operands
        stub(normal, train=False).refresh(True)     # (falls through to the repack: this stub has no library handle)
    pk = stub(normal)
    with torch.no_grad():
        normal[1].add_(1.0)                         # a version moved: the early return must not happen
    with pytest.raises(AttributeError):
        pk.refresh(False)
    pk = stub(normal)
    normal[0].data = torch.ones(3)                  # the reference's EMA: storage replaced, version counter untouched
    with pytest.raises(AttributeError):
        pk.refresh(False)
    pk = stub(normal)
    L.bump_param_epoch()                            # ema_update / FusedAdam: raw-pointer writes
    with pytest.raises(AttributeError):
        pk.refresh(False)
    with pytest.raises(AttributeError):             # inference tensors: never trusted, even with an "equal" key
        stub(frozen).refresh(False)


def test_renderer_built_under_inference_mode_still_describes_itself():
    with torch.inference_mode():
        rend = cl.TensoRFRenderer(syn.default_aabb(), [8, 8, 8], semantic_weight_mode="softmax")
        model = cl.TensorVMSplit([8, 8, 8], num_semantic_classes=4, dim_feature_instance=6, use_semantic_mlp=True,
                                 use_instance_mlp=True, slow_fast_mode=True)
        assert rend._cfg(model, L.HEAD_ALL).n_samples == rend.n_samples
    rend.update_step_ratio(0.3)
    assert rend._cfg(model, 0).n_samples == rend.n_samples and list(rend._cfg(model, 0).aabb_min) == [-1.0, -1.0, -1.0]
