"""Host-side mirror of ``TensoRFRenderer`` (reference model/renderer/panopli_tensoRF_renderer.py:37-300).

Same constructor, buffers (``bbox_aabb, grid_dim, inv_box_extent, units``), attributes
(``step_size, n_samples, step_ratio``) and method signatures as the reference, so the call sites
trainer/train_panopli_tensorf.py:110,131,143 and inference/render_panopli.py:115 work unchanged.
Each of the three render entries is ONE call into libclift_b200.so through an autograd.Function;
chunking (config.chunk) is still honoured by the callers but no longer needed for memory.

RNG parity: the reference draws the per-ray jitter with ``torch.rand_like`` on a CPU tensor and the
random-background coin with ``torch.rand((1,))`` from the CPU default generator, per chunk call
(renderer:807-810, 164).  The same draws are made here, in the same order, and passed to the kernels.
"""
from __future__ import annotations

import collections
import ctypes as C
import warnings
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import lib as L
from .field import TensorVMSplit, param_list

_WORKSPACES = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    """Grow-only per-device scratch (the C ABI never allocates)."""
    ws = _WORKSPACES.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _WORKSPACES.pop(device, None)
        ws = torch.empty((int(nbytes * 1.25) + 1024,), dtype=torch.uint8, device=device)
        _WORKSPACES[device] = ws
    return ws


class _Render(torch.autograd.Function):
    """clift_render_forward / clift_render_backward.  ``params`` are the model parameters in
    ``model.named_parameters()`` order; they are inputs only so autograd routes gradients to them."""

    @staticmethod
    def forward(ctx, renderer, model, rays, jitter, add_bg, heads, want_points, need_grad, *params):
        with L.on(rays.device):         # every launch below targets the device that owns the rays / the model
            return _Render._forward(ctx, renderer, model, rays, jitter, add_bg, heads, want_points, need_grad, *params)

    @staticmethod
    def _forward(ctx, renderer, model, rays, jitter, add_bg, heads, want_points, need_grad, *params):
        lib = L.load()
        dev = rays.device
        # outputs no loss consumes arrive as None in backward (not as zero tensors): the main training pass never uses
        # instance_map (trainer:154), so its backward skips the instance head like the reference's autograd does
        ctx.set_materialize_grads(False)
        renderer._poll_overflow()          # an earlier render whose check was deferred: known by now?
        pk = model.packed(need_grad, params)
        cfg = renderer._cfg(model, heads)
        B = rays.shape[0]
        C_, DI = model.num_semantic_classes, (model.dim_feature_instance or 0)
        new = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        out = L.RenderOut()
        t = {"depth": new(B), "opacity": new(B)}
        if heads & L.HEAD_RGB:
            t["rgb"], t["rgb_raw"] = new(B, 3), new(B, 3)
            t["dist_ray"], t["dist_reg"] = new(B), new(1)
        if heads & L.HEAD_SEMANTIC:
            t["semantic"], t["semantic_raw"] = new(B, C_), new(B, C_)
        if heads & L.HEAD_INSTANCE:
            t["instance"] = new(B, DI)
        if want_points:
            t["points"] = new(B, 3)
        for k, v in t.items():
            setattr(out, k, L.ptr(v))
        out.save_for_backward = 2 if need_grad else 0      # 2: the backward's dL/dZ stash is allocated at backward time
        cap = renderer.max_active(B, need_grad, heads)
        # training renders with a capacity history verify the capacity WITHOUT stalling the host (see _defer_overflow_check)
        deferred = need_grad and renderer.deferred_overflow_check and heads in renderer._active_hist
        # small inference calls (the reference's chunk loop: 2048 rays per call, render_panopli.py:108-121) get room for
        # EVERY sample (24 bytes each): nothing can overflow, so there is nothing to read back and the loop never waits
        exact = (not need_grad and renderer.check_overflow and renderer.max_active_per_ray > 0
                 and B * int(renderer.n_samples) * 24 <= renderer.worst_case_capacity_bytes)
        if exact:
            cap = 0
        while True:
            nbytes = lib.clift_render_workspace_bytes(C.byref(cfg), C.byref(pk.field), B, cap, out.save_for_backward)
            if nbytes < 0:
                L.check(int(nbytes))
            if need_grad:
                # a training forward owns its workspace until its backward ran (several chunks may be in flight,
                # trainer:105-123); the caching allocator makes this a pointer bump after the first step
                ws = torch.empty((int(nbytes),), dtype=torch.uint8, device=dev)
                _WORKSPACES[dev] = ws
            else:
                ws = _workspace(dev, nbytes)
            L.check(lib.clift_render_forward(C.byref(cfg), C.byref(pk.field), L.ptr(rays), L.ptr(jitter), B, int(add_bg),
                                             L.ptr(ws), ws.numel(), cap, C.byref(out), L.stream_ptr(dev)))
            if B == 0 or not renderer.check_overflow or exact:
                break
            if deferred:
                renderer._defer_overflow_check(ws, B, cap, heads)
                break
            n_act, _, overflow, _ = renderer.last_stats(dev)       # one 32-byte D2H (the reference syncs ~10x per call)
            renderer._note_active(heads, n_act, B)
            if not overflow:
                break
            cap = (n_act + 127) // 128 * 128                          # rerun with the exact active-sample count
        ctx.need_grad = need_grad
        if need_grad:
            ctx.renderer, ctx.model, ctx.pk, ctx.cfg = renderer, model, pk, cfg
            ctx.rays, ctx.jitter, ctx.add_bg, ctx.heads, ctx.cap = rays, jitter, add_bg, heads, cap
            # only tensors that are NOT outputs are kept on ctx (an output kept here would be a reference cycle)
            ctx.t = {k: t[k] for k in ("rgb_raw", "semantic_raw", "opacity") if k in t}
            ctx.ws = ws
            ctx.param_names = pk.param_names
        empty = rays.new_zeros((0,))
        res = (t.get("rgb", empty), t.get("semantic", empty), t.get("instance", empty), t["depth"],
               t["dist_reg"].reshape(()) if "dist_reg" in t else empty, t.get("points", empty))
        ctx.mark_non_differentiable(res[3], res[5])
        renderer.last_opacity = t["opacity"]
        return res

    @staticmethod
    def backward(ctx, g_rgb, g_sem, g_ins, _g_depth, g_dist, _g_pts):
        with L.on(ctx.rays.device if ctx.need_grad else torch.cuda.current_device()):
            return _Render._backward(ctx, g_rgb, g_sem, g_ins, _g_depth, g_dist, _g_pts)

    @staticmethod
    def _backward(ctx, g_rgb, g_sem, g_ins, _g_depth, g_dist, _g_pts):
        if not ctx.need_grad:
            raise L.CliftError("backward through a render that was run without need_grad")
        renderer, model, pk = ctx.renderer, ctx.model, ctx.pk
        renderer._poll_overflow()
        if ctx.ws is None:
            raise L.CliftError("backward called twice on the same render (its saved state was consumed)")
        lib = L.load()
        dev = ctx.rays.device
        heads = ctx.heads
        prep = lambda g, on: g.contiguous().float() if (g is not None and on and g.numel() > 0) else None
        g_rgb = prep(g_rgb, heads & L.HEAD_RGB)
        g_sem = prep(g_sem, heads & L.HEAD_SEMANTIC)
        g_ins = prep(g_ins, heads & L.HEAD_INSTANCE)
        g_dist = prep(g_dist.reshape(1) if g_dist is not None and g_dist.numel() == 1 else None, heads & L.HEAD_RGB)
        want_density = (g_rgb is not None or g_dist is not None) and not ctx.renderer._density_frozen(heads)
        fg = pk.prepare_grads(want_density, g_rgb is not None, g_sem is not None, g_ins is not None)
        saved = L.RenderOut()
        for k, v in ctx.t.items():
            setattr(saved, k, L.ptr(v))
        saved.save_for_backward = 2
        B = ctx.rays.shape[0]
        # dL/dZ stash: scratch of this backward only (backwards of a step run one after the other, so the chunks of a
        # step share one such buffer through the caching allocator instead of each forward holding its own)
        zbytes = lib.clift_render_stash_z_bytes(C.byref(ctx.cfg), C.byref(pk.field), B, ctx.cap)
        if zbytes < 0:
            L.check(int(zbytes))
        stash_z = torch.empty((int(zbytes),), dtype=torch.uint8, device=dev)
        saved.stash_z = L.ptr(stash_z)
        L.check(lib.clift_render_backward(C.byref(ctx.cfg), C.byref(pk.field), L.ptr(ctx.rays), L.ptr(ctx.jitter), B,
                                          int(ctx.add_bg), L.ptr(ctx.ws), ctx.ws.numel(), ctx.cap,
                                          C.byref(saved), L.ptr(g_rgb), L.ptr(g_sem), L.ptr(g_ins), L.ptr(g_dist),
                                          C.byref(fg), L.stream_ptr(dev)))
        ctx.ws = None        # consumed
        st = L.stream_ptr(dev)
        # every gradient goes back to checkpoint layout in one launch.  The job list depends only on which gradients flow, so
        # it is built once per combination and replayed into one flat buffer afterwards (_UnpackPlan)
        plan_key = (want_density, g_rgb is not None, g_sem is not None, g_ins is not None)
        plan = pk.unpack_plans.get(plan_key)
        if plan is not None:
            grads = plan.run(lib, dev)
            return (None,) * 8 + tuple(grads.get(n) for n in ctx.param_names)
        batch = L.PackBatch()
        grads = {}
        if want_density:
            gp, gl = pk.unpack_factor_grads("density", batch)
            for i in range(3):
                grads[f"density_plane.{i}"], grads[f"density_line.{i}"] = gp[i], gl[i]
        if g_rgb is not None:
            gp, gl = pk.unpack_factor_grads("appearance", batch)
            for i in range(3):
                grads[f"appearance_plane.{i}"], grads[f"appearance_line.{i}"] = gp[i], gl[i]
            grads["appearance_basis_mat.weight"] = pk.basis.unpack_grads(lib, st, batch)[0]
            for n, g in zip(_mlp_names("render_appearance_mlp.mlp", pk.rgb), pk.rgb.unpack_grads(lib, st, batch)):
                grads[n] = g
        for name, gout in (("semantic", g_sem), ("instance", g_ins)):      # grid-mode heads: their own factors + basis
            if gout is not None and name in pk.grid_basis:
                gp, gl = pk.unpack_factor_grads(name, batch)
                for i in range(3):
                    grads[f"{name}_plane.{i}"], grads[f"{name}_line.{i}"] = gp[i], gl[i]
                grads[f"{name}_basis_mat.weight"] = pk.grid_basis[name].unpack_grads(lib, st, batch)[0]
        if g_sem is not None:
            for n, g in zip(_mlp_names("render_semantic_mlp.mlp", pk.sem), pk.sem.unpack_grads(lib, st, batch)):
                grads[n] = g
        if g_ins is not None and pk.insf is not None:
            for n, g in zip(_mlp_names("render_instance_mlp.mlp", pk.insf), pk.insf.unpack_grads(lib, st, batch)):
                grads[n] = g
            if pk.inss is not None:
                for n, g in zip(_mlp_names("render_instance_mlp.slow_mlp", pk.inss), pk.inss.unpack_grads(lib, st, batch)):
                    grads[n] = g
        pk.unpack_plans[plan_key] = _UnpackPlan(batch, grads)
        batch.run(lib, dev)
        return (None,) * 8 + tuple(grads.get(n) for n in ctx.param_names)


class _UnpackPlan:
    """The layout jobs that bring one combination of gradients back to checkpoint layout, with every destination expressed as
    an offset into ONE flat buffer: a later backward allocates that buffer, views it per parameter and launches the cached
    job table (one table per buffer address the caching allocator hands out) instead of ~45 allocations and job
    descriptors."""

    def __init__(self, batch: "L.PackBatch", grads):
        self.items, offset_of, off = [], {}, 0
        for name, g in grads.items():
            self.items.append((name, tuple(g.shape), g.numel(), off))
            offset_of[g.data_ptr()] = off * 4
            off += (g.numel() + 63) // 64 * 64                   # 256-byte aligned slots
        self.total = off
        self.jobs = [(j, offset_of[j.dst]) for j in batch.jobs]  # (descriptor with the planning run's dst, byte offset)
        self.tiles = int(batch.tiles)
        self.tables = {}

    def run(self, lib, dev):
        flat = torch.empty((self.total,), device=dev, dtype=torch.float32)
        base = flat.data_ptr()
        table = self.tables.get(base)
        if table is None:
            if len(self.tables) >= 8:
                self.tables.clear()
            rows = (L.PackJob * len(self.jobs))()
            for r, (j, off) in zip(rows, self.jobs):
                C.memmove(C.byref(r), C.byref(j), C.sizeof(L.PackJob))
                r.dst = base + off
            table = self.tables[base] = torch.frombuffer(bytearray(bytes(rows)), dtype=torch.uint8).to(dev, non_blocking=True)
        L.check(lib.clift_pack_batch(L.ptr(table), len(self.jobs), self.tiles, L.stream_ptr(dev)))
        return {name: flat[off:off + n].view(shape) for name, shape, n, off in self.items}


def _tensor_key(t: torch.Tensor):
    """(storage address, version) of a buffer; version None = a tensor created under torch.inference_mode(), which keeps no
    version counter - such a buffer is re-read on every call instead of cached."""
    try:
        return t.data_ptr(), t._version
    except RuntimeError:
        return t.data_ptr(), None


def _mlp_names(prefix: str, pm) -> List[str]:
    names = []
    for i, l in enumerate(pm.linears):
        names.append(f"{prefix}.{2 * i}.weight")
        if l.bias is not None:
            names.append(f"{prefix}.{2 * i}.bias")
    return names


class TensoRFRenderer(nn.Module):
    """Drop-in for the reference class of the same name (renderer:37-300)."""

    def __init__(self, bbox_aabb, grid_dim, stop_semantic_grad=True, semantic_weight_mode="none", step_ratio=0.5,
                 distance_scale=25, raymarch_weight_thres=0.0001, alpha_mask_threshold=0.0075, parent_renderer_ref=None,
                 instance_id=0, feature_stop_grad=False, verbose=False):
        super().__init__()
        if not stop_semantic_grad:
            raise L.CliftError("stop_semantic_grad=False is not built (every shipped config uses True, panopli_paper.yaml:35)")
        self._check_weight_mode(semantic_weight_mode)
        self.register_buffer("bbox_aabb", torch.as_tensor(bbox_aabb, dtype=torch.float32).clone())
        self.register_buffer("grid_dim", torch.LongTensor(list(grid_dim)))
        self.register_buffer("inv_box_extent", torch.zeros([3]))
        self.register_buffer("units", torch.zeros([3]))
        self.semantic_weight_mode = semantic_weight_mode
        self.parent_renderer_ref = parent_renderer_ref
        self.step_ratio = step_ratio
        self.distance_scale = distance_scale
        self.raymarch_weight_thres = raymarch_weight_thres
        self.alpha_mask_threshold = alpha_mask_threshold
        self.step_size = None
        self.n_samples = None
        self.stop_semantic_grad = stop_semantic_grad
        self.feature_stop_grad = feature_stop_grad
        self.instance_id = instance_id
        self.verbose = verbose
        # Capacity of the compacted active-sample list, per ray (active = weight > raymarch_weight_thres).  An
        # overflow is detected after the call (check_overflow) and the call is repeated at the exact size.
        self.max_active_per_ray = 192
        self.check_overflow = True
        # Training renders (the ones that size their stash by the capacity) check the capacity without a host sync once a
        # previous render with the same head set has told them what to expect: the 32-byte stats record goes to pinned
        # memory behind the render and is read by a later call (the next render / backward, or
        # synchronize_overflow_checks()).  An overflow found that way cannot be repaired - the truncated maps were already
        # handed out - so it raises (overflow_policy "raise") or warns ("warn") and the capacity history is corrected for
        # the following calls.  False: every render is verified before it returns (one D2H sync per call), and an
        # overflow repeats the call at the exact size.
        self.deferred_overflow_check = True
        self.overflow_policy = "raise"
        # inference calls whose worst case (every sample active) fits this many bytes of records skip the check altogether
        self.worst_case_capacity_bytes = 1 << 30
        self._active_hist = {}            # head set -> decayed maximum of active samples per ray
        self._pending = collections.deque()
        self._pinned = []
        # lib.HEADS_AUTO: tcgen05 tensor-core heads for inference, FP32-FMA heads for training forwards
        self.head_path = L.HEADS_AUTO
        self._active_per_ray = None
        self._host = None
        self.last_opacity = None
        # None: the C ABI's own bound (n_rays * n_samples < 2^31 per call); a smaller value forces the ray-range split
        # of _run (tests use it to cover the path 1600x1600 frames at inference sample counts take)
        self.max_rays_per_call = None
        self.update_step_size(self.grid_dim)

    def __getstate__(self):
        # copy.deepcopy / pickle (ddp_spawn, checkpointing the module object): outstanding capacity checks hold CUDA events and
        # pinned buffers of THIS process - a copy starts with none (the capacity history, plain numbers, travels)
        state = self.__dict__.copy()
        state["_pending"] = collections.deque()
        state["_pinned"] = []
        return state

    # ---- renderer:59-78 ---------------------------------------------------------------------------
    # update_step_size / update_step_ratio / get_target_resolution and the host half of update_bbox_aabb_and_shrink below are
    # deliberate TRANSCRIPTIONS of the reference's host-side bookkeeping (renderer:59-78, 683-713, 756-761): step size, sample
    # count and the shrunk box must come out of the same fp32 tensor operations in the same order, or n_samples / the sample
    # positions stop being bit-exact.  They are a handful of scalar tensor ops; everything per-ray / per-voxel runs in the
    # library.
    def update_step_size(self, grid_dim):
        box_extent = self.bbox_aabb[1] - self.bbox_aabb[0]
        self.grid_dim.data = torch.tensor(grid_dim, device=self.bbox_aabb.device) if isinstance(grid_dim, tuple) else grid_dim
        self.inv_box_extent.data = 2.0 / box_extent
        self.units.data = box_extent / (self.grid_dim - 1 + 1e-3)
        self.step_size = torch.mean(self.units) * self.step_ratio
        box_diag = torch.sqrt(torch.sum(torch.square(box_extent)))
        self.n_samples = int((box_diag / self.step_size).item()) + 1
        self._host = None
        if self.verbose:
            print(f"[{self.instance_id:02d}] aabb {self.bbox_aabb.view(-1).tolist()} grid {self.grid_dim.tolist()} "
                  f"step {float(self.step_size):.6f} samples {self.n_samples}")

    def update_step_ratio(self, step_ratio):
        self.step_ratio = step_ratio
        self.step_size = torch.mean(self.units) * self.step_ratio
        box_extent = self.bbox_aabb[1] - self.bbox_aabb[0]
        box_diag = torch.sqrt(torch.sum(torch.square(box_extent)))
        self.n_samples = int((box_diag / self.step_size).item()) + 1
        self._host = None

    def _apply(self, fn, *a, **k):
        self._host = None
        r = super()._apply(fn, *a, **k)
        if self.step_size is not None and torch.is_tensor(self.step_size):
            self.step_size = fn(self.step_size)
        return r

    def get_target_resolution(self, n_voxels):
        xyz_min, xyz_max = self.bbox_aabb
        voxel_size = ((xyz_max - xyz_min).prod() / n_voxels).pow(1 / 3)
        target_res = ((xyz_max - xyz_min) / voxel_size).long().tolist()
        return tuple(max(x, 1) for x in target_res)

    def max_active(self, n_rays: int, training: bool = False, heads: Optional[int] = None) -> int:
        """<= 0 means worst case (every sample active) to the C ABI.  Training forwards size their activation stash
        by this capacity, so they follow the measured active count of the previous calls with the same head set (decayed
        maximum x1.25 + 8 per ray) instead of the static per-ray bound; an overflow repeats the call at the exact size
        (checked renders) or is reported by a later call (deferred check)."""
        if self.max_active_per_ray <= 0:
            return 0
        per_ray = min(int(self.max_active_per_ray), int(self.n_samples))
        hist = self._active_hist.get(heads, self._active_per_ray)
        if training and self.check_overflow and hist is not None:
            per_ray = min(per_ray, int(hist * 1.25) + 8)
        return max(128, int(n_rays) * per_ray)

    # ---- capacity bookkeeping -------------------------------------------------------------------------
    def _note_active(self, heads: int, n_act: int, n_rays: int) -> None:
        apr = n_act / max(n_rays, 1)
        self._active_per_ray = max(1.0, apr)
        self._active_hist[heads] = max(1.0, apr, 0.9 * self._active_hist.get(heads, 0.0))

    def _defer_overflow_check(self, ws: torch.Tensor, n_rays: int, cap: int, heads: int) -> None:
        """Queue the stats record of the render just launched: device -> pinned host behind the render, no host wait."""
        dev = ws.device
        st = torch.empty((4,), dtype=torch.int64, device=dev)
        L.check(L.load().clift_render_stats(L.ptr(ws), L.ptr(st), L.stream_ptr(dev)))
        host = self._pinned.pop() if self._pinned else torch.empty((4,), dtype=torch.int64).pin_memory()
        host.copy_(st, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        self._pending.append((ev, host, n_rays, cap, heads))

    def _poll_overflow(self, wait: bool = False) -> None:
        """Read the deferred records whose copies have landed (all of them with ``wait``)."""
        while self._pending and (wait or self._pending[0][0].query()):
            ev, host, n_rays, cap, heads = self._pending.popleft()
            if wait:
                ev.synchronize()
            n_act, _, overflow, _ = (int(v) for v in host.tolist())
            self._pinned.append(host)
            self._note_active(heads, n_act, n_rays)
            if overflow:
                self._active_hist[heads] = max(self._active_hist[heads], n_act / max(n_rays, 1))
                msg = (f"a training render of {n_rays} rays had {n_act} active samples but room for {cap}: its maps and "
                       "gradients were computed from the first samples only.  The capacity history now covers this case; "
                       "repeat the step, or set renderer.deferred_overflow_check = False to verify every render before it "
                       "returns (one host sync per call)")
                if self.overflow_policy == "warn":
                    warnings.warn(msg, RuntimeWarning)
                else:
                    self._pending.clear()
                    raise L.CliftError(msg)

    def synchronize_overflow_checks(self) -> None:
        """Wait for every deferred capacity check (end of an epoch / before trusting the last step's numbers)."""
        self._poll_overflow(wait=True)

    @staticmethod
    def _check_weight_mode(mode) -> None:
        # renderer:142-143 composites the semantic / instance maps from the single arg-max sample in this mode; the kernels
        # composite over all active samples only ("softmax" and every other string, which the reference treats as "none")
        if mode == "argmax":
            raise L.CliftError('semantic_weight_mode="argmax" is not built (shipped configs use "softmax", panopli_paper.yaml:19)')

    # ---- descriptor for the C ABI -----------------------------------------------------------------
    def _cfg(self, model: TensorVMSplit, heads: int) -> L.RenderCfg:
        self._check_weight_mode(self.semantic_weight_mode)           # the attribute may have been changed after construction
        if heads & L.HEAD_SEMANTIC:
            # One flag (clift_render_cfg.semantic_softmax) switches both the per-sample Softmax of the semantic MLP
            # (tensoRF.py:594, the model's output_mlp_semantics) and the renderer's log-normalisation (renderer:160-162).
            # Every caller derives both from config.semantic_weight_mode (trainer:54,67; render_panopli.py:76,89); a pair that
            # disagrees would be computed differently from the reference, so it is refused instead.
            act = model.render_semantic_mlp.output_activation
            if not isinstance(act, (nn.Softmax, nn.Identity)):
                raise L.CliftError(f"output_mlp_semantics must be nn.Softmax(dim=-1) or nn.Identity(), got {type(act).__name__}")
            if isinstance(act, nn.Softmax) and act.dim not in (-1, 1):
                raise L.CliftError(f"output_mlp_semantics=nn.Softmax(dim={act.dim}): the class axis is the last one")
            if isinstance(act, nn.Softmax) != (self.semantic_weight_mode == "softmax"):
                raise L.CliftError(f'semantic_weight_mode="{self.semantic_weight_mode}" with output_mlp_semantics='
                                   f"{type(act).__name__}: the reference's callers build Softmax <-> \"softmax\" and Identity <-> "
                                   "anything else together (trainer:54,67), and the kernels switch both with one flag")
        # one D2H of 10 floats per geometry change, not per call.  The key also catches buffers replaced or written behind
        # update_step_size's back (on_load_checkpoint assigns renderer.bbox_aabb directly, trainer:466)
        key = (_tensor_key(self.bbox_aabb), _tensor_key(self.inv_box_extent), id(self.step_size), _tensor_key(self.grid_dim))
        if self._host is None or self._host[3] != key or None in (key[0][1], key[1][1], key[3][1]):
            self._host = (self.bbox_aabb.detach().cpu().tolist(), self.inv_box_extent.detach().cpu().tolist(),
                          float(self.step_size), key, tuple(self.grid_dim.tolist()))
        aabb, inv, step, _, grid = self._host
        cfg = L.RenderCfg()
        L.fill3(cfg.aabb_min, aabb[0])
        L.fill3(cfg.aabb_max, aabb[1])
        L.fill3(cfg.inv_extent, inv)
        cfg.step_size = step
        cfg.n_samples = int(self.n_samples)
        cfg.distance_scale = float(self.distance_scale)
        cfg.weight_thres = float(self.raymarch_weight_thres)
        cfg.semantic_softmax = 1 if self.semantic_weight_mode == "softmax" else 0
        cfg.heads = heads
        cfg.head_path = int(self.head_path)
        if grid != tuple(model.grid_dim()):
            raise L.CliftError(f"renderer.grid_dim {list(grid)} != model factor grid {model.grid_dim()}")
        return cfg

    @staticmethod
    def _density_frozen(heads: int) -> bool:
        # forward_instance_feature / forward_segment_feature evaluate density under no_grad (renderer:187-190, 268-271)
        return not (heads & L.HEAD_RGB)

    def _jitter(self, rays, perturb, is_train) -> Optional[torch.Tensor]:
        if is_train and perturb != 0:
            u = perturb * torch.rand((rays.shape[0], 1))          # CPU default generator, as renderer:808-810
            return u.reshape(-1).to(rays.device, non_blocking=True).contiguous()
        return None

    def _run(self, tensorf, rays, jitter, add_bg, heads, want_points):
        if not rays.is_cuda:
            raise L.CliftError("rays must be a CUDA tensor: this renderer has no CPU path")
        rays = rays.detach().contiguous().float()
        params = param_list(tensorf)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        limit = ((1 << 31) - 1) // max(int(self.n_samples), 1)        # one C call handles n_rays*n_samples < 2^31
        if self.max_rays_per_call is not None:
            limit = max(1, min(limit, int(self.max_rays_per_call)))
        if rays.shape[0] <= limit:
            return _Render.apply(self, tensorf, rays, jitter, add_bg, heads, want_points, need_grad, *params)
        # very large frames (e.g. 1600x1600 at 1024 samples/ray): split into ray ranges, concatenate the per-ray maps,
        # and recombine the per-call distortion means into the mean over all rays
        outs, sizes = [], []
        for b in range(0, rays.shape[0], limit):
            e = min(b + limit, rays.shape[0])
            outs.append(_Render.apply(self, tensorf, rays[b:e].contiguous(), None if jitter is None else jitter[b:e].contiguous(),
                                      add_bg, heads, want_points, need_grad, *params))
            sizes.append(e - b)
        cat = lambda i: torch.cat([o[i] for o in outs]) if outs[0][i].numel() else outs[0][i]
        wts = torch.tensor(sizes, device=rays.device, dtype=torch.float32) / float(rays.shape[0])
        dist = (torch.stack([o[4] for o in outs]) * wts).sum() if outs[0][4].numel() else outs[0][4]
        return cat(0), cat(1), cat(2), cat(3), dist, cat(5)

    # ---- renderer:80-176 --------------------------------------------------------------------------
    def forward(self, tensorf, rays, perturb, white_bg, is_train):
        jitter = self._jitter(rays, perturb, is_train)
        add_bg = bool(white_bg or (is_train and bool(torch.rand((1,)) < 0.5)))     # renderer:164, CPU coin per call
        rgb, sem, ins, depth, dist, _ = self._run(tensorf, rays, jitter, add_bg, self._heads(tensorf), False)
        return rgb, sem, ins, depth, torch.zeros([1, 1], device=rays.device), dist

    @staticmethod
    def _heads(tensorf) -> int:
        h = L.HEAD_RGB | L.HEAD_SEMANTIC
        if tensorf.render_instance_mlp is not None:
            h |= L.HEAD_INSTANCE
        return h

    # ---- renderer:178-217 -------------------------------------------------------------------------
    def forward_instance_feature(self, tensorf, rays, perturb, is_train):
        jitter = self._jitter(rays, perturb, is_train)
        _, _, ins, _, _, pts = self._run(tensorf, rays, jitter, False, L.HEAD_INSTANCE, True)
        return ins, pts

    # ---- renderer:259-300 -------------------------------------------------------------------------
    def forward_segment_feature(self, tensorf, rays, perturb, is_train):
        jitter = self._jitter(rays, perturb, is_train)
        _, seg, _, _, _, _ = self._run(tensorf, rays, jitter, False, L.HEAD_SEMANTIC, False)
        return seg

    # ---- diagnostics ------------------------------------------------------------------------------
    def last_stats(self, device) -> Tuple[int, int, int, int]:
        """(n_active, n_inbox, overflow, n_tiles) of the last render on ``device`` (one small D2H)."""
        ws = _WORKSPACES[torch.device(device) if not isinstance(device, torch.device) else device]
        st = torch.empty((4,), dtype=torch.int64, device=ws.device)
        with L.on(ws.device):
            L.check(L.load().clift_render_stats(L.ptr(ws), L.ptr(st), L.stream_ptr(ws.device)))
        return tuple(int(v) for v in st.cpu().tolist())

    # ---- renderer:668-729: epoch-boundary dense-alpha sweep, bounding box, factor shrink ------------------
    def _lattice(self, device):
        g = self.grid_dim.tolist()
        return [torch.linspace(0, 1, int(n)).to(device).contiguous() for n in g]     # the reference's own linspace values

    @torch.no_grad()
    def get_dense_alpha(self, tensorf):
        """-> (alpha [G0,G1,G2], dense_xyz [G0,G1,G2,3]) as renderer:717-729; alpha comes from clift_dense_alpha."""
        dev = tensorf.density_line[0].device
        pk = tensorf.packed(False)
        cfg = self._cfg(tensorf, 0)
        sx, sy, sz = self._lattice(dev)
        g = [int(v) for v in self.grid_dim.tolist()]
        alpha = torch.empty(g, device=dev)
        with L.on(dev):
            L.check(L.load().clift_dense_alpha(C.byref(cfg), C.byref(pk.field), L.ptr(sx), L.ptr(sy), L.ptr(sz), L.ptr(alpha),
                                               L.stream_ptr(dev)))
        samples = torch.stack(torch.meshgrid(sx, sy, sz, indexing="ij"), -1)
        dense_xyz = self.bbox_aabb[0] * (1 - samples) + self.bbox_aabb[1] * samples
        return alpha, dense_xyz

    @torch.no_grad()
    def alpha_bbox(self, tensorf):
        """(xyz_min, xyz_max, n_valid) of the lattice voxels whose 3x3x3-max-pooled alpha reaches alpha_mask_threshold
        (renderer:668-681): clift_dense_alpha + clift_alpha_bbox, one 7-number readback."""
        dev = tensorf.density_line[0].device
        lib, st = L.load(), L.stream_ptr(dev)
        pk = tensorf.packed(False)
        cfg = self._cfg(tensorf, 0)
        sx, sy, sz = self._lattice(dev)
        g = [int(v) for v in self.grid_dim.tolist()]
        alpha = torch.empty(g, device=dev)
        out = torch.zeros((8,), device=dev)
        scratch = torch.zeros((8,), dtype=torch.int32, device=dev)
        aabb = self._host[0]
        with L.on(dev):
            L.check(lib.clift_dense_alpha(C.byref(cfg), C.byref(pk.field), L.ptr(sx), L.ptr(sy), L.ptr(sz), L.ptr(alpha), st))
            L.check(lib.clift_alpha_bbox(L.ptr(alpha), (C.c_int32 * 3)(*g), L.ptr(sx), L.ptr(sy), L.ptr(sz),
                                         (C.c_float * 3)(*aabb[0]), (C.c_float * 3)(*aabb[1]), float(self.alpha_mask_threshold),
                                         L.ptr(out), out.data_ptr() + 24, L.ptr(scratch), st))
        host = out.cpu()
        n_valid = int(host[6:7].view(torch.int32).item())
        return host[0:3].to(dev), host[3:6].to(dev), n_valid

    @torch.no_grad()
    def update_bbox_aabb_and_shrink(self, tensorf, fractional_lenience=1.0):
        xyz_min, xyz_max, n_valid = self.alpha_bbox(tensorf)
        total_voxels = int(self.grid_dim[0] * self.grid_dim[1] * self.grid_dim[2])
        if n_valid == 0:
            print(f"[{self.instance_id:02d}] no valid voxels found ...")
            return
        # renderer:683-713 on 3-vectors (host-side bookkeeping, same operations in the same order)
        extent = xyz_max - xyz_min
        position = (xyz_min + xyz_max) / 2
        xyz_min_fl = position - (extent * fractional_lenience) / 2
        xyz_max_fl = position + (extent * fractional_lenience) / 2
        box_min, box_max = self.bbox_aabb[0], self.bbox_aabb[1]
        xyz_min = torch.maximum(box_min, xyz_min_fl)
        xyz_max = torch.minimum(box_max, xyz_max_fl)
        if self.parent_renderer_ref is not None:
            box_min, box_max = self.parent_renderer_ref.bbox_aabb[0], self.parent_renderer_ref.bbox_aabb[1]
            xyz_min = torch.maximum(box_min, xyz_min)
            xyz_max = torch.minimum(box_max, xyz_max)
        new_bbox_aabb = torch.stack((xyz_min, xyz_max))
        if self.verbose:
            print(f"[{self.instance_id:02d}] bbox: {xyz_min, xyz_max} alpha rest %%%f" % (n_valid / total_voxels * 100))
        t_l, b_r = (xyz_min - self.bbox_aabb[0]) / self.units, (xyz_max - self.bbox_aabb[0]) / self.units
        t_l, b_r = torch.round(torch.round(t_l)).long(), torch.round(b_r).long() + 1
        b_r = torch.stack([b_r, self.grid_dim]).amin(0)
        new_size = b_r - t_l
        if new_size[0] > 0 and new_size[1] > 0 and new_size[2] > 0:
            tensorf.shrink(t_l, b_r)
            self.bbox_aabb.data = new_bbox_aabb
            self.update_step_size((int(new_size[0]), int(new_size[1]), int(new_size[2])))
