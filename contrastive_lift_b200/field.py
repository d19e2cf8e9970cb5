"""Host-side mirror of ``TensorVMSplit`` (reference model/radiance_field/tensoRF.py:32-315).

Same constructor signature, attribute names and ``state_dict`` keys as the reference, so the
reference's trainer (trainer/train_panopli_tensorf.py:55-65, 99-103, 259, 449-457) and inference
scripts (inference/render_panopli.py:77-98) can construct it, load their checkpoints into it and
hand it to ``TensoRFRenderer``.  The parameters stay ordinary ``nn.Parameter``s in checkpoint layout
(optimizers, EMA and DDP-style all-reduce see what they expect); the kernels read a *packed* copy
(channel-last planes, transposed/padded Linear weights) that is refreshed through
``clift_pack_*`` whenever a parameter's version counter moves.

There is no PyTorch implementation of the field here: every lookup/MLP runs inside
libclift_b200.so.  Configurations outside the compiled envelope raise ``CliftError``.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lib as L

MATRIX_MODE = [[0, 1], [0, 2], [1, 2]]
VECTOR_MODE = [2, 1, 0]


def _sequential(n_in: int, width: int, n_out: int, n_layers: int) -> nn.Sequential:
    layers: List[nn.Module] = [nn.Linear(n_in, width)]
    for _ in range(n_layers - 2):
        layers += [nn.ReLU(inplace=True), nn.Linear(width, width)]
    layers += [nn.ReLU(inplace=True), nn.Linear(width, n_out)]
    return nn.Sequential(*layers)


class _HeadBase(nn.Module):
    """Parameter container with the reference head's attribute names.  The arithmetic of
    ``forward`` lives in the fused CUDA kernels; calling the module directly is not a supported path."""

    def forward(self, *args, **kwargs):  # pragma: no cover - guard
        raise L.CliftError(f"{type(self).__name__} is evaluated inside the fused clift_render_* kernels; "
                           "call TensoRFRenderer.forward / forward_instance_feature / forward_segment_feature")


class MLPRenderFeature(_HeadBase):
    """tensoRF.py:383-418 (parameters only)."""

    def __init__(self, in_channels, out_channels=3, pe_view=2, pe_feat=2, dim_mlp_color=128, output_activation=torch.sigmoid):
        super().__init__()
        self.pe_view, self.pe_feat = pe_view, pe_feat
        self.output_channels = out_channels
        self.view_independent = pe_view == 0 and pe_feat == 0
        self.in_feat_mlp = 2 * pe_view * 3 + 2 * pe_feat * in_channels + in_channels + (3 if not self.view_independent else 0)
        self.output_activation = output_activation
        self.mlp = _sequential(self.in_feat_mlp, dim_mlp_color, out_channels, 3)
        nn.init.constant_(self.mlp[-1].bias, 0)


class MLPRenderInstanceFeature(_HeadBase):
    """tensoRF.py:462-511 (parameters only)."""

    def __init__(self, in_channels, out_channels, num_mlp_layers=5, dim_mlp=256, pe_feat=0,
                 output_activation=nn.Softmax(dim=-1), use_features=False, slow_fast_mode=False):
        super().__init__()
        self.output_channels = out_channels
        self.output_activation = output_activation
        self.pe_feat = pe_feat
        self.use_features = use_features
        self.slow_fast_mode = slow_fast_mode
        self.in_feat_mlp = 2 * pe_feat * in_channels + in_channels + (64 if use_features else 0)
        self.mlp = _sequential(self.in_feat_mlp, dim_mlp, out_channels, num_mlp_layers)
        if slow_fast_mode:
            self.slow_mlp = _sequential(self.in_feat_mlp, dim_mlp, out_channels, num_mlp_layers)


class MLPRenderSemanticFeature(_HeadBase):
    """tensoRF.py:565-594 (parameters only)."""

    def __init__(self, in_channels, out_channels, pe_feat=0, num_mlp_layers=5, dim_mlp=256,
                 output_activation=nn.Identity(), use_features=False):
        super().__init__()
        self.output_channels = out_channels
        self.output_activation = output_activation
        self.pe_feat = pe_feat
        self.use_features = use_features
        self.in_feat_mlp = 2 * pe_feat * in_channels + in_channels + (64 if use_features else 0)
        self.mlp = _sequential(self.in_feat_mlp, dim_mlp, out_channels, num_mlp_layers)


def _param_version(p: torch.Tensor) -> Optional[int]:
    """Autograd version counter of a parameter; None for a tensor created under torch.inference_mode() (it keeps none), which
    makes PackedField.refresh repack on every call instead of trusting a cache it cannot validate."""
    try:
        return p._version
    except RuntimeError:
        return None


def param_list(model: nn.Module) -> List[torch.Tensor]:
    """The parameters of ``model`` in ``named_parameters()`` order (modules in pre-order, shared modules / parameters
    once), without the generator stack of ``nn.Module.named_parameters``: the render entries walk the model on every call."""
    out, seen_m, seen_p = [], set(), set()
    stack = [model]
    while stack:
        m = stack.pop()
        if m is None or id(m) in seen_m:
            continue
        seen_m.add(id(m))
        for q in m._parameters.values():
            if q is not None and id(q) not in seen_p:
                seen_p.add(id(q))
                out.append(q)
        stack.extend(reversed(m._modules.values()))
    return out


class _PackPlan:
    """The library calls of one PackedField refresh, recorded so that the next refresh of the same parameter storage
    replays them (same device tables, same operand buffers) instead of rebuilding ~100 job descriptors on the host.
    Every recorded entry takes the stream as its last argument; it is re-read at replay time."""

    RECORDED = ("clift_pack_batch", "clift_tc16_factor_bound", "clift_pack_linear_tc16_batch", "clift_pack_linear_x16_batch",
                "clift_pack_linear_tc", "clift_pack_linear_tc16")

    def __init__(self, lib, key):
        self._lib, self.key = lib, key
        self.calls: list = []
        self.keep: list = []          # device job tables the recorded pointers refer to

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name not in self.RECORDED:
            return fn

        def call(*args):
            self.calls.append((fn, args[:-1]))
            return fn(*args)
        return call

    def replay(self, stream) -> None:
        for fn, args in self.calls:
            L.check(fn(*args, stream))


def _linears(seq: nn.Sequential) -> List[nn.Linear]:
    return [m for m in seq if isinstance(m, nn.Linear)]


class _PackedMlp:
    """Device buffers of one Linear stack in kernel layout + their gradient mirrors."""

    def __init__(self, linears: Sequence[nn.Linear], device, with_bias: bool = True):
        self.linears = list(linears)
        self.dims = [self.linears[0].in_features] + [l.out_features for l in self.linears]
        self.wt = [torch.zeros((L.k_pad(l.in_features), L.n_pad(l.out_features)), device=device) for l in self.linears]
        self.bias = [torch.zeros((L.n_pad(l.out_features),), device=device) for l in self.linears]
        self.w_dgrad: List[Optional[torch.Tensor]] = [None] * len(self.linears)
        self.w_tc: List[Optional[torch.Tensor]] = [None] * len(self.linears)
        self.w_tc16: List[Optional[torch.Tensor]] = [None] * len(self.linears)
        self.w_t: List[Optional[torch.Tensor]] = [None] * len(self.linears)        # fp32 W^T, source of w_dg16
        self.w_dg16: List[Optional[torch.Tensor]] = [None] * len(self.linears)     # fp16-split operand of the data gradient
        self.w_x16: List[Optional[torch.Tensor]] = [None] * len(self.linears)      # pipelined-kernel stages of 256-wide layers
        self.g_wt: List[Optional[torch.Tensor]] = [None] * len(self.linears)
        self.g_bias: List[Optional[torch.Tensor]] = [None] * len(self.linears)

    def out_bound_ptr(self) -> int:
        """Device address of the bound on |output| of the last layer (header word 3 of its fp16-split operand);
        -1 = no such operand (a stack chained to it stays off the fp16 path)."""
        t = self.w_tc16[-1]
        return -1 if t is None else t.data_ptr() + 12

    def collect(self, batch: "L.PackBatch", training: bool):
        """Queue the fp32 layout jobs of this stack (W^T + bias, and the data-gradient copy in training) on ``batch``."""
        for i, l in enumerate(self.linears):
            batch.linear(l.weight.data, l.bias.data if l.bias is not None else None, self.wt[i], self.bias[i],
                         l.out_features, l.in_features)
            if training:
                if self.w_dgrad[i] is None:
                    self.w_dgrad[i] = torch.zeros((L.k_pad(l.out_features), L.dgrad_pad(l.in_features)), device=self.wt[i].device)
                batch.dgrad(l.weight.data, self.w_dgrad[i], l.out_features, l.in_features)
                if self.w_t[i] is None:
                    self.w_t[i] = torch.empty((l.in_features, l.out_features), device=self.wt[i].device)
                batch.transpose(l.weight.data, self.w_t[i], l.out_features, l.in_features)

    def pack(self, lib, stream, training: bool, in_bound_ptr: Optional[int] = None, in_bound_floor: float = 1.0,
             tc16: Optional["L.Tc16Batch"] = None, chain: int = 0):
        """Tensor-core operands of this stack (the fp32 layouts travel through ``collect``).
        ``in_bound_ptr`` / ``in_bound_floor``: bound on |input| of the first layer for the fp16-split operand chain
        (device scalar and/or constant; xyz, sin/cos and unit directions are bounded by 1)."""
        bound_ptr, floor = in_bound_ptr, in_bound_floor
        for i, l in enumerate(self.linears):
            has_bias = 1 if l.bias is not None else 0
            w, b = L.ptr(l.weight.data), L.ptr(l.bias.data) if l.bias is not None else None
            if not training:    # the 3xTF32 operands serve inference only (PackedField.refresh tracks their staleness)
                nf = lib.clift_tc_weight_floats(l.out_features, l.in_features, has_bias)
                if nf > 0:      # inside the tensor-core envelope: tf32 hi/lo operand
                    if self.w_tc[i] is None:
                        self.w_tc[i] = torch.zeros((nf,), device=self.wt[i].device)
                    L.check(lib.clift_pack_linear_tc(w, b, L.ptr(self.w_tc[i]), l.out_features, l.in_features, stream))
            # fp16-split operand (inference heads AND the training forward, which records the stash from the same kernel)
            nb = lib.clift_tc16_weight_bytes(l.out_features, l.in_features, has_bias)
            if nb > 0 and bound_ptr != -1:   # scales chained layer to layer
                if self.w_tc16[i] is None:
                    self.w_tc16[i] = torch.zeros((nb // 4,), device=self.wt[i].device)
                if tc16 is not None:       # queued: the whole model's operands go out as one batched call
                    tc16.add(l.weight.data, l.bias.data if l.bias is not None else None, self.w_tc16[i], l.out_features,
                             l.in_features, bound_ptr, floor, chain)
                else:
                    L.check(lib.clift_pack_linear_tc16(w, b, L.ptr(self.w_tc16[i]), l.out_features, l.in_features, bound_ptr,
                                                       float(floor), stream))
                bound_ptr, floor = self.w_tc16[i].data_ptr() + 12, 0.0
            else:
                bound_ptr = -1      # chain broken: the rest of this stack stays off the fp16 path
                self.w_tc16[i] = None

    def pack_x16(self, lib, jobs: list):
        """Queue the stages of the pipelined xyz-stack kernel: re-layout of the fp16-split operands just packed
        (``jobs``: [X16Job, ...] of the whole model, run as one clift_pack_linear_x16_batch launch)."""
        for i, l in enumerate(self.linears):
            has_bias = 1 if l.bias is not None else 0
            nb = lib.clift_x16_weight_bytes(l.out_features, l.in_features, has_bias)
            if nb <= 0 or self.w_tc16[i] is None or i + 1 == len(self.linears):
                self.w_x16[i] = None
                continue
            if self.w_x16[i] is None:
                self.w_x16[i] = torch.zeros((nb // 4,), device=self.wt[i].device)
            j = L.X16Job()
            j.w_tc16, j.dst = L.ptr(self.w_tc16[i]), L.ptr(self.w_x16[i])
            j.steps = (l.in_features + 15) // 16 + has_bias
            j.first_block = (jobs[-1].first_block + (jobs[-1].steps * 1024 + 255) // 256) if jobs else 0
            jobs.append(j)

    def pack_dgrad16(self, lib, tc16: "L.Tc16Batch", chain: int):
        """Queue the tensor-core data-gradient operands: clift_pack_linear_tc16 of W^T (no bias; the operand scale of dZ is
        chosen per tile inside the backward kernel, so no bound chain - floor 1)."""
        for i, l in enumerate(self.linears):
            nb = lib.clift_tc16_weight_bytes(l.in_features, l.out_features, 0)
            if nb <= 0 or self.w_t[i] is None:
                self.w_dg16[i] = None
                continue
            if self.w_dg16[i] is None:
                self.w_dg16[i] = torch.zeros((nb // 4,), device=self.wt[i].device)
            tc16.add(self.w_t[i], None, self.w_dg16[i], l.in_features, l.out_features, None, 1.0, chain)

    def fill(self, m: L.Mlp):
        m.n_layers = len(self.linears)
        for i, d in enumerate(self.dims):
            m.dims[i] = d
        for i in range(len(self.linears)):
            m.wt[i] = L.ptr(self.wt[i])
            m.bias[i] = L.ptr(self.bias[i])
            m.w_dgrad[i] = L.ptr(self.w_dgrad[i])
            m.w_tc[i] = L.ptr(self.w_tc[i])
            m.w_tc16[i] = L.ptr(self.w_tc16[i])
            m.w_dg16[i] = L.ptr(self.w_dg16[i])
            m.w_x16[i] = L.ptr(self.w_x16[i])

    def grad_buffers(self, g: L.MlpGrad, want: bool, to_zero: Optional[List[torch.Tensor]] = None):
        """``to_zero``: list the caller clears with one multi-tensor launch (else each buffer is cleared here)."""
        for i in range(len(self.linears)):
            if want:
                if self.g_wt[i] is None:
                    self.g_wt[i] = torch.zeros_like(self.wt[i])
                    self.g_bias[i] = torch.zeros_like(self.bias[i])
                elif to_zero is not None:
                    to_zero += [self.g_wt[i], self.g_bias[i]]
                else:
                    self.g_wt[i].zero_()
                    self.g_bias[i].zero_()
                g.wt[i] = L.ptr(self.g_wt[i])
                g.bias[i] = L.ptr(self.g_bias[i])
            else:
                g.wt[i] = None
                g.bias[i] = None

    def unpack_grads(self, lib, stream, batch: Optional["L.PackBatch"] = None) -> List[torch.Tensor]:
        """-> [dW0, db0, dW1, db1, ...] in nn.Linear layout (``batch``: queued there, valid after batch.run())."""
        out = []
        for i, l in enumerate(self.linears):
            gw = torch.empty_like(l.weight)
            gb = torch.empty_like(l.bias) if l.bias is not None else None
            if batch is not None:
                batch.unlinear(self.g_wt[i], self.g_bias[i], gw, gb, l.out_features, l.in_features)
            else:
                L.check(lib.clift_unpack_linear(L.ptr(self.g_wt[i]), L.ptr(self.g_bias[i]), L.ptr(gw), L.ptr(gb),
                                                l.out_features, l.in_features, stream))
            out.append(gw)
            if gb is not None:
                out.append(gb)
        return out


class TensorVMSplit(nn.Module):
    """Drop-in for the reference class of the same name (tensoRF.py:32-315)."""

    def __init__(self, grid_dim, num_density_comps=(16, 16, 16), num_appearance_comps=(48, 48, 48), num_semantics_comps=None,
                 num_instance_comps=None, dim_appearance=27, dim_semantics=27, dim_instances=27, splus_density_shift=-10,
                 pe_view=2, pe_feat=2, dim_mlp_color=128, dim_mlp_semantics=128, dim_mlp_instance=256, num_semantic_classes=0,
                 dim_feature_instance=None, output_mlp_semantics=torch.nn.Softmax(dim=-1), use_semantic_mlp=False,
                 use_instance_mlp=False, use_feature_reg=False, use_distilled_features_semantic=False,
                 use_distilled_features_instance=False, num_feature_comps=(48, 48, 48), pe_sem=0, pe_ins=0,
                 slow_fast_mode=False, use_proj=False):
        super().__init__()
        if use_distilled_features_semantic or use_distilled_features_instance:
            raise L.CliftError("distilled-feature grids are outside the B200 hot path (off in every shipped contrastive config)")
        if use_proj:
            raise L.CliftError("use_proj (SlowFastProjLayer) is outside the B200 hot path (off in every shipped config)")
        if use_feature_reg:
            raise L.CliftError("use_feature_regularization is not reachable from shipped configs and is not built")
        sem_grid = num_semantics_comps is not None and not use_semantic_mlp
        ins_grid = dim_feature_instance is not None and num_instance_comps is not None and not use_instance_mlp
        if not sem_grid and not use_semantic_mlp:
            raise L.CliftError("TensorVMSplit without a semantic head (use_semantic_mlp=False and no num_semantics_comps) "
                               "is not a configuration the reference's trainer or render scripts build")
        for comps in (num_density_comps, num_appearance_comps) + ((num_semantics_comps,) if sem_grid else ()) + \
                ((num_instance_comps,) if ins_grid else ()):
            if len(set(comps)) != 1:
                raise L.CliftError("per-mode component counts must be equal")
        self.num_density_comps = num_density_comps
        self.num_appearance_comps = num_appearance_comps
        self.num_semantics_comps = num_semantics_comps
        self.num_instance_comps = num_instance_comps
        self.dim_appearance = dim_appearance
        self.dim_semantics = dim_semantics
        self.dim_instances = dim_instances
        self.dim_feature_instance = dim_feature_instance
        ins_out_channels = dim_feature_instance // 2 if slow_fast_mode else dim_feature_instance
        self.num_semantic_classes = num_semantic_classes
        self.splus_density_shift = splus_density_shift
        self.use_semantic_mlp = use_semantic_mlp
        self.use_instance_mlp = use_instance_mlp
        self.slow_fast_mode = slow_fast_mode
        self.use_proj = use_proj
        self.use_feature_reg = False
        self.pe_view, self.pe_feat = pe_view, pe_feat
        self.pe_sem, self.pe_ins = pe_sem, pe_ins
        self.dim_mlp_color = dim_mlp_color
        self.matrix_mode = MATRIX_MODE
        self.vector_mode = VECTOR_MODE
        self.density_plane, self.density_line = self.init_one_svd(num_density_comps, grid_dim, 0.1)
        self.appearance_plane, self.appearance_line = self.init_one_svd(num_appearance_comps, grid_dim, 0.1)
        self.appearance_basis_mat = nn.Linear(sum(num_appearance_comps), dim_appearance, bias=False)
        self.render_appearance_mlp = MLPRenderFeature(dim_appearance, 3, pe_view, pe_feat, dim_mlp_color)
        self.semantic_plane = self.semantic_line = self.semantic_basis_mat = None
        self.instance_plane = self.instance_line = self.instance_basis_mat = None
        self.render_semantic_mlp = self.render_instance_mlp = None
        # tensoRF.py:72-85: a grid-mode head owns a VM factor set + bias-free basis and a 3-layer MLP on the basis feature
        if dim_feature_instance is not None:
            if ins_grid:
                self.instance_plane, self.instance_line = self.init_one_svd(num_instance_comps, grid_dim, 0.1)
                self.instance_basis_mat = nn.Linear(sum(num_instance_comps), dim_instances, bias=False)
                self.render_instance_mlp = MLPRenderInstanceFeature(dim_instances, ins_out_channels, num_mlp_layers=3,
                                                                    dim_mlp=dim_mlp_instance, output_activation=nn.Identity(),
                                                                    slow_fast_mode=slow_fast_mode)
            elif use_instance_mlp:
                self.render_instance_mlp = MLPRenderInstanceFeature(3, ins_out_channels, pe_feat=pe_ins, num_mlp_layers=4,
                                                                    dim_mlp=dim_mlp_instance, output_activation=nn.Identity(),
                                                                    slow_fast_mode=slow_fast_mode)
        if sem_grid:
            self.semantic_plane, self.semantic_line = self.init_one_svd(num_semantics_comps, grid_dim, 0.1)
            self.semantic_basis_mat = nn.Linear(sum(num_semantics_comps), dim_semantics, bias=False)
            self.render_semantic_mlp = MLPRenderSemanticFeature(dim_semantics, num_semantic_classes, num_mlp_layers=3,
                                                                dim_mlp=dim_mlp_semantics, output_activation=output_mlp_semantics)
        else:
            self.render_semantic_mlp = MLPRenderSemanticFeature(3, num_semantic_classes, pe_feat=pe_sem,
                                                                output_activation=output_mlp_semantics)
        self.use_distilled_features_semantic = False
        self.use_distilled_features_instance = False
        self.num_feature_comps = num_feature_comps
        self.feature_plane = self.feature_line = self.feature_basis_mat = self.render_feature_mlp = None
        self._packed: Optional["PackedField"] = None

    # ---- construction helpers (tensoRF.py:99-106) ---------------------------------------------
    def init_one_svd(self, n_components, grid_resolution, scale):
        plane_coef, line_coef = [], []
        for i in range(3):
            vec_id = VECTOR_MODE[i]
            m0, m1 = MATRIX_MODE[i]
            plane_coef.append(nn.Parameter(scale * torch.randn((1, n_components[i], grid_resolution[m1], grid_resolution[m0]))))
            line_coef.append(nn.Parameter(scale * torch.randn((1, n_components[i], grid_resolution[vec_id], 1))))
        return nn.ParameterList(plane_coef), nn.ParameterList(line_coef)

    @property
    def semantic_softmax(self) -> bool:
        return isinstance(self.render_semantic_mlp.output_activation, nn.Softmax)

    def grid_dim(self) -> Tuple[int, int, int]:
        """(gx, gy, gz) read back from the factor shapes (they change under shrink / upsample)."""
        return (self.density_plane[0].shape[3], self.density_plane[0].shape[2], self.density_line[0].shape[2])

    # ---- packed view ---------------------------------------------------------------------------
    def packed(self, training: bool, params: Optional[Sequence[torch.Tensor]] = None) -> "PackedField":
        """``params``: param_list(self) when the caller already walked the model."""
        if self._packed is None or not self._packed.matches(self, params):
            self._packed = PackedField(self)
        self._packed.refresh(training)
        return self._packed

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def __getstate__(self):
        # copy.deepcopy / pickle / torch.save(model) (ddp_spawn, EMA copies): the packed view is a cache full of raw device
        # pointers (ctypes) - it is rebuilt on the next render, never copied
        state = self.__dict__.copy()
        state["_packed"] = None
        return state

    def invalidate_packed(self) -> None:
        """Call after mutating parameters through ``.data`` outside a training render."""
        self._packed = None

    # ---- F1 point-wise entry (tensoRF.py:114-125); parity / dense-alpha use ---------------------
    @torch.no_grad()
    def compute_density(self, xyz_sampled: torch.Tensor) -> torch.Tensor:
        lib = L.load()
        pk = self.packed(False)
        xyz = xyz_sampled.detach().reshape(-1, 3).contiguous().float()
        sigma = torch.empty((xyz.shape[0],), device=xyz.device)
        with L.on(xyz.device):
            L.check(lib.clift_density(C.byref(pk.field), L.ptr(xyz), xyz.shape[0], L.ptr(sigma), L.stream_ptr(xyz.device)))
        return sigma.view(xyz_sampled.shape[:-1])

    # ---- grid surgery (tensoRF.py:158-197): parameter objects are replaced, packed view dropped --
    def _factor_sets(self):
        """(plane list, line list) of every VM factor set this model owns (grid-mode heads add theirs)."""
        sets = [(self.density_plane, self.density_line), (self.appearance_plane, self.appearance_line)]
        if self.semantic_plane is not None:
            sets.append((self.semantic_plane, self.semantic_line))
        if self.instance_plane is not None:
            sets.append((self.instance_plane, self.instance_line))
        return sets

    @torch.no_grad()
    def shrink(self, t_l, b_r):
        for planes, lines in self._factor_sets():
            for i in range(3):
                v = VECTOR_MODE[i]
                m0, m1 = MATRIX_MODE[i]
                lines[i] = nn.Parameter(lines[i].data[..., t_l[v]:b_r[v], :].contiguous())
                planes[i] = nn.Parameter(planes[i].data[..., t_l[m1]:b_r[m1], t_l[m0]:b_r[m0]].contiguous())
        self._packed = None

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        self.appearance_plane, self.appearance_line = self.upsample_plane_line(self.appearance_plane, self.appearance_line, res_target)
        self.density_plane, self.density_line = self.upsample_plane_line(self.density_plane, self.density_line, res_target)
        if self.semantic_plane is not None:
            self.semantic_plane, self.semantic_line = self.upsample_plane_line(self.semantic_plane, self.semantic_line, res_target)
        if self.instance_plane is not None:
            self.instance_plane, self.instance_line = self.upsample_plane_line(self.instance_plane, self.instance_line, res_target)
        self._packed = None

    @torch.no_grad()
    def upsample_plane_line(self, plane_coef, line_coef, res_target):
        # tensoRF.py:190-197: F.interpolate(bilinear, align_corners=True) -> clift_upsample_bilinear
        lib = L.load()

        def resize(t: torch.Tensor, h2: int, w2: int) -> torch.Tensor:
            src = t.data.contiguous()
            home = src.device
            if not src.is_cuda:
                # Lightning restores checkpoints before it moves the module to its device, so on_load_checkpoint
                # (trainer:461-469) resizes a CPU-resident model.  There is still no CPU arithmetic: the factors are staged
                # through the current CUDA device, resized by the same kernel, and returned to where they lived.
                if not torch.cuda.is_available():
                    raise L.CliftError("upsample_volume_grid needs a CUDA device (CPU-resident parameters are staged "
                                       "through it; there is no CPU path)")
                src = src.cuda()
            dst = torch.empty((1, src.shape[1], h2, w2), device=src.device)
            with L.on(src.device):
                L.check(lib.clift_upsample_bilinear(L.ptr(src), L.ptr(dst), src.shape[1], src.shape[2], src.shape[3], h2, w2,
                                                    L.stream_ptr(src.device)))
            return dst if home == dst.device else dst.to(home)

        for i in range(3):
            v = VECTOR_MODE[i]
            m0, m1 = MATRIX_MODE[i]
            plane_coef[i] = nn.Parameter(resize(plane_coef[i], int(res_target[m1]), int(res_target[m0])))
            line_coef[i] = nn.Parameter(resize(line_coef[i], int(res_target[v]), 1))
        return plane_coef, line_coef

    # ---- optimizer groups (tensoRF.py:199-246) --------------------------------------------------
    # The param-group schema (which parameter lists, in which order, with which lr / weight_decay keys) IS the drop-in
    # contract: trainer/__init__.py:134-139 builds its optimizers from these lists and checkpoints store optimizer state by
    # group index, so the four methods below reproduce the reference's groups one for one.
    def get_optimizable_parameters(self, lr_grid, lr_net, weight_decay=0):
        grad_vars = [{'params': self.density_line, 'lr': lr_grid, 'weight_decay': weight_decay},
                     {'params': self.appearance_line, 'lr': lr_grid},
                     {'params': self.density_plane, 'lr': lr_grid, 'weight_decay': weight_decay},
                     {'params': self.appearance_plane, 'lr': lr_grid},
                     {'params': self.appearance_basis_mat.parameters(), 'lr': lr_net},
                     {'params': self.render_appearance_mlp.parameters(), 'lr': lr_net}]
        if self.semantic_plane is not None:
            grad_vars.extend([{'params': self.semantic_plane, 'lr': lr_grid}, {'params': self.semantic_line, 'lr': lr_grid},
                              {'params': self.semantic_basis_mat.parameters(), 'lr': lr_net},
                              {'params': self.render_semantic_mlp.parameters(), 'lr': lr_net}])
        elif self.render_semantic_mlp is not None:
            grad_vars.append({'params': self.render_semantic_mlp.parameters(), 'lr': lr_net})
        return grad_vars

    def get_optimizable_density_parameters(self, lr_grid):
        return [{'params': self.density_line, 'lr': lr_grid}, {'params': self.density_plane, 'lr': lr_grid}]

    def get_optimizable_segment_parameters(self, lr_grid, lr_net, _weight_decay=0):
        if self.semantic_plane is not None:
            return [{'params': self.semantic_plane, 'lr': lr_grid}, {'params': self.semantic_line, 'lr': lr_grid},
                    {'params': self.semantic_basis_mat.parameters(), 'lr': lr_net},
                    {'params': self.render_semantic_mlp.parameters(), 'lr': lr_net}]
        return [{'params': self.render_semantic_mlp.parameters(), 'lr': lr_net}] if self.render_semantic_mlp is not None else []

    def get_optimizable_instance_parameters(self, lr_grid, lr_net, using_DINO=False):
        grad_vars = []
        if self.instance_plane is not None:
            grad_vars.extend([{'params': self.instance_plane, 'lr': lr_grid}, {'params': self.instance_line, 'lr': lr_grid},
                              {'params': self.instance_basis_mat.parameters(), 'lr': lr_net},
                              {'params': self.render_instance_mlp.mlp.parameters(), 'lr': lr_net}])
        elif self.render_instance_mlp is not None:
            grad_vars.append({'params': self.render_instance_mlp.mlp.parameters(), 'lr': lr_net})
        if self.slow_fast_mode and not using_DINO:
            grad_vars.append({'params': self.render_instance_mlp.slow_mlp.parameters(), 'lr': lr_net})
        return grad_vars

    # ---- TV regulariser (tensoRF.py:248-290) via clift_tv_loss -----------------------------------
    def tv_loss_density(self, regularizer=None):
        from .loss import plane_tv
        return sum(plane_tv(p) * 1e-2 for p in self.density_plane)

    def tv_loss_appearance(self, regularizer=None):
        from .loss import plane_tv
        return sum(plane_tv(p) * 1e-2 for p in self.appearance_plane)

    def tv_loss_semantics(self, regularizer=None):
        from .loss import plane_tv
        if self.semantic_plane is None:
            return 0
        return sum(plane_tv(p) * 1e-2 + plane_tv(l) * 1e-3 for p, l in zip(self.semantic_plane, self.semantic_line))

    def tv_loss_instances(self, regularizer=None):
        from .loss import plane_tv
        if self.instance_plane is None:
            return 0
        return sum(plane_tv(p) * 1e-2 + plane_tv(l) * 1e-3 for p, l in zip(self.instance_plane, self.instance_line))

    def total_tv_loss(self, regularizer, config, current_epoch):
        """Same signature as tensoRF.py:281-290.  ``regularizer`` (the reference's TVLoss module) is
        accepted and ignored: the stencil runs in clift_tv_loss.  Semantic/instance planes exist only in
        grid-head mode; their terms switch on at the reference's epochs and are zero otherwise."""
        from .loss import total_tv
        terms = [(p, 1e-2 * config.lambda_tv_density) for p in self.density_plane]
        terms += [(p, 1e-2 * config.lambda_tv_appearance) for p in self.appearance_plane]
        if self.semantic_plane is not None and current_epoch >= config.late_semantic_optimization:
            terms += [(p, 1e-2 * config.lambda_tv_semantics) for p in self.semantic_plane]
            terms += [(l, 1e-3 * config.lambda_tv_semantics) for l in self.semantic_line]
        if self.instance_plane is not None and current_epoch >= config.instance_optimization_epoch:
            terms += [(p, 1e-2 * config.lambda_tv_instances) for p in self.instance_plane]
            terms += [(l, 1e-3 * config.lambda_tv_instances) for l in self.instance_line]
        return total_tv(terms)        # one autograd node, ~20 launches instead of ~10 per plane


class PackedField:
    """Kernel-layout copy of a TensorVMSplit's parameters + the ctypes descriptor handed to the C ABI."""

    def __init__(self, model: TensorVMSplit):
        self.lib = L.load()
        dev = model.density_plane[0].device
        if dev.type != "cuda":
            raise L.CliftError("TensorVMSplit must live on a CUDA device: libclift_b200 has no CPU path")
        self.device = dev
        self.grid = model.grid_dim()
        self.versions: Optional[Tuple[int, ...]] = None
        self.tc_stale = True
        self.trained = False
        self.has_train = self.has_infer = False
        self.model_params = list(model.parameters())
        self.param_names = [n for n, _ in model.named_parameters()]
        self.plans: Dict[bool, _PackPlan] = {}
        self.unpack_plans: dict = {}          # renderer._UnpackPlan per combination of flowing gradients
        mk = lambda plist: [torch.empty((p.shape[2], p.shape[3], p.shape[1]), device=dev) for p in plist]
        mkl = lambda plist: [torch.empty((p.shape[2], p.shape[1]), device=dev) for p in plist]
        self.planes = {"density": mk(model.density_plane), "appearance": mk(model.appearance_plane)}
        self.lines = {"density": mkl(model.density_line), "appearance": mkl(model.appearance_line)}
        self.src = {"density": (model.density_plane, model.density_line),
                    "appearance": (model.appearance_plane, model.appearance_line)}
        # grid-mode heads (tensoRF.py:72-85): their own factor set + basis, same packed layouts as the appearance set
        self.grid_basis: Dict[str, _PackedMlp] = {}
        for name in ("semantic", "instance"):
            planes, lines = getattr(model, f"{name}_plane"), getattr(model, f"{name}_line")
            if planes is None:
                continue
            self.planes[name], self.lines[name] = mk(planes), mkl(lines)
            self.src[name] = (planes, lines)
            self.grid_basis[name] = _PackedMlp([getattr(model, f"{name}_basis_mat")], dev)
        self.basis = _PackedMlp([model.appearance_basis_mat], dev)
        self.tc16_scratch = torch.zeros((8,), device=dev)
        self.grid_scratch = {name: torch.zeros((8,), device=dev) for name in self.grid_basis}   # factor bounds of the grid sets
        self.rgb = _PackedMlp(_linears(model.render_appearance_mlp.mlp), dev)
        self.sem = _PackedMlp(_linears(model.render_semantic_mlp.mlp), dev)
        self.insf = _PackedMlp(_linears(model.render_instance_mlp.mlp), dev) if model.render_instance_mlp is not None else None
        self.inss = (_PackedMlp(_linears(model.render_instance_mlp.slow_mlp), dev)
                     if model.render_instance_mlp is not None and model.slow_fast_mode else None)
        f = L.Field()
        L.fill3(f.grid, self.grid)
        f.density_comps = model.num_density_comps[0]
        f.appearance_comps = model.num_appearance_comps[0]
        f.dim_appearance = model.dim_appearance
        f.pe_view, f.pe_feat, f.pe_sem, f.pe_ins = model.pe_view, model.pe_feat, model.pe_sem, model.pe_ins
        f.num_classes = model.num_semantic_classes
        f.dim_instance = model.render_instance_mlp.output_channels if model.render_instance_mlp is not None else 0
        f.slow_fast = 1 if model.slow_fast_mode else 0
        f.density_shift = float(model.splus_density_shift)
        for i in range(3):
            f.density_plane[i] = L.ptr(self.planes["density"][i])
            f.density_line[i] = L.ptr(self.lines["density"][i])
            f.appearance_plane[i] = L.ptr(self.planes["appearance"][i])
            f.appearance_line[i] = L.ptr(self.lines["appearance"][i])
        for name, gh in (("semantic", f.semantic_grid), ("instance", f.instance_grid)):
            if name not in self.grid_basis:
                gh.comps = 0
                continue
            gh.comps = self.src[name][0][0].shape[1]
            gh.dim = self.grid_basis[name].dims[1]
            for i in range(3):
                gh.plane[i] = L.ptr(self.planes[name][i])
                gh.line[i] = L.ptr(self.lines[name][i])
        self.field = f
        self.grad = L.FieldGrad()
        self.g_planes: Dict[str, List[Optional[torch.Tensor]]] = {k: [None] * 3 for k in self.planes}
        self.g_lines: Dict[str, List[Optional[torch.Tensor]]] = {k: [None] * 3 for k in self.planes}

    def matches(self, model, params: Optional[Sequence[torch.Tensor]] = None) -> bool:
        params = param_list(model) if params is None else params
        return len(params) == len(self.model_params) and all(a is b for a, b in zip(params, self.model_params)) and \
            self.device == model.density_plane[0].device

    def refresh(self, training: bool) -> None:
        # The packed copy is reused while no parameter changed: autograd version counters (optimizer steps, in-place ops),
        # the storage address of every parameter (the reference's EMA assigns ``param.data = ...``, trainer:325-329, which
        # replaces the storage without moving the version counter) and the library's own epoch (ema_update / FusedAdam write
        # through raw pointers).  The two chunk renders of one training step therefore share one packing.  A training
        # packing carries the data-gradient operands, an inference packing the 3xTF32 / pipelined-kernel operands; each is
        # built on first use for a given parameter state.
        ptrs = tuple(p.data_ptr() for p in self.model_params)
        versions = tuple(_param_version(p) for p in self.model_params) + ptrs + (L.param_epoch(),)
        same = versions == self.versions and None not in versions
        if same and (self.has_train if training else self.has_infer):
            return
        if not same:
            self.has_train = self.has_infer = False
        with L.on(self.device):
            plan = self.plans.get(training)
            if plan is not None and plan.key == ptrs:
                # same storage, new values (an optimizer step): the recorded launches read the parameters afresh
                plan.replay(L.stream_ptr(self.device))
                self.versions = versions
                self.tc_stale = training
            else:
                self.plans.pop(training, None)
                plan = _PackPlan(self.lib, ptrs)
                self._refresh(training, versions, plan)
                self.plans[training] = plan           # only a refresh that ran to the end is replayed
        if training:
            self.has_train = True
        else:
            self.has_infer = True

    def _refresh(self, training: bool, versions, plan: "_PackPlan") -> None:
        lib, st = plan, L.stream_ptr(self.device)        # the plan forwards to the library and records the launches
        # every fp32 layout job of the model (factor transposes, W^T + bias, data-gradient copies) in ONE launch
        batch = L.PackBatch()
        for name in self.src:
            planes, lines = self.src[name]
            for i in range(3):
                p, l = planes[i].data, lines[i].data
                batch.plane(p, self.planes[name][i], p.shape[1], p.shape[2], p.shape[3])
                batch.plane(l, self.lines[name][i], l.shape[1], l.shape[2], 1)
        for m in [self.basis, self.rgb, self.sem, self.insf, self.inss] + list(self.grid_basis.values()):
            if m is not None:
                m.collect(batch, training)
        plan.keep.append(batch.run(lib, self.device))
        # fp16-split operand chain: |plane*line| bound -> basis -> (features, dirs, sin/cos) -> rgb stack
        vp3, i3 = C.c_void_p * 3, C.c_int64 * 3
        ap, al = self.planes["appearance"], self.lines["appearance"]
        L.check(lib.clift_tc16_factor_bound(vp3(*[L.ptr(t) for t in ap]), vp3(*[L.ptr(t) for t in al]),
                                            i3(*[t.numel() for t in ap]), i3(*[t.numel() for t in al]),
                                            L.ptr(self.tc16_scratch), st))
        tc16 = L.Tc16Batch()          # chain 0: basis -> rgb stack; chains 1-3: the xyz stacks
        self.basis.pack(lib, st, training, self.tc16_scratch.data_ptr() + 24, 0.0, tc16, 0)
        self.rgb.pack(lib, st, training, self.basis.out_bound_ptr(), 1.0, tc16, 0)
        # grid-mode heads: |plane*line| bound of the head's own factor set -> its basis -> the head's stack(s).  One CTA plans
        # a chain in table order, so a basis sits in the SAME chain as the stacks that read its output bound (the instance
        # basis feeds the fast and the slow net: both follow it in chain 2).
        for chain, (name, m) in enumerate((("semantic", self.sem), ("instance", self.insf), ("instance", self.inss)), start=1):
            if m is None:
                continue
            gb = self.grid_basis.get(name)
            if gb is None:
                m.pack(lib, st, training, None, 1.0, tc16, chain)          # MLP mode: xyz (+ sin/cos), |x| <= 1
                continue
            chain = 1 if name == "semantic" else 2
            if m is not self.inss:
                gp, gl = self.planes[name], self.lines[name]
                L.check(lib.clift_tc16_factor_bound(vp3(*[L.ptr(t) for t in gp]), vp3(*[L.ptr(t) for t in gl]),
                                                    i3(*[t.numel() for t in gp]), i3(*[t.numel() for t in gl]),
                                                    L.ptr(self.grid_scratch[name]), st))
                gb.pack(lib, st, training, self.grid_scratch[name].data_ptr() + 24, 0.0, tc16, chain)
            m.pack(lib, st, training, gb.out_bound_ptr(), 0.0, tc16, chain)     # input = the basis feature
        if training:      # tensor-core data-gradient operands (W^T), one planning chain per stack
            for chain, m in enumerate((self.basis, self.rgb, self.sem, self.insf, self.inss), start=4):
                if m is not None:
                    m.pack_dgrad16(lib, tc16, chain)
        plan.keep.append(tc16.run(lib, self.device))
        # stages of the pipelined kernel for the xyz stacks (MLP-mode heads): inference and training forwards
        jobs: list = []
        for name, m in (("semantic", self.sem), ("instance", self.insf), ("instance", self.inss)):
            if m is not None and name not in self.grid_basis:
                m.pack_x16(lib, jobs)
        if jobs:
            raw = bytes((L.X16Job * len(jobs))(*jobs))
            table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device, non_blocking=True)
            blocks = jobs[-1].first_block + (jobs[-1].steps * 1024 + 255) // 256
            plan.keep.append(table)
            L.check(lib.clift_pack_linear_x16_batch(L.ptr(table), len(jobs), blocks, st))
        f = self.field
        for name, gh in (("semantic", f.semantic_grid), ("instance", f.instance_grid)):
            if name in self.grid_basis:
                gb = self.grid_basis[name]
                gh.basis = L.ptr(gb.wt[0])
                gh.basis_dgrad = L.ptr(gb.w_dgrad[0])
                gh.basis_tc16 = L.ptr(gb.w_tc16[0])
        f.basis = L.ptr(self.basis.wt[0])
        f.basis_dgrad = L.ptr(self.basis.w_dgrad[0])
        f.basis_tc = L.ptr(self.basis.w_tc[0])
        f.basis_tc16 = L.ptr(self.basis.w_tc16[0])
        f.basis_dg16 = L.ptr(self.basis.w_dg16[0])
        self.rgb.fill(f.rgb)
        self.sem.fill(f.semantic)
        if self.insf is not None:
            self.insf.fill(f.instance_fast)
        if self.inss is not None:
            self.inss.fill(f.instance_slow)
        self.versions = versions
        self.tc_stale = training        # training refreshes skip the 3xTF32 operands
        self.trained = self.trained or training

    # ---- gradient side ---------------------------------------------------------------------------
    def prepare_grads(self, want_density: bool, want_rgb: bool, want_sem: bool, want_ins: bool) -> L.FieldGrad:
        g = self.grad
        to_zero: List[torch.Tensor] = []            # cleared with one multi-tensor launch at the end
        wants = {"density": want_density, "appearance": want_rgb, "semantic": want_sem, "instance": want_ins}
        for name in self.planes:
            want = wants[name]
            gh = {"semantic": g.semantic_grid, "instance": g.instance_grid}.get(name)
            for i in range(3):
                if want:
                    if self.g_planes[name][i] is None:
                        self.g_planes[name][i] = torch.zeros_like(self.planes[name][i])
                        self.g_lines[name][i] = torch.zeros_like(self.lines[name][i])
                    else:
                        to_zero += [self.g_planes[name][i], self.g_lines[name][i]]
                gp = L.ptr(self.g_planes[name][i]) if want else None
                gl = L.ptr(self.g_lines[name][i]) if want else None
                if name == "density":
                    g.density_plane[i], g.density_line[i] = gp, gl
                elif name == "appearance":
                    g.appearance_plane[i], g.appearance_line[i] = gp, gl
                else:
                    gh.plane[i], gh.line[i] = gp, gl
            if gh is not None:
                dummy = L.MlpGrad()
                self.grid_basis[name].grad_buffers(dummy, want, to_zero)
                gh.basis = dummy.wt[0]
        dummy = L.MlpGrad()
        self.basis.grad_buffers(dummy, want_rgb, to_zero)
        g.basis = dummy.wt[0]
        self.rgb.grad_buffers(g.rgb, want_rgb, to_zero)
        self.sem.grad_buffers(g.semantic, want_sem, to_zero)
        if self.insf is not None:
            self.insf.grad_buffers(g.instance_fast, want_ins, to_zero)
        if self.inss is not None:
            self.inss.grad_buffers(g.instance_slow, want_ins, to_zero)
        if to_zero:
            torch._foreach_zero_(to_zero)
        return g

    def unpack_factor_grads(self, name: str, batch: Optional["L.PackBatch"] = None) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        lib, st = self.lib, L.stream_ptr(self.device)
        planes, lines = self.src[name]
        gp, gl = [], []
        for i in range(3):
            p, l = planes[i], lines[i]
            a = torch.empty_like(p.data)
            b = torch.empty_like(l.data)
            if batch is not None:
                batch.unplane(self.g_planes[name][i], a, p.shape[1], p.shape[2], p.shape[3])
                batch.unplane(self.g_lines[name][i], b, l.shape[1], l.shape[2], 1)
            else:
                L.check(lib.clift_unpack_plane(L.ptr(self.g_planes[name][i]), L.ptr(a), p.shape[1], p.shape[2], p.shape[3], st))
                L.check(lib.clift_unpack_plane(L.ptr(self.g_lines[name][i]), L.ptr(b), l.shape[1], l.shape[2], 1, st))
            gp.append(a)
            gl.append(b)
        return gp, gl
