"""ctypes binding of libclift_b200.so (include/clift_b200.h).

The library is the product: there is no Python/PyTorch fallback for any compute entry.  If the
shared object is missing or a call fails, an exception is raised - nothing is silently rerouted.
PyTorch is used only for device memory (caching allocator), streams and torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

MAX_LAYERS = 8
ABI_VERSION = 13

HEAD_RGB, HEAD_SEMANTIC, HEAD_INSTANCE, HEAD_ALL = 1, 2, 4, 7
HEADS_AUTO, HEADS_FMA, HEADS_TENSOR, HEADS_TENSOR16 = 0, 1, 2, 3

_fp = C.POINTER(C.c_float)
_vp = C.c_void_p


class Mlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (MAX_LAYERS + 1)),
                ("wt", _vp * MAX_LAYERS), ("bias", _vp * MAX_LAYERS), ("w_dgrad", _vp * MAX_LAYERS), ("w_tc", _vp * MAX_LAYERS),
                ("w_tc16", _vp * MAX_LAYERS), ("w_dg16", _vp * MAX_LAYERS),
                ("w_x16", _vp * MAX_LAYERS)]


class MlpGrad(C.Structure):
    _fields_ = [("wt", _vp * MAX_LAYERS), ("bias", _vp * MAX_LAYERS)]


class GridHead(C.Structure):
    _fields_ = [("comps", C.c_int32), ("dim", C.c_int32), ("plane", _vp * 3), ("line", _vp * 3), ("basis", _vp),
                ("basis_dgrad", _vp), ("basis_tc16", _vp)]


class GridHeadGrad(C.Structure):
    _fields_ = [("plane", _vp * 3), ("line", _vp * 3), ("basis", _vp)]


class Field(C.Structure):
    _fields_ = [("grid", C.c_int32 * 3), ("density_comps", C.c_int32), ("appearance_comps", C.c_int32),
                ("dim_appearance", C.c_int32), ("pe_view", C.c_int32), ("pe_feat", C.c_int32),
                ("pe_sem", C.c_int32), ("pe_ins", C.c_int32), ("num_classes", C.c_int32),
                ("dim_instance", C.c_int32), ("slow_fast", C.c_int32), ("density_shift", C.c_float),
                ("density_plane", _vp * 3), ("density_line", _vp * 3),
                ("appearance_plane", _vp * 3), ("appearance_line", _vp * 3), ("basis", _vp), ("basis_dgrad", _vp), ("basis_tc", _vp),
                ("basis_tc16", _vp), ("basis_dg16", _vp),
                ("rgb", Mlp), ("semantic", Mlp), ("instance_fast", Mlp), ("instance_slow", Mlp),
                ("semantic_grid", GridHead), ("instance_grid", GridHead)]


class FieldGrad(C.Structure):
    _fields_ = [("density_plane", _vp * 3), ("density_line", _vp * 3),
                ("appearance_plane", _vp * 3), ("appearance_line", _vp * 3), ("basis", _vp),
                ("rgb", MlpGrad), ("semantic", MlpGrad), ("instance_fast", MlpGrad), ("instance_slow", MlpGrad),
                ("semantic_grid", GridHeadGrad), ("instance_grid", GridHeadGrad)]


class RenderCfg(C.Structure):
    _fields_ = [("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3), ("inv_extent", C.c_float * 3),
                ("step_size", C.c_float), ("n_samples", C.c_int32), ("distance_scale", C.c_float),
                ("weight_thres", C.c_float), ("semantic_softmax", C.c_int32), ("heads", C.c_int32),
                ("head_path", C.c_int32)]


class RenderOut(C.Structure):
    _fields_ = [("rgb", _vp), ("semantic", _vp), ("instance", _vp), ("depth", _vp), ("opacity", _vp),
                ("dist_reg", _vp), ("rgb_raw", _vp), ("semantic_raw", _vp), ("dist_ray", _vp),
                ("points", _vp), ("weights", _vp), ("save_for_backward", C.c_int32), ("stash_z", _vp)]


class PackJob(C.Structure):
    _fields_ = [("src", _vp), ("dst", _vp), ("s_pitch", C.c_int32), ("s_rows", C.c_int32), ("s_cols", C.c_int32),
                ("d_rows", C.c_int32), ("d_cols", C.c_int32), ("kind", C.c_int32), ("first_tile", C.c_int32),
                ("reserved", C.c_int32)]


class Tc16Job(C.Structure):
    _fields_ = [("w", _vp), ("bias", _vp), ("dst", _vp), ("in_bound", _vp), ("in_bound_floor", C.c_float),
                ("n_out", C.c_int32), ("n_in", C.c_int32), ("chain", C.c_int32), ("first_block", C.c_int32),
                ("reserved", C.c_int32)]


class X16Job(C.Structure):
    _fields_ = [("w_tc16", _vp), ("dst", _vp), ("steps", C.c_int32), ("first_block", C.c_int32)]


class AdamTensor(C.Structure):
    _fields_ = [("param", _vp), ("grad", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp), ("n", C.c_int64)]


MAX_ADAM_GROUPS = 8


class AdamGroup(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("first", C.c_int32), ("count", C.c_int32), ("reserved", C.c_int32), ("step", C.c_int64)]


class EmaPair(C.Structure):
    _fields_ = [("slow", _vp), ("fast", _vp), ("n", C.c_int64)]


class TvJob(C.Structure):
    _fields_ = [("plane_hwc", _vp), ("loss", _vp), ("grad_hwc", _vp), ("comps", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("grad_scale", C.c_float)]


class CliftError(RuntimeError):
    pass


_LIB: Optional[C.CDLL] = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libclift_b200.so")

# name -> (restype, argtypes); every symbol include/clift_b200.h declares
SIGNATURES = {
    "clift_abi_version": (C.c_int32, []),
    "clift_last_error": (C.c_char_p, []),
    "clift_launch_count": (C.c_int64, []),
    "clift_profile_enable": (C.c_int32, [C.c_int32]),
    "clift_profile_stage_ms": (C.c_int32, [_fp]),
    "clift_profile_heads_split_ms": (C.c_int32, [_fp]),
    "clift_debug_last_head_path": (C.c_int32, []),
    "clift_pack_plane": (C.c_int32, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "clift_pack_batch": (C.c_int32, [_vp, C.c_int32, C.c_int32, _vp]),
    "clift_pack_linear_tc16_batch": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "clift_unpack_plane": (C.c_int32, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "clift_pack_linear": (C.c_int32, [_vp, _vp, _vp, _vp, C.c_int32, C.c_int32, _vp]),
    "clift_unpack_linear": (C.c_int32, [_vp, _vp, _vp, _vp, C.c_int32, C.c_int32, _vp]),
    "clift_pack_linear_dgrad": (C.c_int32, [_vp, _vp, C.c_int32, C.c_int32, _vp]),
    "clift_tc_weight_floats": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "clift_pack_linear_tc": (C.c_int32, [_vp, _vp, _vp, C.c_int32, C.c_int32, _vp]),
    "clift_x16_weight_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "clift_pack_linear_x16": (C.c_int32, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "clift_pack_linear_x16_batch": (C.c_int32, [_vp, C.c_int32, C.c_int32, _vp]),
    "clift_debug_tc_gemm": (C.c_int32, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "clift_tc16_weight_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "clift_pack_linear_tc16": (C.c_int32, [_vp, _vp, _vp, C.c_int32, C.c_int32, _vp, C.c_float, _vp]),
    "clift_tc16_factor_bound": (C.c_int32, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64), _vp, _vp]),
    "clift_debug_tc16_gemm": (C.c_int32, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "clift_debug_tc_trace": (C.c_int32, [_vp]),
    "clift_gen_rays": (C.c_int32, [_fp, _fp, C.c_int32, C.c_int32, C.c_float, C.c_float, _vp, _vp, _vp]),
    "clift_sample_points": (C.c_int32, [C.POINTER(RenderCfg), _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "clift_density": (C.c_int32, [C.POINTER(Field), _vp, C.c_int64, _vp, _vp]),
    "clift_render_workspace_bytes": (C.c_int64, [C.POINTER(RenderCfg), C.POINTER(Field), C.c_int64, C.c_int64,
                                                       C.c_int32]),
    "clift_render_stash_z_bytes": (C.c_int64, [C.POINTER(RenderCfg), C.POINTER(Field), C.c_int64, C.c_int64]),
    "clift_render_forward": (C.c_int32, [C.POINTER(RenderCfg), C.POINTER(Field), _vp, _vp, C.c_int64, C.c_int32,
                                         _vp, C.c_int64, C.c_int64, C.POINTER(RenderOut), _vp]),
    "clift_render_stats": (C.c_int32, [_vp, _vp, _vp]),
    "clift_render_backward": (C.c_int32, [C.POINTER(RenderCfg), C.POINTER(Field), _vp, _vp, C.c_int64, C.c_int32,
                                          _vp, C.c_int64, C.c_int64, C.POINTER(RenderOut), _vp, _vp, _vp, _vp,
                                          C.POINTER(FieldGrad), _vp]),
    "clift_slowfast_loss": (C.c_int32, [_vp, _vp, _vp, C.c_int32, C.c_int32, _vp, _vp, _vp]),
    "clift_ema_update": (C.c_int32, [_vp, _vp, C.c_int64, C.c_double, _vp]),
    "clift_ema_update_batch": (C.c_int32, [_vp, C.c_int32, C.c_int64, C.c_double, _vp]),
    "clift_tv_loss_batch": (C.c_int32, [_vp, C.c_int32, C.c_int64, _vp]),
    "clift_adam_step_groups": (C.c_int32, [_vp, C.c_int32, C.c_int64, C.POINTER(AdamGroup), C.c_int32, C.c_float, _vp]),
    "clift_contrastive_loss": (C.c_int32, [_vp, _vp, C.c_int32, C.c_int32, C.c_float, _vp, _vp, _vp]),
    "clift_tv_loss": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, C.c_float, _vp]),
    "clift_adam_step": (C.c_int32, [_vp, C.c_int32, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64,
                                    C.c_float, _vp]),
    "clift_dense_alpha": (C.c_int32, [C.POINTER(RenderCfg), C.POINTER(Field), _vp, _vp, _vp, _vp, _vp]),
    "clift_alpha_bbox": (C.c_int32, [_vp, C.POINTER(C.c_int32), _vp, _vp, _vp, _fp, _fp, C.c_float, _vp, _vp, _vp, _vp]),
    "clift_upsample_bilinear": (C.c_int32, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "clift_assign_centroids": (C.c_int32, [_vp, C.c_int64, C.c_int32, C.c_int32, _vp, C.c_int32, _vp, _vp, _vp]),
    "clift_assign_clusters": (C.c_int32, [_vp, C.c_int64, C.c_int32, C.c_int32, _vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "clift_labels_onehot": (C.c_int32, [_vp, C.c_int64, C.c_int32, _vp, _vp]),
    "clift_allreduce_grads": (C.c_int32, [_vp, _vp, C.c_int64, _vp]),
}


def load() -> C.CDLL:
    """dlopen the in-tree library (never a site-packages copy) and type every entry point."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise CliftError(f"{LIB_PATH} is missing: run `python -m contrastive_lift_b200.build` "
                         "(there is no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    got = lib.clift_abi_version()
    if got != ABI_VERSION:
        raise CliftError(f"libclift_b200.so has ABI {got}, the Python host expects {ABI_VERSION}: rebuild")
    _LIB = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().clift_last_error().decode(errors="replace")
        raise CliftError(f"clift status {rc}: {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CliftError("libclift_b200 takes device pointers: got a CPU tensor (no CPU fallback exists)")
    if not t.is_contiguous():
        raise CliftError("non-contiguous tensor passed to libclift_b200")
    return t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def on(device) -> "torch.cuda.device":
    """Context manager every libclift call site runs under: kernels, cudaFuncSetAttribute and the default-stream pointer all
    refer to the CURRENT device, so it must be the device that owns the tensors (a model on cuda:1 with cuda:0 current
    would otherwise launch on the wrong GPU)."""
    dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
    if dev.type != "cuda":
        raise CliftError(f"libclift_b200 runs on CUDA devices only: got a tensor on {dev} (no CPU fallback exists)")
    return torch.cuda.device(dev)


def launch_count() -> int:
    return int(load().clift_launch_count())


def k_pad(k: int) -> int:
    return (k + 15) // 16 * 16


def n_pad(n: int) -> int:
    return (n + 63) // 64 * 64


def dgrad_pad(n: int) -> int:
    return 64 if n <= 64 else (128 if n <= 128 else 256)


_PARAM_EPOCH = 0


def bump_param_epoch() -> None:
    """Invalidate every cached packed-parameter copy (parameters changed behind autograd's version counters)."""
    global _PARAM_EPOCH
    _PARAM_EPOCH += 1


def param_epoch() -> int:
    return _PARAM_EPOCH


def fill3(arr, values: Sequence) -> None:
    for i in range(3):
        arr[i] = values[i]


class PackBatch:
    """Collects layout jobs (clift_pack_job) and runs them as ONE clift_pack_batch launch: the batched form of
    clift_pack_plane / clift_unpack_plane / clift_pack_linear / clift_unpack_linear / clift_pack_linear_dgrad."""

    def __init__(self):
        self.jobs = []
        self.tiles = 0

    def _add(self, src, dst, s_pitch, s_rows, s_cols, d_rows, d_cols, kind):
        if d_rows <= 0 or d_cols <= 0:
            return
        j = PackJob()
        j.src, j.dst = ptr(src), ptr(dst)
        j.s_pitch, j.s_rows, j.s_cols, j.d_rows, j.d_cols, j.kind = s_pitch, s_rows, s_cols, d_rows, d_cols, kind
        j.first_tile = self.tiles
        self.tiles += ((d_rows + 31) // 32) * ((d_cols + 31) // 32)
        self.jobs.append(j)

    def plane(self, nchw, hwc, comps, h, w):
        self._add(nchw, hwc, h * w, comps, h * w, h * w, comps, 0)

    def unplane(self, hwc, nchw, comps, h, w):
        self._add(hwc, nchw, comps, h * w, comps, comps, h * w, 0)

    def linear(self, w, b, wt, bias_pad, n_out, n_in):
        self._add(w, wt, n_in, n_out, n_in, k_pad(n_in), n_pad(n_out), 0)
        self._add(b, bias_pad, n_out, 1, n_out, 1, n_pad(n_out), 1)

    def unlinear(self, wt, bias_pad, w, b, n_out, n_in):
        self._add(wt, w, n_pad(n_out), n_in, n_out, n_out, n_in, 0)
        if b is not None:
            self._add(bias_pad, b, n_out, 1, n_out, 1, n_out, 1)

    def dgrad(self, w, w_dgrad, n_out, n_in):
        self._add(w, w_dgrad, n_in, n_out, n_in, k_pad(n_out), dgrad_pad(n_in), 1)

    def transpose(self, w, w_t, n_out, n_in):
        """w [n_out][n_in] -> w_t [n_in][n_out] (the source of the tensor-core data-gradient operand)."""
        self._add(w, w_t, n_in, n_out, n_in, n_in, n_out, 0)

    def run(self, lib, device):
        if not self.jobs:
            return
        raw = bytes((PackJob * len(self.jobs))(*self.jobs))
        with on(device):
            table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device, non_blocking=True)
            check(lib.clift_pack_batch(ptr(table), len(self.jobs), int(self.tiles), stream_ptr(device)))
        self.jobs, self.tiles = [], 0
        return table          # a caller that records the launch for replay keeps the job table alive


class Tc16Batch:
    """Collects fp16-split operand jobs (clift_tc16_job) of a model and runs them as ONE clift_pack_linear_tc16_batch call
    (two launches).  ``chain``: layers whose input bound chains through the previous layer's header share a chain id."""

    def __init__(self):
        self.jobs = []
        self.blocks = 0
        self.chains = 0

    def add(self, w, bias, dst, n_out, n_in, in_bound_ptr, floor, chain):
        j = Tc16Job()
        j.w, j.bias, j.dst = ptr(w), ptr(bias), ptr(dst)
        j.in_bound = in_bound_ptr
        j.in_bound_floor = float(floor)
        j.n_out, j.n_in, j.chain = int(n_out), int(n_in), int(chain)
        j.first_block = self.blocks
        n_pad = (n_out + 31) // 32 * 32
        slabs = (n_in + 15) // 16 + (1 if bias is not None else 0)
        self.blocks += (slabs * 16 * n_pad + 255) // 256
        self.chains = max(self.chains, chain + 1)
        self.jobs.append(j)

    def run(self, lib, device):
        if not self.jobs:
            return
        raw = bytes((Tc16Job * len(self.jobs))(*self.jobs))
        with on(device):
            table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device, non_blocking=True)
            check(lib.clift_pack_linear_tc16_batch(ptr(table), len(self.jobs), int(self.chains), int(self.blocks), stream_ptr(device)))
        self.jobs, self.blocks, self.chains = [], 0, 0
        return table
