"""Synthetic scenes, cameras and ray batches.

No real dataset or checkpoint exists offline, so every test, the smoke run and the bench use
a seeded synthetic TensoRF field: factors ~ 0.1*N(0,1) (tensoRF.py:99-106), MLP weights with
nn.Linear's default uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) range, and a *solid ball* written
into density component 0 so that rays terminate (opacity ~0.985, ~12 % active samples); a raw
random field has sigma = softplus(-10) everywhere and no active sample at all.

Parameters are a plain dict keyed by the reference's ``state_dict`` names so the same dict
feeds the reference (load_state_dict), the oracle, and this package's ``TensorVMSplit``.
numpy's PCG64 stream is used (not torch's) so fixtures only need to store a seed.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

MATRIX_MODE = ((0, 1), (0, 2), (1, 2))
VECTOR_MODE = (2, 1, 0)


def _linear(rng: np.random.Generator, n_out: int, n_in: int, bias: bool = True, zero_bias: bool = False):
    bound = 1.0 / math.sqrt(n_in)
    w = rng.uniform(-bound, bound, size=(n_out, n_in)).astype(np.float32)
    if not bias:
        return w, None
    b = np.zeros((n_out,), np.float32) if zero_bias else rng.uniform(-bound, bound, size=(n_out,)).astype(np.float32)
    return w, b


def _mlp(rng, out: Dict[str, torch.Tensor], prefix: str, dims: Sequence[int], zero_last_bias: bool = False):
    for li in range(len(dims) - 1):
        w, b = _linear(rng, dims[li + 1], dims[li], zero_bias=zero_last_bias and li == len(dims) - 2)
        out[f"{prefix}.{2 * li}.weight"] = torch.from_numpy(w)
        out[f"{prefix}.{2 * li}.bias"] = torch.from_numpy(b)


def make_field_params(seed: int, grid_dim: Sequence[int], num_classes: int = 21, max_instances: int = 3,
                      slow_fast: bool = True, density_comps: int = 16, appearance_comps: int = 48,
                      dim_appearance: int = 27, pe_view: int = 2, pe_feat: int = 2, pe_sem: int = 0,
                      pe_ins: int = 0, width_rgb: int = 128, width_sem: int = 256, width_ins: int = 256,
                      ball: Optional[float] = 0.35, ball_gain: float = 3.0,
                      factor_scale: float = 0.1, sem_grid_comps: Optional[int] = None,
                      ins_grid_comps: Optional[int] = None, dim_semantics: int = 27, dim_instances: int = 27,
                      width_sem_grid: int = 128) -> Dict[str, torch.Tensor]:
    """Parameter dict of the default contrastive_lift configuration (MLP heads on xyz).  ``sem_grid_comps`` /
    ``ins_grid_comps`` (e.g. 32) switch that head to grid mode (allgrid.yaml, tensoRF.py:72-84): a VM factor set +
    bias-free basis Linear feeding a 3-layer MLP; their draws come after the default ones, so the default stream is
    unchanged."""
    rng = np.random.default_rng(seed)
    g = list(grid_dim)
    p: Dict[str, torch.Tensor] = {}
    for name, comps in (("density", density_comps), ("appearance", appearance_comps)):
        for i in range(3):
            a, b = MATRIX_MODE[i]
            v = VECTOR_MODE[i]
            plane = (factor_scale * rng.standard_normal((1, comps, g[b], g[a]))).astype(np.float32)
            line = (factor_scale * rng.standard_normal((1, comps, g[v], 1))).astype(np.float32)
            if name == "density" and ball is not None:
                ga = np.exp(-(np.linspace(-1, 1, g[a]) / ball) ** 2)
                gb = np.exp(-(np.linspace(-1, 1, g[b]) / ball) ** 2)
                gv = np.exp(-(np.linspace(-1, 1, g[v]) / ball) ** 2)
                plane[0, 0] = (ball_gain * np.outer(gb, ga)).astype(np.float32)
                line[0, 0, :, 0] = (ball_gain * gv).astype(np.float32)
            p[f"{name}_plane.{i}"] = torch.from_numpy(plane)
            p[f"{name}_line.{i}"] = torch.from_numpy(line)
    w, _ = _linear(rng, dim_appearance, 3 * appearance_comps, bias=False)
    p["appearance_basis_mat.weight"] = torch.from_numpy(w)
    in_rgb = dim_appearance + 3 + 2 * pe_feat * dim_appearance + 2 * pe_view * 3
    _mlp(rng, p, "render_appearance_mlp.mlp", [in_rgb, width_rgb, width_rgb, 3], zero_last_bias=True)
    in_ins = 3 + 2 * pe_ins * 3
    ins_dims = [in_ins, width_ins, width_ins, width_ins, max_instances]
    if ins_grid_comps:       # tensoRF.py:75-77: 3 layers on the 27 basis features
        ins_dims = [dim_instances, width_ins, width_ins, max_instances]
    _mlp(rng, p, "render_instance_mlp.mlp", ins_dims)
    if slow_fast:
        _mlp(rng, p, "render_instance_mlp.slow_mlp", ins_dims)
    in_sem = 3 + 2 * pe_sem * 3
    sem_dims = [in_sem, width_sem, width_sem, width_sem, width_sem, num_classes]
    if sem_grid_comps:       # tensoRF.py:81-85: 3 layers of dim_mlp_semantics
        sem_dims = [dim_semantics, width_sem_grid, width_sem_grid, num_classes]
    _mlp(rng, p, "render_semantic_mlp.mlp", sem_dims)
    for name, comps, dim in (("semantic", sem_grid_comps, dim_semantics), ("instance", ins_grid_comps, dim_instances)):
        if not comps:
            continue
        for i in range(3):
            a, b = MATRIX_MODE[i]
            v = VECTOR_MODE[i]
            p[f"{name}_plane.{i}"] = torch.from_numpy((factor_scale * rng.standard_normal((1, comps, g[b], g[a]))).astype(np.float32))
            p[f"{name}_line.{i}"] = torch.from_numpy((factor_scale * rng.standard_normal((1, comps, g[v], 1))).astype(np.float32))
        w, _ = _linear(rng, dim, 3 * comps, bias=False)
        p[f"{name}_basis_mat.weight"] = torch.from_numpy(w)
    return p


def params_checksum(params: Dict[str, torch.Tensor]) -> float:
    """Order-independent fp64 checksum used by fixtures to detect RNG-stream drift."""
    tot = 0.0
    for k in sorted(params):
        v = params[k].double()
        tot += float(v.sum()) + 0.5 * float((v * v).sum())
    return tot


def ratio_for_samples(aabb: torch.Tensor, grid_dim: Sequence[int], n_samples: int) -> float:
    """the ``step_ratio`` that makes ``TensoRFRenderer.update_step_ratio`` (renderer:59-78) yield ``n_samples``
    samples per ray: n = int(|extent| / (mean(extent / (G - 1 + 1e-3)) * ratio)) + 1."""
    extent = aabb[1] - aabb[0]
    g = torch.as_tensor(list(grid_dim), dtype=torch.long)
    units = extent / (g - 1 + 1e-3)
    diag = torch.sqrt(torch.sum(torch.square(extent)))
    return float(diag / ((n_samples - 0.5) * torch.mean(units)))


def default_aabb() -> torch.Tensor:
    return torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], dtype=torch.float32)


def camera(height: int, width: int, focal_scale: float = 0.8,
           position: Tuple[float, float, float] = (0.0, 0.0, -0.6),
           yaw_deg: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pinhole K (fx=fy=focal_scale*W, centre principal point) and a cam2world looking down +z."""
    k = torch.tensor([[focal_scale * width, 0.0, width / 2.0],
                      [0.0, focal_scale * width, height / 2.0],
                      [0.0, 0.0, 1.0]], dtype=torch.float32)
    c2w = torch.eye(4, dtype=torch.float32)
    if yaw_deg != 0.0:
        a = math.radians(yaw_deg)
        c2w[0, 0], c2w[0, 2], c2w[2, 0], c2w[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
    c2w[:3, 3] = torch.tensor(position, dtype=torch.float32)
    return k, c2w


def random_rays(seed: int, n: int, radius: float = 0.9, near: float = 0.01) -> torch.Tensor:
    """Training-style shuffled rays: origins inside the unit sphere, directions towards the
    scene centre region; includes axis-aligned directions (exact zeros) to exercise renderer:802."""
    rng = np.random.default_rng(seed)
    o = rng.standard_normal((n, 3))
    o = o / np.linalg.norm(o, axis=1, keepdims=True) * rng.uniform(0.3, radius, size=(n, 1))
    tgt = rng.uniform(-0.4, 0.4, size=(n, 3))
    d = tgt - o
    k = max(1, n // 16)
    d[:k] = 0.0
    d[np.arange(k), rng.integers(0, 3, size=k)] = -np.sign(o[np.arange(k), rng.integers(0, 3, size=k)] + 1e-9)
    d[:k] += (np.abs(d[:k]).sum(1, keepdims=True) == 0) * np.array([[0.0, 0.0, 1.0]])
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    o32, d32 = o.astype(np.float32), d.astype(np.float32)
    od = (o32 * d32).sum(1)
    dd = (d32 * d32).sum(1)
    oo = (o32 * o32).sum(1)
    far = ((np.sqrt(od * od + (1.0 - oo) * dd) - od) / dd).astype(np.float32)
    rays = np.concatenate([o32, d32, np.full((n, 1), near, np.float32), far[:, None]], 1)
    return torch.from_numpy(rays)
