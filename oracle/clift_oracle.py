"""CPU oracle for the volumetric-render + contrastive-fusion hot path.

TEST INFRASTRUCTURE ONLY.  This file is a functional, CPU-only restatement (torch fp32
ATen ops on plain tensors, no nn.Module) of the algorithm the reference implements in

    util/ray.py                                   (R1-R4 of SURVEY.md section 8a)
    model/renderer/panopli_tensoRF_renderer.py    (S1-S3, C1-C4)
    model/radiance_field/tensoRF.py               (F1-F3, H1-H3)
    model/loss/loss.py                            (L2, TV)
    trainer/train_panopli_tensorf.py:256-329      (L1, EMA)

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it; the product package ``contrastive_lift_b200`` never does.

Pinning: ``oracle/make_golden.py`` imports the real reference from /root/reference (with
stand-ins for its non-arithmetic imports), runs it and this restatement on the same seeded
inputs, asserts agreement and writes ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py``
re-checks this file against those vectors wherever the repo travels.

Third-party arithmetic: ``torch_efficient_distloss==0.1.3`` (reference requirements.txt:35,
one call site renderer:101) is not installable offline.  ``distortion_loss`` restates its
published O(S) prefix-sum form; it is pinned by the O(S^2) definition
(``distortion_loss_bruteforce``) and an fp64 gradcheck in tests - "parity unpinned" against
the package itself, as DESIGN.md says.

Parameters travel as a dict keyed by the reference's ``state_dict`` names.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# tensoRF.py:61-62 - plane i spans axes MATRIX_MODE[i] (W indexes the first, H the second),
# line i runs along axis VECTOR_MODE[i].
MATRIX_MODE = ((0, 1), (0, 2), (1, 2))
VECTOR_MODE = (2, 1, 0)


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class RenderConfig:
    """Constants of TensoRFRenderer (renderer:39-57) + the head layout of TensorVMSplit."""
    aabb: Tensor                       # (2,3)
    grid_dim: Tuple[int, int, int]
    step_ratio: float = 0.5
    distance_scale: float = 25.0
    weight_thres: float = 1e-4
    density_shift: float = -10.0
    semantic_softmax: bool = True      # semantic_weight_mode == "softmax"
    stop_semantic_grad: bool = True
    pe_view: int = 2
    pe_feat: int = 2
    pe_sem: int = 0
    pe_ins: int = 0
    slow_fast: bool = True
    # derived by step_geometry()
    inv_extent: Optional[Tensor] = None
    units: Optional[Tensor] = None
    step_size: Optional[Tensor] = None  # 0-d tensor like the reference's
    n_samples: int = 0

    def refresh(self) -> "RenderConfig":
        self.inv_extent, self.units, self.step_size, self.n_samples = step_geometry(
            self.aabb, self.grid_dim, self.step_ratio)
        return self


def step_geometry(aabb: Tensor, grid_dim: Sequence[int], step_ratio: float):
    """renderer:59-78 (update_step_size / update_step_ratio)."""
    extent = aabb[1] - aabb[0]
    g = torch.as_tensor(list(grid_dim), dtype=torch.long)
    inv_extent = 2.0 / extent
    units = extent / (g - 1 + 1e-3)
    step = torch.mean(units) * step_ratio
    diag = torch.sqrt(torch.sum(torch.square(extent)))
    n = int((diag / step).item()) + 1
    return inv_extent, units, step, n


def ratio_for_samples(aabb: Tensor, grid_dim: Sequence[int], n_samples: int) -> float:
    """the step_ratio that makes update_step_ratio() yield ``n_samples``."""
    extent = aabb[1] - aabb[0]
    g = torch.as_tensor(list(grid_dim), dtype=torch.long)
    units = extent / (g - 1 + 1e-3)
    diag = torch.sqrt(torch.sum(torch.square(extent)))
    return float(diag / ((n_samples - 0.5) * torch.mean(units)))


# --------------------------------------------------------------------------------------
# R1-R4: ray generation (util/ray.py:8-12,25-31,46-54,81-99; dataset/base.py:211-219)
# --------------------------------------------------------------------------------------
def pixel_directions(height: int, width: int, intrinsics: Tensor) -> Tensor:
    cols = torch.linspace(0, width - 1, width)
    rows = torch.linspace(0, height - 1, height)
    jj, ii = torch.meshgrid(rows, cols, indexing="ij")     # ii = column, jj = row, both (H,W)
    fx, fy, cx, cy = intrinsics[0, 0], intrinsics[1, 1], intrinsics[0, 2], intrinsics[1, 2]
    return torch.stack([(ii - cx) / fx, (jj - cy) / fy, torch.ones_like(ii)], -1)


def world_rays(directions: Tensor, cam2world: Tensor) -> Tuple[Tensor, Tensor]:
    d = directions @ cam2world[:3, :3].T
    d = d / torch.norm(d, dim=-1, keepdim=True)
    o = cam2world[:3, 3].expand(d.shape)
    return o.reshape(-1, 3), d.reshape(-1, 3)


def sphere_far(o: Tensor, d: Tensor, r: float = 1.0) -> Tensor:
    od = torch.sum(o * d, 1)
    dd = torch.sum(d ** 2, 1)
    oo = torch.sum(o ** 2, 1)
    det = od ** 2 + (r ** 2 - oo) * dd
    if not torch.all(det >= 0):
        raise AssertionError("camera outside the unit sphere")
    return (torch.sqrt(det) - od) / dd


def make_rays(height: int, width: int, intrinsics: Tensor, cam2world: Tensor, near: float = 0.01) -> Tensor:
    """[H*W, 8] = [o(3), d(3), near, far]; ray index = row*W + col."""
    o, d = world_rays(pixel_directions(height, width, intrinsics), cam2world)
    far = sphere_far(o, d)
    return torch.cat([o, d, torch.ones_like(far[:, None]) * near, far[:, None]], 1)


# --------------------------------------------------------------------------------------
# S1, S3: sampling (renderer:800-817, 633-634)
# --------------------------------------------------------------------------------------
def sample_points(rays: Tensor, aabb: Tensor, n_samples: int, step_size: Tensor,
                  jitter: Optional[Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
    """``jitter`` is the already-scaled per-ray offset ``perturb * U[0,1)`` of shape (B,1) or None."""
    o, d, near, far = rays[:, 0:3], rays[:, 3:6], rays[:, 6], rays[:, 7]
    vec = torch.where(d == 0, torch.full_like(d, 1e-6), d)
    ra = (aabb[1] - o) / vec
    rb = (aabb[0] - o) / vec
    t_min = torch.minimum(ra, rb).amax(-1).clamp(min=near, max=far)
    idx = torch.arange(n_samples)[None].float()
    if jitter is not None:
        idx = idx.repeat(rays.shape[0], 1) + jitter
    z = t_min[:, None] + step_size * idx
    pts = o[:, None, :] + d[:, None, :] * z[..., None]
    outside = ((aabb[0] > pts) | (pts > aabb[1])).any(dim=-1)
    return pts, z, ~outside


def normalize_points(pts: Tensor, aabb: Tensor, inv_extent: Tensor) -> Tensor:
    return (pts - aabb[0]) * inv_extent - 1


# --------------------------------------------------------------------------------------
# F1-F3: VM field lookup (tensoRF.py:108-156)
# --------------------------------------------------------------------------------------
def _plane_line_coords(xyz: Tensor) -> Tuple[Tensor, Tensor]:
    plane = torch.stack([xyz[..., list(m)] for m in MATRIX_MODE]).view(3, -1, 1, 2)
    line = torch.stack([xyz[..., v] for v in VECTOR_MODE])
    line = torch.stack((torch.zeros_like(line), line), dim=-1).view(3, -1, 1, 2)
    return plane, line


def vm_products(planes: Sequence[Tensor], lines: Sequence[Tensor], xyz: Tensor) -> List[Tensor]:
    """Per mode: (C, N) tensor of plane-tap x line-tap products (library grid_sample path)."""
    cp, cl = _plane_line_coords(xyz)
    out = []
    for i in range(3):
        p = F.grid_sample(planes[i], cp[[i]], align_corners=True).view(-1, xyz.shape[0])
        l = F.grid_sample(lines[i], cl[[i]], align_corners=True).view(-1, xyz.shape[0])
        out.append(p * l)
    return out


def bilinear_taps_explicit(plane: Tensor, x: Tensor, y: Tensor) -> Tensor:
    """Index-level restatement of grid_sample(bilinear, zeros padding, align_corners=True) on a
    (1,C,H,W) image: used to cross-check the library path and to document what the kernels do.
    Returns (C, N)."""
    _, C, H, W = plane.shape
    fx = ((x + 1) / 2) * (W - 1)
    fy = ((y + 1) / 2) * (H - 1)
    x0 = torch.floor(fx)
    y0 = torch.floor(fy)
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = (x1 - fx) * (y1 - fy)
    w_ne = (fx - x0) * (y1 - fy)
    w_sw = (x1 - fx) * (fy - y0)
    w_se = (fx - x0) * (fy - y0)
    img = plane[0]
    out = torch.zeros((C, x.shape[0]), dtype=plane.dtype)
    for xi, yi, w in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
        ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
        xc = xi.clamp(0, W - 1).long()
        yc = yi.clamp(0, H - 1).long()
        out = out + img[:, yc, xc] * (w * ok)[None]
    return out


def vm_products_explicit(planes, lines, xyz: Tensor) -> List[Tensor]:
    out = []
    for i in range(3):
        a, b = MATRIX_MODE[i]
        v = VECTOR_MODE[i]
        p = bilinear_taps_explicit(planes[i], xyz[:, a], xyz[:, b])
        l = bilinear_taps_explicit(lines[i], torch.zeros_like(xyz[:, v]), xyz[:, v])
        out.append(p * l)
    return out


def _factors(params: Params, name: str):
    return [params[f"{name}_plane.{i}"] for i in range(3)], [params[f"{name}_line.{i}"] for i in range(3)]


def density(params: Params, xyz: Tensor, shift: float = -10.0, explicit: bool = False) -> Tensor:
    planes, lines = _factors(params, "density")
    prods = (vm_products_explicit if explicit else vm_products)(planes, lines, xyz)
    feat = torch.zeros((xyz.shape[0],))
    for p in prods:
        feat = feat + torch.sum(p, dim=0)
    return F.softplus(feat + shift)


def vm_feature(params: Params, name: str, xyz: Tensor, explicit: bool = False) -> Tensor:
    planes, lines = _factors(params, name)
    prods = (vm_products_explicit if explicit else vm_products)(planes, lines, xyz)
    return F.linear(torch.cat(prods).T, params[f"{name}_basis_mat.weight"])


# --------------------------------------------------------------------------------------
# H1-H3: MLP heads (tensoRF.py:383-418, 462-511, 565-594)
# --------------------------------------------------------------------------------------
def pos_enc(x: Tensor, freqs: int) -> Tensor:
    bands = 2 ** torch.arange(freqs).float()
    p = (x[..., None] * bands).reshape(x.shape[:-1] + (freqs * x.shape[-1],))
    return torch.cat([torch.sin(p), torch.cos(p)], dim=-1)


def mlp_layers(params: Params, prefix: str) -> List[Tuple[Tensor, Tensor]]:
    out = []
    k = 0
    while f"{prefix}.{k}.weight" in params:
        out.append((params[f"{prefix}.{k}.weight"], params[f"{prefix}.{k}.bias"]))
        k += 2
    return out


def run_mlp(x: Tensor, layers: Sequence[Tuple[Tensor, Tensor]]) -> Tensor:
    for i, (w, b) in enumerate(layers):
        x = F.linear(x, w, b)
        if i + 1 < len(layers):
            x = torch.relu_(x)          # nn.ReLU(inplace=True), tensoRF.py:395,479,575: same values, no second activation buffer
    return x


def rgb_head(params: Params, cfg: RenderConfig, viewdirs: Tensor, feat: Tensor) -> Tensor:
    parts = [feat, viewdirs]
    if cfg.pe_feat > 0:
        parts.append(pos_enc(feat, cfg.pe_feat))
    if cfg.pe_view > 0:
        parts.append(pos_enc(viewdirs, cfg.pe_view))
    return torch.sigmoid(run_mlp(torch.cat(parts, -1), mlp_layers(params, "render_appearance_mlp.mlp")))


def _xyz_input(xyz: Tensor, pe: int) -> Tensor:
    return torch.cat([xyz, pos_enc(xyz, pe)], -1) if pe > 0 else xyz


def semantic_head(params: Params, cfg: RenderConfig, xyz: Tensor) -> Tensor:
    if "semantic_basis_mat.weight" in params:           # grid mode (allgrid.yaml)
        x = vm_feature(params, "semantic", xyz)
        x = _xyz_input(x, 0)
    else:
        x = _xyz_input(xyz, cfg.pe_sem)
    out = run_mlp(x, mlp_layers(params, "render_semantic_mlp.mlp"))
    return torch.softmax(out, -1) if cfg.semantic_softmax else out


def instance_head(params: Params, cfg: RenderConfig, xyz: Tensor) -> Tensor:
    if "instance_basis_mat.weight" in params:
        x = vm_feature(params, "instance", xyz)
    else:
        x = _xyz_input(xyz, cfg.pe_ins)
    out = run_mlp(x, mlp_layers(params, "render_instance_mlp.mlp"))
    if cfg.slow_fast:
        out = torch.cat([out, run_mlp(x, mlp_layers(params, "render_instance_mlp.slow_mlp"))], -1)
    return out


def head_dims(params: Params, cfg: RenderConfig) -> Tuple[int, int]:
    c = params[mlp_layers_last(params, "render_semantic_mlp.mlp")].shape[0]
    d = params[mlp_layers_last(params, "render_instance_mlp.mlp")].shape[0]
    return c, d * (2 if cfg.slow_fast else 1)


def mlp_layers_last(params: Params, prefix: str) -> str:
    k = 0
    while f"{prefix}.{k + 2}.weight" in params:
        k += 2
    return f"{prefix}.{k}.weight"


# --------------------------------------------------------------------------------------
# C1, C2: compositing weights and the distortion regulariser
# --------------------------------------------------------------------------------------
def raw_to_alpha(sigma: Tensor, dist: Tensor):
    """renderer:626-631."""
    alpha = 1.0 - torch.exp(-sigma * dist)
    trans = torch.cumprod(torch.cat([torch.ones(*alpha.shape[:-1], 1), 1.0 - alpha + 1e-10], -1), -1)
    return alpha, alpha * trans[..., :-1], trans[..., -1:]


def distortion_loss(w: Tensor, m: Tensor, interval: Tensor) -> Tensor:
    """mip-NeRF-360 distortion loss, O(S) form of torch_efficient_distloss.eff_distloss:
    mean over rays of  sum_i w_i^2 d_i / 3 + 2 sum_i w_i (m_i W_{<i} - WM_{<i})."""
    uni = (1.0 / 3.0) * (interval * w.pow(2)).sum(dim=-1).mean()
    wm = w * m
    w_cum = w.cumsum(dim=-1)
    wm_cum = wm.cumsum(dim=-1)
    bi = 2.0 * (wm[..., 1:] * w_cum[..., :-1] - w[..., 1:] * wm_cum[..., :-1]).sum(dim=-1).mean()
    return bi + uni


def distortion_loss_bruteforce(w: Tensor, m: Tensor, interval: Tensor) -> Tensor:
    """Definition: sum_ij w_i w_j |m_i - m_j| + 1/3 sum_i w_i^2 d_i, mean over rays."""
    pair = (w[..., :, None] * w[..., None, :] * (m[..., :, None] - m[..., None, :]).abs()).sum((-1, -2))
    return (pair + (w.pow(2) * interval).sum(-1) / 3.0).mean()


# --------------------------------------------------------------------------------------
# C3, C4: the three render entry points (renderer:80-176, 178-217, 259-300)
# --------------------------------------------------------------------------------------
def _march(params: Params, cfg: RenderConfig, rays: Tensor, jitter: Optional[Tensor]):
    pts, z, inbox = sample_points(rays, cfg.aabb, cfg.n_samples, cfg.step_size, jitter)
    dists = torch.cat((z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])), dim=-1)
    xyz = normalize_points(pts, cfg.aabb, cfg.inv_extent)
    sigma = torch.zeros(xyz.shape[:-1])
    if inbox.any():
        sigma[inbox] = density(params, xyz[inbox], cfg.density_shift)
    alpha, weight, _ = raw_to_alpha(sigma, dists * cfg.distance_scale)
    return xyz, z, inbox, dists, sigma, alpha, weight


def render_forward(params: Params, cfg: RenderConfig, rays: Tensor, jitter: Optional[Tensor] = None,
                   add_background: bool = False, detail: bool = False):
    """TensoRFRenderer.forward for one chunk.  ``jitter``: (B,1) = perturb*U or None (inference);
    ``add_background``: the outcome of ``white_bg or (is_train and rand<0.5)`` (renderer:164).
    Returns (rgb, semantics, instances, depth, feats(1,1), dist_reg) [+ a dict of intermediates]."""
    xyz, z, inbox, dists, sigma, alpha, weight = _march(params, cfg, rays, jitter)
    B, S = z.shape
    n_cls, n_ins = head_dims(params, cfg)
    mid = torch.cat(((z[:, 1:] + z[:, :-1]) / 2, z[:, -2:-1] * torch.ones_like(z[:, :1])), dim=-1)
    dist_reg = distortion_loss(weight, mid, dists)
    viewdirs = rays[:, 3:6].view(-1, 1, 3).expand(xyz.shape)
    rgb = torch.zeros((B, S, 3))
    sem = torch.zeros((B, S, n_cls))
    ins = torch.zeros((B, S, n_ins))
    active = weight > cfg.weight_thres
    if active.any():
        xa = xyz[active]
        rgb[active] = rgb_head(params, cfg, viewdirs[active], vm_feature(params, "appearance", xa))
        sem[active] = semantic_head(params, cfg, xa)
        ins[active] = instance_head(params, cfg, xa)
    opacity = torch.sum(weight, -1)
    rgb_map = torch.sum(weight[..., None] * rgb, -2)
    w = weight[..., None]
    if cfg.stop_semantic_grad:
        w = w.detach()
    sem_map = torch.sum(w * sem, -2)
    ins_map = torch.sum(w * ins, -2)
    if cfg.semantic_softmax:
        sem_map = sem_map / (sem_map.sum(-1).unsqueeze(-1) + 1e-8)
        sem_map = torch.log(sem_map + 1e-8)
    if add_background:
        rgb_map = rgb_map + (1.0 - opacity[..., None])
    rgb_map = rgb_map.clamp(0, 1)
    with torch.no_grad():
        depth = torch.sum(weight * z, -1)
    out = (rgb_map, sem_map, ins_map, depth, torch.zeros([1, 1]), dist_reg)
    if detail:
        return out, dict(xyz=xyz, z=z, inbox=inbox, sigma=sigma, alpha=alpha, weight=weight,
                         active=active, opacity=opacity)
    return out


def render_instance_feature(params: Params, cfg: RenderConfig, rays: Tensor, jitter: Optional[Tensor] = None):
    """renderer:178-217: density/weights are constants; only the instance head carries grad."""
    with torch.no_grad():
        xyz, z, inbox, dists, sigma, alpha, weight = _march(params, cfg, rays, jitter)
    B, S = z.shape
    _, n_ins = head_dims(params, cfg)
    ins = torch.zeros((B, S, n_ins))
    active = weight > cfg.weight_thres
    if active.any():
        ins[active] = instance_head(params, cfg, xyz[active])
    ins_map = torch.sum(weight[..., None] * ins, -2)
    with torch.no_grad():
        dist_map = torch.sum(weight * z, -1)
        pts = rays[..., 0:3] + dist_map[..., None] * rays[..., 3:6]
    return ins_map, pts


def render_segment_feature(params: Params, cfg: RenderConfig, rays: Tensor, jitter: Optional[Tensor] = None):
    """renderer:259-300."""
    with torch.no_grad():
        xyz, z, inbox, dists, sigma, alpha, weight = _march(params, cfg, rays, jitter)
    B, S = z.shape
    n_cls, _ = head_dims(params, cfg)
    seg = torch.zeros((B, S, n_cls))
    active = weight > cfg.weight_thres
    if active.any():
        seg[active] = semantic_head(params, cfg, xyz[active])
    seg_map = torch.sum(weight[..., None].detach() * seg, -2)
    if cfg.semantic_softmax:
        seg_map = seg_map / (seg_map.sum(-1).unsqueeze(-1) + 1e-8)
        seg_map = torch.log(seg_map + 1e-8)
    return seg_map


def render_chunked(params: Params, cfg: RenderConfig, rays: Tensor, chunk: int = 2048):
    """Inference plumbing of inference/render_panopli.py:114-120 (chunk loop + cat), no grad."""
    outs = [[], [], [], []]
    with torch.no_grad():
        for i in range(0, rays.shape[0], chunk):
            r = render_forward(params, cfg, rays[i:i + chunk], None, False)
            for k in range(4):
                outs[k].append(r[k])
    return tuple(torch.cat(o, 0) for o in outs)


def density_rgb_only(params: Params, cfg: RenderConfig, rays: Tensor):
    """BASELINE config 1: the density + RGB slice: sample, density, alpha,
    appearance feature, RGB MLP, compositing - called stage by stage like the reference's pieces."""
    with torch.no_grad():
        xyz, z, inbox, dists, sigma, alpha, weight = _march(params, cfg, rays, None)
        B, S = z.shape
        rgb = torch.zeros((B, S, 3))
        active = weight > cfg.weight_thres
        if active.any():
            viewdirs = rays[:, 3:6].view(-1, 1, 3).expand(xyz.shape)
            rgb[active] = rgb_head(params, cfg, viewdirs[active], vm_feature(params, "appearance", xyz[active]))
        rgb_map = torch.sum(weight[..., None] * rgb, -2).clamp(0, 1)
        depth = torch.sum(weight * z, -1)
    return rgb_map, depth


# --------------------------------------------------------------------------------------
# L1, L2: losses (trainer:256-310, 325-329; loss.py:62-82, 9-26)
# --------------------------------------------------------------------------------------
def ema_update(slow: Sequence[Tensor], fast: Sequence[Tensor], momentum: float = 0.9) -> None:
    with torch.no_grad():
        for q, k in zip(fast, slow):
            k.mul_(momentum).add_((1 - momentum) * q)


def slow_fast_loss(features: Tensor, labels: Tensor, confidences: Tensor) -> Tensor:
    """Slow-fast contrastive loss on rendered embeddings [N, 2d] = [fast | slow].
    The EMA of the slow net (trainer:258-259) is a separate call (ema_update)."""
    d = features.shape[-1] // 2
    fast, slow = features.split([d, d], dim=-1)
    slow = slow.detach()
    n = labels.shape[0]
    nf = n // 2
    lab_f, lab_s = labels[:nf], labels[nf:]
    uniq_f, uniq_s = torch.unique(lab_f), torch.unique(lab_s)
    if uniq_f.numel() == 0 or uniq_s.numel() == 0:
        return torch.tensor(0.0)
    cent = torch.stack([slow[nf:][lab_s == l].mean(dim=0) for l in uniq_s])
    both = uniq_f[torch.isin(uniq_f, uniq_s)]
    loss = 0
    for l in both:
        sel = lab_f == l
        dsq = torch.pow(fast[:nf][sel] - cent[uniq_s == l], 2).sum(dim=-1)
        loss = loss + -1.0 * (torch.exp(-dsq / 1.0) * confidences[:nf][sel]).mean()
    if both.shape[0] > 0:
        loss = loss / both.shape[0]
    same = lab_f.unsqueeze(1) == lab_s.unsqueeze(0)
    sim = torch.exp(-torch.cdist(fast[:nf], slow[nf:], p=2) / 1.0)
    logits = torch.exp(sim)
    prob = torch.mul(logits, same).sum(dim=-1) / logits.sum(dim=-1)
    loss = loss + -torch.log(torch.masked_select(prob, prob.ne(0))).mean()
    return loss


def contrastive_loss(features: Tensor, labels: Tensor, temperature: float) -> Tensor:
    n = features.size(0)
    same = labels.view(-1, 1).repeat(1, n).eq_(labels.clone())
    same = same.fill_diagonal_(0, wrap=False)
    dsq = torch.pow(features.unsqueeze(1) - features.unsqueeze(0), 2).sum(dim=-1)
    temp = torch.where(same == 1, torch.ones_like(dsq) * temperature, torch.ones_like(dsq))
    logits = torch.exp(torch.exp(-dsq / temp))
    prob = torch.mul(logits, same).sum(dim=-1) / logits.sum(dim=-1)
    return -torch.masked_select(prob, prob.ne(0)).log().sum() / n


def tv_loss(x: Tensor) -> Tensor:
    """loss.py:9-26 on a (1,C,H,W) plane."""
    n, c, h, w = x.shape
    cnt_h = c * (h - 1) * w + 1e-4
    cnt_w = c * h * (w - 1) + 1e-4
    tv_h = torch.pow(x[:, :, 1:, :] - x[:, :, :h - 1, :], 2).sum()
    tv_w = torch.pow(x[:, :, :, 1:] - x[:, :, :, :w - 1], 2).sum()
    return 2 * (tv_h / cnt_h + tv_w / cnt_w) / n


def total_tv_loss(params: Params, lambda_density: float = 0.1, lambda_appearance: float = 0.01,
                  lambda_semantics: float = 0.02, lambda_instances: float = 0.02) -> Tensor:
    """tensoRF.py:248-290 (past the late_semantic / instance_optimization epochs).  Semantic / instance factor sets
    exist only in grid-head mode; their lines are regularised too (tensoRF.py:264,271)."""
    td = sum(tv_loss(params[f"density_plane.{i}"]) * 1e-2 for i in range(3))
    ta = sum(tv_loss(params[f"appearance_plane.{i}"]) * 1e-2 for i in range(3))
    tot = td * lambda_density + ta * lambda_appearance
    for name, lam in (("semantic", lambda_semantics), ("instance", lambda_instances)):
        if f"{name}_plane.0" in params:
            tot = tot + lam * sum(tv_loss(params[f"{name}_plane.{i}"]) * 1e-2 + tv_loss(params[f"{name}_line.{i}"]) * 1e-3
                                  for i in range(3))
    return tot


def factor_sets(params: Params) -> List[str]:
    """Names of the VM factor sets present in a parameter dict (grid-mode heads add semantic / instance)."""
    return [n for n in ("density", "appearance", "semantic", "instance") if f"{n}_plane.0" in params]


# --------------------------------------------------------------------------------------
# dense-alpha sweep, bounding box, shrink plan, factor upsampling
# --------------------------------------------------------------------------------------
def dense_alpha(params: Params, cfg: RenderConfig) -> Tuple[Tensor, Tensor]:
    """renderer:717-729 + compute_alpha :744-748 -> (alpha [G0,G1,G2], dense_xyz [G0,G1,G2,3])."""
    g = cfg.grid_dim
    samples = torch.stack(torch.meshgrid(torch.linspace(0, 1, g[0]), torch.linspace(0, 1, g[1]), torch.linspace(0, 1, g[2]),
                                         indexing="ij"), -1)
    dense_xyz = cfg.aabb[0] * (1 - samples) + cfg.aabb[1] * samples
    xyz = normalize_points(dense_xyz.view(-1, 3), cfg.aabb, cfg.inv_extent)
    sigma = density(params, xyz, cfg.density_shift).reshape(dense_xyz.shape[:-1])
    alpha = 1 - torch.exp(-sigma * cfg.step_size)
    return alpha, dense_xyz


def alpha_bbox(alpha: Tensor, dense_xyz: Tensor, threshold: float = 0.0075):
    """renderer:669-681: clamp, 3x3x3 max-pool, threshold -> (xyz_min, xyz_max, n_valid); (None, None, 0) if empty."""
    a = F.max_pool3d(alpha.clamp(0, 1)[None, None], kernel_size=3, padding=1, stride=1)[0, 0]
    valid = a >= threshold
    pts = dense_xyz[valid]
    if pts.shape[0] == 0:
        return None, None, 0
    return pts.amin(0), pts.amax(0), int(valid.sum())


def shrink_plan(cfg: RenderConfig, xyz_min: Tensor, xyz_max: Tensor, fractional_lenience: float = 1.0):
    """renderer:683-706 -> (new_aabb (2,3), t_l, b_r) voxel index ranges handed to TensorVMSplit.shrink."""
    extent = xyz_max - xyz_min
    position = (xyz_min + xyz_max) / 2
    lo = torch.maximum(cfg.aabb[0], position - (extent * fractional_lenience) / 2)
    hi = torch.minimum(cfg.aabb[1], position + (extent * fractional_lenience) / 2)
    g = torch.as_tensor(list(cfg.grid_dim), dtype=torch.long)
    t_l, b_r = (lo - cfg.aabb[0]) / cfg.units, (hi - cfg.aabb[0]) / cfg.units
    t_l, b_r = torch.round(torch.round(t_l)).long(), torch.round(b_r).long() + 1
    b_r = torch.stack([b_r, g]).amin(0)
    return torch.stack((lo, hi)), t_l, b_r


def shrink_params(params: Params, t_l: Tensor, b_r: Tensor) -> Params:
    """tensoRF.py:158-177 on the parameter dict (every factor set present; heads untouched)."""
    out = dict(params)
    for name in factor_sets(params):
        for i in range(3):
            v = VECTOR_MODE[i]
            m0, m1 = MATRIX_MODE[i]
            out[f"{name}_line.{i}"] = params[f"{name}_line.{i}"][..., int(t_l[v]):int(b_r[v]), :]
            out[f"{name}_plane.{i}"] = params[f"{name}_plane.{i}"][..., int(t_l[m1]):int(b_r[m1]), int(t_l[m0]):int(b_r[m0])]
    return out


def upsample_params(params: Params, res_target: Sequence[int]) -> Params:
    """tensoRF.py:179-197: bilinear align_corners resize of every plane / line factor."""
    out = dict(params)
    for name in factor_sets(params):
        for i in range(3):
            v = VECTOR_MODE[i]
            m0, m1 = MATRIX_MODE[i]
            out[f"{name}_plane.{i}"] = F.interpolate(params[f"{name}_plane.{i}"], size=(res_target[m1], res_target[m0]),
                                                     mode="bilinear", align_corners=True)
            out[f"{name}_line.{i}"] = F.interpolate(params[f"{name}_line.{i}"], size=(res_target[v], 1), mode="bilinear",
                                                    align_corners=True)
    return out


def target_resolution(aabb: Tensor, n_voxels: int) -> Tuple[int, ...]:
    """renderer:756-761."""
    voxel_size = ((aabb[1] - aabb[0]).prod() / n_voxels).pow(1 / 3)
    return tuple(max(x, 1) for x in ((aabb[1] - aabb[0]) / voxel_size).long().tolist())


# --------------------------------------------------------------------------------------
# Adam (torch.optim.Adam as built by trainer/__init__.py:134-139); rank 4: nearest centroid
# --------------------------------------------------------------------------------------
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
              weight_decay: float = 0.0) -> None:
    """One in-place step of torch's single-tensor Adam (amsgrad=False), restated; pinned against torch.optim.Adam."""
    b1, b2 = betas
    if weight_decay != 0:
        g = g.add(p, alpha=weight_decay)
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def nearest_centroid(features: Tensor, centroids: Tensor) -> Tensor:
    """inference/render_panopli.py:389-396: argmin over torch.cdist (p=2)."""
    return torch.argmin(torch.cdist(features, centroids), dim=-1)


def psnr(x: Tensor, y: Tensor) -> Tensor:
    """util/metrics.py:25-26."""
    return -10 * torch.log10(torch.mean((x - y) ** 2))
