"""Rebuild the inputs of a tests/golden/render_*.npz fixture from its seeds."""
import os

import numpy as np
import torch

from contrastive_lift_b200 import synthetic as syn
from oracle import clift_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MLP_CASES = ("render_a", "render_b", "render_c")
GRID_CASES = ("render_d", "render_e", "render_f")      # grid-mode semantic / instance heads (allgrid.yaml family)
RENDER_CASES = MLP_CASES + GRID_CASES


def grid_comps(fx):
    """(semantic grid comps, instance grid comps) of a fixture; None = that head is an MLP on xyz."""
    g = lambda k: (int(fx[k]) or None) if k in fx.files else None
    return g("sem_grid"), g("ins_grid")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def render_inputs(fx):
    grid = tuple(int(v) for v in fx["grid"])
    sem_grid, ins_grid = grid_comps(fx)
    params = syn.make_field_params(int(fx["seed"]), grid, int(fx["n_cls"]), int(fx["n_ins"]),
                                   slow_fast=bool(fx["slow_fast"]), sem_grid_comps=sem_grid, ins_grid_comps=ins_grid)
    chk = syn.params_checksum(params)
    assert abs(chk - float(fx["params_checksum"])) <= 1e-9 * abs(chk), "numpy RNG stream drifted; regenerate fixtures"
    aabb = torch.from_numpy(fx["aabb"])
    cfg = orc.RenderConfig(aabb=aabb, grid_dim=grid, step_ratio=float(fx["step_ratio"]),
                           semantic_softmax=bool(fx["softmax"]), slow_fast=bool(fx["slow_fast"])).refresh()
    assert cfg.n_samples == int(fx["n_samples"])
    rays = torch.from_numpy(fx["rays"])
    return params, cfg, rays


def train_loss(out, fx, tag, keep=None):
    """The scalar the fixtures' gradients were taken of (oracle/make_golden.py: loss_fn).  ``keep`` (bool per ray)
    drops the per-ray terms of the other rays (the distortion term, which has no activity threshold, stays whole)."""
    tgt = torch.from_numpy(fx[f"{tag}_tgt_rgb"]).to(out[0].device)
    probs = torch.from_numpy(fx[f"{tag}_probs"]).to(out[0].device)
    w_ins = torch.from_numpy(fx[f"{tag}_w_ins"]).to(out[0].device)
    k = torch.ones(out[0].shape[0], device=out[0].device) if keep is None else keep.to(out[0].device).float()
    ce = (-(probs * torch.log_softmax(out[1], -1)).sum(-1) * k).mean()
    return ((((out[0] - tgt) ** 2) * k[:, None]).mean() + 0.37 * out[5] + 0.1 * ce
            + 0.05 * ((out[2] * w_ins).sum(-1) * k).mean())


def grad_digest(g):
    f = g.detach().double().flatten().cpu()
    head = torch.stack([f.sum(), f.abs().sum(), (f * f).sum()])
    return np.concatenate([head.numpy(), f[::97][:512].numpy()])
