"""Contrastive-fusion losses on rendered embeddings, each one fused forward+gradient kernel launch.

    slow_fast_loss   trainer/train_panopli_tensorf.py:256-310 (calculate_instance_clustering_loss, slow_fast branch)
    ema_update       trainer/train_panopli_tensorf.py:325-329 (ema_update_slownet)
    contrastive_loss model/loss/loss.py:62-82
    TVLoss/plane_tv  model/loss/loss.py:9-26

All of them call libclift_b200.so; none has a PyTorch fallback.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.nn as nn

from . import lib as L


class _SlowFast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, labels, confidences):
        lib = L.load()
        f = features.detach().contiguous().float()
        n, w = f.shape
        loss = torch.empty((1,), device=f.device)
        grad = torch.empty_like(f)
        # converted copies stay bound until the launch is enqueued: a temporary released earlier could hand its block to the
        # next conversion and alias the two arguments
        lab = labels.to(f.device).contiguous().long()
        conf = confidences.to(f.device).contiguous().float()
        with L.on(f.device):
            L.check(lib.clift_slowfast_loss(L.ptr(f), L.ptr(lab), L.ptr(conf), n, w // 2, L.ptr(loss), L.ptr(grad),
                                            L.stream_ptr(f.device)))
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def slow_fast_loss(features: torch.Tensor, labels: torch.Tensor, confidences: torch.Tensor) -> torch.Tensor:
    """features [N, 2d] = [fast | slow] as rendered by forward_instance_feature; labels int64 [N]; confidences [N].
    The slow half is treated as detached (trainer:269); gradients reach the fast columns of the first N//2 rows."""
    if features.shape[-1] % 2:
        raise L.CliftError("slow-fast features must be [N, 2d]")
    return _SlowFast.apply(features, labels, confidences)


class _Contrastive(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, labels, temperature):
        lib = L.load()
        f = features.detach().contiguous().float()
        loss = torch.empty((1,), device=f.device)
        grad = torch.empty_like(f)
        lab = labels.to(f.device).contiguous().long()       # bound until the launch is enqueued
        with L.on(f.device):
            L.check(lib.clift_contrastive_loss(L.ptr(f), L.ptr(lab), f.shape[0], f.shape[1], float(temperature), L.ptr(loss),
                                               L.ptr(grad), L.stream_ptr(f.device)))
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def contrastive_loss(features: torch.Tensor, instance_labels: torch.Tensor, temperature: float) -> torch.Tensor:
    """Same signature as model/loss/loss.py:62."""
    return _Contrastive.apply(features, instance_labels, temperature)


@torch.no_grad()
def ema_update(slow_params: Iterable[torch.Tensor], fast_params: Iterable[torch.Tensor], momentum: float = 0.9) -> None:
    """param_k = param_k * momentum + (1 - momentum) * param_q for every pair (trainer:325-329)."""
    lib = L.load()
    for q, k in zip(fast_params, slow_params):
        with L.on(k.device):
            L.check(lib.clift_ema_update(L.ptr(k.data), L.ptr(q.data), k.numel(), float(momentum), L.stream_ptr(k.device)))
    L.bump_param_epoch()     # parameters were mutated through raw pointers: packed copies are stale


def ema_update_slownet(slow_net: nn.Module, fast_net: nn.Module, momentum: float = 0.9) -> None:
    """Module-level form with the reference's argument order (trainer:325)."""
    ema_update(slow_net.parameters(), fast_net.parameters(), momentum)


class _PlaneTV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plane):
        lib = L.load()
        p = plane.detach()
        _, c, h, w = p.shape
        st = L.stream_ptr(p.device)
        hwc = torch.empty((h, w, c), device=p.device)
        src = p.contiguous()                                 # bound until the launches are enqueued
        loss = torch.empty((1,), device=p.device)
        g_hwc = torch.zeros_like(hwc)
        g = torch.empty_like(src)
        with L.on(p.device):
            L.check(lib.clift_pack_plane(L.ptr(src), L.ptr(hwc), c, h, w, st))
            L.check(lib.clift_tv_loss(L.ptr(hwc), c, h, w, L.ptr(loss), L.ptr(g_hwc), 1.0, st))
            L.check(lib.clift_unpack_plane(L.ptr(g_hwc), L.ptr(g), c, h, w, st))
        ctx.save_for_backward(g)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        return g * gout


def plane_tv(plane: torch.Tensor) -> torch.Tensor:
    """TVLoss.forward (loss.py:13-22) of one (1,C,H,W) plane."""
    if plane.shape[0] != 1:
        raise L.CliftError("plane_tv expects a (1,C,H,W) factor plane")
    return _PlaneTV.apply(plane)


class TVLoss(nn.Module):
    """Drop-in for model/loss/loss.py:9-26."""

    def __init__(self, weight=1):
        super().__init__()
        self.weight = weight

    def forward(self, x):
        return self.weight * plane_tv(x)
