"""Contrastive-fusion losses on rendered embeddings, each one fused forward+gradient kernel launch.

    slow_fast_loss   trainer/train_panopli_tensorf.py:256-310 (calculate_instance_clustering_loss, slow_fast branch)
    ema_update       trainer/train_panopli_tensorf.py:325-329 (ema_update_slownet)
    contrastive_loss model/loss/loss.py:62-82
    TVLoss/plane_tv  model/loss/loss.py:9-26

All of them call libclift_b200.so; none has a PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
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
    pairs = [(k.data, q.data) for q, k in zip(fast_params, slow_params)]
    if not pairs:
        return
    dev = pairs[0][0].device
    if any(k.device != dev or q.device != dev for k, q in pairs):
        raise L.CliftError("ema_update: slow and fast parameters must live on one CUDA device")
    # one launch for all pairs; the device table is reused while the parameters keep their storage
    key = tuple((L.ptr(k), L.ptr(q), k.numel()) for k, q in pairs)
    cached = _EMA_TABLES.get(dev)
    if cached is None or cached[0] != key:
        rows = (L.EmaPair * len(key))()
        for r, (pk, pq, n) in zip(rows, key):
            r.slow, r.fast, r.n = pk, pq, n
        table = torch.frombuffer(bytearray(bytes(rows)), dtype=torch.uint8).to(dev, non_blocking=True)
        cached = _EMA_TABLES[dev] = (key, table, max(n for _, _, n in key))
    with L.on(dev):
        L.check(lib.clift_ema_update_batch(L.ptr(cached[1]), len(key), cached[2], float(momentum), L.stream_ptr(dev)))
    L.bump_param_epoch()     # parameters were mutated through raw pointers: packed copies are stale


_EMA_TABLES: dict = {}
_TV_TABLES: dict = {}


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


class _TotalTV(torch.autograd.Function):
    """sum_i coef_i * TV(plane_i) over any number of (1,C,H,W) factor planes / (1,C,L,1) lines with a handful of launches:
    one batched layout job in, one value + one gradient launch per plane (the gradient already scaled by coef_i), one
    batched layout job out; backward is one multi-tensor scale by the upstream scalar."""

    @staticmethod
    def forward(ctx, coefs, *planes):
        lib = L.load()
        dev = planes[0].device
        with L.on(dev):
            st = L.stream_ptr(dev)
            srcs = [p.detach().contiguous() for p in planes]               # bound until the launches are enqueued
            sizes = [p.numel() for p in srcs]
            total = sum(sizes)
            flat = torch.empty((total,), device=dev)                        # packed copies (channel-last)
            g_flat = torch.zeros((total,), device=dev)                      # their gradients (channel-last)
            out_flat = torch.empty((total,), device=dev)                    # ... back in (1,C,H,W) layout, one view per plane
            vals = torch.empty((len(srcs),), device=dev)
            # the three job tables depend only on addresses and shapes: reused while the parameters keep their storage and the
            # caching allocator returns the scratch buffers at the same places (the usual case from the second step on)
            key = (tuple((p.data_ptr(), tuple(p.shape), float(c)) for p, c in zip(srcs, coefs)),
                   flat.data_ptr(), g_flat.data_ptr(), out_flat.data_ptr(), vals.data_ptr())
            tables = _TV_TABLES.get(dev)
            if tables is None or tables[0] != key:
                pack, unpack = L.PackBatch(), L.PackBatch()
                jobs = (L.TvJob * len(srcs))()
                off = 0
                for i, (j, p, n, coef) in enumerate(zip(jobs, srcs, sizes, coefs)):
                    _, c, h, w = p.shape
                    hwc, g_hwc = flat[off:off + n].view(h, w, c), g_flat[off:off + n].view(h, w, c)
                    pack.plane(p, hwc, c, h, w)
                    unpack.unplane(g_hwc, out_flat[off:off + n].view(1, c, h, w), c, h, w)
                    j.plane_hwc, j.loss, j.grad_hwc = L.ptr(hwc), vals.data_ptr() + 4 * i, L.ptr(g_hwc)
                    j.comps, j.h, j.w, j.grad_scale = c, h, w, float(coef)
                    off += n
                up = lambda raw: torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev, non_blocking=True)
                tables = _TV_TABLES[dev] = (key, up(bytes((L.PackJob * len(pack.jobs))(*pack.jobs))), int(pack.tiles),
                                            up(bytes(jobs)), up(bytes((L.PackJob * len(unpack.jobs))(*unpack.jobs))),
                                            int(unpack.tiles), up(bytes((C.c_float * len(srcs))(*[float(c) for c in coefs])))
                                            .view(torch.float32))
            _, t_pack, tiles_pack, t_tv, t_unpack, tiles_unpack, coef_dev = tables
            L.check(lib.clift_pack_batch(L.ptr(t_pack), len(srcs), tiles_pack, st))
            L.check(lib.clift_tv_loss_batch(L.ptr(t_tv), len(srcs), max(sizes), st))     # all values, then all gradients
            L.check(lib.clift_pack_batch(L.ptr(t_unpack), len(srcs), tiles_unpack, st))
            total_loss = (vals * coef_dev).sum()
            grads, off = [], 0
            for p, n in zip(srcs, sizes):
                grads.append(out_flat[off:off + n].view(p.shape))
                off += n
        ctx.grads = grads
        return total_loss

    @staticmethod
    def backward(ctx, gout):
        grads = ctx.grads
        ctx.grads = None
        return (None,) + tuple(torch._foreach_mul(grads, gout))


def total_tv(planes_and_coefs) -> torch.Tensor:
    """sum of coef * TVLoss(plane) over [(plane, coef), ...] (the terms of tensoRF.py:248-290 in one autograd node)."""
    planes = [p for p, _ in planes_and_coefs]
    for p in planes:
        if p.dim() != 4 or p.shape[0] != 1:
            raise L.CliftError("total_tv expects (1,C,H,W) factor planes / (1,C,L,1) lines")
    return _TotalTV.apply([c for _, c in planes_and_coefs], *planes)


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
