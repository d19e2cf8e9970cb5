"""Fused Adam over a parameter group.

Drop-in for the ``torch.optim.Adam`` the reference builds in ``get_optimizer_and_scheduler``
(trainer/__init__.py:134-139; groups from tensoRF.py:199-246, betas trainer:98-103): same constructor
arguments, same ``param_groups`` keys and the same per-parameter state (``step``, ``exp_avg``, ``exp_avg_sq``),
so ``state_dict()`` / ``load_state_dict()`` and LR schedulers are interchangeable with the stock optimizer
and checkpoints keep their format.  ``step()`` issues ONE ``clift_adam_step_groups`` launch for all param groups
(up to 8 groups / step histories per launch) instead of a dozen ATen kernels per tensor.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import lib as L


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *, maximize=False,
                 grad_scale: float = 1.0):
        if amsgrad or maximize:
            raise L.CliftError("FusedAdam: amsgrad / maximize are not built (the reference uses neither)")
        if betas is None:
            betas = (0.9, 0.999)
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self.grad_scale = float(grad_scale)   # e.g. 1/world_size after a SUM all-reduce of the gradient arena

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        keep: List[torch.Tensor] = []      # contiguous gradient copies stay alive until every launch of this step is enqueued
        # segments of the launch under construction: (group, step count, first entry, number of entries); a segment ends when
        # the param group, the step history or the device changes.  Entries are (param, grad, exp_avg, exp_avg_sq, n) rows
        self._entries: list = []
        self._segments: list = []
        self._max_n, self._device = 0, None
        for group in self.param_groups:
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            for p in live:
                if not p.is_cuda:
                    raise L.CliftError("FusedAdam: parameters must live on a CUDA device (no CPU path)")
                if p.grad.is_sparse or p.dtype != torch.float32 or not p.is_contiguous():
                    raise L.CliftError("FusedAdam: dense contiguous fp32 parameters only")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            states = [self.state[p] for p in live]
            steps = [st["step"] for st in states]
            if all(s.device.type == "cpu" for s in steps):
                torch._foreach_add_(steps, 1.0)            # one call for the group's step counters (host tensors, as torch's)
            else:
                for s in steps:
                    s += 1
            entries: list = []
            step = None
            device = None
            max_n = 0
            for p, st, s in zip(live, states, steps):
                t = int(s)
                if step is None:
                    step, device = t, p.device
                elif t != step or p.device != device:      # mixed histories: flush what we have, start a new launch
                    self._launch(lib, entries, max_n, group, step, device)
                    entries, max_n, step, device = [], 0, t, p.device
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                keep.append(g)
                n = p.numel()
                entries.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), n))
                if n > max_n:
                    max_n = n
            self._launch(lib, entries, max_n, group, step, device)
        self._flush(lib)
        # parameters were rewritten through raw pointers: autograd's version counters did not move, so cached packed copies
        # (PackedField.refresh) must be told (ema_update does the same)
        L.bump_param_epoch()
        return loss

    def _launch(self, lib, entries, max_n, group, step, device):
        """Queue one (param group, step count) segment; the launch goes out when the table is full, the device changes, or
        at the end of step()."""
        if not entries:
            return
        if self._segments and (device != self._device or len(self._segments) == L.MAX_ADAM_GROUPS):
            self._flush(lib)
        self._device = device
        self._segments.append((group, int(step), len(self._entries), len(entries)))
        self._entries += entries
        self._max_n = max(self._max_n, int(max_n))

    def _flush(self, lib):
        if not self._segments:
            return
        entries, device = self._entries, self._device
        # the device table of a launch is reused while every pointer in it repeats (parameters and optimizer state never
        # move; the gradients of a pass come back at the same addresses once the caching allocator has settled)
        key = tuple(entries)
        cache = self.__dict__.setdefault("_tables", {})
        table = cache.get(key)
        if table is None:
            if len(cache) >= 8:
                cache.clear()
            rows = (L.AdamTensor * len(entries))()
            for r, e in zip(rows, entries):
                r.param, r.grad, r.exp_avg, r.exp_avg_sq, r.n = e
            with L.on(device):
                table = cache[key] = torch.frombuffer(bytearray(bytes(rows)), dtype=torch.uint8).to(device, non_blocking=True)
        groups = (L.AdamGroup * len(self._segments))()
        for g, (group, step, first, count) in zip(groups, self._segments):
            b1, b2 = group["betas"]
            g.lr, g.beta1, g.beta2, g.eps, g.weight_decay = float(group["lr"]), float(b1), float(b2), float(group["eps"]), \
                float(group["weight_decay"])
            g.first, g.count, g.step = first, count, step
        with L.on(device):
            L.check(lib.clift_adam_step_groups(L.ptr(table), len(entries), self._max_n, groups, len(self._segments),
                                               self.grad_scale, L.stream_ptr(device)))
        self._entries, self._segments, self._max_n = [], [], 0
