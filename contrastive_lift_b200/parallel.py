"""Data-parallel plumbing for the render path: one process per GPU, parameters replicated,
rays sharded, gradients sum-all-reduced once per backward over a persistent flat fp32 arena (the 1/N rides in the optimizer).

This reproduces what the reference gets from Lightning's DDPStrategy (trainer/__init__.py:95-108): each rank renders
its own rays, every ``manual_backward`` (trainer:198, :220) is followed by a mean of the gradients over ranks.  The
path has no other exchange step (rays are independent; the slow-fast loss keeps whole images on one rank, as DDP's
DistributedSampler does), so torch.distributed's NCCL all-reduce over NVLink is the only collective.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


class _NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


_NCCL_LIB = None
_COMMS = {}


def _nccl_lib():
    """ctypes handle of the libnccl this process already uses (PyTorch's bundled copy)."""
    global _NCCL_LIB
    if _NCCL_LIB is None:
        bundled = os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2")
        lib = C.CDLL(bundled if os.path.exists(bundled) else "libnccl.so.2")
        lib.ncclGetUniqueId.argtypes = [C.POINTER(_NcclUniqueId)]
        lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _NcclUniqueId, C.c_int]
        lib.ncclCommDestroy.argtypes = [C.c_void_p]
        lib.ncclGetErrorString.restype = C.c_char_p
        _NCCL_LIB = lib
    return _NCCL_LIB


def nccl_communicator(device: torch.device, group=None) -> int:
    """An ncclComm_t of this rank for ``clift_allreduce_grads`` (the C-ABI collective): torch.distributed does not hand out
    its own communicator, so one is created next to it - rank 0 draws the unique id, torch.distributed broadcasts it, every
    rank calls ncclCommInitRank.  Cached per (device, group); lives until ``destroy_nccl_communicators()``."""
    key = (torch.device(device), id(group))
    if key in _COMMS:
        return _COMMS[key]
    nccl = _nccl_lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = _NcclUniqueId()
    if rank == 0:
        rc = nccl.ncclGetUniqueId(C.byref(uid))
        if rc != 0:
            raise RuntimeError(f"ncclGetUniqueId: {nccl.ncclGetErrorString(rc).decode()}")
    raw = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=device)
    dist.broadcast(raw, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    uid = _NcclUniqueId.from_buffer_copy(bytes(raw.cpu().tolist()))
    comm = C.c_void_p()
    with torch.cuda.device(device):
        rc = nccl.ncclCommInitRank(C.byref(comm), world, uid, rank)
    if rc != 0:
        raise RuntimeError(f"ncclCommInitRank: {nccl.ncclGetErrorString(rc).decode()}")
    _COMMS[key] = comm.value
    return comm.value


def destroy_nccl_communicators() -> None:
    for comm in _COMMS.values():
        _nccl_lib().ncclCommDestroy(C.c_void_p(comm))
    _COMMS.clear()


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of ``n`` rays owned by ``rank`` (ray order is preserved across ranks)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_rays(rays: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    b, e = shard_range(rays.shape[0], rank, world)
    return rays[b:e]


def gather_rays_output(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Inverse of shard_rays for an output map [n_local, K]: all-gather with the shard sizes of shard_range().
    Shards differ by at most one ray; the collective itself runs on equal-size buffers (gloo refuses uneven all_gather
    lists, NCCL falls back to one broadcast per rank for them), the padding row is dropped afterwards."""
    world = dist.get_world_size(group)
    sizes = [e - b for b, e in (shard_range(n_total, r, world) for r in range(world))]
    rows = max(sizes) if sizes else 0
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError(f"gather_rays_output: this rank holds {local.shape[0]} rays, shard_range() assigns it "
                         f"{sizes[dist.get_rank(group)]} of {n_total}")
    send = local.contiguous()
    if send.shape[0] < rows:
        send = torch.cat([send, send.new_zeros((rows - send.shape[0],) + tuple(send.shape[1:]))], 0)
    parts = [send.new_empty(send.shape) for _ in range(world)]
    dist.all_gather(parts, send, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], 0)


def shard_rows_cyclic(rays: torch.Tensor, height: int, width: int, rank: Optional[int] = None,
                      world: Optional[int] = None) -> torch.Tensor:
    """Rows rank, rank + world, rank + 2 world, ... of an [H*W, K] frame (ray index = row * W + col).  A frame's cost is not
    uniform over its rows (the object sits in the middle), so contiguous ranges leave the edge ranks idle while the centre
    ranks work; row-cyclic shards cost the same on every rank.  ``gather_rows_cyclic`` restores the ray order."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return rays.view(height, width, -1)[rank::world].reshape(-1, rays.shape[-1])


def gather_rows_cyclic(local: torch.Tensor, height: int, width: int, group=None) -> torch.Tensor:
    """Inverse of shard_rows_cyclic for an output map [rows_local * W, K] -> [H * W, K] on every rank: ONE all-gather into a
    single buffer (ranks with one row less are padded), then a strided view puts row r * world + k back in place."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    rows = (height + world - 1) // world
    mine = len(range(rank, height, world))
    k = local.shape[-1] if local.dim() > 1 else 1
    if local.shape[0] != mine * width:
        raise ValueError(f"gather_rows_cyclic: this rank holds {local.shape[0]} rays, its cyclic shard has {mine * width}")
    send = local.reshape(mine, width, k)
    if mine < rows:
        send = torch.cat([send, send.new_zeros((rows - mine, width, k))], 0)
    out = send.new_empty((world * rows, width, k))                                # concatenated form (gloo accepts only this)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    full = out.view(world, rows, width, k).permute(1, 0, 2, 3).reshape(rows * world, width, k)[:height]   # row = local_row * world + rank
    full = full.reshape(height * width, k)
    return full if local.dim() > 1 else full[:, 0]


class GradientArena:
    """Persistent flat fp32 arena for the gradient all-reduce of one optimizer's parameters (DDP's bucket, kept across steps).
    On CUDA / NCCL the collective is the library's own C-ABI entry ``clift_allreduce_grads``, stream-ordered on the
    compute stream like every other launch of the path.

    ``reduce()`` gathers the live ``.grad``s into the arena with one multi-tensor copy, issues ONE ``all_reduce(SUM)`` on
    the arena (+ one presence flag per parameter, see below) and scatters the sums back with one multi-tensor copy; the 1/N
    of the mean is left to the optimizer (``FusedAdam(grad_scale=1/N)`` folds it into the update) unless ``average=True``.
    Nothing is allocated after the first call.  ``drain_ms()`` is the summed device time of the ``timed`` reduces since the last
    drain (CUDA events read after the fact, so timing never stalls the host inside a step).

    Presence flags: a parameter whose grad is None here (head not evaluated in this pass - the reason the reference needs
    ``find_unused_parameters=True``, trainer/__init__.py:95-108) contributes zeros and stays None.  DDP would hand every rank
    the reduced gradient if ANY rank had one; here the passes are the same on every rank by construction, so a differing
    None pattern is a bug: the summed flags are checked one call later (no sync on the hot path) and raise."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, c_abi: Optional[bool] = None):
        """``c_abi``: issue the collective through the library's own entry ``clift_allreduce_grads`` on a communicator created
        by ``nccl_communicator`` (default for CUDA parameters under the NCCL backend; CLIFT_ALLREDUCE_TORCH=1 or False: through
        ``torch.distributed.all_reduce``, which is also what CPU / gloo groups use)."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.c_abi = c_abi
        self._comm = None
        if not self.params:
            raise ValueError("GradientArena: no trainable parameters")
        dev = self.params[0].device
        self.sizes = [p.numel() for p in self.params]
        self.total = sum(self.sizes)
        self.flat = torch.zeros((self.total + len(self.params),), dtype=torch.float32, device=dev)
        self.slots = [s.view_as(p) for s, p in zip(self.flat[:self.total].split(self.sizes), self.params)]
        self.flags = self.flat[self.total:]
        self._events: List[tuple] = []
        self._pending_flags = None
        pin = (lambda x: x.pin_memory()) if dev.type == "cuda" else (lambda x: x)
        self._host_flags = pin(torch.zeros((len(self.params),), dtype=torch.float32))
        self._host_sums = pin(torch.zeros((len(self.params),), dtype=torch.float32))     # summed flags of the previous reduce

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def check_patterns(self) -> None:
        """Raise if, in the previous reduce, some ranks had a gradient for a parameter and others did not."""
        if self._pending_flags is None:
            return
        flags, world, ev = self._pending_flags
        self._pending_flags = None
        if ev is not None:
            ev.synchronize()          # the copy was queued a whole step ago: long done, no stall
        bad = [i for i, v in enumerate(flags.tolist()) if v not in (0.0, float(world))]
        if bad:
            raise RuntimeError(f"GradientArena: gradient None-pattern differs across ranks for parameter slots {bad[:8]} "
                               "(every rank must run the same passes; DDP's find_unused_parameters semantics are not emulated)")

    @torch.no_grad()
    def reduce(self, average: bool = False, timed: bool = False) -> int:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return 0
        self.check_patterns()
        world = dist.get_world_size(self.group)
        live = [(s, p.grad) for s, p in zip(self.slots, self.params) if p.grad is not None]
        dead = [s for s, p in zip(self.slots, self.params) if p.grad is None]
        ev = None
        if timed and self.flat.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if dead:
            torch._foreach_zero_(dead)
        if live:
            torch._foreach_copy_([s for s, _ in live], [g for _, g in live])
        self._host_flags.copy_(torch.tensor([0.0 if p.grad is None else 1.0 for p in self.params]))
        self.flags.copy_(self._host_flags, non_blocking=True)
        use_c = self.c_abi
        if use_c is None:
            use_c = self.flat.is_cuda and dist.get_backend(self.group) == "nccl" and os.environ.get("CLIFT_ALLREDUCE_TORCH") != "1"
        if use_c:
            from . import lib as L
            if self._comm is None:
                self._comm = nccl_communicator(self.flat.device, self.group)
            with L.on(self.flat.device):
                L.check(L.load().clift_allreduce_grads(self._comm, L.ptr(self.flat), self.flat.numel(), L.stream_ptr(self.flat.device)))
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if average:
            self.flat[:self.total].mul_(1.0 / world)
        if live:
            torch._foreach_copy_([g for _, g in live], [s for s, _ in live])
        if ev is not None:
            ev[1].record()
            self._events.append(ev)
        if self.flat.is_cuda:         # pinned destination + event: a pageable one would make the copy a host sync
            self._host_sums.copy_(self.flags, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            self._pending_flags = (self._host_sums, world, done)
        else:
            self._pending_flags = (self.flags.clone(), world, None)
        return self.total * 4

    def drain_ms(self) -> float:
        total = 0.0
        for a, b in self._events:
            b.synchronize()
            total += a.elapsed_time(b)
        self._events = []
        return total


_ARENAS = {}


@torch.no_grad()
def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = True) -> int:
    """Sum (then 1/world) every parameter's ``.grad`` across ranks through ONE flat buffer / one collective
    (``GradientArena``, cached per parameter set, so repeated calls allocate nothing).  Returns the bytes reduced."""
    plist: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
    if not plist or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    key = (tuple(id(p) for p in plist), id(group))
    arena = _ARENAS.get(key)
    if arena is None or any(a is not b for a, b in zip(arena.params, plist)):
        arena = _ARENAS[key] = GradientArena(plist, group)
    return arena.reduce(average=average)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Replicate rank ``src``'s parameters and buffers (what DDP does at construction)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
