"""R1-R4: per-frame ray generation on the device (reference util/ray.py:8-12,25-31,46-54,81-99 and the
packing at dataset/base.py:211-219).  One kernel launch per frame; the reference does this on the CPU in
the dataset worker and ships 32 B/ray over PCIe (inference/render_panopli.py:110)."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from . import lib as L


def get_rays(height: int, width: int, intrinsics, cam2world, near: float = 0.01, radius: float = 1.0,
             device="cuda") -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (rays [H*W, 8] = [o(3), d(3), near, far] on ``device``, ray index = row*W + col; bad int32 [1] device flag).
    ``bad`` != 0 when a ray's origin lies outside the scene sphere (the reference asserts, ray.py:96-98); reading it is a
    device sync, so it is left to the caller - ``get_rays_checked`` reads it and raises AssertionError like the reference."""
    lib = L.load()
    dev = torch.device(device)
    k = np.ascontiguousarray(np.asarray(intrinsics, dtype=np.float32).reshape(3, 3))
    c2w = np.ascontiguousarray(np.asarray(cam2world, dtype=np.float32).reshape(4, 4))
    rays = torch.empty((height * width, 8), device=dev)
    bad = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.clift_gen_rays(k.ctypes.data_as(C.POINTER(C.c_float)), c2w.ctypes.data_as(C.POINTER(C.c_float)),
                                   height, width, float(near), float(radius), L.ptr(rays), L.ptr(bad), L.stream_ptr(dev)))
    return rays, bad


def get_rays_checked(height, width, intrinsics, cam2world, near=0.01, radius=1.0, device="cuda") -> torch.Tensor:
    rays, bad = get_rays(height, width, intrinsics, cam2world, near, radius, device)
    if int(bad.item()) != 0:            # one 4-byte D2H; use get_rays() to defer it
        raise AssertionError("rays_intersect_sphere: camera outside the scene sphere (negative determinant)")
    return rays
