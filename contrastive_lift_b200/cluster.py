"""Nearest-centroid assignment of rendered embeddings.

``nearest_centroid`` replaces the ``torch.cdist`` + ``argmin`` pair of inference/render_panopli.py:389-396 (the reference's
own ``assign_clusters`` can keep its bookkeeping and call it); ``assign_clusters`` is the whole of render_panopli.py:371-419
as device kernels behind the same call signature (``clift_assign_clusters`` / ``clift_labels_onehot``).  MeanShift / HDBSCAN
themselves stay the reference's CPU sklearn / hdbscan code (out of scope).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib as L


def nearest_centroid(features: torch.Tensor, centroids: torch.Tensor, return_distance: bool = False):
    """features [n, >=d] (CUDA fp32, rows contiguous), centroids [k, d] -> labels int32 [n] (ties: lowest index)."""
    if not features.is_cuda:
        raise L.CliftError("nearest_centroid: features must be a CUDA tensor (no CPU path)")
    if features.dim() != 2 or features.stride(1) != 1:
        raise L.CliftError("nearest_centroid: features must be [n, d] with unit column stride")
    c = centroids.to(features.device, torch.float32).contiguous()
    n, d = features.shape[0], c.shape[1]
    labels = torch.empty((n,), dtype=torch.int32, device=features.device)
    dist = torch.empty((n,), device=features.device) if return_distance else None
    with L.on(features.device):
        L.check(L.load().clift_assign_centroids(L.ptr(features) if features.is_contiguous() else features.data_ptr(), n, d,
                                                features.stride(0), L.ptr(c), c.shape[0], L.ptr(labels), L.ptr(dist),
                                                L.stream_ptr(features.device)))
    return (labels, dist) if return_distance else labels


def assign_clusters(all_thing_features, all_points_semantics, all_centroids: Dict[int, np.ndarray], device, num_images=None):
    """Call-compatible with inference/render_panopli.py:371 (same arguments, same [num_images, -1, K + 1] fp64 one-hot result),
    computed by ``clift_assign_clusters`` + ``clift_labels_onehot`` on the device: class argmax, nearest centroid inside the
    class's slice of one centroid table, label ranges in ascending class order and the one-hot rows.  The host only builds
    the class -> (first row, row count) table from the ``all_centroids`` dict and reads back one int32 (the label count,
    which sizes the result)."""
    lib = L.load()
    dev = torch.device(device)
    scores = torch.cat([t.to(dev, torch.float32) for t in all_points_semantics], dim=0).contiguous()
    feats = torch.as_tensor(np.ascontiguousarray(all_thing_features), dtype=torch.float32).to(dev)
    n, n_cls, d = feats.shape[0], scores.shape[1], feats.shape[1] - 1
    if scores.shape[0] != n:
        raise L.CliftError(f"assign_clusters: {n} feature rows but {scores.shape[0]} semantic rows")
    first, count, rows = np.zeros(n_cls, np.int32), np.zeros(n_cls, np.int32), []
    for cls in sorted(int(c) for c in all_centroids):
        table = np.asarray(all_centroids[cls], dtype=np.float32).reshape(-1, d)
        if not 0 <= cls < n_cls:
            raise L.CliftError(f"assign_clusters: centroid class {cls} outside the {n_cls} semantic classes")
        first[cls], count[cls] = sum(r.shape[0] for r in rows), table.shape[0]
        rows.append(table)
    cents = torch.from_numpy(np.concatenate(rows, 0) if rows else np.zeros((1, d), np.float32)).to(dev)
    d_first, d_count = torch.from_numpy(first).to(dev), torch.from_numpy(count).to(dev)
    labels = torch.empty((n,), dtype=torch.int32, device=dev)
    scratch = torch.empty((n_cls,), dtype=torch.int32, device=dev)
    stats = torch.empty((2,), dtype=torch.int32, device=dev)
    with L.on(dev):
        L.check(lib.clift_assign_clusters(L.ptr(feats), n, d, feats.stride(0) if n else d + 1, L.ptr(scores), n_cls, L.ptr(cents), L.ptr(d_first),
                                          L.ptr(d_count), L.ptr(labels), L.ptr(scratch), L.ptr(stats), L.stream_ptr(dev)))
        width, missing = (int(v) for v in stats.cpu().tolist())
        if missing:
            raise KeyError(missing - 1)           # a thing class with points but no centroids, as the reference's dict lookup
        onehot = torch.empty((n, width), dtype=torch.float64, device=dev)
        L.check(lib.clift_labels_onehot(L.ptr(labels), n, width, L.ptr(onehot), L.stream_ptr(dev)))
    return onehot.view(num_images, -1, width) if num_images is not None else onehot
