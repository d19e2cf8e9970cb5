"""Nearest-centroid assignment of rendered embeddings.

``nearest_centroid`` replaces the ``torch.cdist`` + ``argmin`` pair of inference/render_panopli.py:389-396;
``assign_clusters`` keeps the reference function's signature and label bookkeeping (render_panopli.py:371-419)
with the distance work on ``clift_assign_centroids``.  MeanShift / HDBSCAN themselves stay the reference's
CPU sklearn / hdbscan code (out of scope).
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
    L.check(L.load().clift_assign_centroids(L.ptr(features) if features.is_contiguous() else features.data_ptr(), n, d,
                                            features.stride(0), L.ptr(c), c.shape[0], L.ptr(labels), L.ptr(dist),
                                            L.stream_ptr(features.device)))
    return (labels, dist) if return_distance else labels


def assign_clusters(all_thing_features, all_points_semantics, all_centroids: Dict[int, np.ndarray], device, num_images=None):
    """render_panopli.py:371-419 with the per-class cdist/argmin on the GPU kernel.  ``all_thing_features`` is the
    numpy array the reference builds ([N, 1+d]; column 0 == -inf marks thing pixels)."""
    sem = torch.cat(all_points_semantics, dim=0).argmax(dim=-1).cpu().numpy()
    thing_mask = all_thing_features[..., 0] == -float("inf")
    features = all_thing_features[thing_mask][:, 1:]
    n_all = all_thing_features.shape[0]
    thing_semantics = sem[thing_mask]
    all_labels = np.zeros(n_all, dtype=np.int32)
    all_thing_labels = np.zeros(features.shape[0], dtype=np.int32)
    max_label = 0
    for thing_cls in np.unique(thing_semantics):
        cls_mask = thing_semantics == thing_cls
        feats = torch.as_tensor(np.ascontiguousarray(features[cls_mask]), dtype=torch.float32).to(device)
        cents = torch.as_tensor(np.asarray(all_centroids[thing_cls]), dtype=torch.float32)
        lab = nearest_centroid(feats, cents).cpu().numpy().astype(np.int64)
        lab[lab != -1] += max_label
        if np.any(lab != -1):
            max_label = lab.max() + 1
        all_thing_labels[cls_mask] = lab
    all_labels[thing_mask] = all_thing_labels
    all_labels[~thing_mask] = -1
    all_labels = all_labels + 1
    num_unique_labels = all_labels.max() + 1
    onehot = np.zeros((n_all, num_unique_labels))
    onehot[np.arange(n_all), all_labels] = 1
    return torch.from_numpy(onehot).view(num_images, -1, num_unique_labels).to(device)
