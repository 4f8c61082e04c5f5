"""Metric arithmetic seam: `distance(u, v, metric)`.

Mirrors fenix.io.coder.distance (src/fenix/io/coder/coder.py:38-50): same metric names, same
conventions (Euclidean distance WITH sqrt; 0.5 - 0.5*cos with eps 1e-12; negated dot), same
ValueError on an unknown name - but the (U x V) distance block is produced by the CUDA
distance-column kernel of libfenix_knn over a temporary device shard holding `v`.
The IVF codebook functions of the reference module (Config/make/load/call) are out of scope.
"""
from __future__ import annotations

import numpy as np

from .. import knn
from . import shards as _shards

LOCATION: str = "codings"


def distance(u, v, metric: str):
    """Distances between every row of `u` (U x D) and every row of `v` (V x D) -> (U x V).

    Accepts numpy arrays or CPU torch tensors and returns the same kind."""
    code = knn.metric_code(metric)
    as_tensor = type(u).__module__.startswith("torch")
    un = np.ascontiguousarray(u.detach().cpu().numpy() if as_tensor else np.asarray(u), dtype=np.float32)
    vn = v.detach().cpu().numpy() if type(v).__module__.startswith("torch") else np.asarray(v)
    vn = np.ascontiguousarray(vn, dtype=np.float32)
    if un.ndim != 2 or vn.ndim != 2 or un.shape[1] != vn.shape[1]:
        raise ValueError(f"distance expects (U,D) and (V,D), got {un.shape} and {vn.shape}")
    corpus = knn.Corpus(_shards.context(_shards.devices()[0]), vn.shape[0], vn.shape[1])
    try:
        corpus.append(vn)
        corpus.finalize()
        out = np.stack([corpus.distances(row, code) for row in un]) if len(un) else np.empty((0, vn.shape[0]), np.float32)
    finally:
        corpus.close()
    if as_tensor:
        import torch

        return torch.from_numpy(out)
    return out


def _unsupported(*_a, **_k):
    raise NotImplementedError("IVF codebooks (coder.make/load/call) are outside the exact k-NN path of this build")


make = load = call = _unsupported


def list(root: str):
    return iter(())


def drop(root: str, name: str) -> None:
    return None
