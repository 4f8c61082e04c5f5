"""CPU oracle for the exact k-NN path of nrlugg/fenix.  TEST INFRASTRUCTURE - NOT PRODUCT CODE.

A restatement (not a copy) of what the reference computes on this path, function by function:

  distance()    <- src/fenix/io/coder/coder.py:38-50      (metric arithmetic)
  from_arrow()  <- src/fenix/io/torch/torch.py:6-10        (values buffer -> (rows, D) tensor)
  call()        <- src/fenix/io/index/index.py:99-111,128-129,133-170
                   (target coercion, schema, per-chunk distance UDF, filter, select_k + take)
  coder_call()  <- src/fenix/io/coder/coder.py:143-194     (composite IVF codes ranked by summed codeword distance)
  ivf_call()    <- src/fenix/io/index/index.py:113-126     (`__CODED_ID__ isin(probe codes)` folded into the filter)

Where the arithmetic really lives: third-party wheels absent from /root/reference -
torch (pinned 2.1.2 in pdm.lock:1772-1773; 2.11.0 in this image) for cdist / normalize / matmul
and pyarrow (pinned 15.0.0, pdm.lock:1244-1245; 24.0.0 here) for select_k_unstable / take /
filter. The restatement calls the same library routines at the same call sites, except L2,
which spells out torch.cdist's matmul path (`_euclidean_dist`: one GEMM with K = D + 2, then
clamp_min(0).sqrt()) so that the arithmetic is visible; tests/test_oracle.py checks it is
bit-identical to torch.cdist here.

Parity pin: the reference's own tests hold no golden vectors for this path (they assert row
count and schema only, tests/test_flight.py:111-114), so the oracle is pinned against OUTPUTS
OF THE LIVE REFERENCE run in the build container: tests/golden/make_golden.py imports
/root/reference/src/fenix, runs fenix.io.index.call on seeded inputs and commits the results
as tests/golden/*.npz; tests/test_oracle.py replays them through this file.
"""
from __future__ import annotations

import warnings
from typing import Optional, Sequence

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import torch
import torch.nn.functional as F

DIST_COL = "__DISTANCE__"
ROW_COL = "__ROW__"
CODE_COL = "__CODED_ID__"


def _cdist_mm(u: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """torch.cdist's `use_mm_for_euclid_dist` path, spelled out (ATen _euclidean_dist)."""
    u_norm = u.pow(2).sum(dim=-1, keepdim=True)
    v_norm = v.pow(2).sum(dim=-1, keepdim=True)
    u_ = torch.cat([u.mul(-2), u_norm, torch.ones_like(u_norm)], dim=-1)
    v_ = torch.cat([v, torch.ones_like(v_norm), v_norm], dim=-1)
    return u_.matmul(v_.transpose(-2, -1)).clamp_min_(0).sqrt_()


def distance(u: torch.Tensor, v: torch.Tensor, metric: str) -> torch.Tensor:
    """coder.py:38-50. u: (U, D), v: (V, D) -> (U, V); smaller is better for every metric."""
    if metric in ("euclidean", "l2"):
        # coder.py:40 torch.cdist(u, v): the mm path is taken when U > 25 or V > 25, else the
        # direct kernel; the reference's chunks (>= 1000 rows) always take the mm path.
        if u.shape[-2] > 25 or v.shape[-2] > 25:
            return _cdist_mm(u, v)
        return torch.cdist(u, v)
    if metric == "cosine":
        # coder.py:43-45
        return 0.5 - 0.5 * F.normalize(u, dim=-1) @ F.normalize(v, dim=-1).transpose(-1, -2)
    if metric in ("dot", "inner_product"):
        # coder.py:48 (unary minus binds to u before the matmul)
        return (-u) @ v.transpose(-1, -2)
    raise ValueError(f"unknown metric {metric!r}")


def from_arrow(x) -> torch.Tensor:
    """torch.py:6-10: zero-copy view of the values buffer (offset / validity ignored)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)  # read-only mmap view, never written
        vals = torch.from_numpy(x.values.to_numpy(zero_copy_only=True))
    if isinstance(x, pa.FixedSizeListScalar):
        return vals
    return vals.view(-1, x.type.list_size)


def _coerce_target(target, typ: pa.DataType) -> pa.FixedSizeListScalar:
    """index.py:101-111."""
    if isinstance(target, pa.ChunkedArray):
        target = target.combine_chunks()
    if isinstance(target, pa.Array):
        target = target.to_numpy()
    if isinstance(target, torch.Tensor):
        target = target.numpy()
    if isinstance(target, np.ndarray):
        target = pa.scalar(target, type=typ)
    return target


def _distance_column(col: pa.ChunkedArray, target: pa.FixedSizeListScalar, metric: str) -> pa.ChunkedArray:
    """index.py:133-162: the scalar UDF is invoked once per chunk of the column."""
    q = from_arrow(target).unsqueeze(0)
    vt = col.type.value_type
    out = []
    for chunk in col.chunks:
        d = distance(q, from_arrow(chunk), metric).squeeze(0).numpy()
        out.append(pa.array(d, type=vt))
    return pa.chunked_array(out, type=vt)


def call(
    data: pa.Table,
    column: str,
    target,
    metric: str,
    select: Optional[Sequence[str]] = None,
    filter: Optional[pc.Expression] = None,
    maxval: Optional[int] = None,
) -> pa.Table:
    """index.py:81-170 for `coding is None`, on an in-memory table."""
    typ = data.schema.field(column).type
    target = _coerce_target(target, typ)
    cols = ([*select] if select is not None else data.column_names) + [DIST_COL]      # :128-129
    data = data.filter(filter) if filter is not None else data                          # :161
    data = data.append_column(DIST_COL, _distance_column(data.column(column), target, metric))  # :162
    data = data.select(cols)                                                            # :163
    if maxval is not None and len(data) > maxval:                                       # :165
        data = data.take(pc.select_k_unstable(data, maxval, [(DIST_COL, "ascending")]))  # :166-167
    return data.combine_chunks()                                                        # :170


def coder_call(target: torch.Tensor, tensor: torch.Tensor, metric: str, maxval: Optional[int] = None) -> torch.Tensor:
    """coder.py:143-194 on tensors: target (T, D), codebooks (n, k, D) -> composite codes ranked by the sum of the
    n codeword distances, best first: (T, maxval) or all (T, k^n). Composite code c uses codeword
    (c // k^(n-1-j)) % k of codebook j (coder.py:171-181: repeat_interleave / repeat index grids)."""
    n, k = tensor.shape[0], tensor.shape[1]
    data = distance(target, tensor.flatten(end_dim=-2), metric).view(-1, n, k)            # :170
    total = torch.tensor(0)
    for j in range(n):                                                                    # :171-181
        grid = torch.arange(0, k).repeat_interleave(k ** (n - j - 1)).repeat(k ** j)
        total = total + data[:, j, grid]
    if maxval is not None:
        return torch.topk(total, maxval, largest=False).indices                           # :184
    return torch.argsort(total, descending=False)                                         # :186


def ivf_call(data: pa.Table, column: str, target, tensor: torch.Tensor, coding_metric: str, metric: Optional[str] = None,
             select: Optional[Sequence[str]] = None, filter: Optional[pc.Expression] = None,
             maxval: Optional[int] = None, probes: Optional[int] = None) -> pa.Table:
    """index.py:81-170 with a coding: `data` already carries `__CODED_ID__` (index.load, :19-35)."""
    typ = data.schema.field(column).type
    scalar = _coerce_target(target, typ)
    if probes is not None:                                                                # :113-126
        if metric is None:
            metric = coding_metric
        codes = coder_call(from_arrow(pa.array([scalar])), tensor, coding_metric, probes)
        mask = pc.field(CODE_COL).isin(pa.array(codes.reshape(-1).numpy()))
        filter = mask if filter is None else (filter & mask)
    return call(data, column, scalar, metric, select=select, filter=filter, maxval=maxval)


def search_rows(data: pa.Table, column: str, target, metric: str, k: Optional[int],
                filter: Optional[pc.Expression] = None) -> tuple[np.ndarray, np.ndarray]:
    """call() with a row-position column attached -> (row positions, distances), reference order."""
    with_rows = data.append_column(ROW_COL, pa.array(np.arange(data.num_rows, dtype=np.int64)))
    out = call(with_rows, column, target, metric, select=[ROW_COL], filter=filter, maxval=k)
    return out.column(ROW_COL).to_numpy(), out.column(DIST_COL).to_numpy()


def canonical(rows: np.ndarray, dist: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Order a result by (distance, row): the comparison form (the reference's own order among
    equal distances is heap pop order, i.e. unspecified)."""
    order = np.lexsort((rows, dist))
    return rows[order], dist[order]


def brute_force_f64(corpus: np.ndarray, queries: np.ndarray, metric: str, k: int) -> tuple[np.ndarray, np.ndarray]:
    """fp64 adjudicator: exact arithmetic of the same definitions, ordered by (distance, row)."""
    x = np.asarray(corpus, dtype=np.float64)
    q = np.atleast_2d(np.asarray(queries, dtype=np.float64))
    if metric in ("euclidean", "l2"):
        d2 = (q * q).sum(1)[:, None] - 2.0 * (q @ x.T) + (x * x).sum(1)[None, :]
        d = np.sqrt(np.maximum(d2, 0.0))
    elif metric == "cosine":
        qn = q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
        xn = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
        d = 0.5 - 0.5 * (qn @ xn.T)
    elif metric in ("dot", "inner_product"):
        d = -(q @ x.T)
    else:
        raise ValueError(f"unknown metric {metric!r}")
    d32 = d.astype(np.float32) + np.float32(0.0)
    kk = min(k, x.shape[0])
    rows = np.empty((q.shape[0], kk), dtype=np.int64)
    dist = np.empty((q.shape[0], kk), dtype=np.float32)
    ar = np.arange(x.shape[0])
    for i in range(q.shape[0]):
        order = np.lexsort((ar, d32[i]))[:kk]
        rows[i], dist[i] = order, d32[i][order]
    return rows, dist
