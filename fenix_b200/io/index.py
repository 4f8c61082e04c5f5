"""Exact search orchestration: the drop-in seam `io.index.call`.

Mirrors fenix.io.index.call (src/fenix/io/index/index.py:81-170) for the `coding is None`
branch: same arguments, same result schema (select columns + `__DISTANCE__` typed as the
column's value type), same `maxval` edge cases (None or >= row count returns every row in
table order with the distance attached, index.py:165), same filter-before-distance semantics
(index.py:161). What changes is where the arithmetic runs: the per-chunk Arrow UDF ->
torch.cdist -> select_k_unstable chain (index.py:133-168) is replaced by one call into
libfenix_knn against a device-resident shard set; only the k winning rows are gathered on the
host. Result rows are ordered by (distance, row position) - the reference's order among equal
distances is unspecified ("unstable").

The IVF branch (`coding`/`probes`, index.py:113-126) is out of scope and raises.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

from .. import knn
from . import shards as _shards
from . import table as _table
from .batcher import MicroBatcher

_batcher = MicroBatcher()

CODE_COL: str = "__CODED_ID__"
DIST_COL: str = "__DISTANCE__"
QUERY_COL: str = "__QUERY__"
LOCATION: str = "indexes"


def _is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "numpy")


def coerce_target(target, dim: int) -> np.ndarray:
    """Query -> float32 array of shape (Q, dim).

    The reference accepts ChunkedArray / Array / Tensor / ndarray / FixedSizeListScalar and casts
    to the column's value type (index.py:101-111). 2-D input (Q queries per call) is this
    build's wire extension; the reference raises on it.
    """
    if isinstance(target, pa.ChunkedArray):
        target = target.combine_chunks()
    if isinstance(target, pa.FixedSizeListScalar):
        target = target.values
    if isinstance(target, pa.FixedSizeListArray):
        flat = target.values.to_numpy(zero_copy_only=False)
        lo = target.offset * target.type.list_size
        target = flat[lo: lo + len(target) * target.type.list_size].reshape(len(target), target.type.list_size)
    elif isinstance(target, pa.Array):
        target = target.to_numpy(zero_copy_only=False)
    if _is_tensor(target):
        target = target.detach().cpu().numpy()
    q = np.asarray(target)
    if q.ndim == 1:
        q = q.reshape(1, -1)
    if q.ndim != 2 or q.shape[1] != dim:
        raise pa.ArrowInvalid(f"target has shape {tuple(np.asarray(target).shape)}; the column holds vectors of size {dim}")
    return np.ascontiguousarray(q, dtype=np.float32)


def _row_mask(data: pa.Table, column: str, filter: pc.Expression) -> np.ndarray:
    """Evaluate the predicate on the host over the scalar columns -> byte mask over rows."""
    n = data.num_rows
    probe = data.drop_columns([column]).append_column("__ROW__", pa.array(np.arange(n, dtype=np.int64)))
    try:
        kept = probe.filter(filter).column("__ROW__")
    except (pa.ArrowInvalid, KeyError):
        # the predicate references the vector column itself
        kept = data.append_column("__ROW__", pa.array(np.arange(n, dtype=np.int64))).filter(filter).column("__ROW__")
    mask = np.zeros(n, dtype=np.uint8)
    mask[kept.to_numpy()] = 1
    return mask


def call(
    root: str,
    coding: str | None,
    source: str | Sequence[str] | pa.Table,
    column: str,
    target,
    metric: str | None = None,
    select: Sequence[str] | None = None,
    filter: pc.Expression | None = None,
    maxval: int | None = None,
    probes: int | None = None,
    precision: int = knn.PREC_FP32,
) -> pa.Table:
    if coding is not None:
        raise NotImplementedError("IVF search (coding/probes) is outside the exact k-NN path of this build")
    if metric is None:
        raise AssertionError("metric is required")
    knn.metric_code(metric)  # ValueError on unknown names, as coder.py:50

    if isinstance(source, pa.Table):
        data = source
        shard = _shards.from_chunks(data.column(column))
        owned = True
    else:
        data = _shards.load_table(root, source)   # raises like table.load when the file is missing
        shard = _shards.get(root, source, column, data)
        owned = False

    try:
        typ = data.schema.field(column).type
        queries = coerce_target(target, typ.list_size)
        batched = queries.shape[0] != 1 or _batched_input(target)

        out_cols = [*select] if select is not None else data.column_names
        mask = _row_mask(data, column, filter) if filter is not None else None
        n_live = int(mask.sum()) if mask is not None else data.num_rows

        if not batched and (maxval is None or n_live <= maxval):
            # every (surviving) row, table order, distance attached (index.py:162-165)
            dist = shard.distances(queries[0], metric)
            out = data.select(out_cols).append_column(DIST_COL, pa.array(dist, type=typ.value_type))
            if mask is not None:
                out = out.filter(pa.array(mask.view(np.bool_)))
            return out.combine_chunks()

        k = n_live if maxval is None else min(int(maxval), n_live)
        if k < 1:
            empty = data.select(out_cols).slice(0, 0).append_column(DIST_COL, pa.array([], type=typ.value_type))
            if batched:
                empty = empty.append_column(QUERY_COL, pa.array([], type=pa.int32()))
            return empty.combine_chunks()
        if not batched and mask is None and not owned and _batcher.enabled:
            # concurrent single-query RPCs against the same table/metric/k share one pass over the corpus
            key = (id(shard), metric, k, precision)
            r1, d1 = _batcher.submit(key, queries[0], lambda qs: shard.search(qs, metric, k, precision, None))
            rows, dist = r1[None, :], d1[None, :]
        else:
            rows, dist = shard.search(queries, metric, k, precision, mask)

        keep = rows.reshape(-1) >= 0
        flat_rows = rows.reshape(-1)[keep]
        out = data.select(out_cols).take(pa.array(flat_rows, type=pa.int64()))
        out = out.append_column(DIST_COL, pa.array(dist.reshape(-1)[keep], type=typ.value_type))
        if batched:
            qid = np.repeat(np.arange(rows.shape[0], dtype=np.int32), rows.shape[1])[keep]
            out = out.append_column(QUERY_COL, pa.array(qid, type=pa.int32()))
        return out.combine_chunks()
    finally:
        if owned:
            shard.close()


def _batched_input(target) -> bool:
    if isinstance(target, pa.ChunkedArray):
        return pa.types.is_fixed_size_list(target.type)
    if isinstance(target, pa.FixedSizeListArray):
        return True
    if isinstance(target, np.ndarray) or _is_tensor(target):
        return target.ndim == 2
    return False


# IVF sidecar management (index.py:19-78) is not part of the exact path.
def load(*_a, **_k):
    raise NotImplementedError("IVF index sidecars are outside the exact k-NN path of this build")


make = load


def list(root: str):
    return iter(())


def drop(root: str, name: str, source: str, column: str) -> None:
    return None
