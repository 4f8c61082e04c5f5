"""Device-resident corpus shards: the Arrow -> pinned host -> HBM handoff.

Replaces fenix.io.torch.from_arrow (src/fenix/io/torch/torch.py:6-10), which re-wraps the
mmap'd values buffer of every chunk on every query, with a copy made ONCE per table version:
each record batch's FixedSizeList values buffer is handed (zero-copy numpy view) to
fx_corpus_append, which stages it through a pinned ring and DMAs it into the shard.
The cache is keyed on (root, sources, column) and validated against (mtime_ns, size) of the
backing files, so do_put / drop-table / remove invalidate it.
"""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import pyarrow as pa

from .. import knn
from . import table as _table

_lock = threading.Lock()
_contexts: dict[int, knn.Context] = {}
_cache: dict[tuple, "ShardSet"] = {}


def devices() -> list[int]:
    """CUDA devices the in-process server shards over (FENIX_DEVICES="0,1,..."; default "0")."""
    spec = os.environ.get("FENIX_DEVICES", "0")
    return [int(tok) for tok in spec.split(",") if tok.strip() != ""]


def context(device: int) -> knn.Context:
    with _lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = _contexts[device] = knn.Context(device)
        return ctx


def chunk_rows(chunk: pa.FixedSizeListArray) -> np.ndarray:
    """Zero-copy (rows, D) float32 view of one chunk's values buffer.

    Unlike torch.py:8-10 (`.values`, which ignores a slice offset) the chunk's offset is
    honoured; validity bitmaps are ignored exactly as the reference does.
    """
    typ = chunk.type
    if not pa.types.is_fixed_size_list(typ):
        raise TypeError(f"vector column must be FixedSizeList<float32>[D], got {typ}")
    d = typ.list_size
    vals = chunk.values.to_numpy(zero_copy_only=True)
    lo = chunk.offset * d
    rows = vals[lo: lo + len(chunk) * d].reshape(len(chunk), d)
    if typ.value_type == pa.float32():
        return rows
    if typ.value_type in (pa.float16(), pa.float64()):
        # shards are float32 (north_star); other float widths are converted once, at upload. The reference computes
        # in the column's own type (index.py:133-159), so float64 columns agree to float32 rounding, not bit for bit.
        return rows.astype(np.float32)
    raise NotImplementedError(f"embedding columns must be float16 / float32 / float64 (got {typ.value_type})")


@dataclass
class ShardSet:
    """Row-sharded device copy of one vector column (one Corpus per device)."""

    dim: int
    n_rows: int
    corpora: list[knn.Corpus] = field(default_factory=list)
    bases: list[int] = field(default_factory=list)
    signature: tuple = ()

    def search(self, queries: np.ndarray, metric: str, k: int, precision: int = knn.PREC_FP32,
               row_mask: Optional[np.ndarray] = None) -> tuple[np.ndarray, np.ndarray]:
        if len(self.corpora) == 1:
            return self.corpora[0].search(queries, metric, k, precision, row_mask)
        return self._search_sharded(queries, metric, k, precision, row_mask)

    def _search_sharded(self, queries, metric, k, precision, row_mask):
        # one thread per device (ctypes releases the GIL); shard-local top-k lists are copied to
        # the first device and merged there by fx_merge_topk (order: distance, then row)
        import torch

        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        n_q = q.shape[0]
        results: list = [None] * len(self.corpora)
        errors: list = []

        def work(i: int) -> None:
            try:
                c = self.corpora[i]
                m = None
                if row_mask is not None:
                    m = row_mask[self.bases[i]: self.bases[i] + c.n_rows]
                results[i] = c.search(q, metric, k, precision, m)
            except BaseException as exc:  # surfaced below
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(self.corpora))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        dev0 = self.corpora[0].ctx
        device = torch.device("cuda", dev0.device)
        rows = torch.from_numpy(np.stack([r[0] for r in results])).to(device)
        dist = torch.from_numpy(np.stack([r[1] for r in results])).to(device)
        out_rows = torch.empty((n_q, k), dtype=torch.int64, device=device)
        out_dist = torch.empty((n_q, k), dtype=torch.float32, device=device)
        torch.cuda.synchronize(device)
        dev0.merge_topk_device(rows.data_ptr(), dist.data_ptr(), len(results), n_q, k,
                               out_rows.data_ptr(), out_dist.data_ptr())
        return out_rows.cpu().numpy(), out_dist.cpu().numpy()

    def distances(self, query: np.ndarray, metric: str) -> np.ndarray:
        parts = [c.distances(query, metric) for c in self.corpora]
        return parts[0] if len(parts) == 1 else np.concatenate(parts)

    def close(self) -> None:
        for c in self.corpora:
            c.close()
        self.corpora = []


def from_chunks(column: pa.ChunkedArray, device_ids: Optional[Sequence[int]] = None) -> ShardSet:
    """Upload a FixedSizeList<float32>[D] column, row-sharded contiguously over the devices."""
    typ = column.type
    if not pa.types.is_fixed_size_list(typ):
        raise TypeError(f"vector column must be FixedSizeList<float32>[D], got {typ}")
    if typ.value_type not in (pa.float16(), pa.float32(), pa.float64()):
        raise NotImplementedError(f"embedding columns must be float16 / float32 / float64 (got {typ.value_type})")
    dim, n = typ.list_size, len(column)
    devs = list(device_ids) if device_ids is not None else devices()
    world = max(1, min(len(devs), max(n, 1)))
    per = -(-n // world) if n else 0
    shard = ShardSet(dim=dim, n_rows=n)
    for r in range(world):
        lo, hi = r * per, min(n, (r + 1) * per)
        shard.bases.append(lo)
        shard.corpora.append(knn.Corpus(context(devs[r]), max(hi - lo, 0), dim, row_base=lo))
    pos = 0
    for chunk in column.chunks:
        rows = chunk_rows(chunk)
        done = 0
        while done < len(rows):
            r = min(pos // per, world - 1) if per else 0
            room = shard.bases[r] + shard.corpora[r].capacity - pos
            take = min(room, len(rows) - done)
            shard.corpora[r].append(rows[done: done + take])
            done += take
            pos += take
    for c in shard.corpora:
        c.finalize()
    return shard


def _signature(root: str, names: Sequence[str]) -> tuple:
    sig = []
    for n in names:
        st = os.stat(_table.path_of(root, n))
        sig.append((n, st.st_mtime_ns, st.st_size))
    return tuple(sig)


_tables: dict[tuple, tuple] = {}


def load_table(root: str, source: str | Sequence[str]) -> pa.Table:
    """`io.table.load` with the parsed (memory-mapped, zero-copy) Table cached per file version: the
    reference re-maps and re-parses the IPC stream on every search (index.py:93-97, ~3.5 % of its query
    time at 100 chunks, more with many chunks)."""
    names = (source,) if isinstance(source, str) else tuple(source)
    key = (os.path.abspath(root), names)
    sig = _signature(root, names)
    with _lock:
        hit = _tables.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
    table = _table.load(root, source)
    with _lock:
        _tables[key] = (sig, table)
    return table


def get(root: str, source: str | Sequence[str], column: str, table: pa.Table) -> ShardSet:
    """Cached shard set for `column` of the named table(s); uploads on first use."""
    names = (source,) if isinstance(source, str) else tuple(source)
    key = (os.path.abspath(root), names, column, tuple(devices()))
    sig = _signature(root, names)
    with _lock:
        hit = _cache.get(key)
        if hit is not None and hit.signature == sig:
            return hit
    fresh = from_chunks(table.column(column))
    fresh.signature = sig
    with _lock:
        old = _cache.get(key)
        _cache[key] = fresh
    if old is not None:
        old.close()
    return fresh


def invalidate(root: Optional[str] = None, name: Optional[str] = None) -> None:
    """Drop cached shards (all, those under `root`, or those that include table `name`)."""
    root = os.path.abspath(root) if root is not None else None
    with _lock:
        doomed = [k for k in _cache if (root is None or k[0] == root) and (name is None or name in k[1])]
        victims = [_cache.pop(k) for k in doomed]
        for k in [k for k in _tables if (root is None or k[0] == root) and (name is None or name in k[1])]:
            _tables.pop(k)
    for v in victims:
        v.close()
