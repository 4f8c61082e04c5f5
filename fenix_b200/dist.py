"""Multi-GPU search: one process per GPU (torch.distributed), one exchange step. Two partitionings:

* `ShardedSearcher` - the corpus is ROW-sharded (the default; large corpora): described below;
* `ReplicaSearcher` - the corpus is REPLICATED and the QUERY batch is split (small corpora / huge batches, SURVEY.md
  section 8e "query-sharding"): each rank answers its contiguous slice of the batch against the whole corpus and the
  slices are all-gathered; nothing to merge.

Row-sharded search:

New relative to the reference (it has no parallelism): rank r owns the
contiguous rows [r*ceil(N/W), min(N,(r+1)*ceil(N/W))) so that global row = base + local row and
"lowest row wins" tie-breaking survives sharding. Every rank answers all Q queries against its
shard (shard-local top-k ordered by (distance, row)), the k*W candidates per query are
all-gathered (NCCL over NVLink on GPUs, gloo on CPU for tests) and merged by fx_merge_topk.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as td

from . import knn


def shard_bounds(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Rows [lo, hi) owned by `rank` under the contiguous ceil split."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = -(-n_rows // world) if n_rows > 0 else 0
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def pack_candidates(rows: torch.Tensor, dist: torch.Tensor) -> torch.Tensor:
    """(rows int64 [Q,k], dist f32 [Q,k]) -> one int64 [2,Q,k] tensor: a single collective."""
    return torch.stack([rows, dist.contiguous().view(torch.int32).to(torch.int64)])


def unpack_candidates(packed: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Inverse of pack_candidates for a gathered [W,2,Q,k] tensor -> rows [W,Q,k], dist [W,Q,k]."""
    rows = packed[:, 0].contiguous()
    dist = packed[:, 1].to(torch.int32).contiguous().view(torch.float32)
    return rows, dist


def gather_candidates(rows: torch.Tensor, dist: torch.Tensor, group=None) -> tuple[torch.Tensor, torch.Tensor]:
    """All-gather every rank's shard-local top-k. Returns rows [W,Q,k], dist [W,Q,k] (list-major,
    the layout fx_merge_topk takes)."""
    world = td.get_world_size(group)
    packed = pack_candidates(rows, dist)
    out = torch.empty((world * packed.shape[0], *packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    td.all_gather_into_tensor(out, packed, group=group)  # concatenation along dim 0 (gloo and nccl agree)
    return unpack_candidates(out.view(world, *packed.shape))


class ShardedSearcher:
    """This rank's shard + the collective merge. All ranks must call `search` together."""

    def __init__(self, corpus: knn.Corpus, group=None) -> None:
        self.corpus = corpus
        self.group = group
        self.world = td.get_world_size(group) if td.is_initialized() else 1
        self.device = torch.device("cuda", corpus.ctx.device)
        self.merge_launches = 0

    def search_device(self, d_queries: torch.Tensor, metric: int, k: int,
                      precision: int = knn.PREC_FP32) -> tuple[torch.Tensor, torch.Tensor]:
        """d_queries: float32 [Q, D] on this rank's GPU. Returns the global (rows, dist) [Q, k]."""
        n_q = d_queries.shape[0]
        rows = torch.empty((n_q, k), dtype=torch.int64, device=self.device)
        dist = torch.empty((n_q, k), dtype=torch.float32, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        self.corpus.search_device(d_queries.data_ptr(), n_q, metric, k, precision, rows.data_ptr(), dist.data_ptr())
        if self.world == 1:
            return rows, dist
        all_rows, all_dist = gather_candidates(rows, dist, self.group)
        out_rows = torch.empty_like(rows)
        out_dist = torch.empty_like(dist)
        torch.cuda.current_stream(self.device).synchronize()
        self.corpus.ctx.merge_topk_device(all_rows.data_ptr(), all_dist.data_ptr(), self.world, n_q, k,
                                          out_rows.data_ptr(), out_dist.data_ptr())
        self.merge_launches += 1
        return out_rows, out_dist

    def search_host(self, h_queries: torch.Tensor, metric: int, k: int, precision: int = knn.PREC_FP32,
                    h_rows: Optional[torch.Tensor] = None, h_dist: Optional[torch.Tensor] = None):
        """End-to-end form: pinned host queries in, pinned host results out (H2D/D2H included)."""
        d_q = h_queries.to(self.device, non_blocking=True)
        rows, dist = self.search_device(d_q, metric, k, precision)
        if h_rows is None:
            return rows.cpu(), dist.cpu()
        h_rows.copy_(rows, non_blocking=True)
        h_dist.copy_(dist, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return h_rows, h_dist


def gather_slices(rows: torch.Tensor, dist: torch.Tensor, n_q: int, group=None) -> tuple[torch.Tensor, torch.Tensor]:
    """All-gather per-rank result slices ([per, k] each, `per` = ceil(n_q / W), short slices padded) into the
    results of the whole batch, [n_q, k], in query order."""
    world = td.get_world_size(group)
    packed = pack_candidates(rows, dist)                                  # [2, per, k]
    out = torch.empty((world * 2, *packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    td.all_gather_into_tensor(out, packed, group=group)
    g_rows, g_dist = unpack_candidates(out.view(world, *packed.shape))    # [W, per, k]
    k = rows.shape[1]
    return g_rows.reshape(-1, k)[:n_q].contiguous(), g_dist.reshape(-1, k)[:n_q].contiguous()


class ReplicaSearcher:
    """Replicated corpus, query batch split over the ranks. All ranks must call `search` together with the SAME
    batch; every rank returns the results of the whole batch."""

    def __init__(self, corpus: knn.Corpus, group=None) -> None:
        self.corpus = corpus
        self.group = group
        self.world = td.get_world_size(group) if td.is_initialized() else 1
        self.rank = td.get_rank(group) if td.is_initialized() else 0
        self.device = torch.device("cuda", corpus.ctx.device)
        self.merge_launches = 0

    def search_device(self, d_queries: torch.Tensor, metric: int, k: int,
                      precision: int = knn.PREC_FP32) -> tuple[torch.Tensor, torch.Tensor]:
        """d_queries: the WHOLE batch, float32 [Q, D] on this rank's GPU. Returns the results of the whole batch."""
        n_q = d_queries.shape[0]
        lo, hi = shard_bounds(n_q, self.world, self.rank)
        return self._search_slice(d_queries[lo:hi].contiguous() if hi > lo else None, n_q, metric, k, precision)

    def _search_slice(self, d_slice: Optional[torch.Tensor], n_q: int, metric: int, k: int, precision: int):
        """Search this rank's slice (already on the device, or None when the slice is empty) and gather the batch."""
        per = -(-n_q // self.world)
        rows = torch.full((per, k), -1, dtype=torch.int64, device=self.device)
        dist = torch.full((per, k), float("inf"), dtype=torch.float32, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        if d_slice is not None and d_slice.shape[0] > 0:
            self.corpus.search_device(d_slice.data_ptr(), d_slice.shape[0], metric, k, precision, rows.data_ptr(), dist.data_ptr())
        if self.world == 1:
            return rows[:n_q], dist[:n_q]
        return gather_slices(rows, dist, n_q, self.group)

    def search_host(self, h_queries: torch.Tensor, metric: int, k: int, precision: int = knn.PREC_FP32,
                    h_rows: Optional[torch.Tensor] = None, h_dist: Optional[torch.Tensor] = None,
                    result_rank: Optional[int] = None):
        """End-to-end form: only this rank's slice of the (pinned) host batch is uploaded; the gathered results are
        copied back to the host on every rank, or on `result_rank` alone."""
        n_q = h_queries.shape[0]
        lo, hi = shard_bounds(n_q, self.world, self.rank)
        d_slice = h_queries[lo:hi].to(self.device, non_blocking=True) if hi > lo else None
        rows, dist = self._search_slice(d_slice, n_q, metric, k, precision)
        if result_rank is not None and self.rank != result_rank:
            torch.cuda.current_stream(self.device).synchronize()
            return None, None
        if h_rows is None:
            return rows.cpu(), dist.cpu()
        h_rows.copy_(rows, non_blocking=True)
        h_dist.copy_(dist, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return h_rows, h_dist
