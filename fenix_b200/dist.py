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
all-gathered over NVLink and merged - inside the C library (fx_search_sharded, its own NCCL communicator; see
include/fenix_knn.h). The torch.distributed helpers below (candidate packing, gathers) serve the replicated-corpus
mode and the CPU (gloo) tests of the host logic.
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


def broadcast_comm_id(make_id, group=None, device: Optional[torch.device] = None) -> bytes:
    """Rank 0 draws the communicator id (`make_id()` -> 128 bytes, knn.Comm.unique_id), every rank receives it over
    the process group that is already up (NCCL on GPUs, gloo in the CPU tests)."""
    rank = td.get_rank(group)
    buf = torch.zeros(knn.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        raw = make_id()
        if len(raw) != knn.COMM_ID_BYTES:
            raise ValueError(f"communicator id must be {knn.COMM_ID_BYTES} bytes")
        buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    if device is not None:
        buf = buf.to(device)
    td.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


class ShardedSearcher:
    """This rank's shard + the exchange. All ranks must call `search_*` together.

    The whole step lives in the C library (fx_search_sharded*): query-slice upload + all-gather, shard search,
    all-gather of the k*W candidates on the library's own NCCL communicator, merge, result copy - one stream, one host
    synchronisation, no Python between the kernels. torch.distributed only carries the communicator id at start-up."""

    def __init__(self, corpus: knn.Corpus, group=None) -> None:
        self.corpus = corpus
        self.group = group
        self.world = td.get_world_size(group) if td.is_initialized() else 1
        self.rank = td.get_rank(group) if td.is_initialized() else 0
        self.device = torch.device("cuda", corpus.ctx.device)
        self.merge_launches = 0
        self.comm = None
        self._out_key, self._out_rows, self._out_dist = None, None, None
        if self.world > 1:
            uid = broadcast_comm_id(knn.Comm.unique_id, group, self.device if td.get_backend(group) == "nccl" else None)
            self.comm = knn.Comm(corpus.ctx, uid, self.world, self.rank)

    def search_device(self, d_queries: torch.Tensor, metric: int, k: int,
                      precision: int = knn.PREC_FP32) -> tuple[torch.Tensor, torch.Tensor]:
        """d_queries: float32 [Q, D], the whole batch, resident on this rank's GPU. Returns the global (rows, dist) [Q, k]
        on the device."""
        n_q = d_queries.shape[0]
        rows, dist = self._out(n_q, k)
        if self.world == 1:
            self.corpus.search_device(d_queries.data_ptr(), n_q, metric, k, precision, rows.data_ptr(), dist.data_ptr())
            return rows, dist
        self.corpus.search_sharded_raw(self.comm, d_queries.data_ptr(), n_q, metric, k, precision, rows.data_ptr(),
                                       dist.data_ptr(), on_device=True)
        self.merge_launches += 1
        return rows, dist

    def _out(self, n_q: int, k: int) -> tuple[torch.Tensor, torch.Tensor]:
        """Result tensors of a device-resident search, reused from call to call (the library synchronises its stream before
        it returns, so the previous result has been consumed or copied by then; callers that keep results clone them)."""
        key = (n_q, k)
        if self._out_key != key:
            self._out_rows = torch.empty((n_q, k), dtype=torch.int64, device=self.device)
            self._out_dist = torch.empty((n_q, k), dtype=torch.float32, device=self.device)
            torch.cuda.current_stream(self.device).synchronize()   # they exist before the library's stream writes them
            self._out_key = key
        return self._out_rows, self._out_dist

    def search_host(self, h_queries: torch.Tensor, metric: int, k: int, precision: int = knn.PREC_FP32,
                    h_rows: Optional[torch.Tensor] = None, h_dist: Optional[torch.Tensor] = None,
                    result_rank: Optional[int] = None):
        """End-to-end form: (pinned) host queries in - every rank uploads 1/W of the batch, the slices are all-gathered
        over NVLink - and (pinned) host results out, on every rank or on `result_rank` alone."""
        n_q = h_queries.shape[0]
        want = result_rank is None or result_rank == self.rank
        if want and h_rows is None:
            h_rows = torch.empty((n_q, k), dtype=torch.int64)
            h_dist = torch.empty((n_q, k), dtype=torch.float32)
        if self.world == 1:
            self.corpus.search_raw(h_queries.data_ptr(), n_q, metric, k, precision, h_rows.data_ptr(), h_dist.data_ptr())
            return h_rows, h_dist
        self.corpus.search_sharded_raw(self.comm, h_queries.data_ptr(), n_q, metric, k, precision,
                                       h_rows.data_ptr() if want else None, h_dist.data_ptr() if want else None)
        if want:
            self.merge_launches += 1
        return (h_rows, h_dist) if want else (None, None)

    def close(self) -> None:
        if self.comm is not None:
            self.comm.close()
            self.comm = None


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
