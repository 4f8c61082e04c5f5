"""world_size-2 gloo test (CPU) of the N>1 host logic: shard bounds, candidate packing, the
all-gather layout fx_merge_topk consumes. The per-shard search is played by the oracle's fp64
brute force and the merge by numpy - both are checkers here, the plumbing is what is tested."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from fenix_b200 import dist as fdist


def test_shard_bounds_cover_rows_contiguously():
    for n in (0, 1, 7, 8, 9, 1000, 1001):
        for w in (1, 2, 3, 8):
            spans = [fdist.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
    with pytest.raises(ValueError):
        fdist.shard_bounds(10, 2, 2)


def test_pack_roundtrip_preserves_bits():
    rows = torch.tensor([[5, -1], [2**40, 3]], dtype=torch.int64)
    dist = torch.tensor([[0.5, float("inf")], [-0.0, 1e-30]], dtype=torch.float32)
    r, d = fdist.unpack_candidates(fdist.pack_candidates(rows, dist).unsqueeze(0))
    assert torch.equal(r[0], rows)
    assert torch.equal(d[0].view(torch.int32), dist.view(torch.int32))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    from oracle import brute_force_f64

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        corpus = rng.standard_normal((1001, 16), dtype=np.float32)
        corpus[900] = corpus[3]  # a cross-shard tie
        queries = np.concatenate([rng.standard_normal((4, 16), dtype=np.float32), corpus[3:4]])
        k = 6
        lo, hi = fdist.shard_bounds(len(corpus), world, rank)
        rows, dist = brute_force_f64(corpus[lo:hi], queries, "l2", k)
        rows = rows + lo
        g_rows, g_dist = fdist.gather_candidates(torch.from_numpy(rows), torch.from_numpy(dist))
        assert g_rows.shape == (world, len(queries), k)
        # checker merge: (distance, row) order over the k*W candidates
        merged = []
        for q in range(len(queries)):
            r = g_rows[:, q].reshape(-1).numpy()
            d = g_dist[:, q].reshape(-1).numpy()
            order = np.lexsort((r, d))[:k]
            merged.append((r[order], d[order]))
        want_rows, want_dist = brute_force_f64(corpus, queries, "l2", k)
        for q in range(len(queries)):
            assert np.array_equal(merged[q][0], want_rows[q]), (rank, q)
            assert np.array_equal(merged[q][1], want_dist[q])
        out[rank] = True
    finally:
        td.destroy_process_group()


def test_two_rank_gloo_gather_and_merge():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _replica_worker(rank, world, port, out):
    from oracle import brute_force_f64

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        corpus = rng.standard_normal((500, 8), dtype=np.float32)
        queries = rng.standard_normal((7, 8), dtype=np.float32)      # 7 queries over 2 ranks: slices of 4 and 3
        k = 5
        per = -(-len(queries) // world)
        lo, hi = fdist.shard_bounds(len(queries), world, rank)
        rows = np.full((per, k), -1, dtype=np.int64)
        dist = np.full((per, k), np.inf, dtype=np.float32)
        rows[: hi - lo], dist[: hi - lo] = brute_force_f64(corpus, queries[lo:hi], "dot", k)
        g_rows, g_dist = fdist.gather_slices(torch.from_numpy(rows), torch.from_numpy(dist), len(queries))
        want_rows, want_dist = brute_force_f64(corpus, queries, "dot", k)
        assert np.array_equal(g_rows.numpy(), want_rows) and np.array_equal(g_dist.numpy(), want_dist)
        out[rank] = True
    finally:
        td.destroy_process_group()


def test_two_rank_gloo_query_sharded_gather():
    """Replicated corpus, split query batch: the gathered slices are the whole batch in query order."""
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_replica_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _id_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        drawn = []

        def make_id():
            drawn.append(rank)
            return bytes(range(128))

        uid = fdist.broadcast_comm_id(make_id)
        out[rank] = (uid == bytes(range(128)), drawn)
    finally:
        td.destroy_process_group()


def test_two_rank_gloo_comm_id_broadcast():
    """The library's communicator id (fx_comm_unique_id) is drawn on rank 0 only and reaches every rank intact."""
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_id_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: (True, [0]), 1: (True, [])}
