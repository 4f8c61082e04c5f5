"""Parity at BASELINE.json's FULL sizes, through the C ABI, inside the GPU test run:

* C3 / C5: 10M x 768 cosine, k = 10 - the 4096-query batch (CTA-pair streaming kernel + sample prepass) and the
  latency-sweep batches 1 and 64 (one-CTA streaming kernel; batch 1 a second time through the captured CUDA graph);
* one 8-GPU shard of C4: 12.5M x 96 inner product, k = 100, 10k queries (resident-query kernel + sample prepass);
* C2: 1M x 128 L2, k = 100, 10k queries.

Each batch is checked like bench.py's `parity` block, with the same code: a sample of the batch against the fp64 CUDA-core
scan over the whole corpus AND against the oracle (the reference's per-chunk torch arithmetic + Arrow select_k_unstable)
over ALL rows, streamed to the host in 65,536-row chunks regenerated from the chunk seeds. Size-independent properties on
the whole batch: sorted by (distance, row), ids in range, no fallback to the scan on Gaussian data.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(built_library):
    import torch

    import bench
    from fenix_b200 import knn

    if torch.cuda.mem_get_info(0)[1] < 100e9:
        pytest.skip("needs a GPU with the HBM of a B200")
    ctx = knn.Context(0)
    yield bench, knn, ctx, torch.device("cuda", 0)
    ctx.close()


def check_batch(bench, knn, corpus, cfg, device, queries=None, n_scan=32, n_oracle=4, expect_variant=None):
    h_q = bench.query_batch(cfg) if queries is None else queries
    k = cfg["k"]
    before = corpus.stats()
    rows, dist = corpus.search(h_q, cfg["metric"], k)
    st = corpus.stats()
    assert st.last_path == 2, "expected the tcgen05 bf16-shadow path"
    if expect_variant is not None:
        assert st.last_variant == expect_variant, (st.last_variant, expect_variant)
    assert st.fallback_queries == before.fallback_queries, "Gaussian data must not need the fp64 scan"
    # size-independent properties over the WHOLE batch
    assert rows.min() >= 0 and rows.max() < cfg["n"]
    assert (np.diff(dist, axis=1) >= 0).all()
    ties = np.diff(dist, axis=1) == 0
    assert (np.diff(rows, axis=1)[ties] > 0).all()

    def scan(qs):
        return corpus.search(qs, cfg["metric"], k, knn.PREC_EXACT_SCAN)

    block = bench.parity_block(cfg, scan, rows, dist, h_q, device, 0, n_scan=n_scan, n_oracle=n_oracle)
    assert block["scan_agrees"], block
    assert block["within_parity_bar"] and block["ok"], block
    return rows, dist, block


def test_c3_and_c5_at_full_size(env):
    bench, knn, ctx, device = env
    cfg = dict(bench.CONFIGS["c3"])
    corpus = bench.build_shard(cfg, ctx, 0, cfg["n"], device)
    try:
        check_batch(bench, knn, corpus, cfg, device, expect_variant=4 | 2)          # CTA pairs + sample prepass
        for b in (64, 1):
            c5 = dict(bench.CONFIGS[f"c5_{b}"], data_num=cfg["num"])
            rows, dist, _ = check_batch(bench, knn, corpus, c5, device, n_scan=8, n_oracle=2)
            if b == 1:   # calls 2 and 3 of the same shape: captured, then replayed as one CUDA graph - same answer
                for _ in range(2):
                    r2, d2 = corpus.search(bench.query_batch(c5), c5["metric"], c5["k"])
                    assert np.array_equal(rows, r2) and np.array_equal(dist, d2)
    finally:
        corpus.close()


def test_c4_shard_at_full_size(env):
    bench, knn, ctx, device = env
    cfg = dict(bench.CONFIGS["c4s"])
    corpus = bench.build_shard(cfg, ctx, 0, cfg["n"], device)
    try:
        check_batch(bench, knn, corpus, cfg, device, n_oracle=2, expect_variant=1 | 2)   # resident-query kernel + sample prepass
    finally:
        corpus.close()


def test_c2_at_full_size(env):
    bench, knn, ctx, device = env
    cfg = dict(bench.CONFIGS["c2"])
    corpus = bench.build_shard(cfg, ctx, 0, cfg["n"], device)
    try:
        check_batch(bench, knn, corpus, cfg, device, expect_variant=1 | 2)
    finally:
        corpus.close()
