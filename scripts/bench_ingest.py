"""Ingest path: Arrow record batches -> fx_corpus_append (pinned ring -> HBM) -> finalize
(norms + bf16 shadow). Prints one JSON line: host-to-shard throughput for a pageable numpy source and for the chunks of
an Arrow table, and the finalize time."""
import json, os, sys, time
import numpy as np
import pyarrow as pa
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenix_b200 import knn
from fenix_b200.io import shards

N, D, CHUNK = 2_000_000, 256, 65_536
rng = np.random.default_rng(7)
corpus = rng.standard_normal((N, D), dtype=np.float32)
ctx = knn.Context(0)
out = {"rows": N, "dim": D, "bytes": corpus.nbytes}
for label in ("warmup", "numpy_pageable"):
    c = knn.Corpus(ctx, N, D)
    t0 = time.perf_counter()
    for lo in range(0, N, CHUNK):
        c.append(corpus[lo: lo + CHUNK])
    t1 = time.perf_counter()
    c.finalize()
    t2 = time.perf_counter()
    out[label] = {"append_s": t1 - t0, "append_GBps": corpus.nbytes / (t1 - t0) / 1e9, "finalize_s": t2 - t1}
    c.close()
batches = []
for lo in range(0, N, CHUNK):
    x = corpus[lo: lo + CHUNK]
    batches.append(pa.record_batch([pa.FixedSizeListArray.from_arrays(pa.array(x.reshape(-1)), D)], names=["vector"]))
table = pa.Table.from_batches(batches)
t0 = time.perf_counter()
s = shards.from_chunks(table.column("vector"))
t1 = time.perf_counter()
out["arrow_chunks_from_chunks"] = {"total_s": t1 - t0, "GBps_incl_finalize": corpus.nbytes / (t1 - t0) / 1e9}
s.close()
del out["warmup"]
print(json.dumps(out))
