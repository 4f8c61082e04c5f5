import os, sys, time, tempfile
import numpy as np, pyarrow as pa
sys.path.insert(0, "/root/repo")
import fenix_b200 as fenix
from fenix_b200 import io as fio, knn
N, D, K = 100_000, 128, 10
rng = np.random.default_rng(1)
corpus = rng.standard_normal((N, D), dtype=np.float32)
batches = [pa.record_batch([pa.array(np.arange(lo, lo + 1000, dtype=np.int64)), pa.FixedSizeListArray.from_arrays(pa.array(corpus[lo:lo+1000].reshape(-1)), D)], names=["id", "vector"]) for lo in range(0, N, 1000)]
root = tempfile.mkdtemp()
fio.table.make(root, "c1", pa.Table.from_batches(batches).to_reader())
qs = rng.standard_normal((300, D), dtype=np.float32)
def timeit(fn, n=200):
    fn(0); fn(1)
    t = []
    for i in range(n):
        t0 = time.perf_counter(); fn(i); t.append(time.perf_counter() - t0)
    t = np.array(t) * 1e3
    return f"p50 {np.median(t):.3f} ms  p99 {np.percentile(t, 99):.3f} ms"
print("index.call select=[id]   ", timeit(lambda i: fio.index.call(root, None, "c1", "vector", qs[i], metric="l2", select=["id"], maxval=K)))
print("index.call default select", timeit(lambda i: fio.index.call(root, None, "c1", "vector", qs[i], metric="l2", maxval=K)))
data = fio.shards.load_table(root, "c1"); shard = fio.shards.get(root, "c1", "vector", data)
print("shard.search              ", timeit(lambda i: shard.search(qs[i:i+1], "l2", K)))
c = shard.corpora[0]
print("device ms of last search  ", c.stats().last_search_ms, "kernel", c.stats().last_main_kernel_ms)
