import os, sys, time, tempfile
import numpy as np, pyarrow as pa
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fenix_b200 as fenix
from fenix_b200.io import index as ix, shards
N, D, K = 100_000, 128, 10
rng = np.random.default_rng(1001)
corpus = rng.standard_normal((N, D), dtype=np.float32)
queries = rng.standard_normal((1000, D), dtype=np.float32)
batches = []
for lo in range(0, N, 1000):
    x = corpus[lo:lo + 1000]
    batches.append(pa.record_batch([pa.array(np.arange(lo, lo + 1000, dtype=np.int64)), pa.FixedSizeListArray.from_arrays(pa.array(x.reshape(-1)), D)], names=["id", "vector"]))
root = tempfile.mkdtemp()
fenix.io.table.make(root, "c1", pa.Table.from_batches(batches).to_reader())
for rep in range(3):
    t0 = time.perf_counter(); out = ix.call(root, None, "c1", "vector", queries[0], metric="l2", select=["id"], maxval=K); t1 = time.perf_counter()
    out = ix.call(root, None, "c1", "vector", queries, metric="l2", select=["id"], maxval=K); t2 = time.perf_counter()
    data = shards.load_table(root, "c1"); sh = shards.get(root, "c1", "vector", data)
    t3 = time.perf_counter(); r, d = sh.search(queries, "l2", K); t4 = time.perf_counter()
    st = sh.corpora[0].stats()
    print(f"rep {rep}: single call {1e3*(t1-t0):.2f} ms, batched call {1e3*(t2-t1):.2f} ms, raw search {1e3*(t4-t3):.2f} ms, device {st.last_search_ms:.3f} ms kernel {st.last_main_kernel_ms:.3f} path {st.last_path} refined {st.refined_queries} fallback {st.fallback_queries}")
