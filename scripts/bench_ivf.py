"""Batched IVF search on one GPU: ONE fx_search_cells launch for the query batch (every query scans the posting lists of
its probe cells) against the per-query loop it replaces (one masked fx_search per query, each a pass over the shard).
Synthetic: N x D Gaussian rows, cells = nearest of C random centroids (host), probes = the P nearest centroids per query."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenix_b200 import knn

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000); ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--cells", type=int, default=1024); ap.add_argument("--probes", type=int, default=8)
ap.add_argument("--queries", type=int, default=256); ap.add_argument("--k", type=int, default=10)
ap.add_argument("--loop-queries", type=int, default=32, help="queries timed through the per-query masked loop")
args = ap.parse_args()
rng = np.random.default_rng(5)
x = rng.standard_normal((args.rows, args.dim), dtype=np.float32)
cent = rng.standard_normal((args.cells, args.dim), dtype=np.float32)
q = rng.standard_normal((args.queries, args.dim), dtype=np.float32)


def nearest(a, b, p):   # the p nearest rows of b for every row of a (L2), host
    out = np.empty((len(a), p), np.int64)
    for lo in range(0, len(a), 65536):
        d = -2.0 * a[lo:lo + 65536] @ b.T + (b * b).sum(1)[None, :]
        out[lo:lo + 65536] = np.argpartition(d, p - 1, axis=1)[:, :p] if p < len(b) else np.argsort(d, axis=1)
    return out


cell = nearest(x, cent, 1)[:, 0]
probes = nearest(q, cent, args.probes).astype(np.int32)
ctx = knn.Context(0)
c = knn.Corpus(ctx, args.rows, args.dim); c.append(x); c.finalize()
c.set_cells(cell)
sizes = np.bincount(cell, minlength=args.cells)
probed_rows = float(sizes[probes].sum(1).mean())
for _ in range(3):
    rows, dist = c.search_cells(q, "l2", args.k, probes)
t = []
for _ in range(10):
    t0 = time.perf_counter(); rows, dist = c.search_cells(q, "l2", args.k, probes); t.append(time.perf_counter() - t0)
one = float(np.median(t))
dev_ms = c.stats().last_search_ms
# the loop it replaces
sub = min(args.loop_queries, args.queries)
masks = [np.isin(cell, probes[i]).astype(np.uint8) for i in range(sub)]
c.search(q[0], "l2", args.k, row_mask=masks[0])
t0 = time.perf_counter()
same = True
for i in range(sub):
    r, d = c.search(q[i], "l2", args.k, row_mask=masks[i])
    same = same and np.array_equal(r[0], rows[i]) and np.array_equal(d[0], dist[i])
loop = (time.perf_counter() - t0) / sub * args.queries
print(json.dumps({"workload": f"batched IVF: {args.rows} x {args.dim} L2, {args.cells} cells, {args.probes} probes, k={args.k}, {args.queries} queries per batch",
                  "probed_rows_per_query": probed_rows, "selectivity": probed_rows / args.rows,
                  "one_launch_ms_per_batch": one * 1e3, "one_launch_device_ms": dev_ms, "one_launch_qps": args.queries / one,
                  "probed_bytes_per_s_GB": probed_rows * args.dim * 4 * args.queries / (dev_ms * 1e-3) / 1e9,
                  "masked_loop_ms_per_batch (extrapolated from %d queries)" % sub: loop * 1e3, "masked_loop_qps": args.queries / loop,
                  "speedup": loop / one, "identical_results": bool(same)}))
