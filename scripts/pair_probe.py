"""CTA-pair (cta_group::2) streaming kernel against the fp64 scan and the one-CTA kernel on a few shapes.
Run it under `timeout`: a mis-synchronised pair would hang, not fail."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from fenix_b200 import knn  # noqa: E402
from fenix_b200.csrc.build import build  # noqa: E402

build()
ctx = knn.Context(0)
rng = np.random.default_rng(5)
ok = True
for n, d, nq, k, metric in [(20_000, 256, 300, 10, "dot"), (50_000, 768, 130, 10, "cosine"), (30_000, 200, 257, 37, "l2"),
                            (200_000, 768, 1024, 10, "cosine"), (40_000, 320, 513, 100, "l2")]:
    x = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((nq, d), dtype=np.float32)
    c = knn.Corpus(ctx, n, d)
    c.append(x)
    c.finalize()
    ctx.set_option("FENIX_TC_PAIR", 1)
    t0 = time.perf_counter()
    rows, dist = c.search(q, metric, k)
    dt = time.perf_counter() - t0
    st = c.stats()
    ctx.set_option("FENIX_TC_PAIR", 0)
    rows1, dist1 = c.search(q, metric, k)
    st1 = c.stats()
    ctx.set_option("FENIX_TC_PAIR", None)
    rows_s, dist_s = c.search(q, metric, k, knn.PREC_EXACT_SCAN)
    same = np.array_equal(rows, rows_s) and np.array_equal(dist, dist_s)
    same1 = np.array_equal(rows1, rows_s) and np.array_equal(dist1, dist_s)
    mask = (np.arange(n) % 3 != 0).astype(np.uint8)
    ctx.set_option("FENIX_TC_PAIR", 1)
    rows_m, dist_m = c.search(q, metric, k, row_mask=mask)
    ctx.set_option("FENIX_TC_PAIR", None)
    stm = c.stats()
    rows_ms, dist_ms = c.search(q, metric, k, knn.PREC_EXACT_SCAN, row_mask=mask)
    same_m = np.array_equal(rows_m, rows_ms) and np.array_equal(dist_m, dist_ms)
    print(f"n={n} d={d} nq={nq} k={k} {metric}: pair variant={st.last_variant} path={st.last_path} equal_scan={same} "
          f"kernel {st.last_main_kernel_ms:.3f} ms ({dt * 1e3:.1f} ms call) | one-CTA variant={st1.last_variant} equal_scan={same1} "
          f"kernel {st1.last_main_kernel_ms:.3f} ms | masked variant={stm.last_variant} equal_scan={same_m} "
          f"fallback={st.fallback_queries} refined={st.refined_queries}", flush=True)
    ok = ok and same and same1 and same_m and (st.last_variant & 4) == 4 and (st1.last_variant & 4) == 0
    c.close()
ctx.close()
print("PAIR PROBE", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
