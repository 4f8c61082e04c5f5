import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenix_b200 import knn
ctx = knn.Context(0)
for (n, d, nq, k) in [(50_000, 128, 200, 10), (100_000, 768, 300, 10)]:
    rng = np.random.default_rng(d)
    x = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((nq, d), dtype=np.float32)
    c = knn.Corpus(ctx, n, d); c.append(x); c.finalize()
    r_ex, _ = c.search(q, "dot", k, knn.PREC_EXACT_SCAN)
    r_tf, _ = c.search(q, "dot", k, knn.PREC_TF32)
    missing, found = [], []
    for a, b in zip(r_ex, r_tf):
        sb = set(b.tolist())
        for r in a:
            (found if r in sb else missing).append(int(r))
    missing, found = np.array(missing), np.array(found)
    print(f"n={n} d={d}: found {len(found)} missing {len(missing)}")
    for name, arr in (("found", found), ("missing", missing)):
        if len(arr) == 0: continue
        tile = arr // 256
        print(f"  {name}: tile%2 hist {np.bincount(tile % 2, minlength=2)}  half hist {np.bincount((arr % 256) // 128, minlength=2)} "
              f" tile%4 hist {np.bincount(tile % 4, minlength=4)} col%32 hist(first 8) {np.bincount(arr % 32, minlength=32)[:8]}")
    # per-query: does tf32 return garbage rows?
    print("  example exact:", r_ex[0], " tf32:", r_tf[0])
    c.close()
