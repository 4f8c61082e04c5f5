import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from fenix_b200 import knn
rng = np.random.default_rng(1)
x = rng.standard_normal((100_000, 128), dtype=np.float32)
ctx = knn.Context(0); c = knn.Corpus(ctx, len(x), 128); c.append(x); c.finalize()
q = rng.standard_normal((8, 128), dtype=np.float32)
for nq in (1, 8):
    for prec, name in ((knn.PREC_FP32, "direct"), (knn.PREC_EXACT_SCAN, "scan")):
        for _ in range(5):
            c.search(q[:nq], "l2", 10, prec)
        print(name, nq, "queries: device", c.stats().last_search_ms * 1e3, "us")
ctx.set_option("FENIX_DEBUG_DIRECT", 1)
for i in range(4):
    c.search(q[:1], "l2", 10)
c.search(q[:8], "l2", 10)
ctx.set_option("FENIX_DEBUG_DIRECT", None)
