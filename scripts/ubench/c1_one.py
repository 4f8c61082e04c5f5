import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from fenix_b200 import knn
rng = np.random.default_rng(1)
x = rng.standard_normal((100_000, 128), dtype=np.float32)
ctx = knn.Context(0); c = knn.Corpus(ctx, len(x), 128); c.append(x); c.finalize()
q = rng.standard_normal((8, 128), dtype=np.float32)
for _ in range(6):
    c.search(q[:1], "l2", 10)
print(c.stats().last_search_ms * 1e3)
