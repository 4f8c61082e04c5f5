import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenix_b200 import knn
ctx = knn.Context(0)
rng = np.random.default_rng(91)
corpus = rng.standard_normal((30000, 96), dtype=np.float32)
corpus[5000:5400] = corpus[17]
corpus[20000:20050] = corpus[17] + 1e-4
queries = np.concatenate([corpus[17:18], rng.standard_normal((40, 96), dtype=np.float32)])
c = knn.Corpus(ctx, len(corpus), 96); c.append(corpus); c.finalize()
for metric in ("l2", "cosine", "dot"):
    for nq in (1, 41):
        b = c.stats()
        rows, dist = c.search(queries[:nq], metric, 10, knn.PREC_FP32)
        a = c.stats()
        print(metric, nq, "refined", a.refined_queries - b.refined_queries, "fallback", a.fallback_queries - b.fallback_queries, rows[0][:6], dist[0][:4])
