import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenix_b200 import knn
ctx = knn.Context(0)
kind = "bf16" if os.environ.get("FENIX_DEBUG_BF16") else "tf32"
for (n, d) in [(8192, 128), (8192, 32), (8192, 64), (8192, 768), (8192, 100)]:
    rng = np.random.default_rng(d)
    x = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((128, d), dtype=np.float32)
    c = knn.Corpus(ctx, n, d); c.append(x); c.finalize()
    s = c.debug_scores(q, "dot")
    ref = q.astype(np.float64) @ x[:256].astype(np.float64).T
    err = np.abs(s - ref)
    bound = np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(x[:256], axis=1)[None, :]
    print(f"{kind} d={d}: max abs err {err.max():.4g}  max err/(|q||x|) {np.max(err/bound):.3g}  frac bad(>2e-2*bound) {(err > 2e-2*bound).mean():.3f}")
    c.close()
