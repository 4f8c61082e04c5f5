import time, sys, os
sys.path.insert(0, os.getcwd())
t0 = time.time()
from fenix_b200 import knn
import numpy as np
t1 = time.time()
ctx = knn.Context(0)
t2 = time.time()
x = np.random.default_rng(0).standard_normal((100000, 128), dtype=np.float32)
c = knn.Corpus(ctx, len(x), 128); c.append(x); c.finalize()
t3 = time.time()
r = c.search(x[:1], "l2", 10)
t4 = time.time()
r = c.search(x[:100], "l2", 10)
t5 = time.time()
print(f"import {t1-t0:.2f}s fx_init {t2-t1:.2f}s upload+finalize {t3-t2:.2f}s first direct search {t4-t3:.3f}s first tc search {t5-t4:.3f}s")
