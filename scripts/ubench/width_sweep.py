"""Main filter kernel time of the CTA-pair streaming kernel against the one-CTA streaming kernel over row widths (1M rows,
4096 queries, cosine, k = 10): where does cta_group::2 start to pay?"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from fenix_b200 import knn
ctx = knn.Context(0)
dev = torch.device("cuda", 0)
n, nq, k = 1_000_000, 4096, 10
for d in (200, 256, 320, 384, 512, 768):
    g = torch.Generator(device=dev).manual_seed(d)
    c = knn.Corpus(ctx, n, d)
    for lo in range(0, n, 250_000):
        x = torch.randn((250_000, d), generator=g, device=dev)
        torch.cuda.synchronize()
        c.append_device(x.data_ptr(), 250_000)
    c.finalize()
    q = torch.randn((nq, d), generator=g, device=dev)
    rows = torch.empty((nq, k), dtype=torch.int64, device=dev); dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    m = knn.metric_code("cosine")
    res = {}
    for name, pair in (("pair", None), ("one-CTA", 0)):
        ctx.set_option("FENIX_TC_PAIR", pair)
        ts = []
        for i in range(6):
            c.search_device(q.data_ptr(), nq, m, k, knn.PREC_FP32, rows.data_ptr(), dist.data_ptr())
            st = c.stats()
            if i >= 2: ts.append(st.last_main_kernel_ms)
        res[name] = (float(np.mean(ts)), st.last_variant, rows.clone())
    ctx.set_option("FENIX_TC_PAIR", None)
    fl = 2.0 * nq * n * d
    print(f"D={d}: pair {res['pair'][0]:.3f} ms ({fl / res['pair'][0] / 1e9:.0f} TFLOP/s, variant {res['pair'][1]}) | one-CTA {res['one-CTA'][0]:.3f} ms "
          f"({fl / res['one-CTA'][0] / 1e9:.0f} TFLOP/s, variant {res['one-CTA'][1]}) | same ids {torch.equal(res['pair'][2], res['one-CTA'][2])}", flush=True)
    c.close()
