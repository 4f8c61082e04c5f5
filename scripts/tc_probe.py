"""GPU probe: tensor-core filter path vs the fp64 scan on a few shapes (prints path/fallback/timing)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fenix_b200 import knn

shapes = [(50_000, 128, 200, 10, "l2"), (50_000, 128, 200, 10, "cosine"), (50_000, 128, 200, 10, "dot"),
          (200_000, 96, 1000, 100, "dot"), (100_000, 768, 300, 10, "cosine"), (1_000_000, 128, 2048, 100, "l2"),
          (300_000, 100, 64, 10, "l2"), (100_003, 36, 7, 5, "cosine")]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
ctx = knn.Context(0)
for (n, d, nq, k, metric) in shapes:
    g = torch.Generator(device="cuda").manual_seed(n + d)
    x = torch.randn((n, d), generator=g, device="cuda")
    q = torch.randn((nq, d), generator=g, device="cuda").cpu().numpy()
    c = knn.Corpus(ctx, n, d)
    torch.cuda.synchronize()
    c.append_device(x.data_ptr(), n)
    c.finalize()
    t0 = time.perf_counter()
    r1, d1 = c.search(q, metric, k, knn.PREC_FP32)
    t1 = time.perf_counter()
    st = c.stats()
    r1b, d1b = c.search(q, metric, k, knn.PREC_FP32)
    st_b = c.stats()
    r2, d2 = c.search(q[: min(nq, 64)], metric, k, knn.PREC_EXACT_SCAN)
    m = min(nq, 64)
    same = np.array_equal(r1[:m], r2) and np.array_equal(d1[:m], d2)
    rt, dt = c.search(q, metric, k, knn.PREC_TF32)
    recall = np.mean([len(set(a) & set(b)) / k for a, b in zip(r1, rt)])
    print(f"n={n} d={d} q={nq} k={k} {metric}: path={st.last_path} fallback={st.fallback_queries}/{nq} "
          f"first={1e3*(t1-t0):.1f}ms search_ms={st_b.last_search_ms:.3f} kernel_ms={st_b.last_main_kernel_ms:.3f} "
          f"equal_scan={same} tf32_recall={recall:.4f}", flush=True)
    if not same:
        bad = [i for i in range(m) if not np.array_equal(r1[i], r2[i])]
        print("  mismatching queries:", bad[:10])
        i = bad[0] if bad else 0
        print("  tc  :", r1[i][:10], d1[i][:5])
        print("  scan:", r2[i][:10], d2[i][:5])
    c.close()
ctx.close()
