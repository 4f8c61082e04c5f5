"""fx_merge_topk on one GPU: W sorted lists of k (distance, row) pairs per query -> top-k; shared-memory sort against the
rank merge (FENIX_MERGE_RANK_MIN), the C3 and C4 exchange shapes at 8 GPUs."""
import sys, os, time
import torch
sys.path.insert(0, os.getcwd())
from fenix_b200 import knn
ctx = knn.Context(0)
dev = torch.device("cuda", 0)
for (W, nq, k) in ((8, 4096, 10), (8, 10000, 100), (2, 10000, 100), (8, 64, 10)):
    g = torch.Generator(device=dev).manual_seed(1)
    dist = torch.rand((W, nq, k), generator=g, device=dev).sort(dim=2).values.contiguous()
    rows = (torch.arange(W, device=dev)[:, None, None] * 1_000_000 + torch.randint(0, 1_000_000, (W, nq, k), generator=g, device=dev)).to(torch.int64)
    rows = rows.sort(dim=2).values.contiguous()   # (ties are impossible here; any order of rows inside a list is fine as long as keys are sorted)
    out_r = torch.empty((nq, k), dtype=torch.int64, device=dev); out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    res = {}
    for name, knob in (("sort", 1 << 30), ("rank", 1)):
        ctx.set_option("FENIX_MERGE_RANK_MIN", knob)
        for _ in range(3):
            ctx.merge_topk_device(rows.data_ptr(), dist.data_ptr(), W, nq, k, out_r.data_ptr(), out_d.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        for _ in range(10):
            ctx.merge_topk_device(rows.data_ptr(), dist.data_ptr(), W, nq, k, out_r.data_ptr(), out_d.data_ptr())
        torch.cuda.synchronize()
        res[name] = ((time.perf_counter() - t0) / 10 * 1e3, out_r.clone(), out_d.clone())
    same = torch.equal(res["sort"][1], res["rank"][1]) and torch.equal(res["sort"][2], res["rank"][2])
    print(f"W={W} queries={nq} k={k}: sort {res['sort'][0]:.3f} ms, rank {res['rank'][0]:.3f} ms per call (incl. sync), identical {same}")
ctx.set_option("FENIX_MERGE_RANK_MIN", None)
