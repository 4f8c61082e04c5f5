"""torchrun --nproc-per-node N scripts/dist_check.py : the sharded search of the C library (fx_search_sharded: query
all-gather, shard search, candidate all-gather over its own NCCL communicator, merge; one host synchronisation) equals the
unsharded search - device-resident and host-buffer forms, ragged batches, a row mask, an adversarial case that forces the
certificate-failure path (second exchange) on some ranks only, and k above the shared-memory merge limit."""
import os, sys
import numpy as np
import torch
import torch.distributed as td
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenix_b200 import knn
from fenix_b200.dist import ShardedSearcher, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
td.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(5)
n, d, nq = 200_003, 128, 301
corpus = rng.standard_normal((n, d), dtype=np.float32)
corpus[150_000] = corpus[7]
# 1,500 exact copies of one row inside the LAST rank's shard: queries near it overflow the refinement tier there
# (certificate failure -> fp64 scan on that rank only), which exercises the collective second exchange
corpus[n - 2000: n - 500] = corpus[n - 2001]
queries = np.concatenate([rng.standard_normal((nq - 2, d), dtype=np.float32), corpus[7:8], corpus[n - 2001: n - 2000] + np.float32(1e-3)])
ctx = knn.Context(local)
lo, hi = shard_bounds(n, world, rank)
shard = knn.Corpus(ctx, hi - lo, d, row_base=lo); shard.append(corpus[lo:hi]); shard.finalize()
searcher = ShardedSearcher(shard)
d_q = torch.from_numpy(queries).to(dev)
h_q = torch.from_numpy(queries).pin_memory()
whole = None
if rank == 0:
    whole = knn.Corpus(ctx, n, d); whole.append(corpus); whole.finalize()
ok = True
for metric, k in (("l2", 10), ("cosine", 100), ("dot", 10), ("l2", 1500)):
    m = knn.metric_code(metric)
    rows, dist = searcher.search_device(d_q, m, k)
    h_rows, h_dist = searcher.search_host(h_q, m, k, result_rank=0)
    st = shard.stats()
    if rank == 0:
        want_rows, want_dist = whole.search(queries, metric, k)
        same_dev = np.array_equal(rows.cpu().numpy(), want_rows) and np.array_equal(dist.cpu().numpy(), want_dist)
        same_host = np.array_equal(h_rows.numpy(), want_rows) and np.array_equal(h_dist.numpy(), want_dist)
        print(f"world={world} {metric} k={k}: device form == unsharded: {same_dev}; host form == unsharded: {same_host}; "
              f"exchange {st.last_exchange_ms:.3f} ms", flush=True)
        ok &= same_dev and same_host
    print(f"  rank {rank}: fallback_queries={st.fallback_queries} refined_queries={st.refined_queries}", flush=True)
# a handful of queries: every rank answers its shard with ONE direct-scan launch (direct_scan.cuh), then the same exchange
for nq_small in (1, 5):
    m = knn.metric_code("l2")
    rows, dist = searcher.search_device(d_q[:nq_small].contiguous(), m, 10)
    st = shard.stats()
    if rank == 0:
        want_rows, want_dist = whole.search(queries[:nq_small], "l2", 10)
        same = np.array_equal(rows.cpu().numpy(), want_rows) and np.array_equal(dist.cpu().numpy(), want_dist)
        print(f"world={world} l2 k=10, {nq_small} queries (shard path {st.last_path}): == unsharded: {same}", flush=True)
        ok &= same and st.last_path == 3
td.barrier()
if rank == 0:
    whole.close()
searcher.close()
shard.close(); ctx.close()
td.destroy_process_group()
print(f"rank {rank} DIST CHECK {'OK' if ok else 'FAILED'}", flush=True)
sys.exit(0 if ok else 1)
