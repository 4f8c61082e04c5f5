"""torchrun --nproc-per-node N scripts/dist_check.py : sharded search (NCCL all-gather + merge) == unsharded."""
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
n, d, nq = 200_003, 128, 300
corpus = rng.standard_normal((n, d), dtype=np.float32)
corpus[150_000] = corpus[7]
queries = np.concatenate([rng.standard_normal((nq - 1, d), dtype=np.float32), corpus[7:8]])
ctx = knn.Context(local)
lo, hi = shard_bounds(n, world, rank)
shard = knn.Corpus(ctx, hi - lo, d, row_base=lo); shard.append(corpus[lo:hi]); shard.finalize()
searcher = ShardedSearcher(shard)
d_q = torch.from_numpy(queries).to(dev)
ok = True
for metric, k in (("l2", 10), ("cosine", 100), ("dot", 10)):
    rows, dist = searcher.search_device(d_q, knn.metric_code(metric), k)
    if rank == 0:
        whole = knn.Corpus(ctx, n, d); whole.append(corpus); whole.finalize()
        want_rows, want_dist = whole.search(queries, metric, k)
        same = np.array_equal(rows.cpu().numpy(), want_rows) and np.array_equal(dist.cpu().numpy(), want_dist)
        print(f"world={world} {metric} k={k}: sharded == unsharded: {same}", flush=True)
        ok &= same
        whole.close()
td.barrier()
shard.close(); ctx.close()
td.destroy_process_group()
sys.exit(0 if ok else 1)
