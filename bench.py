#!/usr/bin/env python
"""bench.py - exact k-NN QPS of the B200 path next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic queries against the resident
corpus. Default workload: BASELINE.json configs[2] "10M x 768 cosine, k=10, 4096-query batch,
row-sharded at 1/2/4/8 B200" - the shape north_star's target is stated on ("exact k-NN at
10M x 768 on 1 B200, near-linear QPS scaling to 8 GPUs"); it fits one GPU (30.7 GB), so the same
total workload runs at every N (strong scaling). Other configs: --config c2|c4|c5_B.
Launched by torchrun for N > 1 (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (N rows, D, metric, k, Q, cfg#)
    "c2": dict(n=1_000_000, d=128, metric="l2", k=100, q=10_000, num=2,
               label="SIFT-1M-shape 1M x 128 fp32 L2, k=100, 10k-query batch"),
    "c3": dict(n=10_000_000, d=768, metric="cosine", k=10, q=4096, num=3,
               label="10M x 768 cosine, k=10, 4096-query batch, row-sharded"),
    "c4": dict(n=100_000_000, d=96, metric="inner_product", k=100, q=10_000, num=4,
               label="100M x 96 inner product (Deep-100M shape), k=100, 10k-query batch"),
}
CONFIGS["c4s"] = dict(n=12_500_000, d=96, metric="inner_product", k=100, q=10_000, num=4,
                      label="one 8-GPU shard of C4: 12.5M x 96 inner product, k=100, 10k-query batch")
for _b in (1, 2, 4, 8, 16, 32, 64):
    CONFIGS[f"c5_{_b}"] = dict(n=10_000_000, d=768, metric="cosine", k=10, q=_b, num=5,
                               label=f"latency sweep batch {_b}, k=10 over 10M x 768")
GEN_CHUNK = 65_536  # rows per seeded generation chunk


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        p["_source"] = "measured"
        return p
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}


def ncu_traffic(config: str, world: int):
    """DRAM bytes (read + write) of the dominant kernel per launch, from the committed `ncu --set full` capture of
    this configuration on one GPU (profiles/ncu_traffic.json, written by scripts/summarise_profiles.py), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if world != 1 or not os.path.exists(path):
        return None, "no single-GPU ncu capture for this configuration"
    rec = json.load(open(path)).get(config)
    if not rec:
        return None, "no single-GPU ncu capture for this configuration"
    return rec["dram_bytes_per_launch"], f"profiles/{rec['capture']} (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"


def query_batch(cfg: dict) -> np.ndarray:
    rng = np.random.default_rng(2000 + cfg["num"])
    return rng.standard_normal((cfg["q"], cfg["d"]), dtype=np.float32)


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path, host cores only
# ------------------------------------------------------------------------------------------
def cpu_corpus_sample(cfg: dict, rows: int) -> np.ndarray:
    """First `rows` rows of the synthetic corpus, generated on the host (seed 1000 + cfg#)."""
    out = np.empty((rows, cfg["d"]), dtype=np.float32)
    for ci, lo in enumerate(range(0, rows, GEN_CHUNK)):
        rng = np.random.default_rng([1000 + cfg["num"], ci])
        hi = min(rows, lo + GEN_CHUNK)
        out[lo:hi] = rng.standard_normal((GEN_CHUNK, cfg["d"]), dtype=np.float32)[: hi - lo]
    return out


def time_reference(cfg: dict, steps: int, warmup: int, budget_s: float = 25.0) -> dict:
    """Times oracle.call (the restated fenix.io.index.call: per-chunk torch distance + Arrow
    select_k + take) one query at a time, as the reference serves them, on a bounded row sample;
    per-query cost is linear in N, so QPS at the full N = sample QPS * sample_rows / N."""
    import pyarrow as pa
    import torch

    from oracle import call as oracle_call

    sample_rows = min(cfg["n"], max(GEN_CHUNK, int(2.5e8 // cfg["d"])))  # ~1 GB of floats
    corpus = cpu_corpus_sample(cfg, sample_rows)
    batches = []
    for lo in range(0, sample_rows, GEN_CHUNK):
        x = corpus[lo: lo + GEN_CHUNK]
        vec = pa.FixedSizeListArray.from_arrays(pa.array(x.reshape(-1)), list_size=cfg["d"])
        batches.append(pa.record_batch([pa.array(np.arange(lo, lo + len(x), dtype=np.int64)), vec], names=["id", "vector"]))
    table = pa.Table.from_batches(batches)
    queries = query_batch(cfg)
    k = cfg["k"]

    def one(i: int) -> None:
        oracle_call(table, "vector", queries[i % len(queries)], cfg["metric"], select=["id"], maxval=k)

    t0 = time.perf_counter()
    one(0)
    per_query = time.perf_counter() - t0
    q_per_step = max(1, int(budget_s / max(per_query, 1e-6) / max(steps + warmup, 1)))
    q_per_step = min(q_per_step, 64)
    n = 0
    for _ in range(warmup):
        for _ in range(q_per_step):
            one(n); n += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        for _ in range(q_per_step):
            one(n); n += 1
    dt = time.perf_counter() - t0
    sample_qps = steps * q_per_step / dt
    scale = sample_rows / cfg["n"]
    # best case for the CPU: the same arithmetic (oracle.distance = the reference's torch formulas) as ONE batched call
    # over the whole sample instead of one call per chunk and query, plus torch.topk - separates library speed from
    # the reference's per-chunk, per-query overhead
    from oracle import distance as oracle_distance
    qb = torch.from_numpy(queries[: min(len(queries), 256)])
    xs = torch.from_numpy(corpus)
    t0 = time.perf_counter()
    torch.topk(oracle_distance(qb, xs, cfg["metric"]), min(k, sample_rows), largest=False)
    best_dt = time.perf_counter() - t0
    best_qps = len(qb) / best_dt * scale
    return dict(
        value=sample_qps * scale, unit="queries/s", cores=torch.get_num_threads(), kind="port",
        sample=(f"{steps} steps x {q_per_step} sequential single-query oracle.call (restated fenix.io.index.call, "
                f"{GEN_CHUNK}-row chunks, select=['id']) on the first {sample_rows} of {cfg['n']} rows; "
                f"measured {sample_qps:.3f} q/s on the sample, scaled by {scale:.4g} (cost linear in N)"),
        ms_per_step=dt / steps * 1e3, host_cpus=os.cpu_count(), arrow_threads=pa.cpu_count(),
        best_case={"value": best_qps, "unit": "queries/s",
                   "what": f"one batched torch call ({len(qb)} queries x {sample_rows} rows, the reference's formulas) + torch.topk, "
                           f"scaled by {scale:.4g}: library speed without the reference's per-chunk, per-query overhead"},
    )


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML from a thread, ~4 ms period, so
    even a 40 ms region gets samples; falls back to one nvidia-smi query when NVML is unavailable)."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu: int) -> None:
        self.gpu = gpu
        self.sm, self.bits, self.max_mhz = [], 0, None
        self._stop = None
        self._thread = None
        self._nvml = None

    def _loop(self, handle) -> None:
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(handle))
            except Exception:
                break
            self._stop.wait(0.004)

    def start(self) -> None:
        import threading

        try:
            import pynvml as nv

            nv.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: map the CUDA ordinal to the NVML device through its UUID
            import torch

            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            try:
                handle = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                handle = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self._nvml = nv
            self._stop = threading.Event()
            self._thread = threading.Thread(target=self._loop, args=(handle,), daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None

    def stop(self) -> dict:
        if self._nvml is None:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(t) for t in out.strip().split(",")[:2]]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": ["sampled once after the region (NVML unavailable)"], "samples": 1}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self._stop.set()
        self._thread.join(timeout=2)
        reasons = sorted(name for name, bit in self.REASONS.items() if self.bits & bit)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def build_shard(cfg: dict, ctx, lo: int, hi: int, device):
    """Rows [lo, hi) of the synthetic corpus, generated on the GPU in seeded 65,536-row chunks
    (content is independent of the world size) and appended device-to-device."""
    import torch

    from fenix_b200 import knn

    c = knn.Corpus(ctx, hi - lo, cfg["d"], row_base=lo)
    gen = torch.Generator(device=device)
    first, last = lo // GEN_CHUNK, (max(hi, lo + 1) - 1) // GEN_CHUNK
    for ci in range(first, last + 1):
        gen.manual_seed((1000 + cfg["num"]) * 1_000_003 + ci)
        block = torch.randn((GEN_CHUNK, cfg["d"]), generator=gen, device=device, dtype=torch.float32)
        a, b = max(lo, ci * GEN_CHUNK), min(hi, (ci + 1) * GEN_CHUNK)
        if b <= a:
            continue
        piece = block[a - ci * GEN_CHUNK: b - ci * GEN_CHUNK].contiguous()
        torch.cuda.synchronize(device)
        c.append_device(piece.data_ptr(), b - a)
    c.finalize()
    return c


def run_ours(args, cfg: dict) -> dict:
    import torch
    import torch.distributed as td

    from fenix_b200 import knn
    from fenix_b200.csrc.build import build as build_lib
    from fenix_b200.dist import ReplicaSearcher, ShardedSearcher, shard_bounds

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the search path has no CPU fallback")
    if rank == 0:
        build_lib()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        td.init_process_group("nccl", device_id=device)
        td.barrier()

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize(device)

    ctx = knn.Context(local)
    # N > 1: row-shard the corpus, or - small corpora with big batches - replicate it and split the query batch
    par = args.parallelism
    if par == "auto":
        par = "queries" if 8.0 * cfg["n"] * cfg["d"] <= 8e9 and cfg["q"] // max(world, 1) >= 256 else "rows"
    by_queries = par == "queries" and world > 1
    lo, hi = (0, cfg["n"]) if by_queries else shard_bounds(cfg["n"], world, rank)
    t_build = time.perf_counter()
    corpus = build_shard(cfg, ctx, lo, hi, device)
    t_build = time.perf_counter() - t_build
    searcher = ReplicaSearcher(corpus) if by_queries else ShardedSearcher(corpus)
    metric, k, n_q, d = knn.metric_code(cfg["metric"]), cfg["k"], cfg["q"], cfg["d"]
    prec = {"fp32": knn.PREC_FP32, "tf32": knn.PREC_TF32, "scan": knn.PREC_EXACT_SCAN}[args.precision]

    h_q = torch.from_numpy(query_batch(cfg)).pin_memory()
    d_q = h_q.to(device)
    h_rows = torch.empty((n_q, k), dtype=torch.int64).pin_memory()
    h_dist = torch.empty((n_q, k), dtype=torch.float32).pin_memory()

    # ---- device-resident timing (value) ----
    for _ in range(args.warmup):
        searcher.search_device(d_q, metric, k, prec)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    st0 = corpus.stats()
    kernel_ms, search_ms = [], []
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rows, dist = searcher.search_device(d_q, metric, k, prec)
        st = corpus.stats()
        kernel_ms.append(st.last_main_kernel_ms)
        search_ms.append(st.last_search_ms)
    barrier()
    elapsed = time.perf_counter() - t0
    st1 = corpus.stats()
    clocks = sampler.stop() if rank == 0 else None
    launches = st1.kernel_launches - st0.kernel_launches + (args.steps if world > 1 and not by_queries else 0)   # + merge kernels
    fallback = st1.fallback_queries - st0.fallback_queries
    refined = st1.refined_queries - st0.refined_queries

    # ---- end to end with HOST buffers: pinned queries in, pinned results out, copies inside the timed region.
    # N = 1: the reference-facing C-ABI call itself (fx_search on host pointers); N > 1: the sharded searcher
    # (H2D on every rank, shard search, NCCL all-gather, merge, D2H).
    def e2e_step():
        if world == 1:
            corpus.search_raw(h_q.data_ptr(), n_q, metric, k, prec, h_rows.data_ptr(), h_dist.data_ptr())
        else:
            if by_queries:   # each rank uploads its slice of the batch; the gathered result is read back on rank 0
                searcher.search_host(h_q, metric, k, prec, h_rows, h_dist, result_rank=0)
            else:
                searcher.search_host(h_q, metric, k, prec, h_rows, h_dist)

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    if world == 1:   # the two paths must agree (same kernels, different plumbing)
        assert torch.equal(h_rows, rows.cpu()) and torch.equal(h_dist, dist.cpu()), "e2e result differs from the device-resident result"

    def rank_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    elapsed = rank_max(elapsed)
    e2e_elapsed = rank_max(e2e_elapsed)
    k_ms = rank_max(float(np.mean(kernel_ms)))
    s_ms = rank_max(float(np.mean(search_ms)))

    # ---- optional recall of the tf32 mode against the exact result ----
    recall = None
    path = st1.last_path            # 0 scan, 1 TF32 filter, 2 bf16 filter
    if args.precision == "fp32" and path >= 1:
        approx = knn.PREC_BF16 if path == 2 else knn.PREC_TF32
        r_t, _ = searcher.search_device(d_q, metric, k, approx)
        a, b = rows.cpu().numpy(), r_t.cpu().numpy()
        recall = float(np.mean([len(set(x) & set(y)) / k for x, y in zip(a, b)]))

    out = None
    if rank == 0:
        pk = peaks()
        n_shard = hi - lo
        n_q_rank = -(-n_q // world) if by_queries else n_q     # queries one launch of this rank answers
        flops = 2.0 * n_q_rank * n_shard * d
        elem = 2.0 if path == 2 else 4.0     # the bf16 filter streams the bf16 shadow, otherwise the fp32 rows
        bytes_alg = elem * n_shard * d + 4.0 * n_q_rank * d + 12.0 * n_q_rank * k
        tensor_bound = path >= 1 and n_q_rank >= (420 if path == 2 else 210)
        if tensor_bound:
            achieved = flops / (k_ms * 1e-3) / 1e12
            peak = pk["bf16_tflops_sustained"]
            roof = dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=None,
                        peak_source=f"{pk['_source']} bf16 dense sustained (cuBLAS 8192^3 loop)" +
                                    ("" if path == 2 else "; this launch issues kind::tf32 MMAs whose nominal rate is half of "
                                     "bf16, so frac <= ~0.5 by construction"))
            if path == 1:
                roof["frac_of_tf32_nominal_half"] = achieved / (peak / 2)
        else:
            achieved = bytes_alg / (k_ms * 1e-3) / 1e9
            peak = pk["hbm_gbs"]
            roof = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None,
                        peak_source=f"{pk['_source']} copy bandwidth")
        variant = int(getattr(st1, "last_variant", 0))
        kernel_name = {0: "exact_scan_kernel", 1: "knn_tc_filter_kernel<epilogue, tf32> (streaming)",
                       2: "knn_rq_filter_kernel<epilogue> (resident-query, bf16)" if variant & 1
                       else "knn_tc_filter_kernel<epilogue, bf16> (streaming)"}[path]
        roof["traffic"], roof["traffic_source"] = ncu_traffic(args.config if not (args.rows or args.queries) else "", world)
        roof.update(kernel=kernel_name, sample_prepass=bool(variant & 2),
                    kernel_ms=k_ms, search_device_ms=s_ms,
                    algorithmic=dict(flops=flops, bytes=bytes_alg, per="launch (one query batch against this rank's shard)"))
        out = {
            "metric": "knn_qps", "value": n_q / (elapsed / args.steps), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.precision != "tf32" else "tf32", "data": "synthetic",
            "config": {
                "workload": f"{args.config}: {cfg['label']}", "n_rows": cfg["n"], "dim": d, "metric": cfg["metric"], "k": k,
                "queries_per_step": n_q, "precision_mode": args.precision, "parallelism": (f"replicated corpus, query batch split x{world}" if by_queries else f"row-shard x{world}"),
                "l2_policy": "corpus shard is larger than L2 (126 MB), no flush needed" if 4.0 * n_shard * d > 2.5e8
                else "corpus shard fits in L2: steady-state (warm L2) timing",
                "path": {0: "fp64 exact scan (CUDA cores)", 1: "tcgen05 TF32 filter + fp64 rerank + certificate",
                         2: "tcgen05 bf16-shadow filter + fp64 rerank + certificate"}[path],
                "refined_queries": int(refined), "fallback_queries": int(fallback), "corpus_build_s": t_build,
            },
            "clocks": clocks,
            "e2e": {"value": n_q / (e2e_elapsed / args.steps), "unit": "queries/s",
                    "api": "fx_search (C ABI, pinned host buffers)" if world == 1 else f"fenix_b200.dist.{type(searcher).__name__}.search_host",
                    "h2d_bytes_per_step": int(h_q.numel() * 4) * (1 if by_queries else world), "d2h_bytes_per_step": int(n_q * k * 12),
                    "ms_per_step": e2e_elapsed / args.steps * 1e3},
            "gpu_launches": int(launches),
            "roofline": roof,
            # `value` is the slower of the two clocks: the host clock around the K steps (barrier + device synchronise on
            # both sides, max over ranks) also pays the library's host-side work between its kernels; the CUDA-event time
            # is recorded by the library on its own stream around each search (max over ranks, mean over steps)
            "timing": {"value_from": "host clock between device-synchronised barriers, max over ranks",
                       "cuda_event_ms_per_step": s_ms, "host_clock_ms_per_step": elapsed / args.steps * 1e3},
        }
        if recall is not None:
            out["approx_mode_recall_at_k"] = {"mode": "bf16" if path == 2 else "tf32", "recall": recall}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = time_reference(cfg, steps=3, warmup=1, budget_s=20.0)
    corpus.close()
    ctx.close()
    if world > 1:
        td.barrier()
        td.destroy_process_group()
    return out


def run_reference(args, cfg: dict):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return None
    base = time_reference(cfg, steps=args.steps, warmup=args.warmup, budget_s=float(os.environ.get("FENIX_BENCH_BUDGET_S", 120.0)))
    return {
        "impl": "reference", "metric": "knn_qps", "value": base["value"], "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfg['label']}", "n_rows": cfg["n"], "dim": cfg["d"], "metric": cfg["metric"],
                   "k": cfg["k"], "queries_per_step": cfg["q"], "parallelism": "host cores only (rank 0)"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "best_case")},
        "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32", "scan"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parallelism", default="auto", choices=["auto", "rows", "queries"],
                    help="N > 1: row-shard the corpus, or replicate it and split the query batch (auto: replicate when the "
                         "corpus and its shadows take <= 8 GB and every rank still gets >= 256 queries)")
    ap.add_argument("--rows", type=int, default=0, help="tuning only: override the corpus row count of the config")
    ap.add_argument("--queries", type=int, default=0, help="tuning only: override the query batch of the config")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.rows or args.queries:
        cfg["n"] = args.rows or cfg["n"]
        cfg["q"] = args.queries or cfg["q"]
        cfg["label"] += f" [TUNING OVERRIDE rows={cfg['n']} queries={cfg['q']}: not a benchmark configuration]"
    out = run_reference(args, cfg) if args.impl == "reference" else run_ours(args, cfg)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
