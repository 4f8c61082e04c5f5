"""BASELINE.json configs[0]: exact L2 k=10 over 100k x 128 fp32, 1k queries through the Flight server.

Runs the fenix_b200 server + client over loopback (sequential single-query RPCs, as the reference serves
them, then the batched wire extension), and the same server class with io.index.call swapped for the oracle
port of the reference's CPU path (so both arms pay the same RPC cost). Prints one JSON line per arm."""
import argparse, json, os, sys, time, tempfile, shutil
import numpy as np
import pyarrow as pa
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--devices", type=int, default=1, help="GPUs the ONE server process row-shards the table over (FENIX_DEVICES=0..N-1): "
                "fx_group_search - a worker thread, a context and an NCCL communicator per device inside the library")
ap.add_argument("--rows", type=int, default=100_000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--metric", default="l2")
ap.add_argument("--batch", type=int, default=1000, help="queries per RPC of the batched wire extension arm")
ap.add_argument("--chunk", type=int, default=1000, help="rows per record batch of the table")
ap.add_argument("--no-reference", action="store_true")
args = ap.parse_args()
os.environ["FENIX_DEVICES"] = ",".join(str(i) for i in range(args.devices))
import fenix_b200 as fenix
from fenix_b200 import io as fio

N, D, K, NQ, CHUNK, METRIC = args.rows, args.dim, 10, max(1000, args.batch), args.chunk, args.metric
rng = np.random.default_rng(1001)
queries = np.random.default_rng(2001).standard_normal((NQ, D), dtype=np.float32)


def gen_batches():
    for lo in range(0, N, CHUNK):
        x = rng.standard_normal((min(CHUNK, N - lo), D), dtype=np.float32)
        yield pa.record_batch([pa.array(np.arange(lo, lo + len(x), dtype=np.int64)),
                               pa.FixedSizeListArray.from_arrays(pa.array(x.reshape(-1)), D)], names=["id", "vector"])


schema = pa.schema({"id": pa.int64(), "vector": pa.list_(pa.float32(), D)})
root = tempfile.mkdtemp(prefix="fenix_c1_")
port = 9311
server = fenix.Server(root, "127.0.0.1", port)
client = fenix.Flight("127.0.0.1", port)
t0 = time.perf_counter()
client.make_table("c1", pa.RecordBatchReader.from_batches(schema, gen_batches()))
t_put = time.perf_counter() - t0
t0 = time.perf_counter()
client.search(queries[0], "c1", "vector", METRIC, select=["id"], maxval=K)
t_first = time.perf_counter() - t0
print(json.dumps({"arm": "ingest", "devices": args.devices, "rows": N, "dim": D, "do_put_s": t_put,
                  "first_search_after_put_ms": t_first * 1e3,
                  "note": "do_put uploads the vector column(s) to the device shard(s) before it returns (FENIX_UPLOAD_ON_PUT=0: lazily, by the first search); "
                          "the first search only builds the bf16 shadow its metric streams"}), flush=True)

def run(label, n_q, fn):
    fn(0)
    lat = []
    t0 = time.perf_counter()
    for i in range(n_q):
        t1 = time.perf_counter(); fn(i); lat.append(time.perf_counter() - t1)
    dt = time.perf_counter() - t0
    lat = np.array(lat) * 1e3
    print(json.dumps({"arm": label, "queries": n_q, "qps": n_q / dt, "p50_ms": float(np.median(lat)), "p99_ms": float(np.percentile(lat, 99))}), flush=True)

ours = {}
def ours_single(i):
    ours[i] = client.search(queries[i], "c1", "vector", METRIC, select=["id"], maxval=K)
run("fenix_b200 Flight, 1 query per RPC (GPU)", NQ, ours_single)
run("fenix_b200 Flight, 1 query per RPC, default select (the 10 winning vectors are gathered and returned)", NQ,
    lambda i: client.search(queries[i], "c1", "vector", METRIC, maxval=K))
qb = queries[: args.batch]
client.search(qb, "c1", "vector", METRIC, select=["id"], maxval=K)   # warm-up (scratch growth for the batch)
t0 = time.perf_counter()
for _ in range(5):
    out = client.search(qb, "c1", "vector", METRIC, select=["id"], maxval=K)
dt = (time.perf_counter() - t0) / 5
print(json.dumps({"arm": f"fenix_b200 Flight, batched wire extension: {args.batch} queries in one RPC ({args.devices} GPU(s))",
                  "queries": args.batch, "qps": args.batch / dt, "ms_per_rpc": dt * 1e3}), flush=True)
# the batched answer against the sequential single-query answers (same server)
same_b = sum(out.filter(pa.compute.field("__QUERY__") == i).column("id").to_pylist() == ours[i].column("id").to_pylist() for i in range(0, min(args.batch, NQ), 41))
print(json.dumps({"batched_vs_single": f"{same_b}/{len(range(0, min(args.batch, NQ), 41))} sampled queries identical"}), flush=True)

# concurrent clients: 16 threads, each with its own connection, sequential single-query RPCs
import threading
import pyarrow.compute  # noqa: F401
def concurrent(label, n_threads=16, per_thread=125):
    clients = [fenix.Flight("127.0.0.1", port) for _ in range(n_threads)]
    for c in clients:
        c.search(queries[0], "c1", "vector", METRIC, select=["id"], maxval=K)
    res = {}
    def work(t):
        for j in range(per_thread):
            i = t * per_thread + j
            res[i] = clients[t].search(queries[i % NQ], "c1", "vector", METRIC, select=["id"], maxval=K)
    b0, r0 = fio.index._batcher.batches, fio.index._batcher.requests
    ths = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    nb, nr = fio.index._batcher.batches - b0, fio.index._batcher.requests - r0
    ok = all(sorted(res[i].column("id").to_pylist()) == sorted(ours[i % NQ].column("id").to_pylist()) for i in range(0, n_threads * per_thread, 37))
    print(json.dumps({"arm": label, "queries": n_threads * per_thread, "qps": n_threads * per_thread / dt,
                      "gpu_batches": nb, "mean_batch": (nr / nb if nb else None), "same_as_sequential": ok}), flush=True)
wait = fio.index._batcher.max_wait
fio.index._batcher.max_wait = 0.0
concurrent("fenix_b200 Flight, 16 concurrent clients, micro-batching OFF")
fio.index._batcher.max_wait = wait
concurrent("fenix_b200 Flight, 16 concurrent clients, micro-batching ON (300 us window)")

# The same with the 16 clients in their OWN processes: the threads above share this process's GIL with the server's
# handler threads (client-side pickling / Arrow glue is ~150 us of Python per request), so they measure the interpreter, not
# the server. Separate client processes leave the server's GIL to the server.
import subprocess
def concurrent_procs(label, n_procs=16, per_proc=250):
    q_path = os.path.join(root, "_queries.npy")
    np.save(q_path, queries)
    b0, r0 = fio.index._batcher.batches, fio.index._batcher.requests
    start = time.time() + 8.0           # (every child has imported pyarrow and warmed its connection by then)
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_flight_client.py")
    procs = [subprocess.Popen([sys.executable, script, str(port), q_path, METRIC, str(K), str(per_proc), repr(start)],
                              stdout=subprocess.PIPE, text=True) for _ in range(n_procs)]
    spans = []
    for p_ in procs:
        out_, _ = p_.communicate(timeout=300)
        a_, e_ = out_.split()[-2:]
        spans.append((float(a_), float(e_)))
    nb, nr = fio.index._batcher.batches - b0, fio.index._batcher.requests - r0
    dt = max(e for _, e in spans) - min(s_ for s_, _ in spans)
    print(json.dumps({"arm": label, "queries": n_procs * per_proc, "qps": n_procs * per_proc / dt,
                      "gpu_batches": nb, "mean_batch": (nr / nb if nb else None),
                      "late_starters": sum(1 for s_, _ in spans if s_ > start + 0.05)}), flush=True)

fio.index._batcher.max_wait = 0.0
concurrent_procs("fenix_b200 Flight, 16 client PROCESSES, micro-batching OFF")
fio.index._batcher.max_wait = wait
concurrent_procs("fenix_b200 Flight, 16 client PROCESSES, micro-batching ON (300 us window)")

# reference arm: same server, CPU path of the reference (oracle port) behind io.index.call
from oracle import call as oracle_call
real_call = fio.index.call
def cpu_call(root_, coding, source, column, target, metric=None, select=None, filter=None, maxval=None, probes=None):
    data = fio.table.load(root_, source)
    return oracle_call(data, column, target, metric, select=select, filter=filter, maxval=maxval)
fio.index.call = cpu_call
ref = {}
def ref_single(i):
    ref[i] = client.search(queries[i], "c1", "vector", METRIC, select=["id"], maxval=K)
n_ref = 0 if args.no_reference else (100 if N * D <= 2e7 else 8)
if n_ref:
    run("reference CPU path (oracle port) behind the same Flight server", n_ref, ref_single)
fio.index.call = real_call
if n_ref:
    same = sum(ours[i].column("id").to_pylist() == ref[i].column("id").to_pylist() for i in range(n_ref))
    print(json.dumps({"parity": f"{same}/{n_ref} queries return identical ids in identical order (served GPU path vs the reference's CPU path behind the same server)"}), flush=True)
client.remove(); server.shutdown(); shutil.rmtree(root, ignore_errors=True)
