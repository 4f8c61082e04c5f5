"""One sequential single-query Flight client in its own process (scripts/bench_flight.py spawns 16 of them):
argv = port, queries.npy, metric, k, requests, start (epoch seconds). Prints "first-request-start last-request-end"."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fenix_b200 as fx

port, q_path, metric, k, n_req, start = int(sys.argv[1]), sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6])
qs = np.load(q_path)
c = fx.Flight("127.0.0.1", port)
for j in range(5):
    c.search(qs[j], "c1", "vector", metric, select=["id"], maxval=k)
while time.time() < start:
    time.sleep(0.001)
t0 = time.time()
for j in range(n_req):
    c.search(qs[j % len(qs)], "c1", "vector", metric, select=["id"], maxval=k)
print(t0, time.time(), flush=True)
