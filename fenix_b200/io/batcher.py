"""Server-side micro-batching of concurrent single-query searches.

The reference answers one query per `do_exchange` RPC (flight.py:62-77) and so pays a full pass over the
corpus per query. Flight handlers run on gRPC pool threads; when several single-query searches against the
same (table, column, metric, k) are in flight at once, the first one in becomes the leader, waits a few
hundred microseconds for peers, and issues ONE batched `fx_search` for all of them - the corpus is streamed
once for the whole group. A lone request never waits (no peers in flight => run immediately).
Tuning: FENIX_MICROBATCH_US (max wait, default 300; 0 disables), FENIX_MICROBATCH_MAX (default 256).
"""
from __future__ import annotations

import os
import threading
import time
from typing import Callable, Hashable

import numpy as np


class _Request:
    __slots__ = ("query", "done", "result", "error")

    def __init__(self, query: np.ndarray) -> None:
        self.query = query
        self.done = None        # an Event, created by a follower (under the batcher's lock) before it waits
        self.result = None
        self.error = None


_PROMOTED = object()   # marker handed to a follower that must take over as leader of the remaining group


class MicroBatcher:
    def __init__(self, max_wait_us: float | None = None, max_batch: int | None = None) -> None:
        self.max_wait = (float(os.environ.get("FENIX_MICROBATCH_US", 300)) if max_wait_us is None else max_wait_us) * 1e-6
        self.max_batch = int(os.environ.get("FENIX_MICROBATCH_MAX", 256)) if max_batch is None else max_batch
        self._lock = threading.Lock()
        self._pending: dict[Hashable, list[_Request]] = {}
        self._inflight = 0
        self.batches = 0
        self.requests = 0

    @property
    def enabled(self) -> bool:
        return self.max_wait > 0 and self.max_batch > 1

    def submit(self, key: Hashable, query: np.ndarray,
               runner: Callable[[np.ndarray], tuple[np.ndarray, np.ndarray]]) -> tuple[np.ndarray, np.ndarray]:
        """Run `runner(queries[B, D]) -> (rows[B, k], dist[B, k])` for this query together with any concurrent
        requests under the same key; returns this query's (rows[k], dist[k])."""
        req = _Request(query)
        with self._lock:
            self._inflight += 1
            group = self._pending.get(key)
            leader = group is None
            if leader:
                group = self._pending[key] = []
            group.append(req)
            if not leader:
                req.done = threading.Event()
            lonely = self._inflight == 1
        try:
            if not leader:
                req.done.wait()
                if req.error is _PROMOTED:       # the previous leader's batch was full: lead what is left
                    req.error = None
                    req.done = None
                    leader, lonely = True, False
            if leader:
                if not lonely:
                    deadline = time.perf_counter() + self.max_wait
                    while time.perf_counter() < deadline:
                        with self._lock:
                            if len(self._pending.get(key, ())) >= self.max_batch:
                                break
                        time.sleep(50e-6)
                with self._lock:
                    group = self._pending.pop(key)
                    # max_batch is a cap, not only a wake-up condition: the overflow of a burst stays pending under a
                    # new leader (its first member, woken below), so scratch buffers never grow past the cap
                    batch, rest = group[: self.max_batch], group[self.max_batch:]
                    if rest:
                        self._pending[key] = rest
                        promoted = rest[0]
                    else:
                        promoted = None
                if promoted is not None:
                    promoted.error = _PROMOTED
                    promoted.done.set()
                try:
                    rows, dist = runner(batch[0].query[None, :] if len(batch) == 1 else np.stack([r.query for r in batch]))
                    for i, r in enumerate(batch):
                        r.result = (rows[i], dist[i])
                except BaseException as exc:  # every member of the batch sees the failure
                    for r in batch:
                        r.error = exc
                finally:
                    with self._lock:
                        self.batches += 1
                        self.requests += len(batch)
                    for r in batch:
                        if r.done is not None:
                            r.done.set()
        finally:
            with self._lock:
                self._inflight -= 1
        if req.error is not None:
            raise req.error
        return req.result
