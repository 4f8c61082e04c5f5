"""CPU tests of the host-side mirror of the reference interface (no device work)."""
import os
import pickle

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import fenix_b200 as fenix
from conftest import table_of
from fenix_b200.io import index as ix
from fenix_b200.io import shards


def test_public_surface_matches_reference():
    # src/fenix/__init__.py:1-2 and the Flight client methods (flight.py:137-292)
    for name in ("Flight", "Server", "io"):
        assert hasattr(fenix, name)
    for name in ("make_table", "read_table", "drop_table", "make_index", "sync_index", "drop_index", "search", "remove"):
        assert callable(getattr(fenix.Flight, name))
    import inspect

    sig = inspect.signature(fenix.Flight.search)
    assert list(sig.parameters) == ["self", "target", "source", "column", "metric", "coding", "select", "filter", "maxval", "probes"]
    sig = inspect.signature(fenix.io.index.call)
    assert list(sig.parameters)[:10] == ["root", "coding", "source", "column", "target", "metric", "select", "filter", "maxval", "probes"]
    assert ix.DIST_COL == "__DISTANCE__" and ix.CODE_COL == "__CODED_ID__"


def test_table_roundtrip_and_chunking(tmp_path):
    corpus = np.random.default_rng(0).standard_normal((2500, 12), dtype=np.float32)
    src = table_of(corpus, 1000)
    out = fenix.io.table.make(str(tmp_path), "test/table", src.to_reader())
    assert out == src
    assert os.path.exists(tmp_path / "sources" / "test" / "table.arrow")
    loaded = fenix.io.table.load(str(tmp_path), "test/table")
    assert loaded.column("vector").num_chunks == 3  # writer's batch boundaries survive
    both = fenix.io.table.load(str(tmp_path), ["test/table", "test/table"])
    assert both.num_rows == 5000
    assert sorted(fenix.io.table.list(str(tmp_path))) == ["test/table"]
    fenix.io.table.drop(str(tmp_path), "test/table")
    assert list(fenix.io.table.list(str(tmp_path))) == []


def test_chunk_rows_is_zero_copy_and_honours_offset():
    corpus = np.arange(40, dtype=np.float32).reshape(10, 4)
    arr = table_of(corpus, 10).column("vector").chunk(0)
    view = shards.chunk_rows(arr)
    assert view.shape == (10, 4) and np.array_equal(view, corpus)
    assert not view.flags.owndata
    sl = arr.slice(3, 4)
    assert np.array_equal(shards.chunk_rows(sl), corpus[3:7])


def test_chunk_rows_converts_other_float_widths():
    """float16 / float64 columns become float32 rows at upload (one conversion per table version); anything else
    is refused."""
    for typ, npt in ((pa.float64(), np.float64), (pa.float16(), np.float16)):
        x = np.arange(12, dtype=npt).reshape(3, 4)
        arr = pa.FixedSizeListArray.from_arrays(pa.array(x.reshape(-1), type=typ), 4)
        rows = shards.chunk_rows(arr.slice(1))
        assert rows.dtype == np.float32 and np.array_equal(rows, x[1:].astype(np.float32))
    ints = pa.FixedSizeListArray.from_arrays(pa.array(np.arange(8, dtype=np.int32)), 4)
    with pytest.raises(NotImplementedError):
        shards.chunk_rows(ints)


def test_coerce_target_forms():
    q = np.arange(6, dtype=np.float64)
    typ = pa.list_(pa.float32(), 6)
    forms = [q, q.astype(np.float32), pa.array(q), pa.chunked_array([pa.array(q[:2]), pa.array(q[2:])]),
             pa.scalar(q.astype(np.float32), type=typ)]
    import torch

    forms.append(torch.from_numpy(q))
    for f in forms:
        out = ix.coerce_target(f, 6)
        assert out.dtype == np.float32 and out.shape == (1, 6) and np.array_equal(out[0], q.astype(np.float32))
    batch = ix.coerce_target(np.ones((3, 6)), 6)
    assert batch.shape == (3, 6)
    with pytest.raises(pa.ArrowInvalid):
        ix.coerce_target(np.ones(5), 6)


def test_take_rows_over_column_kinds_and_slices():
    """The gather out of cached chunk views (numeric and fixed-size-list columns without nulls) and the one-row-slice
    fallback (nulls, strings, nested lists) both equal Table.take - also on a sliced table, whose chunks carry offsets."""
    rng = np.random.default_rng(3)
    n, chunk = 900, 50
    batches = []
    for lo in range(0, n, chunk):
        ids = np.arange(lo, lo + chunk, dtype=np.int64)
        vec = pa.FixedSizeListArray.from_arrays(pa.array(rng.standard_normal(chunk * 4).astype(np.float32)), 4)
        ivec = pa.FixedSizeListArray.from_arrays(pa.array(rng.integers(0, 9, chunk * 3).astype(np.int32)), 3)
        nullable = pa.array([None if i % 7 == 0 else int(i) for i in ids], type=pa.int32())
        text = pa.array([f"row{i}" for i in ids])
        f64 = pa.array(rng.standard_normal(chunk))
        ragged = pa.array([[int(i)] * (i % 3) for i in ids], type=pa.list_(pa.int64()))
        batches.append(pa.record_batch([pa.array(ids), vec, ivec, nullable, text, f64, ragged],
                                       names=["id", "vector", "ivec", "nullable", "text", "f64", "ragged"]))
    full = pa.Table.from_batches(batches)
    for t in (full, full.slice(37, 700)):
        rows = np.array([0, 5, t.num_rows - 1, 49, 50, 51, 333, 5], dtype=np.int64)
        got = ix.take_rows(t, t.column_names, rows)
        want = t.take(pa.array(rows))
        assert got.schema == want.schema
        assert got.combine_chunks() == want.combine_chunks()


def test_take_rows_equals_table_take():
    rng = np.random.default_rng(0)
    corpus = rng.standard_normal((1000, 6), dtype=np.float32)
    t = table_of(corpus, 64)
    for rows in ([5, 999, 0, 64, 63, 128, 5], [], [7], list(range(990, 1000))):
        rows = np.array(rows, dtype=np.int64)
        for cols in (["id", "vector"], ["vector"], ["id"]):
            got = ix.take_rows(t, cols, rows)
            want = t.select(cols).take(pa.array(rows, type=pa.int64()))
            assert got.schema == want.schema
            assert got.combine_chunks() == want.combine_chunks()
    big = rng.integers(0, 1000, 500)          # many rows: falls back to Table.take
    assert ix.take_rows(t, ["id", "vector"], big).combine_chunks() == t.take(pa.array(big)).combine_chunks()


def test_row_mask_follows_expression():
    corpus = np.zeros((10, 4), dtype=np.float32)
    t = table_of(corpus, 4)
    mask = ix._row_mask(t, "vector", pc.field("id") >= 6)
    assert mask.tolist() == [0] * 6 + [1] * 4


def test_ivf_search_needs_its_table_and_codebook(tmp_path):
    """`coding=` loads the table, the codebook and the sidecar like index.py:93-95; missing files surface as they do
    in the reference (FileNotFoundError), before anything touches the device."""
    with pytest.raises(FileNotFoundError):
        fenix.io.index.call(str(tmp_path), "some-coding", "t", "vector", np.zeros(4), metric="l2", probes=4)


def test_unknown_metric_is_value_error(tmp_path):
    with pytest.raises(ValueError):
        fenix.io.index.call(str(tmp_path), None, "t", "vector", np.zeros(4), metric="manhattan")


def test_search_command_wire_format(monkeypatch):
    """The descriptor is the reference's pickled dict (flight.py:258-271) and the request stream a
    one-column table named `target` (flight.py:279)."""
    import pyarrow.flight as fl

    seen = {}

    class FakeWriter:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def begin(self, schema):
            seen["schema"] = schema

        def write_table(self, t):
            seen["table"] = t

        def done_writing(self):
            pass

    class FakeReader:
        def read_all(self):
            return pa.table({"ok": [1]})

    class FakeConn:
        def do_exchange(self, descriptor):
            seen["descriptor"] = descriptor
            return FakeWriter(), FakeReader()

        def close(self):
            pass

    client = fenix.Flight()
    client.__dict__["conn"] = FakeConn()
    flt = pc.field("id") > 3
    client.search(np.arange(4, dtype=np.float32), "s", "vector", "l2", select=["id"], filter=flt, maxval=7)
    cmd = pickle.loads(seen["descriptor"].command)
    assert set(cmd) == {"coding", "source", "column", "metric", "select", "filter", "maxval", "probes"}
    assert cmd["source"] == "s" and cmd["maxval"] == 7 and cmd["coding"] is None
    assert pickle.loads(cmd["filter"]).equals(flt)
    assert seen["table"].column_names == ["target"] and seen["table"].num_rows == 4
    with pytest.raises(AssertionError):
        client.search(np.zeros(4), "s", "vector", "manhattan")
    # wire extension: 2-D target -> FixedSizeList column with Q rows
    client.search(np.zeros((3, 4), dtype=np.float32), "s", "vector", "l2", maxval=2)
    assert pa.types.is_fixed_size_list(seen["table"].schema.field("target").type) and seen["table"].num_rows == 3


def test_micro_batcher_coalesces_concurrent_requests():
    """Concurrent submissions under one key run as ONE batched call; a lone request runs at once."""
    import threading
    import time

    from fenix_b200.io.batcher import MicroBatcher

    calls = []

    def runner(qs):
        calls.append(len(qs))
        time.sleep(0.02)  # a slow "search" so that peers pile up behind the first leader
        return qs.sum(axis=1, keepdims=True).astype(np.int64), qs[:, :1].astype(np.float32)

    mb = MicroBatcher(max_wait_us=20000, max_batch=64)
    t0 = time.perf_counter()
    r, d = mb.submit("k", np.full(4, 7.0, np.float32), runner)
    assert time.perf_counter() - t0 < 0.035 and r[0] == 28 and calls == [1]   # lonely => no waiting
    out = {}

    def client(i):
        out[i] = mb.submit("k", np.full(4, float(i), np.float32), runner)

    threads = [threading.Thread(target=client, args=(i,)) for i in range(24)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert all(out[i][0][0] == 4 * i and out[i][1][0] == float(i) for i in range(24))
    assert sum(calls) == 25 and len(calls) < 12, calls     # far fewer calls than requests
    # failures propagate to every member of the batch
    def boom(qs):
        raise RuntimeError("device lost")
    with pytest.raises(RuntimeError):
        mb.submit("x", np.zeros(4, np.float32), boom)


def test_micro_batcher_caps_the_batch():
    """max_batch is a cap: a burst larger than it is served in several batches, the remainder under a promoted leader."""
    import threading
    import time

    from fenix_b200.io.batcher import MicroBatcher

    calls = []

    def runner(qs):
        calls.append(len(qs))
        time.sleep(0.01)
        return qs.sum(axis=1, keepdims=True).astype(np.int64), qs[:, :1].astype(np.float32)

    mb = MicroBatcher(max_wait_us=30000, max_batch=4)
    out = {}

    def client(i):
        out[i] = mb.submit("k", np.full(4, float(i), np.float32), runner)

    threads = [threading.Thread(target=client, args=(i,)) for i in range(19)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=20)
    assert not any(t.is_alive() for t in threads), "a request was never answered"
    assert all(out[i][0][0] == 4 * i for i in range(19))
    assert sum(calls) == 19 and max(calls) <= 4, calls


class _FakeCorpus:
    closed = 0

    def close(self):
        type(self).closed += 1


def test_shard_cache_is_single_flight_and_never_closes_a_leased_set(tmp_path, monkeypatch):
    """Concurrent cold requests share ONE upload; a set replaced by a new table version is retired and only closed
    when its last lease is returned (a closed set must never reach a request thread)."""
    import threading
    import time

    root = str(tmp_path)
    data = table_of(np.arange(40, dtype=np.float32).reshape(10, 4), 5)
    fenix.io.table.make(root, "t", data.to_reader())
    builds = []

    def fake_from_chunks(column, device_ids=None):
        builds.append(1)
        time.sleep(0.05)      # an upload takes a while: the other threads arrive meanwhile
        return shards.ShardSet(dim=4, n_rows=len(column), corpora=[_FakeCorpus()], bases=[0])

    monkeypatch.setattr(shards, "from_chunks", fake_from_chunks)
    _FakeCorpus.closed = 0
    shards.invalidate(root)
    table = shards.load_table(root, "t")
    got = []

    def request():
        got.append(shards.get(root, "t", "vector", table))

    threads = [threading.Thread(target=request) for _ in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert len(builds) == 1 and len({id(g) for g in got}) == 1
    first = got[0]
    assert first.corpora and first._users == 6
    # a new version of the table arrives while those six searches are still running
    fenix.io.table.make(root, "t", table_of(np.ones((7, 4), np.float32), 7).to_reader())
    assert _FakeCorpus.closed == 0 and first.corpora, "a leased set was closed under its users"
    with shards.get(root, "t", "vector", shards.load_table(root, "t")) as second:
        assert second is not first and second.n_rows == 7 and len(builds) == 2
    for _ in range(5):
        first.release()
    assert _FakeCorpus.closed == 0
    first.release()                       # the last lease: now the retired set is freed
    assert _FakeCorpus.closed == 1 and first.corpora == []
    shards.invalidate(root)               # idle sets are closed at once
    assert _FakeCorpus.closed == 2


def test_arrow_make_replaces_the_file_atomically(tmp_path):
    """A table already memory-mapped keeps reading its old version while (and after) a new one is written."""
    from fenix_b200.io import arrow as fx_arrow

    path = str(tmp_path / "sources" / "t.arrow")
    old = fx_arrow.make(path, table_of(np.arange(32, dtype=np.float32).reshape(8, 4), 4).to_reader())
    new = fx_arrow.make(path, table_of(np.full((3, 4), 9, np.float32), 3).to_reader())
    assert old.num_rows == 8 and old.column("id").to_pylist() == list(range(8))       # old mapping intact
    assert old.column("vector").chunk(1).values.to_numpy()[-1] == 31.0
    assert new.num_rows == 3 and fx_arrow.load(path).num_rows == 3
    assert os.listdir(os.path.dirname(path)) == ["t.arrow"]                            # no temp file left behind


def test_warm_spec(tmp_path, monkeypatch):
    """FENIX_WARM="table:column,table2": the named columns (or every vector column) are uploaded ahead of the first search."""
    root = str(tmp_path)
    fenix.io.table.make(root, "a", table_of(np.zeros((4, 4), np.float32), 4).to_reader())
    fenix.io.table.make(root, "b", table_of(np.zeros((6, 4), np.float32), 6).to_reader())
    seen = []

    def fake_from_chunks(column, device_ids=None):
        seen.append(len(column))
        return shards.ShardSet(dim=4, n_rows=len(column), corpora=[_FakeCorpus()], bases=[0])

    monkeypatch.setattr(shards, "from_chunks", fake_from_chunks)
    shards.invalidate(root)
    assert shards.warm(root, "a:vector, b, missing:vector") == ["a:vector", "b:vector"]
    assert seen == [4, 6]
    assert shards.warm(root, "a:vector") == ["a:vector"] and seen == [4, 6]           # cached: no second upload
    shards.invalidate(root)


def test_cells_csr_groups_rows_by_cell():
    """The inverted index fx_corpus_set_cells takes: rows grouped by cell, ascending inside a cell, empty cells allowed."""
    from fenix_b200 import knn

    cell = np.array([2, 0, 2, 5, 0, 2], dtype=np.int64)
    inv, off = knn.cells_csr(cell)
    assert inv.dtype == np.int32 and off.dtype == np.int64
    assert off.tolist() == [0, 2, 2, 5, 5, 5, 6]
    assert inv.tolist() == [1, 4, 0, 2, 5, 3]
    inv0, off0 = knn.cells_csr(np.zeros(0, np.int64))
    assert inv0.size == 0 and off0.tolist() == [0]
    with pytest.raises(ValueError):
        knn.cells_csr(np.array([0, -1]))


def test_batched_ivf_host_glue_maps_probe_codes_to_cells():
    """io.index._search_cells: composite probe codes -> dense cell numbers of the codes present in the sidecar (absent and
    repeated codes become -1), the inverted index installed once per (coding, sidecar version), `more_than` short-circuit."""
    calls = {"set": 0, "search": []}

    class FakeCorpus:
        n_rows = 8

        def set_cells(self, dense):
            calls["set"] += 1
            calls["dense"] = np.asarray(dense).tolist()
            return int(np.max(dense)) + 1

        def search_cells(self, queries, metric, k, probes, mask):
            calls["search"].append((np.asarray(probes).tolist(), k, None if mask is None else np.asarray(mask).tolist()))
            return np.zeros((len(queries), k), np.int64), np.zeros((len(queries), k), np.float32)

    class FakeShard:
        import threading as _t
        corpora = [FakeCorpus()]
        _cells = None
        _cells_lock = _t.Lock()

    from fenix_b200 import knn

    codes = np.array([70, 10, 70, 33, 10, 10, 70, 99], dtype=np.int64)       # cells: 10 -> 0, 33 -> 1, 70 -> 2, 99 -> 3
    queries = np.zeros((2, 4), np.float32)
    probe_codes = np.array([[70, 5, 70], [99, 33, 10]], dtype=np.int64)     # 5 is in no row; 70 repeats
    shard = FakeShard()
    out = ix._search_cells(shard, ("cb", 1), codes, queries, "l2", 3, probe_codes, None, knn.PREC_FP32)
    assert out is not None and calls["set"] == 1 and calls["dense"] == [2, 0, 2, 1, 0, 0, 2, 3]
    probes, k, mask = calls["search"][-1]
    assert k == 3 and mask is None
    assert sorted(probes[0]) == [-1, -1, 2] and probes[1] == [3, 1, 0]
    # same key: the inverted index is reused; another key replaces it
    ix._search_cells(shard, ("cb", 1), codes, queries, "l2", 3, probe_codes, None, knn.PREC_FP32)
    assert calls["set"] == 1
    ix._search_cells(shard, ("cb", 2), codes, queries, "l2", 3, probe_codes, None, knn.PREC_FP32)
    assert calls["set"] == 2
    # more_than: the first query's probed cells hold 3 rows (cell 2) -> nothing to select when maxval >= 3
    n_before = len(calls["search"])
    assert ix._search_cells(shard, ("cb", 2), codes, queries[:1], "l2", 3, probe_codes[:1], None, knn.PREC_FP32, more_than=3) is None
    assert len(calls["search"]) == n_before
    assert ix._search_cells(shard, ("cb", 2), codes, queries[:1], "l2", 2, probe_codes[:1], None, knn.PREC_FP32, more_than=2) is not None
    # shapes the one-launch path does not take
    assert ix._search_cells(shard, ("cb", 2), codes, queries, "l2", None, probe_codes, None, knn.PREC_FP32) is None
    assert ix._search_cells(shard, ("cb", 2), codes, queries, "l2", 3, probe_codes, None, knn.PREC_BF16) is None
