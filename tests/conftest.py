import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
METRICS = ["cosine", "dot", "inner_product", "l2", "euclidean"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_library():
    """Build libfenix_knn.so in-tree (nvcc cross-compiles without a GPU)."""
    from fenix_b200.csrc.build import build

    return build()


def golden_cases():
    return sorted(f[len("ref_"):-len(".npz")] for f in os.listdir(GOLDEN) if f.startswith("ref_") and f.endswith(".npz"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))


def table_of(corpus, chunk):
    import pyarrow as pa

    n, d = corpus.shape
    batches = []
    for lo in range(0, max(n, 1), chunk):
        x = np.ascontiguousarray(corpus[lo: lo + chunk])
        vec = pa.FixedSizeListArray.from_arrays(pa.array(x.reshape(-1)), list_size=d)
        ids = pa.array(np.arange(lo, lo + len(x), dtype=np.int64))
        batches.append(pa.record_batch([ids, vec], names=["id", "vector"]))
    schema = pa.schema({"id": pa.int64(), "vector": pa.list_(pa.float32(), d)})
    return pa.Table.from_batches(batches, schema)


def golden_filter(mod):
    import pyarrow.compute as pc

    if not mod:
        return None
    return (pc.field("id") - (pc.field("id") / mod) * mod) == 0


def assert_same_neighbours(got_rows, got_dist, ref_rows, ref_dist, corpus=None, query=None, metric=None,
                           rtol=1e-5):
    """Parity bar (BASELINE.json north_star): ids bit-exact under (distance, row) order, distances
    within 1e-5 relative. Both sides are first put in canonical order. An id mismatch is accepted
    only when it is a genuine tie inside the tolerance band: the two rows' reference distances
    differ by less than rtol (the reference's own fp32 GEMM-form rounding noise)."""
    from oracle import canonical

    got_rows, got_dist = canonical(np.asarray(got_rows), np.asarray(got_dist))
    ref_rows, ref_dist = canonical(np.asarray(ref_rows), np.asarray(ref_dist))
    assert got_rows.shape == ref_rows.shape, (got_rows.shape, ref_rows.shape)
    scale = np.maximum(np.abs(ref_dist), 1e-30)
    floor = 0.0
    if corpus is not None and query is not None and metric in ("l2", "euclidean"):
        # absolute floor for GEMM-form cancellation in the REFERENCE: d = sqrt(d2), d2 carries
        # ~eps32 * (|q|^2 + |x|^2) absolute error
        q2 = float(np.dot(query.astype(np.float64), query.astype(np.float64)))
        x2 = float((corpus.astype(np.float64) ** 2).sum(1).max())
        floor = np.sqrt(2.0 ** -22 * (q2 + x2))
    tol = np.maximum(rtol * scale, rtol * floor * 10 + 0.0)
    if metric in ("cosine",):
        tol = np.maximum(tol, 2e-7)  # 0.5 - 0.5*cos: fp32 rounding of the reference near 0/0.5
    if metric in ("dot", "inner_product") and corpus is not None and query is not None:
        # fp32 accumulation noise of the reference's sgemv when the dot product cancels to ~0
        qn = float(np.linalg.norm(query.astype(np.float64)))
        xn = float(np.sqrt((corpus.astype(np.float64) ** 2).sum(1).max()))
        tol = np.maximum(tol, 2.0 ** -22 * qn * xn)
    bad = np.abs(got_dist - ref_dist) > tol
    assert not bad.any(), f"distance mismatch: got {got_dist[bad][:5]} ref {ref_dist[bad][:5]} tol {tol[bad][:5]}"
    diff = got_rows != ref_rows
    if diff.any():
        # permitted only as a permutation inside tie bands of the reference distances
        assert sorted(got_rows.tolist()) == sorted(ref_rows.tolist()) or _boundary_tie(got_rows, ref_rows, got_dist, ref_dist, tol), \
            f"id mismatch: got {got_rows[diff][:8]} ref {ref_rows[diff][:8]}"
        for i in np.nonzero(diff)[0]:
            j = np.nonzero(ref_rows == got_rows[i])[0]
            if len(j):
                assert abs(ref_dist[j[0]] - ref_dist[i]) <= tol[i] * 2, f"id swap outside tie band at rank {i}"


def _boundary_tie(got_rows, ref_rows, got_dist, ref_dist, tol):
    # differing membership is only legal at the k-th boundary when the boundary distances tie
    return abs(float(got_dist[-1]) - float(ref_dist[-1])) <= float(tol[-1]) * 2
