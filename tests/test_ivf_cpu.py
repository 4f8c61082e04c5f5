"""IVF pre-filter without a GPU: the oracle's restatement of coder.call /
index.call(coding, probes) against the committed outputs of the live reference (tests/golden/ivf_*.npz,
provenance tests/golden/make_golden_ivf.py), and the host-side bookkeeping of the product code."""
import os

import numpy as np
import pyarrow as pa
import pytest
import torch

from conftest import GOLDEN, golden_filter, table_of
from oracle import CODE_COL, coder_call, ivf_call

IVF_CASES = sorted(f[len("ivf_"):-len(".npz")] for f in os.listdir(GOLDEN) if f.startswith("ivf_"))


def load_ivf(name):
    return np.load(os.path.join(GOLDEN, f"ivf_{name}.npz"))


def test_ivf_goldens_exist():
    assert len(IVF_CASES) >= 3


@pytest.mark.parametrize("case", IVF_CASES)
def test_oracle_ranks_composite_codes_like_the_reference(case):
    g = load_ivf(case)
    tensor = torch.from_numpy(g["tensor"])
    metric = str(g["coding_metric"])
    ranked = coder_call(torch.from_numpy(g["queries"]), tensor, metric, None).numpy()
    assert np.array_equal(ranked, g["ranked"])
    # the sidecar: every row's best composite code, assigned record batch by record batch (index.py:46-52)
    chunk, codes = int(g["chunk"]), []
    for lo in range(0, len(g["corpus"]), chunk):
        codes.append(coder_call(torch.from_numpy(g["corpus"][lo: lo + chunk]), tensor, metric, 1)[:, 0].numpy())
    assert np.array_equal(np.concatenate(codes), g["codes"])


@pytest.mark.parametrize("case", IVF_CASES)
def test_oracle_ivf_search_matches_the_reference(case):
    g = load_ivf(case)
    tensor = torch.from_numpy(g["tensor"])
    data = table_of(g["corpus"], int(g["chunk"])).append_column(CODE_COL, pa.array(g["codes"]))
    flt = golden_filter(int(g["filter_mod"]))
    for qi, q in enumerate(g["queries"]):
        for p in g["probes"]:
            for m in (None, "l2", "cosine", "dot"):
                got = ivf_call(data, "vector", q, tensor, str(g["coding_metric"]), metric=m, select=["id"], filter=flt,
                               maxval=int(g["k"]), probes=int(p))
                key = f"{m or 'default'}:{qi}:{int(p)}"
                assert np.array_equal(got.column("id").to_numpy(), g[key + ":id"]), key
                assert np.array_equal(got.column("__DISTANCE__").to_numpy(), g[key + ":dist"]), key


def test_composite_sums_follow_the_reference_layout():
    """Product-side bookkeeping (no device needed): codebook 0 is the most significant digit of a composite code
    and the sums accumulate in float32 in codebook order (coder.py:171-181)."""
    from fenix_b200.io.coder import composite_sums

    rng = np.random.default_rng(3)
    for n, k in ((1, 7), (2, 5), (3, 4)):
        d = rng.standard_normal((6, n, k)).astype(np.float32)
        got = composite_sums(d)
        want = torch.tensor(0)
        for j in range(n):
            grid = torch.arange(0, k).repeat_interleave(k ** (n - j - 1)).repeat(k ** j)
            want = want + torch.from_numpy(d)[:, j, grid]
        assert got.dtype == np.float32 and np.array_equal(got, want.numpy())


def test_index_sidecar_paths_and_listing(tmp_path):
    from fenix_b200.io import arrow as fx_arrow
    from fenix_b200.io import index as fx_index

    root = str(tmp_path)
    path = fx_index.sidecar_path(root, "cb", "test/table", "vector")
    assert path == os.path.join(root, "indexes", "test/table", "vector", "cb.arrow")
    batch = pa.record_batch([pa.array([1, 2, 3], type=pa.int64())], names=[fx_index.CODE_COL])
    fx_arrow.make(path, pa.RecordBatchReader.from_batches(batch.schema, [batch]))
    assert [*fx_index.list(root)] == [os.path.join("test/table", "vector", "cb")]
    fx_index.drop(root, "cb", "test/table", "vector")
    assert [*fx_index.list(root)] == []
