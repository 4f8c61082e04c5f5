"""Generate tests/golden/*.npz by running the LIVE reference (read-only tree at
/root/reference) in the build container. The reference cannot travel to the GPU box, so its
outputs are committed as small fixtures; this script is the provenance.

    python tests/golden/make_golden.py           # rewrites tests/golden/ref_*.npz

Each case stores the seeded inputs (corpus, queries), the call arguments and, per metric, what
`fenix.io.index.call` returned: the `id` column (= row positions, tests/test_flight.py:29-31
style) and `__DISTANCE__`, in the reference's own output order.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = ["cosine", "dot", "inner_product", "l2", "euclidean"]


def table_of(corpus: np.ndarray, chunk: int) -> pa.Table:
    n, d = corpus.shape
    batches = []
    for lo in range(0, n, chunk):
        x = np.ascontiguousarray(corpus[lo: lo + chunk])
        vec = pa.FixedSizeListArray.from_arrays(pa.array(x.reshape(-1)), list_size=d)
        ids = pa.array(np.arange(lo, lo + len(x), dtype=np.int64))
        batches.append(pa.record_batch([ids, vec], names=["id", "vector"]))
    return pa.Table.from_batches(batches)


def cases() -> dict:
    out = {}
    rng = np.random.default_rng(1234)
    # 1. gaussian, several queries
    out["gauss"] = dict(corpus=rng.standard_normal((4096, 64), dtype=np.float32),
                        queries=rng.standard_normal((6, 64), dtype=np.float32), chunk=512, k=10, filter_mod=0)
    # 2. the reference tests' clustered batches (tests/test_flight.py:21-22) with uniform queries (:104)
    parts = []
    for _ in range(3):
        x = rng.standard_normal((1000, 32), dtype=np.float32)
        parts.append(x + 10 * x[0, :])
    out["clustered"] = dict(corpus=np.concatenate(parts), queries=rng.random((4, 32), dtype=np.float32),
                            chunk=1000, k=10, filter_mod=0)
    # 3. duplicated rows (ties) + query equal to the duplicated row (distance 0 for l2)
    c = rng.standard_normal((1000, 16), dtype=np.float32)
    c[100:150] = c[7]
    out["ties"] = dict(corpus=c, queries=np.stack([c[7], c[500]]), chunk=250, k=8, filter_mod=0)
    # 4. odd widths (not multiples of 32 / 4) and k = 100
    out["d100"] = dict(corpus=rng.standard_normal((2500, 100), dtype=np.float32),
                       queries=rng.standard_normal((3, 100), dtype=np.float32), chunk=700, k=100, filter_mod=0)
    out["d7"] = dict(corpus=rng.standard_normal((600, 7), dtype=np.float32),
                     queries=rng.standard_normal((3, 7), dtype=np.float32), chunk=600, k=5, filter_mod=0)
    # 5. maxval >= N and maxval None: all rows, table order
    out["all_rows"] = dict(corpus=rng.standard_normal((300, 24), dtype=np.float32),
                           queries=rng.standard_normal((2, 24), dtype=np.float32), chunk=128, k=-1, filter_mod=0)
    out["k_ge_n"] = dict(corpus=rng.standard_normal((40, 24), dtype=np.float32),
                         queries=rng.standard_normal((2, 24), dtype=np.float32), chunk=16, k=64, filter_mod=0)
    # 6. predicate filter: id % 3 == 0 survive
    out["filtered"] = dict(corpus=rng.standard_normal((3000, 48), dtype=np.float32),
                           queries=rng.standard_normal((3, 48), dtype=np.float32), chunk=1000, k=7, filter_mod=3)
    # 7. a zero vector in the corpus and a zero query (cosine eps branch)
    z = rng.standard_normal((512, 20), dtype=np.float32)
    z[33] = 0
    out["zeros"] = dict(corpus=z, queries=np.stack([np.zeros(20, np.float32), z[5]]), chunk=512, k=6, filter_mod=0)
    return out


def main() -> None:
    sys.path.insert(0, REF_SRC)
    sys.dont_write_bytecode = True
    import fenix  # the live reference
    import torch

    root = tempfile.mkdtemp(prefix="fenix_golden_")
    try:
        for name, case in cases().items():
            table = table_of(case["corpus"], case["chunk"])
            fenix.io.table.make(root, name, table.to_reader())
            k = None if case["k"] < 0 else case["k"]
            flt = None
            if case["filter_mod"]:
                m = case["filter_mod"]
                flt = (pc.field("id") - (pc.field("id") / m) * m) == 0  # integer division: id % m == 0
            payload = dict(corpus=case["corpus"], queries=case["queries"], chunk=np.int64(case["chunk"]),
                           k=np.int64(case["k"]), filter_mod=np.int64(case["filter_mod"]))
            for metric in METRICS:
                for qi, q in enumerate(case["queries"]):
                    res = fenix.io.index.call(root, None, name, "vector", q, metric=metric, select=["id"],
                                              filter=flt, maxval=k)
                    assert res.column_names == ["id", "__DISTANCE__"]
                    payload[f"{metric}:{qi}:id"] = res.column("id").to_numpy()
                    payload[f"{metric}:{qi}:dist"] = res.column("__DISTANCE__").to_numpy()
            np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), **payload)
            print(name, {k_: v.shape for k_, v in payload.items() if k_.startswith("l2:0")})
        with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as f:
            f.write("generated by tests/golden/make_golden.py from the live reference at /root/reference\n")
            f.write(f"torch {torch.__version__}, pyarrow {pa.__version__}, numpy {np.__version__}, "
                    f"torch threads {torch.get_num_threads()}\n")
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
