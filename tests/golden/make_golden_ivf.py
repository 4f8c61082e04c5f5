"""Generate tests/golden/ivf_*.npz by running the LIVE reference's IVF path (read-only tree at /root/reference)
in the build container; the reference cannot travel to the GPU box, so its outputs are committed as small
fixtures and this script is their provenance.

    python tests/golden/make_golden_ivf.py

Two shims for library drift (the reference pins torch 2.1.2, this image has 2.11): `torch.load` defaults to
weights_only=True since 2.6 and refuses the pyarrow DataType the reference stores beside its codebook, and
`torch.compile` (coder.py:94) needs a working inductor C++ toolchain, which this container lacks - the eager
function it wraps is the reference's own arithmetic. Neither touches the reference's files.

Per case the fixture holds the corpus, the codebook the reference trained (its training is unseeded, so the
tensor itself is the fixture), the `__CODED_ID__` sidecar `index.make` wrote, the ranked probe codes of
`coder.call`, and what `index.call(coding=, probes=)` returned for every (metric, query, probes).
"""
from __future__ import annotations

import functools
import os
import shutil
import sys
import tempfile

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import table_of  # noqa: E402

CASES = {
    # name: (rows, dim, chunk, coding metric, codebook_size, num_codebooks, k, probes list, filter_mod)
    "l2_2x8": (6000, 32, 1000, "l2", 8, 2, 10, (1, 4, 16), 0),
    "cosine_1x32": (5000, 24, 700, "cosine", 32, 1, 7, (2, 8), 0),
    "dot_3x4_filtered": (4000, 16, 512, "dot", 4, 3, 5, (3, 12), 3),
}


def main() -> None:
    sys.path.insert(0, REF_SRC)
    sys.dont_write_bytecode = True
    import torch

    torch.load = functools.partial(torch.load, weights_only=False)
    torch.compile = lambda fn, **_kw: fn
    import fenix  # the live reference

    rng = np.random.default_rng(777)
    root = tempfile.mkdtemp(prefix="fenix_golden_ivf_")
    try:
        for name, (n, d, chunk, metric, ksize, nbooks, k, probe_list, fmod) in CASES.items():
            corpus = rng.standard_normal((n, d), dtype=np.float32)
            queries = rng.standard_normal((4, d), dtype=np.float32)
            fenix.io.table.make(root, name, table_of(corpus, chunk).to_reader())
            config = dict(metric=metric, codebook_size=ksize, num_codebooks=nbooks, batch_size=64, num_epochs=2)
            coding = fenix.io.coder.make(root, "cb_" + name, name, "vector", config)
            joined = fenix.io.index.make(root, "cb_" + name, name, "vector")
            flt = None
            if fmod:
                flt = (pc.field("id") - (pc.field("id") / fmod) * fmod) == 0
            payload = dict(corpus=corpus, queries=queries, chunk=np.int64(chunk), k=np.int64(k), filter_mod=np.int64(fmod),
                           tensor=coding["tensor"].numpy(), codebook_size=np.int64(ksize), num_codebooks=np.int64(nbooks),
                           coding_metric=np.array(metric), codes=joined.column("__CODED_ID__").to_numpy(),
                           probes=np.array(probe_list, dtype=np.int64),
                           ranked=fenix.io.coder.call(queries, ("" + root, "cb_" + name), None))
            for qi, q in enumerate(queries):
                for p in probe_list:
                    for m in (None, "l2", "cosine", "dot"):
                        res = fenix.io.index.call(root, "cb_" + name, name, "vector", q, metric=m, select=["id"],
                                                  filter=flt, maxval=k, probes=p)
                        key = f"{m or 'default'}:{qi}:{p}"
                        payload[key + ":id"] = res.column("id").to_numpy()
                        payload[key + ":dist"] = res.column("__DISTANCE__").to_numpy()
            np.savez_compressed(os.path.join(HERE, f"ivf_{name}.npz"), **payload)
            print(name, payload["tensor"].shape, np.bincount(payload["codes"]).tolist()[:12], payload["ranked"].shape)
        with open(os.path.join(HERE, "PROVENANCE.txt"), "a") as f:
            f.write(f"ivf_*.npz: tests/golden/make_golden_ivf.py, live reference, torch {torch.__version__} "
                    "(torch.load weights_only=False, torch.compile -> eager)\n")
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
