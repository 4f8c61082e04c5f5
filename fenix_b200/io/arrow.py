"""Arrow IPC stream files: the on-disk corpus format shared with the reference.

Mirrors fenix.io.arrow (src/fenix/io/arrow/arrow.py:6-21): `make` writes one IPC *stream*
message per incoming record batch, `load` memory-maps the file and returns a zero-copy Table
whose chunk boundaries are the writer's batch boundaries.
"""
from __future__ import annotations

import os

import pyarrow as pa


def load(path: str) -> pa.Table:
    with pa.memory_map(path, "rb") as mm:
        return pa.ipc.open_stream(mm).read_all()


def make(path: str, data: pa.RecordBatchReader) -> pa.Table:
    if not path.endswith(".arrow"):
        raise AssertionError(f"table files end in .arrow: {path}")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with pa.OSFile(path, "wb") as sink, pa.ipc.new_stream(sink, data.schema) as out:
        for batch in data:
            out.write_batch(batch)
    return load(path)
