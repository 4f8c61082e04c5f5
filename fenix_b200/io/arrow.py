"""Arrow IPC stream files: the on-disk corpus format shared with the reference.

Mirrors fenix.io.arrow (src/fenix/io/arrow/arrow.py:6-21): `make` writes one IPC *stream*
message per incoming record batch, `load` memory-maps the file and returns a zero-copy Table
whose chunk boundaries are the writer's batch boundaries. (One difference: the reference truncates and rewrites
the file in place; here the new version replaces the old one atomically.)
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
    # Written beside the target and renamed over it: tables already memory-mapped by running searches (the parsed-table
    # cache keeps them across requests) stay valid until their last reader lets go, instead of seeing truncated pages.
    tmp = f"{path}.{os.getpid()}.{id(data):x}.tmp"
    try:
        with pa.OSFile(tmp, "wb") as sink, pa.ipc.new_stream(sink, data.schema) as out:
            for batch in data:
                out.write_batch(batch)
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)
    return load(path)
