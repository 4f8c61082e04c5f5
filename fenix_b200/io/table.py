"""Named tables = <root>/sources/<name>.arrow (mirrors fenix.io.table, table.py:9-56).

Only the read side (`load`, `join`) is on the search path; the write side exists so that the
Flight handlers (`do_put`, `drop-table`) behave as the reference's do and so that device
shards cached for a table are dropped when the file changes.
"""
from __future__ import annotations

import os
from typing import Iterator, Literal, Sequence

import pyarrow as pa

from . import arrow as _arrow

LOCATION: str = "sources"


def path_of(root: str, name: str) -> str:
    return os.path.join(root, LOCATION, name + ".arrow")


def load(root: str, name: str | Sequence[str]) -> pa.Table:
    if isinstance(name, str):
        return _arrow.load(path_of(root, name))
    if not isinstance(name, Sequence):
        raise AssertionError("source must be a table name or a sequence of names")
    return join(*(load(root, n) for n in name))


def make(root: str, name: str, data: pa.RecordBatchReader) -> pa.Table:
    from . import shards  # local import: shards imports this module

    table = _arrow.make(path_of(root, name), data)
    shards.invalidate(root, name)
    return table


def join(*data: pa.Table, axis: Literal[0, 1] = 0) -> pa.Table:
    if len(data) == 1:
        return data[0]
    if axis == 0:
        return pa.concat_tables(data)
    if axis == 1:
        return pa.table({c: t.column(c) for t in data for c in t.column_names})
    raise ValueError(f"axis must be 0 or 1, got {axis}")


def list(root: str) -> Iterator[str]:
    base = os.path.join(root, LOCATION)
    for dirpath, _dirs, files in os.walk(base):
        for f in sorted(files):
            if f.endswith(".arrow"):
                yield os.path.relpath(os.path.join(dirpath, f), base).removesuffix(".arrow")


def drop(root: str, name: str) -> None:
    from . import shards

    p = path_of(root, name)
    if os.path.exists(p):
        os.unlink(p)
    shards.invalidate(root, name)
