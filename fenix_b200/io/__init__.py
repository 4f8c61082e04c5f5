"""Host-side mirror of the reference's `fenix.io` namespace for the exact-search path
(src/fenix/io/__init__.py:1-2): `io.index.call` is the seam the Flight handler dispatches to."""
from . import arrow, batcher, coder, index, shards, table
from .index import call

__all__ = ["arrow", "batcher", "coder", "index", "shards", "table", "call"]
