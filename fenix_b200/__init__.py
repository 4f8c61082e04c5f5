"""fenix_b200 - B200-native exact k-NN search path behind nrlugg/fenix's Python/Flight API.

Same public surface as the reference package for this path (src/fenix/__init__.py:1-2):
`Flight`, `Server`, `io`. The arithmetic runs in hand-written sm_100a CUDA behind the C ABI
of include/fenix_knn.h (`fenix_b200.knn`); there is no CPU fallback.
"""
from . import io, knn
from .flight import Flight, Server

__version__ = "0.1.0"
__all__ = ["Flight", "Server", "io", "knn"]
