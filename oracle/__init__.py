"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's exact-search path.

Nothing under fenix_b200/ may import this package. Allowed users: tests/, the smoke() check in
__graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py.
"""
from .fenix_oracle import (  # noqa: F401
    DIST_COL, brute_force_f64, call, canonical, distance, from_arrow, search_rows,
)
