"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's search path (exact, and the IVF pre-filter).

Nothing under fenix_b200/ may import this package. Allowed users: tests/, the smoke() check in
__graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py.
"""
from .fenix_oracle import (  # noqa: F401
    CODE_COL, DIST_COL, brute_force_f64, call, canonical, coder_call, distance, from_arrow, ivf_call, search_rows,
)
