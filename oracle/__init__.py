"""CPU oracle for the goFeature search path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package (see oracle/gf_oracle.c for the full statement and the
parity pin status).
"""
from . import lib, model  # noqa: F401
from .lib import MODE_CANONICAL, MODE_F64, MODE_SEQ32  # noqa: F401
