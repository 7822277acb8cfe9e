"""mom6_b200: B200-native (sm_100a) implementation of MOM6's split-explicit dycore hot path.

The product is the C-ABI library mom6_b200/libmom6cu.so (include/mom6cu.h); this package is the
thin Python harness used by tests and bench.py.
"""
from .api import Context, Mom6cuError, make_domain  # noqa: F401
