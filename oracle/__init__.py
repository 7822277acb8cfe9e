"""TEST INFRASTRUCTURE ONLY: ctypes binding of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product package mom6_b200 never does.
"""
from .pyoracle import *  # noqa: F401,F403
