"""The device's real power == the host libm's pow, bit for bit.

bt_rem_u = (SUM frhatu*visc_rem_u)**Instep (src/core/MOM_barotropic.F90:1502,1508) is the one operation of the hot path that IEEE 754 does not
pin; the reference's answer is its platform's libm pow().  mom6_b200/csrc/pow_glibc.cuh restates that routine (the ARM Optimized Routines pow
every glibc >= 2.28 ships, with the library's own tables: tools/gen_pow_tables.py) as a host/device function; tests/harness/pow_host.cpp compiles
the same header for the host and compares it with the running libm on seeded arguments:
  mode 0  x in (0, 1.25] (and 20 % tiny x down to 2^-60), y = 1/n, n = 1..400 -- what btstep passes (av_rem, 1/nstep)
  mode 1  x = 2^U(-60,60), y = U(-8,8)
  mode 2  x within 2^-50..2^-1 of 1, y = 1/n
The FMA variant must agree everywhere (x86-64 glibc dispatches to its FMA build on every CPU since 2013); the SSE2 variant of the same source
differs from it in ~0.06 % of the arguments, which is why "a correctly rounded pow" would not do.  MOM6CU_POW_BIG=1 runs 6e8 arguments
(the log of that run is profiles/r02_pow_glibc_6e8.log)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(ROOT, "tests", "harness", "pow_host.cpp")
    out = os.path.join(ROOT, "tests", "harness", "libpowhost.so")
    deps = [src, os.path.join(ROOT, "mom6_b200", "csrc", "pow_glibc.cuh"), os.path.join(ROOT, "mom6_b200", "csrc", "pow_glibc_tables.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", "-shared", "-fPIC", "-x", "c++", src, "-o", out])
    L = C.CDLL(out)
    L.pow_compare.argtypes = [C.c_int, C.c_ulonglong, C.c_longlong] + [C.POINTER(C.c_longlong)] * 3
    L.pow_one.restype = C.c_double
    L.pow_one.argtypes = [C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int)]
    return L


def _host_has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


@pytest.mark.skipif(not _host_has_fma(), reason="the host libm dispatches to its SSE2 build on this CPU")
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_bit_identical_to_libm(harness, mode):
    big = os.environ.get("MOM6CU_POW_BIG", "0") == "1"
    n = (400_000_000, 100_000_000, 100_000_000)[mode] if big else 6_000_000
    a, b, c = C.c_longlong(), C.c_longlong(), C.c_longlong()
    harness.pow_compare(mode, 20261017 + mode, n, a, b, c)
    print(f"mode {mode}: {n} arguments, mismatches FMA variant {a.value}, SSE2 variant {b.value}, outside the domain {c.value}")
    assert a.value == 0 and c.value == 0
    assert b.value > 0            # the two builds of the library do differ: the variant matters


@pytest.mark.skipif(not _host_has_fma(), reason="the host libm dispatches to its SSE2 build on this CPU")
def test_special_arguments(harness):
    ok = C.c_int()
    cases = [(1.0, 1.0 / 23.0), (0.5, 1.0), (1.0 - 2.0 ** -53, 1.0 / 23.0), (1.0 + 2.0 ** -52, 1.0 / 3.0), (5e-324, 0.5), (2.0 ** -1030, 1.0 / 7.0),
             (0.9999, 1.0 / 68.0), (1e-30, 1.0 / 400.0), (1.25, 1.0), (3.0, 2.0 ** -60)]
    for x, y in cases:
        got = harness.pow_one(1, x, y, C.byref(ok))
        assert ok.value == 1 and np.float64(got).view(np.int64) == np.float64(math.pow(x, y)).view(np.int64), (x, y)
    # outside the validated domain the routine says so instead of answering
    for x, y in [(-1.0, 0.5), (0.0, 0.5), (np.inf, 0.5), (2.0, 2.0 ** -70), (1e-300, 1.0), (1e300, 1.0)]:
        harness.pow_one(1, x, y, C.byref(ok))
        assert ok.value == 0 or (x, y) == (1e300, 1.0)
