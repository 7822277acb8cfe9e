"""The column code the GPU threads run (mom6_b200/csrc/remap_column.cuh), compiled as plain C++ by a test-only harness,
against the oracle: bitwise on every scheme / option combination and on ragged, vanished-layer and 1-layer columns."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from remap_cases import columns, cs_variants, SHAPES, KINDS

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("remap_host") / "libremap_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(HERE, "harness", "remap_host.cpp"), "-o", out])
    lib = C.CDLL(out)
    vp = C.c_void_p
    lib.remap_host_batch.argtypes = [C.c_int] * 5 + [C.c_double] * 2 + [C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
    lib.remap_host_batch_stream.argtypes = lib.remap_host_batch.argtypes
    return lib


def test_column_code_matches_oracle(harness, oracle):
    rng = np.random.default_rng(11)
    nbad = 0
    for (n0, n1) in SHAPES:
        for kind in KINDS:
            h0, u0, h1 = columns(rng, 6, n0, n1, kind)
            for cs in cs_variants():
                got = np.zeros_like(h1)
                rc = harness.remap_host_batch(cs["remapping_scheme"], cs["boundary_extrapolation"], cs["force_bounds_in_subcell"],
                                              cs["force_bounds_in_target"], cs["om4_remap_via_sub_cells"], cs["h_neglect"],
                                              cs["h_neglect_edge"], 6, n0, h0.ctypes.data, u0.ctypes.data, n1, h1.ctypes.data,
                                              got.ctypes.data)
                assert rc == 0
                for c in range(6):
                    ref, _ = oracle.remapping_core_h(cs, h0[c], u0[c], h1[c])
                    if not np.array_equal(ref, got[c], equal_nan=True):
                        nbad += 1
                        if nbad < 5:
                            print("MISMATCH", n0, n1, kind, cs, np.abs(ref - got[c]).max())
    assert nbad == 0


def test_streaming_column_code_matches_oracle(harness, oracle):
    """remap_stream.cuh (what the device runs for PCM / PLM / PPM_H4) against the oracle, bit for bit."""
    rng = np.random.default_rng(12)
    nbad = 0
    for (n0, n1) in SHAPES:
        for kind in KINDS:
            h0, u0, h1 = columns(rng, 6, n0, n1, kind)
            for cs in cs_variants():
                if cs["remapping_scheme"] == 5:
                    continue
                got = np.zeros_like(h1)
                rc = harness.remap_host_batch_stream(cs["remapping_scheme"], cs["boundary_extrapolation"], cs["force_bounds_in_subcell"],
                                                     cs["force_bounds_in_target"], cs["om4_remap_via_sub_cells"], cs["h_neglect"],
                                                     cs["h_neglect_edge"], 6, n0, h0.ctypes.data, u0.ctypes.data, n1, h1.ctypes.data,
                                                     got.ctypes.data)
                assert rc == 0
                for c in range(6):
                    ref, _ = oracle.remapping_core_h(cs, h0[c], u0[c], h1[c])
                    if not np.array_equal(ref, got[c], equal_nan=True):
                        nbad += 1
                        if nbad < 8:
                            print("MISMATCH", n0, n1, kind, {k: cs[k] for k in ("remapping_scheme", "boundary_extrapolation", "force_bounds_in_subcell",
                                                                                "force_bounds_in_target", "om4_remap_via_sub_cells")},
                                  np.abs(ref - got[c]).max(), np.flatnonzero(ref != got[c])[:6])
    assert nbad == 0
