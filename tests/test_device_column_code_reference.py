"""The device's column / cell code, built for the host, against digests of the REFERENCE'S OWN outputs.

csrc/mle_column.cuh, thickdiff_column.cuh and hordiff_cell.cuh are the functions the CUDA kernels of mixedlayer_restrat,
thickness_diffuse and tracer_hordiff call per column or cell; tests/test_mle.py, test_thickness_diffuse.py and test_tracer_hordiff.py
compile them with g++ (same -ffp-contract=off arithmetic) and drive them the way the kernels' launchers do.  Here those host builds
run the reference-digest cases of tests/refcases.py and must reproduce tests/golden/reference_f90_digests.json -- what the
reference's Fortran produced under oracle/f90run -- bit for bit.  No GPU and no reference tree needed: this is the part of
"device == reference" that can be checked anywhere (the kernels around the column code are checked on the B200 by
tests/test_reference_golden.py)."""
import json
import os

import pytest

import refcases
from test_mle import _run_device_code_on_host as mle_host_run, mle_host  # noqa: F401
from test_thickness_diffuse import _run_device_code_on_host as td_host_run, td_host  # noqa: F401
from test_tracer_hordiff import _run_device_code_on_host as hd_host_run, hd_host  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
WANT = json.load(open(os.path.join(HERE, "golden", "reference_f90_digests.json")))


def _names(stage):
    return sorted(n for n, c in refcases.CASES.items() if c["stage"] == stage)


def _check(name, dom, a, cs):
    c = refcases.CASES[name]
    got = refcases.collect(dom, c["outputs"], a, cs)
    assert sorted(got) == WANT[name]["outputs"], name
    assert refcases.digest(got) == WANT[name]["digest"], name


@pytest.mark.parametrize("name", _names("mixedlayer_restrat"))
def test_mle_column_code_matches_reference_digest(mle_host, name):   # noqa: F811
    dom, grid, gv, cs, a = refcases.build(name)
    out = mle_host_run(mle_host, dom, grid, gv, cs, a)
    _check(name, dom, dict(a, h=out["h"], uhtr=out["uhtr"], vhtr=out["vhtr"]),
           dict(cs, MLD_filtered=out["MLD_filtered"], MLD_filtered_slow=out["MLD_filtered_slow"]))


@pytest.mark.parametrize("name", _names("thickness_diffuse"))
def test_thickness_diffuse_column_code_matches_reference_digest(td_host, name):   # noqa: F811
    dom, grid, gv, cs, a = refcases.build(name)
    out = td_host_run(td_host, dom, grid, gv, cs, a)
    _check(name, dom, dict(a, **out), cs)


@pytest.mark.parametrize("name", _names("tracer_hordiff"))
def test_tracer_hordiff_cell_code_matches_reference_digest(hd_host, name):   # noqa: F811
    dom, grid, gv, cs, a = refcases.build(name)
    n, out = hd_host_run(hd_host, dom, grid, gv, cs, a)
    _check(name, dom, out, cs)
