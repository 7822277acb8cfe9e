"""The C++ oracle against the REFERENCE'S OWN FORTRAN, executed here.

There is no Fortran compiler in this image, so oracle/f90run translates the reference's .F90 sources (read where they lie under
/root/reference; nothing is copied) into Python that evaluates every expression in binary64 in source order, and these tests
run the translated reference routine and the oracle on copies of the same seeded inputs.  Every output must agree BIT FOR BIT
over the computational domain -- with one stated exception: Fortran leaves MAX/MIN of equal operands of opposite sign
processor-dependent, so -0.0 and +0.0 compare equal.

Skipped where /root/reference does not exist (the GPU box); tests/test_reference_golden.py then still checks the oracle
against digests of these same reference runs (tests/golden/reference_f90_digests.json, made by tests/golden/make_reference_digests.py)."""
import numpy as np
import pytest

from mom6_b200 import synthetic
from oracle import f90run

pytestmark = pytest.mark.skipif(not f90run.available(), reason="the reference tree is not present")

import os  # noqa: E402

from refcases import CASES, SLOW, run_case  # noqa: E402

LIVE = sorted(n for n in CASES if n not in SLOW or os.environ.get("F90RUN_SLOW", "0") == "1")


@pytest.mark.parametrize("name", LIVE)
def test_oracle_matches_translated_reference(oracle, name):
    ref_out, orc_out = run_case(oracle, name, want_ref=True)
    bad = []
    for k in ref_out:
        r, o = ref_out[k], orc_out[k]
        if not np.array_equal((r + 0.0).view(np.int64), (o + 0.0).view(np.int64)):
            bad.append(f"{k}: {np.count_nonzero(r != o)} of {r.size} differ, max |d| = {np.nanmax(np.abs(r - o)):.3e}")
        assert np.count_nonzero(r) > 0 or k.startswith("zero_ok:"), f"{name}: reference output {k} is identically zero"
    assert not bad, f"{name}: " + "; ".join(bad)


def test_efp_sums_match_the_translated_reference(oracle):
    """reproducing_sum_2d / _3d, real_to_EFP, EFP_to_real, EFP_plus / EFP_minus (through the overloaded operators) and EFP_real_diff of
    the reference's own src/framework/MOM_coms.F90, executed by the translator, against oracle/efp.cpp: the sums, the per-layer sums
    and the six integers of the extended-fixed-point representation, exactly."""
    from mom6_b200.api import make_domain
    from oracle.f90run.rt import FArray
    M = f90run.load(["src/framework/MOM_coms.F90"])["mom_coms"]
    dom = make_domain(12, 10, nk=3)
    r = np.random.default_rng(5)
    shp = (dom.jed - dom.jsd + 1, dom.ied - dom.isd + 1)
    a2 = np.ascontiguousarray(r.standard_normal(shp) * 10.0 ** r.integers(-12, 12, size=shp))
    a3 = np.ascontiguousarray(r.standard_normal((3,) + shp) * 1e3)
    w = dict(isr=dom.isc - dom.isd + 1, ier=dom.iec - dom.isd + 1, jsr=dom.jsc - dom.jsd + 1, jer=dom.jec - dom.jsd + 1)
    for ov in (True, False):
        o = oracle.reproducing_sum(dom, a2, overflow_check=ov, want_efp=True, **w)
        efp = M["_new_efp_type"]()
        s = M["reproducing_sum_2d"](FArray.from_numpy(a2, (1, 1)), w["isr"], w["ier"], w["jsr"], w["jer"], efp_sum=efp, overflow_check=ov)
        assert o["sum"] == s and [int(x) for x in o["EFP_sum"]] == efp.v.tolist() and any(efp.v.tolist())
    o = oracle.reproducing_sum(dom, a3, want_sums=True, want_efp=True, **w)
    sums, efp = FArray.alloc("r", [(1, 3)]), M["_new_efp_type"]()
    s = M["reproducing_sum_3d"](FArray.from_numpy(a3, (1, 1, 1)), w["isr"], w["ier"], w["jsr"], w["jer"], sums=sums, efp_sum=efp)
    assert o["sum"] == s and np.array_equal(o["sums"], np.array(sums.tolist())) and [int(x) for x in o["EFP_sum"]] == efp.v.tolist()
    for x, y in ((-1234.56789e8, 3.25), (7.0e-20, -5.5e30), (0.1, 0.2)):
        e1, e2 = M["real_to_efp"](x), M["real_to_efp"](y)
        assert [int(v) for v in oracle.efp_op("from_real", x)] == e1.v.tolist()
        assert oracle.efp_op("to_real", e1.v.tolist()) == M["efp_to_real"](e1)
        assert [int(v) for v in oracle.efp_op("plus", e1.v.tolist(), e2.v.tolist())] == (e1 + e2).v.tolist()
        assert [int(v) for v in oracle.efp_op("minus", e1.v.tolist(), e2.v.tolist())] == (e1 - e2).v.tolist()
        assert oracle.efp_op("diff", e1.v.tolist(), e2.v.tolist()) == M["efp_real_diff"](e1, e2)
