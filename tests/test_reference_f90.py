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
