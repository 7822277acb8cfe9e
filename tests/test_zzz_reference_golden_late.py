"""The device legs of the reference-digest cases that were added or changed after the round's last GPU run (tests/refcases.py LATE):
the same comparison as tests/test_reference_golden.py::test_device_matches_reference_digest -- the CUDA path against the sha256 of
what the reference's own Fortran produced -- kept in a file that sorts last, so that under `pytest -x` a disagreement in a case
nobody has yet seen on a B200 cannot hide the results of the other GPU test files.  For the same reason all of them run inside ONE
test that goes through every case and reports every disagreement together, instead of stopping at the first."""
import pytest

import refcases
from test_reference_golden import device_against_digest



@pytest.mark.gpu
def test_device_matches_reference_digest_late_cases(ctx_factory):
    bad, seen = [], 0
    for name in sorted(refcases.LATE):
        seen += 1
        try:
            device_against_digest(ctx_factory, name)
        except Exception as e:   # noqa: BLE001 -- an assertion (digest differs) or an error of the C ABI: both are reported, with the case
            bad.append(f"{name}: {type(e).__name__}: {str(e).splitlines()[0][:200] if str(e) else ''}")
    assert seen == len(refcases.LATE)
    assert not bad, f"{len(bad)} of {seen} late cases disagree with the reference digest:\n  " + "\n  ".join(bad)


def test_late_and_early_cases_partition_the_cases():
    assert refcases.LATE <= set(refcases.CASES) and len(refcases.LATE) + len(set(refcases.CASES) - refcases.LATE) == len(refcases.CASES)
