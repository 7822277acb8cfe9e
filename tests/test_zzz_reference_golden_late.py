"""The device legs of the reference-digest cases that were added or changed after the round's last GPU run (tests/refcases.py LATE):
the same comparison as tests/test_reference_golden.py::test_device_matches_reference_digest -- the CUDA path against the sha256 of
what the reference's own Fortran produced -- kept in a file that sorts last, so that under `pytest -x` a disagreement in a case
nobody has yet seen on a B200 cannot hide the results of the other GPU test files."""
import pytest

import refcases
from test_reference_golden import device_against_digest


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(refcases.LATE))
def test_device_matches_reference_digest_late_cases(ctx_factory, name):
    device_against_digest(ctx_factory, name)
