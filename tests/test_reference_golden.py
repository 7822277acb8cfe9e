"""The oracle (CPU) and the CUDA path (GPU) against digests of the REFERENCE'S OWN outputs.

tests/golden/reference_f90_digests.json holds, for every case of tests/refcases.py, the sha256 of what the reference's Fortran
source produced when executed by oracle/f90run in the build container (tests/golden/make_reference_digests.py).  These tests
need neither the reference tree nor, on the GPU side, the oracle: the device path is compared with the reference's answer
directly, bit for bit (-0.0 folded onto +0.0, see tests/test_reference_f90.py)."""
import json
import os

import numpy as np
import pytest

import refcases

HERE = os.path.dirname(os.path.abspath(__file__))
WANT = json.load(open(os.path.join(HERE, "golden", "reference_f90_digests.json")))


def test_every_case_has_a_digest():
    assert sorted(WANT) == sorted(refcases.CASES)


@pytest.mark.parametrize("name", sorted(refcases.CASES))
def test_oracle_matches_reference_digest(oracle, name):
    inputs = refcases.build(name)
    got = refcases.run_oracle(oracle, name, inputs)
    assert sorted(got) == WANT[name]["outputs"], name
    assert refcases.digest(got) == WANT[name]["digest"], name
    assert not refcases.untouched(name, inputs, got), f"{name}: outputs come back as they went in"


def device_against_digest(ctx_factory, name):
    if name in refcases.DEVICE_REFUSES:
        from mom6_b200.api import Mom6cuError
        with pytest.raises(Mom6cuError, match="rc=3"):
            refcases.run_device(ctx_factory, name, refcases.build(name))
        return
    got = refcases.run_device(ctx_factory, name, refcases.build(name))
    assert sorted(got) == WANT[name]["outputs"], name
    assert refcases.digest(got) == WANT[name]["digest"], name


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(set(refcases.CASES) - refcases.LATE))
def test_device_matches_reference_digest(ctx_factory, name):
    device_against_digest(ctx_factory, name)
