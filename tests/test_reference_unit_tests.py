"""The reference's OWN unit tests, executed through oracle/f90run.

remapping_unit_tests (src/ALE/MOM_remapping.F90:2072-2943, the driver of config_src/drivers/unit_tests/test_MOM_remapping.F90) and
mixedlayer_restrat_unit_tests (src/parameterizations/lateral/MOM_mixed_layer_restrat.F90:2000-2100) hold several hundred
known-answer checks written by the reference's authors.  They return .true. on any failure.  Running them through the
Fortran-subset translator checks the TRANSLATOR (expression order, integer/real rules, array sections, argument association,
type-bound and elemental procedures) against values the reference itself asserts -- the same translator that
tests/test_reference_f90.py uses to arbitrate the oracle."""
import os

import pytest

from oracle import f90run

pytestmark = pytest.mark.skipif(not f90run.available(), reason="the reference tree is not present")


def test_reference_remapping_unit_tests_pass():
    from oracle.f90run import rt, stages
    files = list(stages.ALE_FILES[1:16]) + ["src/framework/numerical_testing_type.F90", "src/ALE/MOM_hybgen_remap.F90"]
    files += sorted("src/ALE/" + f for f in os.listdir(os.path.join(f90run.REFERENCE_ROOT, "src/ALE")) if f.startswith("Recon1d"))
    R = f90run.load(files)
    rt.LENIENT_READS = True   # Recon1d_PPM_hybgen.F90:133 reads u(n+1) and discards it
    try:
        failed = R["mom_remapping"]["remapping_unit_tests"](False, num_comp_samp=3)
    finally:
        rt.LENIENT_READS = False
    assert failed is False


def test_reference_mixedlayer_restrat_unit_tests_pass():
    from oracle.f90run import stages
    R = f90run.load(list(stages.MLE_FILES))
    assert R["mom_mixed_layer_restrat"]["mixedlayer_restrat_unit_tests"](False) is False


def test_reference_remapping_unit_tests_detect_a_one_ulp_class_error():
    """negative control: the same run reports failure when one edge value of edge_values_explicit_h4 is off by 1e-13 (relative)"""
    from oracle.f90run import rt, stages
    files = list(stages.ALE_FILES[1:16]) + ["src/framework/numerical_testing_type.F90", "src/ALE/MOM_hybgen_remap.F90"]
    files += sorted("src/ALE/" + f for f in os.listdir(os.path.join(f90run.REFERENCE_ROOT, "src/ALE")) if f.startswith("Recon1d"))
    R = f90run.load(files)
    f0 = R["regrid_edge_values"]["edge_values_explicit_h4"]

    def perturbed(*a, **k):
        r = f0(*a, **k)
        e = a[3] if len(a) > 3 else k.get("edge_val")
        e.v[0] = e.v[0] * (1.0 + 1e-13)
        return r
    for ns in R.values():
        if ns.get("edge_values_explicit_h4") is f0:
            ns["edge_values_explicit_h4"] = perturbed
    rt.LENIENT_READS = True
    try:
        assert R["mom_remapping"]["remapping_unit_tests"](False, num_comp_samp=3) is True
    finally:
        rt.LENIENT_READS = False
