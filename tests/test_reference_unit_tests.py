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


def test_reference_eos_consistency_tests():
    """test_EOS_consistency (src/equation_of_state/MOM_EOS.F90:2302-2660), called as EOS_unit_tests calls it (:2078-2131) for the two
    equations of state of the frozen option set: the published density at T = 25, S = 35, p = 1e7, rho * spv = 1, the reference-density
    offsets, every first and second derivative against finite differences of three orders, and the analytic against the quadrature
    layer-mean specific volume.  LINEAR passes; WRIGHT with USE_WRIGHT_2ND_DERIV_BUG passes everything except drho_dT_dT -- the failure
    the reference itself documents at :2081-2083 ("a known failure") -- and passes outright without the bug flag."""
    from oracle.f90run import rt, stages
    R = f90run.load(list(stages.EOS_FILES), extra_stubs=dict(stdout=6, stderr=0))
    M = R["mom_eos"]

    def run(name, check, **init):
        rt.UNITS[0], rt.UNITS[6] = [], []
        try:
            E = M["_new_eos_type"]()
            M["eos_manual_init"](E, **init)
            failed = M["test_eos_consistency"](25.0, 35.0, 1.0e7, E, True, name, rho_check=check * E.kg_m3_to_r, avg_sv_check=True)
            return failed, [ln for ln in rt.UNITS[0] if "disagree" in ln], len(rt.UNITS[6])
        finally:
            del rt.UNITS[0], rt.UNITS[6]

    failed, bad, nok = run("LINEAR", 1028.0, form_of_eos=M["eos_linear"], rho_t0_s0=1000.0, drho_dt=-0.2, drho_ds=0.8, drho_dp=5.0e-7)
    assert failed is False and not bad and nok >= 13
    failed, bad, nok = run("WRIGHT", 1027.54303596346, form_of_eos=M["eos_wright"], use_wright_2nd_deriv_bug=True)
    # (the translator evaluates .AND. left to right and stops at the first false operand, so the checks after the failing one are skipped)
    assert failed is True and len(bad) == 1 and "WRIGHT drho_dT_dT" in bad[0] and nok >= 8
    failed, bad, nok = run("WRIGHT", 1027.54303596346, form_of_eos=M["eos_wright"], use_wright_2nd_deriv_bug=False)
    assert failed is False and not bad
    # negative control: a check value off in the 12th digit is reported
    failed, bad, nok = run("LINEAR", 1028.0 * (1.0 + 1.0e-11), form_of_eos=M["eos_linear"], rho_t0_s0=1000.0, drho_dt=-0.2, drho_ds=0.8, drho_dp=5.0e-7)
    assert failed is True
