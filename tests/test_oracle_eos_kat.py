"""Known-answer tests that PIN the oracle's equation of state to the reference's own check values.

Reference: EOS_unit_tests, src/equation_of_state/MOM_EOS.F90:2036-2150 -- for each equation of state it calls test_EOS_consistency
(:2302-2660) at T=25 degC, S=35 ppt, p=1e7 Pa with a published density `rho_check`:
    EOS_WRIGHT  (use_Wright_2nd_deriv_bug)  rho_check = 1027.54303596346   (:2078-2079)
    EOS_LINEAR  (Rho_T0_S0=1000, dRho_dT=-0.2, dRho_dS=0.8, dRho_dp=5e-7)  rho_check = 1028.0   (:2129-2131)
with tolerance tol*rho, tol = 1000*epsilon (:2469-2473), the agreement of the density with and without a reference density (:2500-2502),
and the first derivatives against 4th-order centred differences at two step sizes with the convergence criterion of check_FD (:2628-2660).
The same checks are run here on the oracle's restatement (oracle/eos.hpp) and on the copies oracle/mle.cpp and oracle/thickdiff.cpp
carry; the device code is compared bit for bit with these functions by the GPU parity tests of the stages that use them."""
import numpy as np
import pytest

EPS = np.finfo(np.float64).eps
T0, S0, P0 = 25.0, 35.0, 1.0e7
WRIGHT, LINEAR = 3, 1
LIN = (1000.0, -0.2, 0.8, 5.0e-7)
CHECK = {WRIGHT: 1027.54303596346, LINEAR: 1028.0}


def _first_deriv(R, dx):
    """first_deriv, order = 4 (:2583-2584); R[-3:3] stored at index n + 3."""
    return (8.0 * (R[4] - R[2]) - (R[5] - R[1])) / (12.0 * dx)


def _check_fd(val, fd, tol, order=4):
    """check_FD (:2641)."""
    return abs(fd[0] - val) < (1.2 * abs(fd[1] - val) / 2 ** order + abs(tol))


@pytest.mark.parametrize("form", [WRIGHT, LINEAR])
def test_density_matches_the_reference_check_value(oracle, form):
    lin = LIN if form == LINEAR else None
    tol = 1000.0 * EPS
    rho_ref = 1000.0
    rho = oracle.eos_eval("rho", form, T0, S0, P0, lin4=lin)
    anom = oracle.eos_eval("anom", form, T0, S0, P0, rho_ref=rho_ref, lin4=lin)
    assert abs(CHECK[form] - (rho_ref + anom)) < tol * (rho_ref + anom)          # :2473
    assert abs(rho - (rho_ref + anom)) < tol * rho                               # :2502
    assert abs(CHECK[form] - rho) < tol * rho
    # the copy in oracle/mle.cpp (density_elem)
    assert oracle.eos_eval("rho", form, T0, S0, P0, lin4=lin, impl="mle") == rho


@pytest.mark.parametrize("form", [WRIGHT, LINEAR])
def test_density_derivatives_pass_the_reference_consistency_test(oracle, form):
    lin = LIN if form == LINEAR else None
    tol = 1000.0 * EPS
    rho_ref = 1000.0
    r_tol = 50.0 * 10.0 * EPS                                                    # :2399
    dT, dS = 0.1, 0.5                                                            # :2395-2396
    count_fac = 18.0 / 12.0                                                      # :2525
    fdT, fdS = [], []
    for n in (1, 2):
        RT = [oracle.eos_eval("anom", form, T0 + n * dT * i, S0, P0, rho_ref=rho_ref, lin4=lin) for i in range(-3, 4)]
        RS = [oracle.eos_eval("anom", form, T0, S0 + n * dS * j, P0, rho_ref=rho_ref, lin4=lin) for j in range(-3, 4)]
        fdT.append(_first_deriv(RT, n * dT)); fdS.append(_first_deriv(RS, n * dS))
    drho_dT = oracle.eos_eval("drho_dT", form, T0, S0, P0, lin4=lin)
    drho_dS = oracle.eos_eval("drho_dS", form, T0, S0, P0, lin4=lin)
    assert _check_fd(drho_dT, fdT, tol * abs(drho_dT) + count_fac * r_tol / dT)  # :2533-2534
    assert _check_fd(drho_dS, fdS, tol * abs(drho_dS) + count_fac * r_tol / dS)  # :2535-2536
    assert drho_dT < 0.0 < drho_dS
    # the copy in oracle/thickdiff.cpp (calculate_density_derivs)
    assert oracle.eos_eval("drho_dT", form, T0, S0, P0, lin4=lin, impl="thickdiff") == drho_dT
    assert oracle.eos_eval("drho_dS", form, T0, S0, P0, lin4=lin, impl="thickdiff") == drho_dS


@pytest.mark.parametrize("form", [WRIGHT, LINEAR])
def test_unit_rescaling_wrapper_is_exact_for_powers_of_two(oracle, form):
    """calculate_density_1d (:308-354) with EOS%kg_m3_to_R etc. set to powers of two gives the unscaled answer times the scale, bit for bit
    (the reference's dimensional-consistency criterion)."""
    lin = LIN if form == LINEAR else None
    kg_m3_to_R, RL2_T2_to_Pa, C_to_degC, S_to_ppt = 2.0 ** 4, 2.0 ** -7, 2.0 ** 3, 2.0 ** -2
    sc = (kg_m3_to_R, RL2_T2_to_Pa, C_to_degC, S_to_ppt)
    r = np.random.default_rng(7)
    for _ in range(50):
        T, S, p = r.uniform(-2, 30), r.uniform(30, 38), r.uniform(0, 5e7)
        rho = oracle.eos_eval("rho", form, T, S, p, lin4=lin)
        assert oracle.eos_eval("rho", form, T / C_to_degC, S / S_to_ppt, p / RL2_T2_to_Pa, lin4=lin, scales=sc) == kg_m3_to_R * rho
        an = oracle.eos_eval("anom", form, T, S, p, rho_ref=1035.0, lin4=lin)
        assert oracle.eos_eval("anom", form, T / C_to_degC, S / S_to_ppt, p / RL2_T2_to_Pa, rho_ref=1035.0 * kg_m3_to_R, lin4=lin,
                               scales=sc) == kg_m3_to_R * an
        d = oracle.eos_eval("drho_dT", form, T, S, p, lin4=lin)
        assert oracle.eos_eval("drho_dT", form, T / C_to_degC, S / S_to_ppt, p / RL2_T2_to_Pa, lin4=lin, scales=sc) == kg_m3_to_R * C_to_degC * d


def test_wright_anomaly_form_agrees_with_density_minus_reference(oracle):
    """density_anomaly_elem_buggy_Wright (:102-130) is the algebraic rearrangement of density - rho_ref: equal to roundoff of rho."""
    r = np.random.default_rng(11)
    for _ in range(200):
        T, S, p = r.uniform(-2, 30), r.uniform(0, 40), r.uniform(0, 6e7)
        rho = oracle.eos_eval("rho", WRIGHT, T, S, p)
        an = oracle.eos_eval("anom", WRIGHT, T, S, p, rho_ref=1035.0)
        assert abs((rho - 1035.0) - an) < 8 * EPS * rho
