// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Equation-of-state elements shared by the pressure-force restatement:
//   EOS_WRIGHT ("buggy" Wright 1997 fit): /root/reference/src/equation_of_state/MOM_EOS_Wright.F90
//     coefficients :23-37, density_elem_buggy_Wright :80-97, density_anomaly_elem_buggy_Wright :102-130,
//     calculate_density_derivs_elem_buggy_Wright :178-206
//   EOS_LINEAR: MOM_EOS_linear.F90 density_elem_linear :60-68, density_anomaly_elem_linear :74-84, derivs :131-132
//   the unit-rescaling wrapper calculate_density_1d, MOM_EOS.F90:308-354 (EOS%RL2_T2_to_Pa, C_to_degC, S_to_ppt, kg_m3_to_R)
// PARITY: PINNED -- tests/test_oracle_eos_kat.py checks these against the reference's own check values
// (EOS_unit_tests, MOM_EOS.F90:2075-2131: rho(T=25, S=35, p=1e7) = 1027.54303596346 for WRIGHT, 1028.0 for LINEAR with
// Rho_T0_S0=1000, dRho_dT=-0.2, dRho_dS=0.8, dRho_dp=5e-7) and re-runs test_EOS_consistency's finite-difference checks (:2302-2560).
#ifndef MOM6_ORACLE_EOS_HPP
#define MOM6_ORACLE_EOS_HPP
#include "../include/mom6cu.h"

namespace orc {

namespace wright {
const double a0 = 7.057924e-4, a1 = 3.480336e-7, a2 = -1.112733e-7;
const double b0 = 5.790749e8, b1 = 3.516535e6, b2 = -4.002714e4, b3 = 2.084372e2, b4 = 5.944068e5, b5 = -9.643486e3;
const double c0 = 1.704853e5, c1 = 7.904722e2, c2 = -7.984422, c3 = 5.140652e-2, c4 = -2.302158e2, c5 = -3.079464;
}  // namespace wright

struct EOSp {
  int form;
  double Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp;
  // EOS_type unit conversion factors (MOM_EOS.F90:140-150); 1 in an unscaled run
  double kg_m3_to_R = 1.0, R_to_kg_m3 = 1.0, RL2_T2_to_Pa = 1.0, C_to_degC = 1.0, S_to_ppt = 1.0;
  bool unscaled() const { return RL2_T2_to_Pa == 1.0 && R_to_kg_m3 == 1.0 && C_to_degC == 1.0 && S_to_ppt == 1.0; }
};

// density_elem in mks units [kg m-3]
inline double density_mks(const EOSp& E, double T, double S, double p) {
  if (E.form == MOM6CU_EOS_LINEAR) return E.Rho_T0_S0 + E.dRho_dT * T + E.dRho_dS * S + E.dRho_dp * p;  // MOM_EOS_linear.F90:66
  using namespace wright;
  const double al0 = (a0 + a1 * T) + a2 * S;  // MOM_EOS_Wright.F90:91-94
  const double p0 = (b0 + b4 * S) + T * (b1 + T * (b2 + b3 * T) + b5 * S);
  const double lambda = (c0 + c4 * S) + T * (c1 + T * (c2 + c3 * T) + c5 * S);
  return (p + p0) / (lambda + al0 * (p + p0));
}

// density_anomaly_elem in mks units [kg m-3]
inline double density_anomaly_mks(const EOSp& E, double T, double S, double pressure, double rho_ref) {
  if (E.form == MOM6CU_EOS_LINEAR) return (E.Rho_T0_S0 - rho_ref) + ((E.dRho_dT * T + E.dRho_dS * S) + E.dRho_dp * pressure);  // :81-82
  using namespace wright;
  const double pa_000 = (b0 * (1.0 - a0 * rho_ref) - rho_ref * c0);  // MOM_EOS_Wright.F90:118-128
  const double al_TS = a1 * T + a2 * S;
  const double al0 = a0 + al_TS;
  const double p_TSp = pressure + (b4 * S + T * (b1 + (T * (b2 + b3 * T) + b5 * S)));
  const double lam_TS = c4 * S + T * (c1 + (T * (c2 + c3 * T) + c5 * S));
  return (pa_000 + (p_TSp - rho_ref * (p_TSp * al0 + (b0 * al_TS + lam_TS)))) / ((c0 + lam_TS) + al0 * (b0 + p_TSp));
}

inline void density_derivs_mks(const EOSp& E, double T, double S, double p, double& drho_dT, double& drho_dS) {
  if (E.form == MOM6CU_EOS_LINEAR) { drho_dT = E.dRho_dT; drho_dS = E.dRho_dS; return; }  // MOM_EOS_linear.F90:131-132
  using namespace wright;
  const double al0 = (a0 + a1 * T) + a2 * S;  // MOM_EOS_Wright.F90:193-204
  const double p0 = (b0 + b4 * S) + T * (b1 + T * ((b2 + b3 * T)) + b5 * S);
  const double lambda = (c0 + c4 * S) + T * (c1 + T * ((c2 + c3 * T)) + c5 * S);
  double I_denom2 = 1.0 / (lambda + al0 * (p + p0));
  I_denom2 = I_denom2 * I_denom2;
  drho_dT = I_denom2 * (lambda * (b1 + T * (2.0 * b2 + 3.0 * b3 * T) + b5 * S) -
                        (p + p0) * ((p + p0) * a1 + (c1 + T * (c2 * 2.0 + c3 * 3.0 * T) + c5 * S)));
  drho_dS = I_denom2 * (lambda * (b4 + b5 * T) - (p + p0) * ((p + p0) * a2 + (c4 + c5 * T)));
}

// calculate_density_1d for one point, MOM_EOS.F90:308-354: inputs and result in the model's (possibly rescaled) units.
inline double calculate_density(const EOSp& E, double T, double S, double pressure, const double* rho_ref) {
  double rho;
  if (E.unscaled()) {
    rho = rho_ref ? density_anomaly_mks(E, T, S, pressure, *rho_ref) : density_mks(E, T, S, pressure);
  } else {
    const double pres = E.RL2_T2_to_Pa * pressure, Ta = E.C_to_degC * T, Sa = E.S_to_ppt * S;
    rho = rho_ref ? density_anomaly_mks(E, Ta, Sa, pres, E.R_to_kg_m3 * (*rho_ref)) : density_mks(E, Ta, Sa, pres);
  }
  const double rho_scale = E.kg_m3_to_R;
  if (rho_scale != 1.0) rho = rho_scale * rho;
  return rho;
}

// calculate_density_derivs_1d for one point, MOM_EOS.F90:783-835
inline void calculate_density_derivs(const EOSp& E, double T, double S, double pressure, double& drho_dT, double& drho_dS) {
  if (E.RL2_T2_to_Pa == 1.0 && E.C_to_degC == 1.0 && E.S_to_ppt == 1.0) density_derivs_mks(E, T, S, pressure, drho_dT, drho_dS);
  else density_derivs_mks(E, E.C_to_degC * T, E.S_to_ppt * S, E.RL2_T2_to_Pa * pressure, drho_dT, drho_dS);
  const double rho_scale = E.kg_m3_to_R;
  const double dRdT_scale = rho_scale * E.C_to_degC, dRdS_scale = rho_scale * E.S_to_ppt;
  if (dRdT_scale != 1.0 || dRdS_scale != 1.0) { drho_dT = dRdT_scale * drho_dT; drho_dS = dRdS_scale * drho_dS; }
}

}  // namespace orc
#endif
