// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of PressureForce_FV_Bouss, /root/reference/src/core/MOM_PressureForce_FV.F90:947-2017, with
//   int_density_dz -> analytic_int_density_dz (src/core/MOM_density_integrals.F90:42-103, src/equation_of_state/MOM_EOS.F90:1384-1499)
//   int_density_dz_linear (src/equation_of_state/MOM_EOS_linear.F90:275-440)
//   int_density_dz_wright (src/equation_of_state/MOM_EOS_Wright.F90:389-655), density_elem / calculate_density_derivs_elem (:80, :178)
//   Set_pbce_Bouss (src/core/MOM_PressureForce_Montgomery.F90:649-748)
//   RECONSTRUCT_FOR_PRESSURE: TS_PLM_edge_values / TS_PPM_edge_values (src/ALE/MOM_ALE.F90:1495-1660) and the quadrature integrals
//   int_density_dz_generic_plm (src/core/MOM_density_integrals.F90:418-868) / int_density_dz_generic_ppm (:874-1310)
// Frozen options: no tides/SAL, no Stanley term, no intxpa corrections/resets, nk_rho_varies = 0.
// The EOS elements (oracle/eos.hpp) are PINNED to the reference's check values (tests/test_oracle_eos_kat.py); the analytic layer
// integrals are checked against Boole quadrature of that pinned density; the routine as a whole has no vector in the reference.
#include "oracle.h"
#include "ogrid.hpp"
#include "eos.hpp"
#include <cmath>
#include <vector>
#include <omp.h>

using namespace orc;

namespace {

using namespace orc::wright;

inline double max3(double a, double b, double c) { return fmax2(fmax2(a, b), c); }

inline double density(const EOSp& E, double T, double S, double p) { return calculate_density(E, T, S, p, nullptr); }
inline void density_derivs(const EOSp& E, double T, double S, double p, double& drho_dT, double& drho_dS) {
  calculate_density_derivs(E, T, S, p, drho_dT, drho_dS);
}

struct IntArgs {
  const OGrid* G; V2 T, S, z_t, z_b, dpa, intz_dpa, intx_dpa, inty_dpa, bathyT, SSH, Z_0p;
  double rho_ref, rho_0, G_e, dz_neglect; int MassWghtInterp;
  // unit conversion factors handed down by analytic_int_density_dz (MOM_EOS.F90:1440-1470); all 1 in an unscaled run
  double rho_scale = 1.0, pres_scale = 1.0, temp_scale = 1.0, saln_scale = 1.0;
};

// the weights of the mass-weighted interpolation shared by both EOS forms
inline void mass_weights(double hWght_in, double hL, double hR, double& hWt_LL, double& hWt_LR, double& hWt_RR, double& hWt_RL) {
  const double r = (hL - hR) / (hL + hR);
  const double hWght = hWght_in * (r * r);
  const double iDenom = 1.0 / (hWght * (hR + hL) + hL * hR);
  hWt_LL = (hWght * hL + hR * hL) * iDenom; hWt_LR = (hWght * hR) * iDenom;
  hWt_RR = (hWght * hR + hR * hL) * iDenom; hWt_RL = (hWght * hL) * iDenom;
}

// int_density_dz_linear, MOM_EOS_linear.F90:275-440
void int_density_dz_linear(const IntArgs& A, const EOSp& E) {
  const OGrid& G = *A.G;
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  const double C1_6 = 1.0 / 6.0, C1_90 = 1.0 / 90.0;
  const double G_e = A.G_e, GxRho = A.G_e * A.rho_0, rho_ref = A.rho_ref;
  const double Rho_T0_S0 = E.Rho_T0_S0, dRho_dT = E.dRho_dT, dRho_dS = E.dRho_dS, dRho_dp = E.dRho_dp;
  const bool do_massWeight = (A.MassWghtInterp & 1) != 0, top_massWeight = (A.MassWghtInterp & 2) != 0;
  const V2 &T = A.T, &S = A.S, &z_t = A.z_t, &z_b = A.z_b, &z0pres = A.Z_0p, &dpa = A.dpa;
  for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) {
    const double dz = z_t(i, j) - z_b(i, j);
    const double p_ave = -GxRho * (0.5 * (z_t(i, j) + z_b(i, j)) - z0pres(i, j));
    const double rho_anom = (Rho_T0_S0 - rho_ref) + dRho_dT * T(i, j) + dRho_dS * S(i, j) + dRho_dp * p_ave;
    dpa(i, j) = G_e * rho_anom * dz;
    A.intz_dpa(i, j) = 0.5 * G_e * (rho_anom - C1_6 * dRho_dp * (GxRho * dz)) * (dz * dz);
  }
  for (int dir = 0; dir < 2; ++dir) {
    const int di = dir == 0 ? 1 : 0, dj = dir == 0 ? 0 : 1;
    const int jlo = dir == 0 ? js : Jsq, jhi = dir == 0 ? je : Jeq, ilo = dir == 0 ? Isq : is, ihi = dir == 0 ? Ieq : ie;
    const V2& out = dir == 0 ? A.intx_dpa : A.inty_dpa;
    for (int j = jlo; j <= jhi; ++j) for (int i = ilo; i <= ihi; ++i) {
      const int ip = i + di, jp = j + dj;
      double hWght = 0.0;
      if (do_massWeight) hWght = max3(0., -A.bathyT(i, j) - z_t(ip, jp), -A.bathyT(ip, jp) - z_t(i, j));
      if (top_massWeight) hWght = max3(hWght, z_b(ip, jp) - A.SSH(i, j), z_b(i, j) - A.SSH(ip, jp));
      if (hWght <= 0.0) {
        const double dzL = z_t(i, j) - z_b(i, j), dzR = z_t(ip, jp) - z_b(ip, jp);
        double p_ave = -GxRho * (0.5 * (z_t(i, j) + z_b(i, j)) - z0pres(i, j));
        const double raL = (Rho_T0_S0 - rho_ref) + ((dRho_dT * T(i, j) + dRho_dS * S(i, j)) + dRho_dp * p_ave);
        p_ave = -GxRho * (0.5 * (z_t(ip, jp) + z_b(ip, jp)) - z0pres(ip, jp));
        const double raR = (Rho_T0_S0 - rho_ref) + ((dRho_dT * T(ip, jp) + dRho_dS * S(ip, jp)) + dRho_dp * p_ave);
        out(i, j) = G_e * C1_6 * ((dzL * (2.0 * raL + raR)) + (dzR * (2.0 * raR + raL)));
      } else {
        const double hL = (z_t(i, j) - z_b(i, j)) + A.dz_neglect, hR = (z_t(ip, jp) - z_b(ip, jp)) + A.dz_neglect;
        double hWt_LL, hWt_LR, hWt_RR, hWt_RL;
        mass_weights(hWght, hL, hR, hWt_LL, hWt_LR, hWt_RR, hWt_RL);
        double intz[6];
        intz[1] = dpa(i, j); intz[5] = dpa(ip, jp);
        for (int m = 2; m <= 4; ++m) {
          const double wt_L = 0.25 * (double)(5 - m), wt_R = 1.0 - wt_L;
          const double wtT_L = (wt_L * hWt_LL) + (wt_R * hWt_RL), wtT_R = (wt_L * hWt_LR) + (wt_R * hWt_RR);
          const double dz = (wt_L * (z_t(i, j) - z_b(i, j))) + (wt_R * (z_t(ip, jp) - z_b(ip, jp)));
          const double p_ave = -GxRho * ((wt_L * (0.5 * (z_t(i, j) + z_b(i, j)) - z0pres(i, j))) +
                                         (wt_R * (0.5 * (z_t(ip, jp) + z_b(ip, jp)) - z0pres(ip, jp))));
          const double rho_anom = (Rho_T0_S0 - rho_ref) + ((dRho_dT * ((wtT_L * T(i, j)) + (wtT_R * T(ip, jp))) +
                                                            dRho_dS * ((wtT_L * S(i, j)) + (wtT_R * S(ip, jp)))) + dRho_dp * p_ave);
          intz[m] = G_e * rho_anom * dz;
        }
        out(i, j) = C1_90 * (7.0 * (intz[1] + intz[5]) + 32.0 * (intz[2] + intz[4]) + 12.0 * intz[3]);
      }
    }
  }
}

// int_density_dz_wright, MOM_EOS_Wright.F90:389-655
void int_density_dz_wright(const IntArgs& A) {
  const OGrid& G = *A.G;
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  const double C1_3 = 1.0 / 3.0, C1_7 = 1.0 / 7.0, C1_9 = 1.0 / 9.0, C1_90 = 1.0 / 90.0;
  // :497-534 (the scale factors are always present in the call from analytic_int_density_dz when any of them differs from 1;
  // multiplying by factors that equal 1 changes nothing, so one code path restates both calls)
  const double GxRho = A.pres_scale * A.G_e * A.rho_0, Pa_to_RL2_T2 = 1.0 / A.pres_scale;
  const double g_Earth = (A.pres_scale * A.G_e) * A.rho_scale;
  const double rho_ref_mks = A.rho_ref / A.rho_scale, I_Rho = A.rho_scale / A.rho_0;
  double a1s = a1, a2s = a2, b1s = b1, b2s = b2, b3s = b3, b4s = b4, b5s = b5, c1s = c1, c2s = c2, c3s = c3, c4s = c4, c5s = c5;
  if (A.temp_scale != 1.0) {
    const double ts = A.temp_scale;
    a1s = a1s * ts;
    b1s = b1s * ts; b2s = b2s * (ts * ts); b3s = b3s * (ts * ts * ts); b5s = b5s * ts;
    c1s = c1s * ts; c2s = c2s * (ts * ts); c3s = c3s * (ts * ts * ts); c5s = c5s * ts;
  }
  if (A.saln_scale != 1.0) {
    a2s = a2s * A.saln_scale;
    b4s = b4s * A.saln_scale; b5s = b5s * A.saln_scale;
    c4s = c4s * A.saln_scale; c5s = c5s * A.saln_scale;
  }
  const bool do_massWeight = (A.MassWghtInterp & 1) != 0, top_massWeight = (A.MassWghtInterp & 2) != 0;
  const V2 &T = A.T, &S = A.S, &z_t = A.z_t, &z_b = A.z_b, &z0pres = A.Z_0p, &dpa = A.dpa;
  A2 al0_2d = G.aH(), p0_2d = G.aH(), lambda_2d = G.aH();
  for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) {
    al0_2d(i, j) = (a0 + a1s * T(i, j)) + a2s * S(i, j);
    p0_2d(i, j) = (b0 + b4s * S(i, j)) + T(i, j) * (b1s + T(i, j) * ((b2s + b3s * T(i, j))) + b5s * S(i, j));
    lambda_2d(i, j) = (c0 + c4s * S(i, j)) + T(i, j) * (c1s + T(i, j) * ((c2s + c3s * T(i, j))) + c5s * S(i, j));
    const double al0 = al0_2d(i, j), p0 = p0_2d(i, j), lambda = lambda_2d(i, j);
    const double dz = z_t(i, j) - z_b(i, j);
    const double p_ave = -GxRho * (0.5 * (z_t(i, j) + z_b(i, j)) - z0pres(i, j));
    const double I_al0 = 1.0 / al0;
    const double I_Lzz = 1.0 / (p0 + (lambda * I_al0) + p_ave);
    const double eps = 0.5 * GxRho * dz * I_Lzz, eps2 = eps * eps;
    const double rho_anom = (p0 + p_ave) * (I_Lzz * I_al0) - rho_ref_mks;
    const double rem = I_Rho * (lambda * (I_al0 * I_al0)) * eps2 * (C1_3 + eps2 * (0.2 + eps2 * (C1_7 + C1_9 * eps2)));
    dpa(i, j) = Pa_to_RL2_T2 * (g_Earth * rho_anom * dz - 2.0 * eps * rem);
    A.intz_dpa(i, j) = Pa_to_RL2_T2 * (0.5 * g_Earth * rho_anom * (dz * dz) - dz * (1.0 + eps) * rem);
  }
  for (int dir = 0; dir < 2; ++dir) {
    const int di = dir == 0 ? 1 : 0, dj = dir == 0 ? 0 : 1;
    const int jlo = dir == 0 ? js : Jsq, jhi = dir == 0 ? je : Jeq, ilo = dir == 0 ? Isq : is, ihi = dir == 0 ? Ieq : ie;
    const V2& out = dir == 0 ? A.intx_dpa : A.inty_dpa;
    for (int j = jlo; j <= jhi; ++j) for (int i = ilo; i <= ihi; ++i) {
      const int ip = i + di, jp = j + dj;
      double hWght = 0.0;
      if (do_massWeight) hWght = max3(0., -A.bathyT(i, j) - z_t(ip, jp), -A.bathyT(ip, jp) - z_t(i, j));
      if (top_massWeight) hWght = max3(hWght, z_b(ip, jp) - A.SSH(i, j), z_b(i, j) - A.SSH(ip, jp));
      double hWt_LL, hWt_LR, hWt_RR, hWt_RL;
      if (hWght > 0.) {
        const double hL = (z_t(i, j) - z_b(i, j)) + A.dz_neglect, hR = (z_t(ip, jp) - z_b(ip, jp)) + A.dz_neglect;
        mass_weights(hWght, hL, hR, hWt_LL, hWt_LR, hWt_RR, hWt_RL);
      } else { hWt_LL = 1.0; hWt_LR = 0.0; hWt_RR = 1.0; hWt_RL = 0.0; }
      double intz[6];
      intz[1] = dpa(i, j); intz[5] = dpa(ip, jp);
      for (int m = 2; m <= 4; ++m) {
        const double wt_L = 0.25 * (double)(5 - m), wt_R = 1.0 - wt_L;
        const double wtT_L = (wt_L * hWt_LL) + (wt_R * hWt_RL), wtT_R = (wt_L * hWt_LR) + (wt_R * hWt_RR);
        const double al0 = (wtT_L * al0_2d(i, j)) + (wtT_R * al0_2d(ip, jp));
        const double p0 = (wtT_L * p0_2d(i, j)) + (wtT_R * p0_2d(ip, jp));
        const double lambda = (wtT_L * lambda_2d(i, j)) + (wtT_R * lambda_2d(ip, jp));
        const double dz = (wt_L * (z_t(i, j) - z_b(i, j))) + (wt_R * (z_t(ip, jp) - z_b(ip, jp)));
        const double p_ave = -GxRho * ((wt_L * (0.5 * (z_t(i, j) + z_b(i, j)) - z0pres(i, j))) +
                                       (wt_R * (0.5 * (z_t(ip, jp) + z_b(ip, jp)) - z0pres(ip, jp))));
        const double I_al0 = 1.0 / al0;
        const double I_Lzz = 1.0 / (p0 + (lambda * I_al0) + p_ave);
        const double eps = 0.5 * GxRho * dz * I_Lzz, eps2 = eps * eps;
        intz[m] = Pa_to_RL2_T2 * (g_Earth * dz * ((p0 + p_ave) * (I_Lzz * I_al0) - rho_ref_mks) -
                                  2.0 * eps * I_Rho * (lambda * (I_al0 * I_al0)) * eps2 * (C1_3 + eps2 * (0.2 + eps2 * (C1_7 + C1_9 * eps2))));
      }
      out(i, j) = C1_90 * (7.0 * (intz[1] + intz[5]) + 32.0 * (intz[2] + intz[4]) + 12.0 * intz[3]);
    }
  }
}


// int_density_dz_generic_plm (MOM_density_integrals.F90:418-868, ppm = false) and int_density_dz_generic_ppm (:874-1310, ppm = true)
// for layer k: 5-point Boole quadrature in the vertical of the density anomaly of linear / parabolic T,S profiles, and a 3 x 5 point
// quadrature along each face.  No Stanley SGS terms.  Tm, Sm are the layer means (tv%T, tv%S), used by the parabolic form only.
struct GenArgs {
  const OGrid* G; int k;
  V3 T_t, T_b, S_t, S_b, Tm, Sm, e;
  V2 dpa, intz_dpa, intx_dpa, inty_dpa, bathyT, Z_0p;
  double rho_ref, rho_0, G_e, dz_subroundoff, h_nv;
  int MassWghtInterp, MassWghtInterpVanOnly, use_inaccurate_form;
};

void int_density_dz_generic(const GenArgs& A, const EOSp& EOS, bool ppm) {
  const OGrid& G = *A.G;
  const int Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB, k = A.k;
  const V3 &T_t = A.T_t, &T_b = A.T_b, &S_t = A.S_t, &S_b = A.S_b, &e = A.e;
  const V2 &z0pres = A.Z_0p, &bathyT = A.bathyT, &dpa = A.dpa;
  const double C1_90 = 1.0 / 90.0, G_e = A.G_e, rho_ref = A.rho_ref, dz_subroundoff = A.dz_subroundoff;
  const double GxRho = A.G_e * A.rho_0;
  const double massWeightToggle = (A.MassWghtInterp & 1) ? 1. : 0., TopWeightToggle = (A.MassWghtInterp & 2) ? 1. : 0.;
  const double massWeightNVonlyToggle = A.MassWghtInterpVanOnly ? 0. : 1.;
  const double h_nonvanished = A.h_nv;
  const bool use_rho_ref = ppm ? true : !A.use_inaccurate_form;   // the parabolic form always passes rho_ref to the EOS (:1059)
  double wt_t[6], wt_b[6];
  for (int n = 1; n <= 5; ++n) { wt_t[n] = 0.25 * (double)(5 - n); wt_b[n] = 1.0 - wt_t[n]; }
  auto dens = [&](double T, double S, double p) { return use_rho_ref ? calculate_density(EOS, T, S, p, &rho_ref) : calculate_density(EOS, T, S, p, nullptr); };

  // 1. vertical integrals (:563-614 / :1030-1077)
  for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) {
    double s6 = 0., t6 = 0.;
    if (ppm) {
      s6 = 3.0 * (2.0 * A.Sm(i, j, k) - (S_t(i, j, k) + S_b(i, j, k)));
      t6 = 3.0 * (2.0 * A.Tm(i, j, k) - (T_t(i, j, k) + T_b(i, j, k)));
    }
    const double dz = e(i, j, k) - e(i, j, k + 1);
    double r5[6], u5[6];
    for (int n = 1; n <= 5; ++n) {
      const double p5 = -GxRho * ((e(i, j, k) - z0pres(i, j)) - 0.25 * (double)(n - 1) * dz);
      double S5, T5;
      if (ppm) {
        S5 = wt_t[n] * S_t(i, j, k) + wt_b[n] * (S_b(i, j, k) + s6 * wt_t[n]);
        T5 = wt_t[n] * T_t(i, j, k) + wt_b[n] * (T_b(i, j, k) + t6 * wt_t[n]);
      } else {
        S5 = wt_t[n] * S_t(i, j, k) + wt_b[n] * S_b(i, j, k);
        T5 = wt_t[n] * T_t(i, j, k) + wt_b[n] * T_b(i, j, k);
      }
      r5[n] = dens(T5, S5, p5);
      u5[n] = r5[n] - rho_ref;
    }
    if (use_rho_ref) {
      const double rho_anom = C1_90 * (7.0 * (r5[1] + r5[5]) + 32.0 * (r5[2] + r5[4]) + 12.0 * r5[3]);
      dpa(i, j) = G_e * dz * rho_anom;
      A.intz_dpa(i, j) = 0.5 * G_e * (dz * dz) * (rho_anom - C1_90 * (16.0 * (r5[4] - r5[2]) + 7.0 * (r5[5] - r5[1])));
    } else {
      const double rho_anom = C1_90 * (7.0 * (r5[1] + r5[5]) + 32.0 * (r5[2] + r5[4]) + 12.0 * r5[3]) - rho_ref;
      dpa(i, j) = G_e * dz * rho_anom;
      A.intz_dpa(i, j) = 0.5 * G_e * (dz * dz) * (rho_anom - C1_90 * (16.0 * (u5[4] - u5[2]) + 7.0 * (u5[5] - u5[1])));
    }
  }
  // 2./3. horizontal integrals along x (dir 0: :617-744 / :1080-1198) and y (dir 1: :747-866 / :1201-1308)
  for (int dir = 0; dir < 2; ++dir) {
    const int di = dir == 0 ? 1 : 0, dj = dir == 0 ? 0 : 1;
    const int jlo = dir == 0 ? G.jsc : Jsq, jhi = dir == 0 ? G.jec : Jeq, ilo = dir == 0 ? Isq : G.isc, ihi = dir == 0 ? Ieq : G.iec;
    const V2& out = dir == 0 ? A.intx_dpa : A.inty_dpa;
    for (int j = jlo; j <= jhi; ++j) for (int i = ilo; i <= ihi; ++i) {
      const int ip = i + di, jp = j + dj;
      double hWght = massWeightToggle * max3(0., -bathyT(i, j) - e(ip, jp, k), -bathyT(ip, jp) - e(i, j, k));
      const double hWghtTop = TopWeightToggle * max3(0., e(ip, jp, k + 1) - e(i, j, 1), e(i, j, k + 1) - e(ip, jp, 1));
      hWght = fmax2(hWght, hWghtTop);
      if (((e(i, j, k) - e(i, j, k + 1)) > h_nonvanished) && ((e(ip, jp, k) - e(ip, jp, k + 1)) > h_nonvanished))
        hWght = massWeightNVonlyToggle * hWght;
      double Ttl, Tbl, Tml = 0., Ttr, Tbr, Tmr = 0., Stl, Sbl, Sml = 0., Str, Sbr, Smr = 0.;
      if (hWght > 0.) {
        const double hL = (e(i, j, k) - e(i, j, k + 1)) + dz_subroundoff;
        const double hR = (e(ip, jp, k) - e(ip, jp, k + 1)) + dz_subroundoff;
        const double r = (hL - hR) / (hL + hR);
        hWght = hWght * (r * r);
        const double iDenom = 1. / (hWght * (hR + hL) + hL * hR);
        Ttl = ((hWght * hR) * T_t(ip, jp, k) + (hWght * hL + hR * hL) * T_t(i, j, k)) * iDenom;
        Ttr = ((hWght * hL) * T_t(i, j, k) + (hWght * hR + hR * hL) * T_t(ip, jp, k)) * iDenom;
        Tbl = ((hWght * hR) * T_b(ip, jp, k) + (hWght * hL + hR * hL) * T_b(i, j, k)) * iDenom;
        Tbr = ((hWght * hL) * T_b(i, j, k) + (hWght * hR + hR * hL) * T_b(ip, jp, k)) * iDenom;
        Stl = ((hWght * hR) * S_t(ip, jp, k) + (hWght * hL + hR * hL) * S_t(i, j, k)) * iDenom;
        Str = ((hWght * hL) * S_t(i, j, k) + (hWght * hR + hR * hL) * S_t(ip, jp, k)) * iDenom;
        Sbl = ((hWght * hR) * S_b(ip, jp, k) + (hWght * hL + hR * hL) * S_b(i, j, k)) * iDenom;
        Sbr = ((hWght * hL) * S_b(i, j, k) + (hWght * hR + hR * hL) * S_b(ip, jp, k)) * iDenom;
        if (ppm) {
          Tml = ((hWght * hR) * A.Tm(ip, jp, k) + (hWght * hL + hR * hL) * A.Tm(i, j, k)) * iDenom;
          Tmr = ((hWght * hL) * A.Tm(i, j, k) + (hWght * hR + hR * hL) * A.Tm(ip, jp, k)) * iDenom;
          Sml = ((hWght * hR) * A.Sm(ip, jp, k) + (hWght * hL + hR * hL) * A.Sm(i, j, k)) * iDenom;
          Smr = ((hWght * hL) * A.Sm(i, j, k) + (hWght * hR + hR * hL) * A.Sm(ip, jp, k)) * iDenom;
        }
      } else {
        Ttl = T_t(i, j, k); Tbl = T_b(i, j, k); Ttr = T_t(ip, jp, k); Tbr = T_b(ip, jp, k);
        Stl = S_t(i, j, k); Sbl = S_b(i, j, k); Str = S_t(ip, jp, k); Sbr = S_b(ip, jp, k);
        if (ppm) { Tml = A.Tm(i, j, k); Tmr = A.Tm(ip, jp, k); Sml = A.Sm(i, j, k); Smr = A.Sm(ip, jp, k); }
      }
      double intz[6];
      intz[1] = dpa(i, j); intz[5] = dpa(ip, jp);
      for (int m = 2; m <= 4; ++m) {
        const double w_left = wt_t[m], w_right = wt_b[m];
        const double dz_x = (w_left * (e(i, j, k) - e(i, j, k + 1))) + (w_right * (e(ip, jp, k) - e(ip, jp, k + 1)));
        double T15[6], S15[6], p15[6], r15[6];
        p15[1] = -GxRho * ((w_left * (e(i, j, k) - z0pres(i, j))) + (w_right * (e(ip, jp, k) - z0pres(ip, jp))));
        for (int n = 2; n <= 5; ++n) p15[n] = p15[n - 1] + GxRho * 0.25 * dz_x;
        if (ppm) {
          const double T_top = (w_left * Ttl) + (w_right * Ttr), T_mn = (w_left * Tml) + (w_right * Tmr), T_bot = (w_left * Tbl) + (w_right * Tbr);
          const double S_top = (w_left * Stl) + (w_right * Str), S_mn = (w_left * Sml) + (w_right * Smr), S_bot = (w_left * Sbl) + (w_right * Sbr);
          const double s6 = 3.0 * (2.0 * S_mn - (S_top + S_bot));
          const double t6 = 3.0 * (2.0 * T_mn - (T_top + T_bot));
          for (int n = 1; n <= 5; ++n) {
            S15[n] = wt_t[n] * S_top + wt_b[n] * (S_bot + s6 * wt_t[n]);
            T15[n] = wt_t[n] * T_top + wt_b[n] * (T_bot + t6 * wt_t[n]);
          }
        } else {
          T15[1] = (w_left * Ttl) + (w_right * Ttr); T15[5] = (w_left * Tbl) + (w_right * Tbr);
          S15[1] = (w_left * Stl) + (w_right * Str); S15[5] = (w_left * Sbl) + (w_right * Sbr);
          for (int n = 2; n <= 4; ++n) {
            S15[n] = wt_t[n] * S15[1] + wt_b[n] * S15[5];
            T15[n] = wt_t[n] * T15[1] + wt_b[n] * T15[5];
          }
        }
        for (int n = 1; n <= 5; ++n) r15[n] = dens(T15[n], S15[n], p15[n]);
        if (use_rho_ref) intz[m] = (G_e * dz_x * (C1_90 * (7.0 * (r15[1] + r15[5]) + 32.0 * (r15[2] + r15[4]) + 12.0 * r15[3])));
        else intz[m] = (G_e * dz_x * (C1_90 * (7.0 * (r15[1] + r15[5]) + 32.0 * (r15[2] + r15[4]) + 12.0 * r15[3]) - rho_ref));
      }
      out(i, j) = C1_90 * (7.0 * (intz[1] + intz[5]) + 32.0 * (intz[2] + intz[4]) + 12.0 * intz[3]);
    }
  }
}

// TS_PLM_edge_values / TS_PPM_edge_values, MOM_ALE.F90:1495-1660 (answer_date >= 20190101: h_neglect = h_neglect_edge = GV%H_subroundoff)
void ts_edge_values(const OGrid& G, int scheme, bool bdry_extrap, double H_subroundoff, const V3& h, const V3& T, const V3& S, const V3& T_t,
                    const V3& T_b, const V3& S_t, const V3& S_b) {
  const int nz = G.ke;
#pragma omp parallel for
  for (int j = G.jsc - 1; j <= G.jec + 1; ++j) {
    std::vector<double> hc(nz + 2), q(nz + 2), qt(nz + 2), qb(nz + 2);
    for (int i = G.isc - 1; i <= G.iec + 1; ++i) {
      for (int k = 1; k <= nz; ++k) hc[k] = h(i, j, k);
      for (int f = 0; f < 2; ++f) {   // salinity first, then temperature (the order of the reference; the fields are independent)
        const V3& Q = f == 0 ? S : T; const V3& Qt = f == 0 ? S_t : T_t; const V3& Qb = f == 0 ? S_b : T_b;
        for (int k = 1; k <= nz; ++k) q[k] = Q(i, j, k);
        if (scheme == 1) ale_plm_edge_values_column(nz, hc.data(), q.data(), bdry_extrap, H_subroundoff, qt.data(), qb.data());
        else ale_ppm_edge_values_column(nz, hc.data(), q.data(), bdry_extrap, H_subroundoff, H_subroundoff, qt.data(), qb.data());
        for (int k = 1; k <= nz; ++k) { Qt(i, j, k) = qt[k]; Qb(i, j, k) = qb[k]; }
      }
    }
  }
}

}  // namespace

extern "C" int oracle_pressure_force(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV,
                                     const mom6cu_pressureforce_cs* CS, const mom6cu_pressureforce_args* A, int nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  if (CS->unsupported || !GV->Boussinesq) return 3;
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  const int nz = G.ke;
  const bool use_EOS = CS->EOS_form != MOM6CU_EOS_NONE, use_p_atm = A->p_atm != nullptr;
  EOSp E = {CS->EOS_form, CS->Rho_T0_S0, CS->dRho_dT, CS->dRho_dS, CS->dRho_dp};
  if (CS->kg_m3_to_R != 0.0) { E.kg_m3_to_R = CS->kg_m3_to_R; E.R_to_kg_m3 = 1.0 / CS->kg_m3_to_R; }
  if (CS->RL2_T2_to_Pa != 0.0) E.RL2_T2_to_Pa = CS->RL2_T2_to_Pa;
  if (CS->C_to_degC != 0.0) E.C_to_degC = CS->C_to_degC;
  if (CS->S_to_ppt != 0.0) E.S_to_ppt = CS->S_to_ppt;
  // use_ALE = CS%reconstruct .and. use_EOS with an ALE control structure (:1120-1122); Recon_Scheme 1 = PLM, 2 = PPM (:2181)
  const bool use_ALE = CS->reconstruct && use_EOS;
  if (use_ALE && CS->Recon_Scheme > 0 && (CS->Recon_Scheme > 2 || CS->ALE_answer_date < 20190101 || nz < (CS->Recon_Scheme == 2 ? 4 : 2))) return 3;
  const V3 h = G.H3(A->h), PFu = G.U3(A->PFu), PFv = G.V3_(A->PFv);
  V3 T, S, pbce;
  if (use_EOS) { T = G.H3(A->T); S = G.H3(A->S); }
  if (A->pbce) pbce = G.H3(A->pbce);
  V2 p_atm, eta;
  if (use_p_atm) p_atm = G.H(A->p_atm);
  if (A->eta) eta = G.H(A->eta);
  // :1126-1144
  const double h_neglect = GV->H_subroundoff, dz_neglect = CS->dZ_subroundoff;
  const double I_Rho0 = 1.0 / GV->Rho0, G_Rho0 = GV->g_Earth / GV->Rho0, GxRho0 = GV->g_Earth * GV->Rho0;
  const double rho_ref = CS->rho_ref;
  double rho0_int_density, rho0_set_pbce, GxRho_ref, I_g_rho;
  if (CS->rho_ref_bug) { rho0_int_density = rho_ref; rho0_set_pbce = rho_ref; GxRho_ref = GxRho0; I_g_rho = 1.0 / (rho_ref * GV->g_Earth); }
  else { rho0_int_density = GV->Rho0; rho0_set_pbce = GV->Rho0; GxRho_ref = GV->g_Earth * rho_ref; I_g_rho = 1.0 / (GV->Rho0 * GV->g_Earth); }

  A3 e(G.isd, G.ied, G.jsd, G.jed, nz + 1), pa(G.isd, G.ied, G.jsd, G.jed, nz + 1);
  A3 dpa(G.isd, G.ied, G.jsd, G.jed, nz), intz_dpa(G.isd, G.ied, G.jsd, G.jed, nz);
  A3 intx_pa(G.isd - 1, G.ied, G.jsd, G.jed, nz + 1), intx_dpa(G.isd - 1, G.ied, G.jsd, G.jed, nz);
  A3 inty_pa(G.isd, G.ied, G.jsd - 1, G.jed, nz + 1), inty_dpa(G.isd, G.ied, G.jsd - 1, G.jed, nz);
  A2 Z_0p = G.aH();
  auto plane = [&](const V3& a, int k) { return V2(a.p + (size_t)(k - 1) * a.ni * a.nj, a.ilo, a.ilo + a.ni - 1, a.jlo, a.jlo + a.nj - 1); };

  for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) e(i, j, nz + 1) = -G.bathyT(i, j);  // :1150-1152
  for (int j = Jsq; j <= Jeq + 1; ++j) for (int k = nz; k >= 1; --k) for (int i = Isq; i <= Ieq + 1; ++i)
    e(i, j, k) = e(i, j, k + 1) + h(i, j, k) * GV->H_to_Z;  // :1200-1202
  // :1252-1276
  for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) {
    if (use_p_atm) pa(i, j, 1) = GxRho_ref * (e(i, j, 1) - CS->Z_ref) + p_atm(i, j);
    else pa(i, j, 1) = GxRho_ref * (e(i, j, 1) - CS->Z_ref);
    if (CS->use_SSH_in_Z0p && use_p_atm) Z_0p(i, j) = e(i, j, 1) + p_atm(i, j) * I_g_rho;
    else if (CS->use_SSH_in_Z0p) Z_0p(i, j) = e(i, j, 1);
    else Z_0p(i, j) = CS->Z_ref;
  }
  // :1241-1250: sub-layer T,S profiles for the pressure integrals
  A3 T_t(G.isd, G.ied, G.jsd, G.jed, use_ALE ? nz : 1), T_b(G.isd, G.ied, G.jsd, G.jed, use_ALE ? nz : 1);
  A3 S_t(G.isd, G.ied, G.jsd, G.jed, use_ALE ? nz : 1), S_b(G.isd, G.ied, G.jsd, G.jed, use_ALE ? nz : 1);
  if (use_ALE && CS->Recon_Scheme > 0) ts_edge_values(G, CS->Recon_Scheme, CS->boundary_extrap != 0, GV->H_subroundoff, h, T, S, T_t, T_b, S_t, S_b);
#pragma omp parallel for
  for (int k = 1; k <= nz; ++k) {  // :1278-1337
    if (use_EOS) {
      IntArgs I = {&G, plane(T, k), plane(S, k), plane(e, k), plane(e, k + 1), plane(dpa, k), plane(intz_dpa, k), plane(intx_dpa, k),
                   plane(inty_dpa, k), G.bathyT, plane(e, 1), Z_0p, rho_ref, rho0_int_density, GV->g_Earth, dz_neglect, CS->MassWghtInterp};
      if (use_ALE && CS->Recon_Scheme > 0) {  // :1286-1300
        GenArgs Gn = {&G, k, T_t, T_b, S_t, S_b, T, S, e, plane(dpa, k), plane(intz_dpa, k), plane(intx_dpa, k), plane(inty_dpa, k), G.bathyT, Z_0p,
                      rho_ref, rho0_int_density, GV->g_Earth, dz_neglect, GV->H_to_Z * CS->h_nonvanished, CS->MassWghtInterp,
                      CS->MassWghtInterpVanOnly, CS->use_inaccurate_pgf_rho_anom};
        int_density_dz_generic(Gn, E, CS->Recon_Scheme == 2);
      } else if (CS->EOS_form == MOM6CU_EOS_LINEAR) {  // analytic_int_density_dz, MOM_EOS.F90:1440-1455
        EOSp El = E;
        El.Rho_T0_S0 = E.kg_m3_to_R * E.Rho_T0_S0; El.dRho_dT = (E.kg_m3_to_R * E.C_to_degC) * E.dRho_dT;
        El.dRho_dS = (E.kg_m3_to_R * E.S_to_ppt) * E.dRho_dS; El.dRho_dp = (E.kg_m3_to_R * E.RL2_T2_to_Pa) * E.dRho_dp;
        int_density_dz_linear(I, El);
      } else {
        I.rho_scale = E.kg_m3_to_R; I.pres_scale = E.RL2_T2_to_Pa; I.temp_scale = E.C_to_degC; I.saln_scale = E.S_to_ppt;
        int_density_dz_wright(I);
      }
      if (GV->Z_to_H != 1.0) for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) intz_dpa(i, j, k) = intz_dpa(i, j, k) * GV->Z_to_H;
    } else {
      A2 dz_geo = G.aH();
      const double Rlay = CS->Rlay[k - 1];
      for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) {
        dz_geo(i, j) = GV->g_Earth * GV->H_to_Z * h(i, j, k);
        dpa(i, j, k) = (Rlay - rho_ref) * dz_geo(i, j);
        intz_dpa(i, j, k) = 0.5 * (Rlay - rho_ref) * dz_geo(i, j) * h(i, j, k);
      }
      for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) intx_dpa(I, j, k) = 0.5 * (Rlay - rho_ref) * (dz_geo(I, j) + dz_geo(I + 1, j));
      for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) inty_dpa(i, J, k) = 0.5 * (Rlay - rho_ref) * (dz_geo(i, J) + dz_geo(i, J + 1));
    }
  }
  for (int k = 1; k <= nz; ++k) for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i)
    pa(i, j, k + 1) = pa(i, j, k) + dpa(i, j, k);  // :1340-1345
  // :1538-1558
  for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) intx_pa(I, j, 1) = 0.5 * (pa(I, j, 1) + pa(I + 1, j, 1));
  for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) inty_pa(i, J, 1) = 0.5 * (pa(i, J, 1) + pa(i, J + 1, 1));
  for (int k = 1; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) intx_pa(I, j, k + 1) = intx_pa(I, j, k) + intx_dpa(I, j, k);
  for (int k = 1; k <= nz; ++k) for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) inty_pa(i, J, k + 1) = inty_pa(i, J, k) + inty_dpa(i, J, k);
  // :1795-1813
#pragma omp parallel for
  for (int k = 1; k <= nz; ++k) {
    for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I)
      PFu(I, j, k) = (((pa(I, j, k) * h(I, j, k) + intz_dpa(I, j, k)) - (pa(I + 1, j, k) * h(I + 1, j, k) + intz_dpa(I + 1, j, k))) +
                      ((h(I + 1, j, k) - h(I, j, k)) * intx_pa(I, j, k) - (e(I + 1, j, k + 1) - e(I, j, k + 1)) * intx_dpa(I, j, k) * GV->Z_to_H)) *
                     ((2.0 * I_Rho0 * G.IdxCu(I, j)) / ((h(I, j, k) + h(I + 1, j, k)) + h_neglect));
    for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i)
      PFv(i, J, k) = (((pa(i, J, k) * h(i, J, k) + intz_dpa(i, J, k)) - (pa(i, J + 1, k) * h(i, J + 1, k) + intz_dpa(i, J + 1, k))) +
                      ((h(i, J + 1, k) - h(i, J, k)) * inty_pa(i, J, k) - (e(i, J + 1, k + 1) - e(i, J, k + 1)) * inty_dpa(i, J, k) * GV->Z_to_H)) *
                     ((2.0 * I_Rho0 * G.IdyCv(i, J)) / ((h(i, J, k) + h(i, J + 1, k)) + h_neglect));
  }
  if (CS->GFS_scale < 1.0) {  // :1843-1875
    A2 dM = G.aH();
    for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) {
      if (use_EOS) {
        const double rho_in_situ = density(E, T(i, j, 1), S(i, j, 1), use_p_atm ? p_atm(i, j) : 0.0);
        dM(i, j) = (CS->GFS_scale - 1.0) * (G_Rho0 * rho_in_situ) * (e(i, j, 1) - CS->Z_ref);
      } else dM(i, j) = (CS->GFS_scale - 1.0) * (G_Rho0 * CS->Rlay[0]) * (e(i, j, 1) - CS->Z_ref);
    }
    for (int k = 1; k <= nz; ++k) {
      for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) PFu(I, j, k) = PFu(I, j, k) - (dM(I + 1, j) - dM(I, j)) * G.IdxCu(I, j);
      for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) PFv(i, J, k) = PFv(i, J, k) - (dM(i, J + 1) - dM(i, J)) * G.IdyCv(i, J);
    }
  }
  if (A->pbce) {  // Set_pbce_Bouss, MOM_PressureForce_Montgomery.F90:685-745
    const double Rho0xG = rho0_set_pbce * GV->g_Earth;
    for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) {
      if (use_EOS) {
        const double Ihtot = GV->H_to_Z / ((e(i, j, 1) - e(i, j, nz + 1)) + dz_neglect);
        double press = -Rho0xG * (e(i, j, 1) - CS->Z_ref);
        const double rho_in_situ = density(E, T(i, j, 1), S(i, j, 1), press);
        pbce(i, j, 1) = G_Rho0 * (CS->GFS_scale * rho_in_situ) * GV->H_to_Z;
        for (int k = 2; k <= nz; ++k) {
          press = -Rho0xG * (e(i, j, k) - CS->Z_ref);
          const double T_int = 0.5 * (T(i, j, k - 1) + T(i, j, k)), S_int = 0.5 * (S(i, j, k - 1) + S(i, j, k));
          double dR_dT, dR_dS;
          density_derivs(E, T_int, S_int, press, dR_dT, dR_dS);
          pbce(i, j, k) = pbce(i, j, k - 1) + G_Rho0 * ((e(i, j, k) - e(i, j, nz + 1)) * Ihtot) *
                          (dR_dT * (T(i, j, k) - T(i, j, k - 1)) + dR_dS * (S(i, j, k) - S(i, j, k - 1)));
        }
      } else {
        const double Ihtot = 1.0 / ((e(i, j, 1) - e(i, j, nz + 1)) + dz_neglect);
        pbce(i, j, 1) = CS->g_prime[0] * GV->H_to_Z;
        for (int k = 2; k <= nz; ++k) pbce(i, j, k) = pbce(i, j, k - 1) + (CS->g_prime[k - 1] * GV->H_to_Z) * ((e(i, j, k) - e(i, j, nz + 1)) * Ihtot);
      }
    }
  }
  if (A->eta) for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) eta(i, j) = e(i, j, 1) * GV->Z_to_H;  // :1885-1887
  return 0;
}

// Accessors for the known-answer tests of the equation of state (tests/test_oracle_eos_kat.py): which = 0 density, 1 density anomaly
// from rho_ref, 2 drho_dT, 3 drho_dS; scales = {kg_m3_to_R, RL2_T2_to_Pa, C_to_degC, S_to_ppt} or NULL.
extern "C" double oracle_eos_eval(int which, int form, const double* lin4, const double* scales, double T, double S, double p, double rho_ref) {
  EOSp E = {form, lin4 ? lin4[0] : 0., lin4 ? lin4[1] : 0., lin4 ? lin4[2] : 0., lin4 ? lin4[3] : 0.};
  if (scales) { E.kg_m3_to_R = scales[0]; E.R_to_kg_m3 = 1.0 / scales[0]; E.RL2_T2_to_Pa = scales[1]; E.C_to_degC = scales[2]; E.S_to_ppt = scales[3]; }
  if (which == 0) return calculate_density(E, T, S, p, nullptr);
  if (which == 1) return calculate_density(E, T, S, p, &rho_ref);
  double dT, dS;
  calculate_density_derivs(E, T, S, p, dT, dS);
  return which == 2 ? dT : dS;
}
