// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of PressureForce_FV_Bouss, /root/reference/src/core/MOM_PressureForce_FV.F90:947-2017, with
//   int_density_dz -> analytic_int_density_dz (src/core/MOM_density_integrals.F90:42-103, src/equation_of_state/MOM_EOS.F90:1384-1499)
//   int_density_dz_linear (src/equation_of_state/MOM_EOS_linear.F90:275-440)
//   int_density_dz_wright (src/equation_of_state/MOM_EOS_Wright.F90:389-655), density_elem / calculate_density_derivs_elem (:80, :178)
//   Set_pbce_Bouss (src/core/MOM_PressureForce_Montgomery.F90:649-748)
// Frozen options: no tides/SAL, no Stanley term, no intxpa corrections/resets, nk_rho_varies = 0, layer-constant T,S.
#include "oracle.h"
#include "ogrid.hpp"
#include <cmath>
#include <omp.h>

using namespace orc;

namespace {

// Wright (1997) "buggy" fit used by EOS_WRIGHT, MOM_EOS_Wright.F90:23-37
const double a0 = 7.057924e-4, a1 = 3.480336e-7, a2 = -1.112733e-7;
const double b0 = 5.790749e8, b1 = 3.516535e6, b2 = -4.002714e4, b3 = 2.084372e2, b4 = 5.944068e5, b5 = -9.643486e3;
const double c0 = 1.704853e5, c1 = 7.904722e2, c2 = -7.984422, c3 = 5.140652e-2, c4 = -2.302158e2, c5 = -3.079464;

inline double max3(double a, double b, double c) { return fmax2(fmax2(a, b), c); }

struct EOSp { int form; double Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp; };

inline double density(const EOSp& E, double T, double S, double p) {
  if (E.form == MOM6CU_EOS_LINEAR) return E.Rho_T0_S0 + E.dRho_dT * T + E.dRho_dS * S + E.dRho_dp * p;  // MOM_EOS_linear.F90:66
  const double al0 = (a0 + a1 * T) + a2 * S;  // MOM_EOS_Wright.F90:91-94
  const double p0 = (b0 + b4 * S) + T * (b1 + T * (b2 + b3 * T) + b5 * S);
  const double lambda = (c0 + c4 * S) + T * (c1 + T * (c2 + c3 * T) + c5 * S);
  return (p + p0) / (lambda + al0 * (p + p0));
}
inline void density_derivs(const EOSp& E, double T, double S, double p, double& drho_dT, double& drho_dS) {
  if (E.form == MOM6CU_EOS_LINEAR) { drho_dT = E.dRho_dT; drho_dS = E.dRho_dS; return; }  // MOM_EOS_linear.F90:131-132
  const double al0 = (a0 + a1 * T) + a2 * S;  // MOM_EOS_Wright.F90:193-204
  const double p0 = (b0 + b4 * S) + T * (b1 + T * ((b2 + b3 * T)) + b5 * S);
  const double lambda = (c0 + c4 * S) + T * (c1 + T * ((c2 + c3 * T)) + c5 * S);
  double I_denom2 = 1.0 / (lambda + al0 * (p + p0));
  I_denom2 = I_denom2 * I_denom2;
  drho_dT = I_denom2 * (lambda * (b1 + T * (2.0 * b2 + 3.0 * b3 * T) + b5 * S) -
                        (p + p0) * ((p + p0) * a1 + (c1 + T * (c2 * 2.0 + c3 * 3.0 * T) + c5 * S)));
  drho_dS = I_denom2 * (lambda * (b4 + b5 * T) - (p + p0) * ((p + p0) * a2 + (c4 + c5 * T)));
}

struct IntArgs {
  const OGrid* G; V2 T, S, z_t, z_b, dpa, intz_dpa, intx_dpa, inty_dpa, bathyT, SSH, Z_0p;
  double rho_ref, rho_0, G_e, dz_neglect; int MassWghtInterp;
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

// int_density_dz_wright, MOM_EOS_Wright.F90:389-655 (no unit rescaling: rho_scale, pres_scale, temp_scale, saln_scale absent)
void int_density_dz_wright(const IntArgs& A) {
  const OGrid& G = *A.G;
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  const double C1_3 = 1.0 / 3.0, C1_7 = 1.0 / 7.0, C1_9 = 1.0 / 9.0, C1_90 = 1.0 / 90.0;
  const double GxRho = A.G_e * A.rho_0, g_Earth = A.G_e, Pa_to_RL2_T2 = 1.0, rho_ref_mks = A.rho_ref, I_Rho = 1.0 / A.rho_0;
  const bool do_massWeight = (A.MassWghtInterp & 1) != 0, top_massWeight = (A.MassWghtInterp & 2) != 0;
  const V2 &T = A.T, &S = A.S, &z_t = A.z_t, &z_b = A.z_b, &z0pres = A.Z_0p, &dpa = A.dpa;
  A2 al0_2d = G.aH(), p0_2d = G.aH(), lambda_2d = G.aH();
  for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) {
    al0_2d(i, j) = (a0 + a1 * T(i, j)) + a2 * S(i, j);
    p0_2d(i, j) = (b0 + b4 * S(i, j)) + T(i, j) * (b1 + T(i, j) * ((b2 + b3 * T(i, j))) + b5 * S(i, j));
    lambda_2d(i, j) = (c0 + c4 * S(i, j)) + T(i, j) * (c1 + T(i, j) * ((c2 + c3 * T(i, j))) + c5 * S(i, j));
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

}  // namespace

extern "C" int oracle_pressure_force(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV,
                                     const mom6cu_pressureforce_cs* CS, const mom6cu_pressureforce_args* A, int nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  if (CS->unsupported || !GV->Boussinesq) return 3;
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  const int nz = G.ke;
  const bool use_EOS = CS->EOS_form != MOM6CU_EOS_NONE, use_p_atm = A->p_atm != nullptr;
  const EOSp E = {CS->EOS_form, CS->Rho_T0_S0, CS->dRho_dT, CS->dRho_dS, CS->dRho_dp};
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
#pragma omp parallel for
  for (int k = 1; k <= nz; ++k) {  // :1278-1337
    if (use_EOS) {
      IntArgs I = {&G, plane(T, k), plane(S, k), plane(e, k), plane(e, k + 1), plane(dpa, k), plane(intz_dpa, k), plane(intx_dpa, k),
                   plane(inty_dpa, k), G.bathyT, plane(e, 1), Z_0p, rho_ref, rho0_int_density, GV->g_Earth, dz_neglect, CS->MassWghtInterp};
      if (CS->EOS_form == MOM6CU_EOS_LINEAR) int_density_dz_linear(I, E); else int_density_dz_wright(I);
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
