// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of /root/reference/src/parameterizations/lateral/MOM_mixed_layer_restrat.F90: mixedlayer_restrat :149-186 ->
// mixedlayer_restrat_OM4 :189-714 (Boussinesq; MLE_USE_PBL_MLD or detect_mld :1503-1569; no Stanley variance; constant front length), mu :717-751;
// find_ustar_mech_forcing (src/core/MOM_forcing_type.F90:1236-1296, the forces%ustar / H_T_units branch :1270-1272);
// density_elem of EOS_LINEAR (src/equation_of_state/MOM_EOS_linear.F90:60-68) and EOS_WRIGHT (MOM_EOS_Wright.F90:80-97).
// PARITY: mu is PINNED by the reference's unit test (mixedlayer_restrat_unit_tests :2014-2041), see tests/test_mle.py;
// the routine as a whole is PINNED BY A REFERENCE RUN: the reference's own mixedlayer_restrat, executed by oracle/f90run, agrees bit for
// bit on 4 option sets (tests/test_reference_f90.py, mixedlayer_restrat/*).
#include "oracle.h"
#include "ogrid.hpp"
#include <cmath>
#include <vector>

using namespace orc;

namespace {
const double a0 = 7.057924e-4, a1 = 3.480336e-7, a2 = -1.112733e-7;
const double b0 = 5.790749e8, b1 = 3.516535e6, b2 = -4.002714e4, b3 = 2.084372e2, b4 = 5.944068e5, b5 = -9.643486e3;
const double c0 = 1.704853e5, c1 = 7.904722e2, c2 = -7.984422, c3 = 5.140652e-2, c4 = -2.302158e2, c5 = -3.079464;

inline double density(const mom6cu_mle_cs* E, double T, double S, double p) {
  if (E->EOS_form == MOM6CU_EOS_LINEAR) return E->Rho_T0_S0 + E->dRho_dT * T + E->dRho_dS * S + E->dRho_dp * p;
  const double al0 = (a0 + a1 * T) + a2 * S;
  const double p0 = (b0 + b4 * S) + T * (b1 + T * (b2 + b3 * T) + b5 * S);
  const double lambda = (c0 + c4 * S) + T * (c1 + T * (c2 + c3 * T) + c5 * S);
  return (p + p0) / (lambda + al0 * (p + p0));
}
}  // namespace

// mu :717-751
extern "C" double oracle_mle_mu(double sigma, double dh) {
  double mu = fmax2(0., (1. - (2. * sigma + 1.) * (2. * sigma + 1.)) * (1. + (5. / 21.) * ((2. * sigma + 1.) * (2. * sigma + 1.))));
  const double xp = fmax2(0., fmin2(1., (-sigma - 0.5) * 2. / (1. + 2. * dh)));
  const double dd = std::pow(fmax2(1. - (xp * xp) * (3. - 2. * xp), 0.), 1. + 2. * dh);
  const double bottop = 0.5 * (1. - std::copysign(1., sigma + 0.5));
  mu = fmax2(mu, dd * bottop);
  return mu;
}

// mixedlayer_restrat_OM4 :189-714.  Returns 0 or 3 (an option outside the frozen set) / 2 (a FATAL of the routine).
extern "C" int oracle_mixedlayer_restrat(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV, mom6cu_mle_cs* CS,
                                         double* hp, double* uhtrp, double* vhtrp, const double* Tp, const double* Sp,
                                         const double* ustarp, double dt, const double* h_MLDp, const double* Rd_dx_hp) {
  if (!GV->Boussinesq || CS->use_Bodner || CS->use_Stanley_ML || CS->fl_from_file) return 3;
  if (CS->EOS_form != MOM6CU_EOS_LINEAR && CS->EOS_form != MOM6CU_EOS_WRIGHT) return 2;  // "An equation of state must be used with this module."
  if (CS->front_length > 0. && !Rd_dx_hp) return 2;  // "The resolution argument, Rd/dx, was not associated."
  if (!(CS->MLE_density_diff > 0.) && (!CS->MLE_use_PBL_MLD || !h_MLDp)) return 2;  // "No MLD to use for MLE parameterization."
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, nz = G.ke;
  const V3 h = G.H3(hp), uhtr = G.U3(uhtrp), vhtr = G.V3_(vhtrp), T = G.H3((double*)Tp), S = G.H3((double*)Sp);
  V2 h_MLD; if (h_MLDp) h_MLD = G.H((double*)h_MLDp);
  const V2 ustar = G.H((double*)ustarp), MLD_filtered = G.H(CS->MLD_filtered),
           MLD_filtered_slow = G.H(CS->MLD_filtered_slow);
  V2 Rd_dx_h; if (Rd_dx_hp) Rd_dx_h = G.H((double*)Rd_dx_hp);
  A3 uhml(G.isd - 1, G.ied, G.jsd, G.jed, nz), vhml(G.isd, G.ied, G.jsd - 1, G.jed, nz), h_avail(G.isd, G.ied, G.jsd, G.jed, nz);
  A2 U_star_2d = G.aH(), MLD_fast = G.aH(), htot_fast = G.aH(), Rml_av_fast = G.aH(), MLD_slow = G.aH(), mle_fl_2d = G.aH(),
     htot_slow = G.aH(), Rml_av_slow = G.aH();
  std::vector<double> a(nz + 1), b(nz + 1);
  const double h_min = 0.5 * GV->Angstrom_H;
  const double vonKar_x_pi2 = CS->vonKar * 9.8696;
  // find_ustar(forces, tv, U_star_2d, G, GV, US, halo=1, H_T_units=.true.)
  for (int j = js - 1; j <= je + 1; ++j) for (int i = is - 1; i <= ie + 1; ++i) U_star_2d(i, j) = GV->Z_to_H * ustar(i, j);
  if (CS->MLE_density_diff > 0.) {  // detect_mld :1503-1569 (sigma-0; no Stanley variance)
    std::vector<double> rhoSurf(G.ied + 2), deltaRhoAtKm1(G.ied + 2), deltaRhoAtK(G.ied + 2), dK(G.ied + 2), dKm1(G.ied + 2);
    for (int j = js - 1; j <= je + 1; ++j) {
      for (int i = is - 1; i <= ie + 1; ++i) {
        dK[i] = 0.5 * h(i, j, 1);
        rhoSurf[i] = density(CS, T(i, j, 1), S(i, j, 1), 0.);
        deltaRhoAtK[i] = 0.;
        MLD_fast(i, j) = 0.;
      }
      for (int k = 2; k <= nz; ++k) {
        for (int i = is - 1; i <= ie + 1; ++i) {
          dKm1[i] = dK[i];
          dK[i] = dK[i] + 0.5 * (h(i, j, k) + h(i, j, k - 1));
          deltaRhoAtKm1[i] = deltaRhoAtK[i];
          deltaRhoAtK[i] = density(CS, T(i, j, k), S(i, j, k), 0.);
        }
        for (int i = is - 1; i <= ie + 1; ++i) deltaRhoAtK[i] = deltaRhoAtK[i] - rhoSurf[i];
        for (int i = is - 1; i <= ie + 1; ++i) {
          const double ddRho = deltaRhoAtK[i] - deltaRhoAtKm1[i];
          if ((MLD_fast(i, j) == 0.) && (ddRho > 0.) && (deltaRhoAtKm1[i] < CS->MLE_density_diff) && (deltaRhoAtK[i] >= CS->MLE_density_diff)) {
            const double aFac = (CS->MLE_density_diff - deltaRhoAtKm1[i]) / ddRho;
            MLD_fast(i, j) = dK[i] * aFac + dKm1[i] * (1. - aFac);
          }
        }
      }
      for (int i = is - 1; i <= ie + 1; ++i) {
        MLD_fast(i, j) = CS->MLE_MLD_stretch * MLD_fast(i, j);
        if ((MLD_fast(i, j) == 0.) && (deltaRhoAtK[i] < CS->MLE_density_diff)) MLD_fast(i, j) = dK[i];
      }
    }
  } else {
    for (int j = js - 1; j <= je + 1; ++j) for (int i = is - 1; i <= ie + 1; ++i) MLD_fast(i, j) = CS->MLE_MLD_stretch * h_MLD(i, j);
  }
  if (CS->MLE_MLD_decay_time > 0.) {  // :316-328
    const double aFac = CS->MLE_MLD_decay_time / (dt + CS->MLE_MLD_decay_time);
    const double bFac = dt / (dt + CS->MLE_MLD_decay_time);
    for (int j = js - 1; j <= je + 1; ++j) for (int i = is - 1; i <= ie + 1; ++i) {
      MLD_filtered(i, j) = fmax2(MLD_fast(i, j), bFac * MLD_fast(i, j) + aFac * MLD_filtered(i, j));
      MLD_fast(i, j) = MLD_filtered(i, j);
    }
  }
  if (CS->MLE_MLD_decay_time2 > 0.) {  // :331-346
    const double aFac = CS->MLE_MLD_decay_time2 / (dt + CS->MLE_MLD_decay_time2);
    const double bFac = dt / (dt + CS->MLE_MLD_decay_time2);
    for (int j = js - 1; j <= je + 1; ++j) for (int i = is - 1; i <= ie + 1; ++i) {
      MLD_filtered_slow(i, j) = fmax2(MLD_fast(i, j), bFac * MLD_fast(i, j) + aFac * MLD_filtered_slow(i, j));
      MLD_slow(i, j) = MLD_filtered_slow(i, j);
    }
  } else {
    for (int j = js - 1; j <= je + 1; ++j) for (int i = is - 1; i <= ie + 1; ++i) MLD_slow(i, j) = MLD_fast(i, j);
  }
  const double I4dt = 0.25 / dt;
  const double g_Rho0 = GV->H_to_Z * GV->g_Earth / GV->Rho0;
  const double h_neglect = GV->H_subroundoff;
  bool res_upscale;
  if (CS->front_length > 0.) {
    res_upscale = true;
    for (int j = js - 1; j <= je + 1; ++j) for (int i = is - 1; i <= ie + 1; ++i) mle_fl_2d(i, j) = CS->front_length;
  } else res_upscale = false;

  std::vector<double> rho_ml(G.ied + 2), Rml_int_fast(G.ied + 2), Rml_int_slow(G.ied + 2);
  for (int j = js - 1; j <= je + 1; ++j) {  // :375-412
    for (int i = is - 1; i <= ie + 1; ++i) { htot_fast(i, j) = 0.0; Rml_int_fast[i] = 0.0; htot_slow(i, j) = 0.0; Rml_int_slow[i] = 0.0; }
    bool keep_going = true;
    for (int k = 1; k <= nz; ++k) {
      for (int i = is - 1; i <= ie + 1; ++i) h_avail(i, j, k) = fmax2(I4dt * G.areaT(i, j) * (h(i, j, k) - GV->Angstrom_H), 0.0);
      if (keep_going) {
        for (int i = is - 1; i <= ie + 1; ++i) rho_ml[i] = density(CS, T(i, j, k), S(i, j, k), 0.0);
        bool line_is_empty = true;
        for (int i = is - 1; i <= ie + 1; ++i) {
          if (htot_fast(i, j) < MLD_fast(i, j)) {
            const double dh = fmin2(h(i, j, k), MLD_fast(i, j) - htot_fast(i, j));
            Rml_int_fast[i] = Rml_int_fast[i] + dh * rho_ml[i];
            htot_fast(i, j) = htot_fast(i, j) + dh;
            line_is_empty = false;
          }
          if (htot_slow(i, j) < MLD_slow(i, j)) {
            const double dh = fmin2(h(i, j, k), MLD_slow(i, j) - htot_slow(i, j));
            Rml_int_slow[i] = Rml_int_slow[i] + dh * rho_ml[i];
            htot_slow(i, j) = htot_slow(i, j) + dh;
            line_is_empty = false;
          }
        }
        if (line_is_empty) keep_going = false;
      }
    }
    for (int i = is - 1; i <= ie + 1; ++i) {
      Rml_av_fast(i, j) = -(g_Rho0 * Rml_int_fast[i]) / (htot_fast(i, j) + h_neglect);
      Rml_av_slow(i, j) = -(g_Rho0 * Rml_int_slow[i]) / (htot_slow(i, j) + h_neglect);
    }
  }

  const double tail = CS->MLE_tail_dh;
  // ---- U-points :464-540
  for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
    const int i = I;
    const double u_star = fmax2(CS->ustar_min, 0.5 * (U_star_2d(i, j) + U_star_2d(i + 1, j)));
    const double absf = 0.5 * (std::fabs(G.CoriolisBu(I, j - 1)) + std::fabs(G.CoriolisBu(I, j)));
    const double lfront = 0.5 * (mle_fl_2d(i, j) + mle_fl_2d(i + 1, j));
    double I_LFront = 0.0; if (lfront != 0.0) I_LFront = 1.0 / lfront;
    double res_scaling_fac = 0.0;
    if (res_upscale) res_scaling_fac = (std::sqrt(0.5 * ((G.dxCu(I, j) * G.dxCu(I, j)) + (G.dyCu(I, j) * G.dyCu(I, j)))) * I_LFront) *
                                       fmin2(1., 0.5 * (Rd_dx_h(i, j) + Rd_dx_h(i + 1, j)));
    double h_vel = 0.5 * ((htot_fast(i, j) + htot_fast(i + 1, j)) + h_neglect);
    double mom_mixrate = vonKar_x_pi2 * (u_star * u_star) / (absf * (h_vel * h_vel) + 4.0 * (h_vel + h_neglect) * u_star);
    double timescale = 0.0625 * (absf + 2.0 * mom_mixrate) / ((absf * absf) + (mom_mixrate * mom_mixrate));
    timescale = timescale * CS->ml_restrat_coef;
    if (res_upscale) timescale = timescale * res_scaling_fac;
    double uDml = timescale * G.mask2dCu(I, j) * G.dyCu(I, j) * G.IdxCu(I, j) * (Rml_av_fast(i + 1, j) - Rml_av_fast(i, j)) * (h_vel * h_vel);
    h_vel = 0.5 * ((htot_slow(i, j) + htot_slow(i + 1, j)) + h_neglect);
    mom_mixrate = vonKar_x_pi2 * (u_star * u_star) / (absf * (h_vel * h_vel) + 4.0 * (h_vel + h_neglect) * u_star);
    timescale = 0.0625 * (absf + 2.0 * mom_mixrate) / ((absf * absf) + (mom_mixrate * mom_mixrate));
    timescale = timescale * CS->ml_restrat_coef2;
    if (res_upscale) timescale = timescale * res_scaling_fac;
    double uDml_slow = timescale * G.mask2dCu(I, j) * G.dyCu(I, j) * G.IdxCu(I, j) * (Rml_av_slow(i + 1, j) - Rml_av_slow(i, j)) * (h_vel * h_vel);
    if (uDml + uDml_slow == 0.) {
      for (int k = 1; k <= nz; ++k) uhml(I, j, k) = 0.0;
    } else {
      const double IhTot = 2.0 / ((htot_fast(i, j) + htot_fast(i + 1, j)) + h_neglect);
      const double IhTot_slow = 2.0 / ((htot_slow(i, j) + htot_slow(i + 1, j)) + h_neglect);
      double zpa = 0.0, zpb = 0.0;
      for (int k = 1; k <= nz; ++k) {
        const double hAtVel = 0.5 * (h(i, j, k) + h(i + 1, j, k));
        a[k] = oracle_mle_mu(zpa, tail);
        zpa = zpa - (hAtVel * IhTot);
        a[k] = a[k] - oracle_mle_mu(zpa, tail);
        if (a[k] * uDml > 0.0) { if (a[k] * uDml > h_avail(i, j, k)) uDml = h_avail(i, j, k) / a[k]; }
        else if (a[k] * uDml < 0.0) { if (-a[k] * uDml > h_avail(i + 1, j, k)) uDml = -h_avail(i + 1, j, k) / a[k]; }
      }
      for (int k = 1; k <= nz; ++k) {
        const double hAtVel = 0.5 * (h(i, j, k) + h(i + 1, j, k));
        b[k] = oracle_mle_mu(zpb, tail);
        zpb = zpb - (hAtVel * IhTot_slow);
        b[k] = b[k] - oracle_mle_mu(zpb, tail);
        if (b[k] * uDml_slow > 0.0) {
          if (b[k] * uDml_slow > h_avail(i, j, k) - a[k] * uDml) uDml_slow = fmax2(0., h_avail(i, j, k) - a[k] * uDml) / b[k];
        } else if (b[k] * uDml_slow < 0.0) {
          if (-b[k] * uDml_slow > h_avail(i + 1, j, k) + a[k] * uDml) uDml_slow = -fmax2(0., h_avail(i + 1, j, k) + a[k] * uDml) / b[k];
        }
      }
      for (int k = 1; k <= nz; ++k) {
        uhml(I, j, k) = a[k] * uDml + b[k] * uDml_slow;
        uhtr(I, j, k) = uhtr(I, j, k) + uhml(I, j, k) * dt;
      }
    }
  }
  // ---- V-points :543-621
  for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
    const int j = J;
    const double u_star = fmax2(CS->ustar_min, 0.5 * (U_star_2d(i, j) + U_star_2d(i, j + 1)));
    const double lfront = 0.5 * (mle_fl_2d(i, j) + mle_fl_2d(i, j + 1));
    double I_LFront = 0.0; if (lfront != 0.0) I_LFront = 1.0 / lfront;
    const double absf = 0.5 * (std::fabs(G.CoriolisBu(i - 1, J)) + std::fabs(G.CoriolisBu(i, J)));
    double res_scaling_fac = 0.0;
    if (res_upscale) res_scaling_fac = (std::sqrt(0.5 * ((G.dxCv(i, J) * G.dxCv(i, J)) + (G.dyCv(i, J) * G.dyCv(i, J)))) * I_LFront) *
                                       fmin2(1., 0.5 * (Rd_dx_h(i, j) + Rd_dx_h(i, j + 1)));
    double h_vel = 0.5 * ((htot_fast(i, j) + htot_fast(i, j + 1)) + h_neglect);
    double mom_mixrate = vonKar_x_pi2 * (u_star * u_star) / (absf * (h_vel * h_vel) + 4.0 * (h_vel + h_neglect) * u_star);
    double timescale = 0.0625 * (absf + 2.0 * mom_mixrate) / ((absf * absf) + (mom_mixrate * mom_mixrate));
    timescale = timescale * CS->ml_restrat_coef;
    if (res_upscale) timescale = timescale * res_scaling_fac;
    double vDml = timescale * G.mask2dCv(i, J) * G.dxCv(i, J) * G.IdyCv(i, J) * (Rml_av_fast(i, j + 1) - Rml_av_fast(i, j)) * (h_vel * h_vel);
    h_vel = 0.5 * ((htot_slow(i, j) + htot_slow(i, j + 1)) + h_neglect);
    mom_mixrate = vonKar_x_pi2 * (u_star * u_star) / (absf * (h_vel * h_vel) + 4.0 * (h_vel + h_neglect) * u_star);
    timescale = 0.0625 * (absf + 2.0 * mom_mixrate) / ((absf * absf) + (mom_mixrate * mom_mixrate));
    timescale = timescale * CS->ml_restrat_coef2;
    if (res_upscale) timescale = timescale * res_scaling_fac;
    double vDml_slow = timescale * G.mask2dCv(i, J) * G.dxCv(i, J) * G.IdyCv(i, J) * (Rml_av_slow(i, j + 1) - Rml_av_slow(i, j)) * (h_vel * h_vel);
    if (vDml + vDml_slow == 0.) {
      for (int k = 1; k <= nz; ++k) vhml(i, J, k) = 0.0;
    } else {
      const double IhTot = 2.0 / ((htot_fast(i, j) + htot_fast(i, j + 1)) + h_neglect);
      const double IhTot_slow = 2.0 / ((htot_slow(i, j) + htot_slow(i, j + 1)) + h_neglect);
      double zpa = 0.0, zpb = 0.0;
      for (int k = 1; k <= nz; ++k) {
        const double hAtVel = 0.5 * (h(i, j, k) + h(i, j + 1, k));
        a[k] = oracle_mle_mu(zpa, tail);
        zpa = zpa - (hAtVel * IhTot);
        a[k] = a[k] - oracle_mle_mu(zpa, tail);
        if (a[k] * vDml > 0.0) { if (a[k] * vDml > h_avail(i, j, k)) vDml = h_avail(i, j, k) / a[k]; }
        else if (a[k] * vDml < 0.0) { if (-a[k] * vDml > h_avail(i, j + 1, k)) vDml = -h_avail(i, j + 1, k) / a[k]; }
      }
      for (int k = 1; k <= nz; ++k) {
        const double hAtVel = 0.5 * (h(i, j, k) + h(i, j + 1, k));
        b[k] = oracle_mle_mu(zpb, tail);
        zpb = zpb - (hAtVel * IhTot_slow);
        b[k] = b[k] - oracle_mle_mu(zpb, tail);
        if (b[k] * vDml_slow > 0.0) {
          if (b[k] * vDml_slow > h_avail(i, j, k) - a[k] * vDml) vDml_slow = fmax2(0., h_avail(i, j, k) - a[k] * vDml) / b[k];
        } else if (b[k] * vDml_slow < 0.0) {
          if (-b[k] * vDml_slow > h_avail(i, j + 1, k) + a[k] * vDml) vDml_slow = -fmax2(0., h_avail(i, j + 1, k) + a[k] * vDml) / b[k];
        }
      }
      for (int k = 1; k <= nz; ++k) {
        vhml(i, J, k) = a[k] * vDml + b[k] * vDml_slow;
        vhtr(i, J, k) = vhtr(i, J, k) + vhml(i, J, k) * dt;
      }
    }
  }
  // ---- :623-627
  for (int j = js; j <= je; ++j) for (int k = 1; k <= nz; ++k) for (int i = is; i <= ie; ++i) {
    h(i, j, k) = h(i, j, k) - dt * G.IareaT(i, j) * ((uhml(i, j, k) - uhml(i - 1, j, k)) + (vhml(i, j, k) - vhml(i, j - 1, k)));
    if (h(i, j, k) < h_min) h(i, j, k) = h_min;
  }
  return 0;
}

// The density_elem restatement this file uses, for the known-answer test of the equation of state (tests/test_oracle_eos_kat.py).
extern "C" double oracle_mle_eos_density(int form, const double* lin4, double T, double S, double p) {
  mom6cu_mle_cs E = {};
  E.EOS_form = form;
  if (lin4) { E.Rho_T0_S0 = lin4[0]; E.dRho_dT = lin4[1]; E.dRho_dS = lin4[2]; E.dRho_dp = lin4[3]; }
  return density(&E, T, S, p);
}
