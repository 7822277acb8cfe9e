// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of btstep, /root/reference/src/core/MOM_barotropic.F90:455-2172, with its helpers
// btstep_find_Cor :2836-2914, btstep_ubt_from_layer :3388-3428, btstep_layer_accel :3432-3504,
// set_local_BT_cont_types :4876-5003, and of btcalc :4360-4605 and bt_mass_source :5243-5296.
// The substep loop itself is oracle_btstep_timeloop (bt_timeloop.cpp).
// Frozen options: USE_BT_CONT_TYPE=True, LINEARIZED_BT_CORIOLIS=True, no OBCs, no SAL/tides,
// no dynamic_psurf, no linear wave drag / filters, INTEGRAL_BT_CONTINUITY=False, NONLIN_BT_STRESS=False,
// ADJUST_BT_CONT=False, no gradual ICs, no eta_PF_start, answer_date >= 20190101, single tile (halo updates are
// the cyclic wraps of mpp_update_domains on one PE).
#include "oracle.h"
#include "ogrid.hpp"
#include <algorithm>
#include <cmath>
#include <vector>
#include <omp.h>

using namespace orc;

namespace {
enum { FA_EE = 0, FA_E0, FA_W0, FA_WW, UBT_WW, UBT_EE, CRV_W, CRV_E, UH_WW, UH_EE };
inline double find_uhbt(double u, const double* BTC) {  // :4610-4631
  if (u == 0.0) return 0.0;
  if (u < BTC[UBT_EE]) return (u - BTC[UBT_EE]) * BTC[FA_EE] + BTC[UH_EE];
  if (u < 0.0) return u * (BTC[FA_E0] + BTC[CRV_E] * (u * u));
  if (u <= BTC[UBT_WW]) return u * (BTC[FA_W0] + BTC[CRV_W] * (u * u));
  return (u - BTC[UBT_WW]) * BTC[FA_WW] + BTC[UH_WW];
}
}  // namespace

extern "C" int oracle_btstep(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV,
                             const mom6cu_barotropic_cs* CS, const mom6cu_btstep_args* A, int nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  if (CS->unsupported || CS->adjust_BT_cont || !A->BT_cont) return 3;  // ADJUST_BT_CONT (:1180-1198) is not restated
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  const int nz = G.ke;
  const int isdw = d->isdw, iedw = d->iedw, jsdw = d->jsdw, jedw = d->jedw;
  auto WH = [&](const double* p) { return V2((double*)p, isdw, iedw, jsdw, jedw); };
  auto WU = [&](const double* p) { return V2((double*)p, isdw - 1, iedw, jsdw, jedw); };
  auto WV = [&](const double* p) { return V2((double*)p, isdw, iedw, jsdw - 1, jedw); };
  auto WQ = [&](const double* p) { return V2((double*)p, isdw - 1, iedw, jsdw - 1, jedw); };
  const double dt = A->dt;
  const double subroundoff = 1e-30;  // :445
  const double Idt = 1.0 / dt;
  const V3 U_in = G.U3(A->U_in), V_in = G.V3_(A->V_in), bc_accel_u = G.U3(A->bc_accel_u), bc_accel_v = G.V3_(A->bc_accel_v);
  const V3 pbce = G.H3(A->pbce), U_Cor = G.U3(A->U_Cor), V_Cor = G.V3_(A->V_Cor);
  const V3 visc_rem_u = G.U3(A->visc_rem_u), visc_rem_v = G.V3_(A->visc_rem_v);
  const V3 accel_layer_u = G.U3(A->accel_layer_u), accel_layer_v = G.V3_(A->accel_layer_v);
  const V3 frhatu = G.U3(CS->frhatu), frhatv = G.V3_(CS->frhatv);
  const V2 eta_in = G.H(A->eta_in), eta_PF_in = G.H(A->eta_PF_in), eta_out = G.H(A->eta_out);
  const V2 taux = G.U(A->taux), tauy = G.V(A->tauy), uhbtav = G.U(A->uhbtav), vhbtav = G.V(A->vhbtav);
  const V2 eta_cor = G.H(CS->eta_cor), IDatu = G.U(CS->IDatu), IDatv = G.V(CS->IDatv);
  const V2 CSubtav = G.U(CS->ubtav), CSvbtav = G.V(CS->vbtav);
  const V2 CS_IareaT = WH(CS->IareaT), CS_bathyT = WH(CS->bathyT), CS_IdxCu = WU(CS->IdxCu), CS_IdyCv = WV(CS->IdyCv);
  const V2 q_D = WQ(CS->q_D), D_u_Cor = WU(CS->D_u_Cor), D_v_Cor = WV(CS->D_v_Cor);
  const V2 ua_polarity = WH(CS->ua_polarity), va_polarity = WH(CS->va_polarity);
  const V2 OBCmask_u = WU(CS->OBCmask_u), OBCmask_v = WV(CS->OBCmask_v);
  const bool find_etaav = A->etaav != nullptr;
  const bool add_uh0 = A->uh0 != nullptr;
  if (add_uh0 && !(A->vh0 && A->u_uh0 && A->v_vh0)) return 2;  // :776-778 FATAL

  // :765-802
  const int stencil = std::max(1, CS->min_stencil);
  int num_cycles = 1;
  if (CS->use_wide_halos) num_cycles = std::min((is - isdw) / stencil, (js - jsdw) / stencil);
  const int isvf = is - (num_cycles - 1) * stencil, ievf = ie + (num_cycles - 1) * stencil;
  const int jsvf = js - (num_cycles - 1) * stencil, jevf = je + (num_cycles - 1) * stencil;
  const int nstep = (int)std::ceil(dt / CS->dtbt - 0.0001);
  const double Instep = 1.0 / (double)nstep;
  const double dtbt = dt * Instep;

  // wide work arrays (:569-673)
  A2 q(isdw - 1, iedw, jsdw - 1, jedw), DCor_u(isdw - 1, iedw, jsdw, jedw), DCor_v(isdw, iedw, jsdw - 1, jedw);
  A2 gtot_E(isdw, iedw, jsdw, jedw), gtot_W(isdw, iedw, jsdw, jedw), gtot_N(isdw, iedw, jsdw, jedw), gtot_S(isdw, iedw, jsdw, jedw);
  A2 eta(isdw, iedw, jsdw, jedw), eta_PF(isdw, iedw, jsdw, jedw), eta_src(isdw, iedw, jsdw, jedw);
  A2 eta_sum(isdw, iedw, jsdw, jedw), eta_wtd(isdw, iedw, jsdw, jedw);
  A2 Cor_ref_u(isdw - 1, iedw, jsdw, jedw), BT_force_u(isdw - 1, iedw, jsdw, jedw), ubt(isdw - 1, iedw, jsdw, jedw);
  A2 Datu(isdw - 1, iedw, jsdw, jedw), bt_rem_u(isdw - 1, iedw, jsdw, jedw), uhbt0(isdw - 1, iedw, jsdw, jedw);
  A2 uhbt(isdw - 1, iedw, jsdw, jedw), u_accel_bt(isdw - 1, iedw, jsdw, jedw);
  A2 Cor_ref_v(isdw, iedw, jsdw - 1, jedw), BT_force_v(isdw, iedw, jsdw - 1, jedw), vbt(isdw, iedw, jsdw - 1, jedw);
  A2 Datv(isdw, iedw, jsdw - 1, jedw), bt_rem_v(isdw, iedw, jsdw - 1, jedw), vhbt0(isdw, iedw, jsdw - 1, jedw);
  A2 vhbt(isdw, iedw, jsdw - 1, jedw), v_accel_bt(isdw, iedw, jsdw - 1, jedw);
  std::vector<double> f4u_s((size_t)4 * DCor_u.size(), 0.0), f4v_s((size_t)4 * DCor_v.size(), 0.0);
  VM2 f_4_u(f4u_s.data(), 4, isdw - 1, iedw, jsdw, jedw), f_4_v(f4v_s.data(), 4, isdw, iedw, jsdw - 1, jedw);
  std::vector<double> bu_s((size_t)10 * DCor_u.size(), 0.0), bv_s((size_t)10 * DCor_v.size(), 0.0);
  VM2 BTCL_u(bu_s.data(), 10, isdw - 1, iedw, jsdw, jedw), BTCL_v(bv_s.data(), 10, isdw, iedw, jsdw - 1, jedw);
  // G-sized work arrays
  A2 ubt_Cor = G.aU(), vbt_Cor = G.aV(), av_rem_u = G.aU(), av_rem_v = G.aV(), ubt_wtd = G.aU(), vbt_wtd = G.aV();
  A2 Iwt_u_tot = G.aU(), Iwt_v_tot = G.aV(), e_anom = G.aH();
  A3 wt_u(G.isd - 1, G.ied, G.jsd, G.jed, nz), wt_v(G.isd, G.ied, G.jsd - 1, G.jed, nz);
  auto halo_wide = [&](const V2& f, int st) { oracle_fill_halo_2d(d, f.p, st, 1); };
  auto halo_G = [&](const V2& f, int st) { oracle_fill_halo_2d(d, f.p, st, 0); };

  // LINEARIZED_BT_CORIOLIS :868-881
  for (int J = jsvf - 2; J <= jevf + 1; ++J) for (int I = isvf - 2; I <= ievf + 1; ++I) q(I, J) = q_D(I, J);
  for (int j = jsvf - 1; j <= jevf + 1; ++j) for (int I = isvf - 2; I <= ievf + 1; ++I) DCor_u(I, j) = D_u_Cor(I, j);
  for (int J = jsvf - 2; J <= jevf + 1; ++J) for (int i = isvf - 1; i <= ievf + 1; ++i) DCor_v(i, J) = D_v_Cor(i, J);
  // :938-965 zeroing is done by construction; :997-1003 copy the input arrays
  for (int j = G.jsd; j <= G.jed; ++j) for (int i = G.isd; i <= G.ied; ++i) { eta(i, j) = eta_in(i, j); eta_PF(i, j) = eta_PF_in(i, j); }

  // :1011-1034 weights
  for (int k = 1; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
    double visc_rem = fmin2(visc_rem_u(I, j, k), 1.);
    visc_rem = fmax2(visc_rem, 1. - 0.5 * Instep / (visc_rem + subroundoff));
    visc_rem = fmax2(visc_rem, 0.);
    wt_u(I, j, k) = frhatu(I, j, k) * visc_rem;
  }
  for (int k = 1; k <= nz; ++k) for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
    double visc_rem = fmin2(visc_rem_v(i, J, k), 1.);
    visc_rem = fmax2(visc_rem, 1. - 0.5 * Instep / (visc_rem + subroundoff));
    visc_rem = fmax2(visc_rem, 0.);
    wt_v(i, J, k) = frhatv(i, J, k) * visc_rem;
  }
  if (!CS->wt_uv_bug) {  // :1036-1058
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) Iwt_u_tot(I, j) = wt_u(I, j, 1);
    for (int k = 2; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I)
      Iwt_u_tot(I, j) = Iwt_u_tot(I, j) + wt_u(I, j, k);
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I)
      if (std::fabs(Iwt_u_tot(I, j)) > 0.0) Iwt_u_tot(I, j) = G.mask2dCu(I, j) / Iwt_u_tot(I, j);
    for (int k = 1; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I)
      wt_u(I, j, k) = wt_u(I, j, k) * Iwt_u_tot(I, j);
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) Iwt_v_tot(i, J) = wt_v(i, J, 1);
    for (int k = 2; k <= nz; ++k) for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i)
      Iwt_v_tot(i, J) = Iwt_v_tot(i, J) + wt_v(i, J, k);
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i)
      if (std::fabs(Iwt_v_tot(i, J)) > 0.0) Iwt_v_tot(i, J) = G.mask2dCv(i, J) / Iwt_v_tot(i, J);
    for (int k = 1; k <= nz; ++k) for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i)
      wt_v(i, J, k) = wt_v(i, J, k) * Iwt_v_tot(i, J);
  }
  // :1062-1073 reference velocities of the Coriolis terms
  for (int j = js; j <= je; ++j) for (int k = 1; k <= nz; ++k) for (int I = is - 1; I <= ie; ++I)
    ubt_Cor(I, j) = ubt_Cor(I, j) + wt_u(I, j, k) * U_Cor(I, j, k);
  for (int J = js - 1; J <= je; ++J) for (int k = 1; k <= nz; ++k) for (int i = is; i <= ie; ++i)
    vbt_Cor(i, J) = vbt_Cor(i, J) + wt_v(i, J, k) * V_Cor(i, J, k);
  // :1079-1090 gtot
  for (int j = js; j <= je; ++j) for (int k = 1; k <= nz; ++k) for (int I = is - 1; I <= ie; ++I) {
    gtot_E(I, j) = gtot_E(I, j) + pbce(I, j, k) * wt_u(I, j, k);
    gtot_W(I + 1, j) = gtot_W(I + 1, j) + pbce(I + 1, j, k) * wt_u(I, j, k);
  }
  for (int J = js - 1; J <= je; ++J) for (int k = 1; k <= nz; ++k) for (int i = is; i <= ie; ++i) {
    gtot_N(i, J) = gtot_N(i, J) + pbce(i, J, k) * wt_v(i, J, k);
    gtot_S(i, J + 1) = gtot_S(i, J + 1) + pbce(i, J + 1, k) * wt_v(i, J, k);
  }
  const double dgeo_de = 1.0 + CS->G_extra;  // :1117

  // set_local_BT_cont_types :4876-5003 with halo = 1+ievf-ie (:1132)
  {
    const mom6cu_bt_cont* B = A->BT_cont;
    const int hs = std::max(1 + ievf - ie, 0);
    A2 uBT_EE(isdw - 1, iedw, jsdw, jedw), uBT_WW(isdw - 1, iedw, jsdw, jedw), FA_u_EE(isdw - 1, iedw, jsdw, jedw),
        FA_u_E0(isdw - 1, iedw, jsdw, jedw), FA_u_W0(isdw - 1, iedw, jsdw, jedw), FA_u_WW(isdw - 1, iedw, jsdw, jedw);
    A2 vBT_NN(isdw, iedw, jsdw - 1, jedw), vBT_SS(isdw, iedw, jsdw - 1, jedw), FA_v_NN(isdw, iedw, jsdw - 1, jedw),
        FA_v_N0(isdw, iedw, jsdw - 1, jedw), FA_v_S0(isdw, iedw, jsdw - 1, jedw), FA_v_SS(isdw, iedw, jsdw - 1, jedw);
    const V2 b_uEE = G.U(B->uBT_EE), b_uWW = G.U(B->uBT_WW), b_FEE = G.U(B->FA_u_EE), b_FE0 = G.U(B->FA_u_E0),
             b_FW0 = G.U(B->FA_u_W0), b_FWW = G.U(B->FA_u_WW);
    const V2 b_vNN = G.V(B->vBT_NN), b_vSS = G.V(B->vBT_SS), b_FNN = G.V(B->FA_v_NN), b_FN0 = G.V(B->FA_v_N0),
             b_FS0 = G.V(B->FA_v_S0), b_FSS = G.V(B->FA_v_SS);
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
      uBT_EE(I, j) = b_uEE(I, j); uBT_WW(I, j) = b_uWW(I, j);
      FA_u_EE(I, j) = b_FEE(I, j); FA_u_E0(I, j) = b_FE0(I, j); FA_u_W0(I, j) = b_FW0(I, j); FA_u_WW(I, j) = b_FWW(I, j);
    }
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
      vBT_NN(i, J) = b_vNN(i, J); vBT_SS(i, J) = b_vSS(i, J);
      FA_v_NN(i, J) = b_FNN(i, J); FA_v_N0(i, J) = b_FN0(i, J); FA_v_S0(i, J) = b_FS0(i, J); FA_v_SS(i, J) = b_FSS(i, J);
    }
    for (const V2* f : {(const V2*)&uBT_EE, (const V2*)&uBT_WW, (const V2*)&FA_u_EE, (const V2*)&FA_u_E0, (const V2*)&FA_u_W0, (const V2*)&FA_u_WW}) halo_wide(*f, 1);
    for (const V2* f : {(const V2*)&vBT_NN, (const V2*)&vBT_SS, (const V2*)&FA_v_NN, (const V2*)&FA_v_N0, (const V2*)&FA_v_S0, (const V2*)&FA_v_SS}) halo_wide(*f, 2);
    const double C1_3 = 1.0 / 3.0;
    for (int j = js - hs; j <= je + hs; ++j) for (int I = is - hs - 1; I <= ie + hs; ++I) {
      double* b = BTCL_u.at(I, j);
      b[FA_EE] = FA_u_EE(I, j); b[FA_E0] = FA_u_E0(I, j); b[FA_W0] = FA_u_W0(I, j); b[FA_WW] = FA_u_WW(I, j);
      b[UBT_EE] = 1.0 * uBT_EE(I, j); b[UBT_WW] = 1.0 * uBT_WW(I, j);
      b[UH_EE] = b[UBT_EE] * (C1_3 * (2.0 * b[FA_E0] + b[FA_EE]));
      b[UH_WW] = b[UBT_WW] * (C1_3 * (2.0 * b[FA_W0] + b[FA_WW]));
      b[CRV_E] = 0.0; b[CRV_W] = 0.0;
      if (std::fabs(b[UBT_WW]) > 0.0) b[CRV_W] = (C1_3 * (b[FA_WW] - b[FA_W0])) / (b[UBT_WW] * b[UBT_WW]);
      if (std::fabs(b[UBT_EE]) > 0.0) b[CRV_E] = (C1_3 * (b[FA_EE] - b[FA_E0])) / (b[UBT_EE] * b[UBT_EE]);
    }
    for (int J = js - hs - 1; J <= je + hs; ++J) for (int i = is - hs; i <= ie + hs; ++i) {
      double* b = BTCL_v.at(i, J);
      b[FA_EE] = FA_v_NN(i, J); b[FA_E0] = FA_v_N0(i, J); b[FA_W0] = FA_v_S0(i, J); b[FA_WW] = FA_v_SS(i, J);
      b[UBT_EE] = 1.0 * vBT_NN(i, J); b[UBT_WW] = 1.0 * vBT_SS(i, J);
      b[UH_EE] = b[UBT_EE] * (C1_3 * (2.0 * b[FA_E0] + b[FA_EE]));
      b[UH_WW] = b[UBT_WW] * (C1_3 * (2.0 * b[FA_W0] + b[FA_WW]));
      b[CRV_E] = 0.0; b[CRV_W] = 0.0;
      if (std::fabs(b[UBT_WW]) > 0.0) b[CRV_W] = (C1_3 * (b[FA_WW] - b[FA_W0])) / (b[UBT_WW] * b[UBT_WW]);
      if (std::fabs(b[UBT_EE]) > 0.0) b[CRV_E] = (C1_3 * (b[FA_EE] - b[FA_E0])) / (b[UBT_EE] * b[UBT_EE]);
    }
  }

  // :1155-1238 reference transports
  if (add_uh0) {
    const V3 uh0 = G.U3(A->uh0), vh0 = G.V3_(A->vh0), u_uh0 = G.U3(A->u_uh0), v_vh0 = G.V3_(A->v_vh0);
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) { uhbt(I, j) = 0.0; ubt(I, j) = 0.0; }
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) { vhbt(i, J) = 0.0; vbt(i, J) = 0.0; }
    if (CS->visc_rem_u_uh0) {
      for (int j = js; j <= je; ++j) for (int k = 1; k <= nz; ++k) for (int I = is - 1; I <= ie; ++I) {
        uhbt(I, j) = uhbt(I, j) + uh0(I, j, k);
        ubt(I, j) = ubt(I, j) + wt_u(I, j, k) * u_uh0(I, j, k);
      }
      for (int J = js - 1; J <= je; ++J) for (int k = 1; k <= nz; ++k) for (int i = is; i <= ie; ++i) {
        vhbt(i, J) = vhbt(i, J) + vh0(i, J, k);
        vbt(i, J) = vbt(i, J) + wt_v(i, J, k) * v_vh0(i, J, k);
      }
    } else {
      for (int j = js; j <= je; ++j) for (int k = 1; k <= nz; ++k) for (int I = is - 1; I <= ie; ++I) {
        uhbt(I, j) = uhbt(I, j) + uh0(I, j, k);
        ubt(I, j) = ubt(I, j) + frhatu(I, j, k) * u_uh0(I, j, k);
      }
      for (int J = js - 1; J <= je; ++J) for (int k = 1; k <= nz; ++k) for (int i = is; i <= ie; ++i) {
        vhbt(i, J) = vhbt(i, J) + vh0(i, J, k);
        vbt(i, J) = vbt(i, J) + frhatv(i, J, k) * v_vh0(i, J, k);
      }
    }
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) uhbt0(I, j) = uhbt(I, j) - find_uhbt(ubt(I, j), BTCL_u.at(I, j));
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) vhbt0(i, J) = vhbt(i, J) - find_uhbt(vbt(i, J), BTCL_v.at(i, J));
  }
  // btstep_ubt_from_layer :3388-3428
  ubt.fill(0.0); vbt.fill(0.0);
  for (int j = js; j <= je; ++j) for (int k = 1; k <= nz; ++k) for (int I = is - 1; I <= ie; ++I)
    ubt(I, j) = ubt(I, j) + wt_u(I, j, k) * U_in(I, j, k);
  for (int J = js - 1; J <= je; ++J) for (int k = 1; k <= nz; ++k) for (int i = is; i <= ie; ++i)
    vbt(i, J) = vbt(i, J) + wt_v(i, J, k) * V_in(i, J, k);
  for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) if (std::fabs(ubt(I, j)) < CS->vel_underflow) ubt(I, j) = 0.0;
  for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) if (std::fabs(vbt(i, J)) < CS->vel_underflow) vbt(i, J) = 0.0;
  uhbt.fill(0.0); vhbt.fill(0.0); u_accel_bt.fill(0.0); v_accel_bt.fill(0.0);

  // :1258-1330 vertical average forcing
  for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
    if (G.mask2dCu(I, j) > 0.0) BT_force_u(I, j) = taux(I, j) * GV->RZ_to_H * IDatu(I, j) * visc_rem_u(I, j, 1);
    else BT_force_u(I, j) = 0.0;
  }
  for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
    if (G.mask2dCv(i, J) > 0.0) BT_force_v(i, J) = tauy(i, J) * GV->RZ_to_H * IDatv(i, J) * visc_rem_v(i, J, 1);
    else BT_force_v(i, J) = 0.0;
  }
  if (A->taux_bot && A->tauy_bot) {
    const V2 taux_bot = G.U(A->taux_bot), tauy_bot = G.V(A->tauy_bot);
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) if (G.mask2dCu(I, j) > 0.0)
      BT_force_u(I, j) = BT_force_u(I, j) - taux_bot(I, j) * GV->RZ_to_H * IDatu(I, j);
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) if (G.mask2dCv(i, J) > 0.0)
      BT_force_v(i, J) = BT_force_v(i, J) - tauy_bot(i, J) * GV->RZ_to_H * IDatv(i, J);
  }
  for (int j = js; j <= je; ++j) for (int k = 1; k <= nz; ++k) for (int I = Isq; I <= Ieq; ++I)
    BT_force_u(I, j) = BT_force_u(I, j) + wt_u(I, j, k) * bc_accel_u(I, j, k);
  for (int J = Jsq; J <= Jeq; ++J) for (int k = 1; k <= nz; ++k) for (int i = is; i <= ie; ++i)
    BT_force_v(i, J) = BT_force_v(i, J) + wt_v(i, J, k) * bc_accel_v(i, J, k);

  // btstep_find_Cor :2866-2895
  if (CS->Sadourny) {
    for (int J = jsvf - 1; J <= jevf; ++J) for (int i = isvf - 1; i <= ievf + 1; ++i) {
      f_4_v(1, i, J) = OBCmask_v(i, J) * DCor_u(i - 1, J) * q(i - 1, J);
      f_4_v(2, i, J) = OBCmask_v(i, J) * DCor_u(i, J) * q(i, J);
      f_4_v(4, i, J) = OBCmask_v(i, J) * DCor_u(i, J + 1) * q(i, J);
      f_4_v(3, i, J) = OBCmask_v(i, J) * DCor_u(i - 1, J + 1) * q(i - 1, J);
    }
    for (int j = jsvf - 1; j <= jevf + 1; ++j) for (int I = isvf - 1; I <= ievf; ++I) {
      f_4_u(4, I, j) = OBCmask_u(I, j) * DCor_v(I + 1, j) * q(I, j);
      f_4_u(3, I, j) = OBCmask_u(I, j) * DCor_v(I, j) * q(I, j);
      f_4_u(1, I, j) = OBCmask_u(I, j) * DCor_v(I, j - 1) * q(I, j - 1);
      f_4_u(2, I, j) = OBCmask_u(I, j) * DCor_v(I + 1, j - 1) * q(I, j - 1);
    }
  } else {
    for (int J = jsvf - 1; J <= jevf; ++J) for (int i = isvf - 1; i <= ievf + 1; ++i) {
      f_4_v(1, i, J) = OBCmask_v(i, J) * DCor_u(i - 1, J) * ((q(i, J) + q(i - 1, J - 1)) + q(i - 1, J)) / 3.0;
      f_4_v(2, i, J) = OBCmask_v(i, J) * DCor_u(i, J) * (q(i, J) + (q(i - 1, J) + q(i, J - 1))) / 3.0;
      f_4_v(4, i, J) = OBCmask_v(i, J) * DCor_u(i, J + 1) * (q(i, J) + (q(i - 1, J) + q(i, J + 1))) / 3.0;
      f_4_v(3, i, J) = OBCmask_v(i, J) * DCor_u(i - 1, J + 1) * ((q(i, J) + q(i - 1, J + 1)) + q(i - 1, J)) / 3.0;
    }
    for (int j = jsvf - 1; j <= jevf + 1; ++j) for (int I = isvf - 1; I <= ievf; ++I) {
      f_4_u(4, I, j) = OBCmask_u(I, j) * DCor_v(I + 1, j) * (q(I, j) + (q(I + 1, j) + q(I, j - 1))) / 3.0;
      f_4_u(3, I, j) = OBCmask_u(I, j) * DCor_v(I, j) * (q(I, j) + (q(I - 1, j) + q(I, j - 1))) / 3.0;
      f_4_u(1, I, j) = OBCmask_u(I, j) * DCor_v(I, j - 1) * ((q(I, j) + q(I - 1, j - 1)) + q(I, j - 1)) / 3.0;
      f_4_u(2, I, j) = OBCmask_u(I, j) * DCor_v(I + 1, j - 1) * ((q(I, j) + q(I + 1, j - 1)) + q(I, j - 1)) / 3.0;
    }
  }
  // :1436-1441 halo updates, :1444-1447 polarity
  halo_wide(gtot_E, 0); halo_wide(gtot_N, 0); halo_wide(gtot_W, 0); halo_wide(gtot_S, 0);
  halo_G(ubt_Cor, 1); halo_G(vbt_Cor, 2);
  for (int j = jsvf - 1; j <= jevf + 1; ++j) for (int i = isvf - 1; i <= ievf + 1; ++i) {
    if (ua_polarity(i, j) < 0.0) std::swap(gtot_E(i, j), gtot_W(i, j));
    if (va_polarity(i, j) < 0.0) std::swap(gtot_N(i, j), gtot_S(i, j));
  }
  for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I)
    Cor_ref_u(I, j) = (((f_4_u(4, I, j) * vbt_Cor(I + 1, j)) + (f_4_u(1, I, j) * vbt_Cor(I, j - 1))) +
                       ((f_4_u(3, I, j) * vbt_Cor(I, j)) + (f_4_u(2, I, j) * vbt_Cor(I + 1, j - 1))));
  for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i)
    Cor_ref_v(i, J) = -1.0 * (((f_4_v(1, i, J) * ubt_Cor(i - 1, J)) + (f_4_v(4, i, J) * ubt_Cor(i, J + 1))) +
                              ((f_4_v(2, i, J) * ubt_Cor(i, J)) + (f_4_v(3, i, J) * ubt_Cor(i - 1, J + 1))));
  // :1476-1509 viscous remnant of the barotropic velocities
  for (int j = js; j <= je; ++j) for (int k = 1; k <= nz; ++k) for (int I = is - 1; I <= ie; ++I)
    av_rem_u(I, j) = av_rem_u(I, j) + frhatu(I, j, k) * visc_rem_u(I, j, k);
  for (int J = js - 1; J <= je; ++J) for (int k = 1; k <= nz; ++k) for (int i = is; i <= ie; ++i)
    av_rem_v(i, J) = av_rem_v(i, J) + frhatv(i, J, k) * visc_rem_v(i, J, k);
  if (CS->strong_drag) {
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I)
      bt_rem_u(I, j) = G.mask2dCu(I, j) * ((nstep * av_rem_u(I, j)) / (1.0 + (nstep - 1) * av_rem_u(I, j)));
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i)
      bt_rem_v(i, J) = G.mask2dCv(i, J) * ((nstep * av_rem_v(i, J)) / (1.0 + (nstep - 1) * av_rem_v(i, J)));
  } else {
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
      bt_rem_u(I, j) = 0.0;
      if (G.mask2dCu(I, j) * av_rem_u(I, j) > 0.0) bt_rem_u(I, j) = G.mask2dCu(I, j) * std::pow(av_rem_u(I, j), Instep);
    }
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
      bt_rem_v(i, J) = 0.0;
      if (G.mask2dCv(i, J) * av_rem_v(i, J) > 0.0) bt_rem_v(i, J) = G.mask2dCv(i, J) * std::pow(av_rem_v(i, J), Instep);
    }
  }
  // :1549-1587 mass source
  if (CS->bound_BT_corr) {
    if (CS->BT_cont_bounds) {
      for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) if (G.mask2dT(i, j) > 0.0) {
        if (eta_cor(i, j) > 0.0) {
          const double u_max_cor = G.dxT(i, j) * (CS->maxCFL_BT_cont * Idt);
          const double v_max_cor = G.dyT(i, j) * (CS->maxCFL_BT_cont * Idt);
          const double eta_cor_max = dt * (CS_IareaT(i, j) *
              (((find_uhbt(u_max_cor, BTCL_u.at(i, j)) + uhbt0(i, j)) - (find_uhbt(-u_max_cor, BTCL_u.at(i - 1, j)) + uhbt0(i - 1, j))) +
               ((find_uhbt(v_max_cor, BTCL_v.at(i, j)) + vhbt0(i, j)) - (find_uhbt(-v_max_cor, BTCL_v.at(i, j - 1)) + vhbt0(i, j - 1)))));
          eta_cor(i, j) = fmin2(eta_cor(i, j), fmax2(0.0, eta_cor_max));
        } else {
          double Htot = eta(i, j);
          if (GV->Boussinesq) Htot = CS_bathyT(i, j) * GV->Z_to_H + eta(i, j);
          eta_cor(i, j) = fmax2(eta_cor(i, j), -fmax2(0.0, Htot));
        }
      }
    } else {
      const V2 eta_cor_bound = G.H(CS->eta_cor_bound);
      for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i)
        if (std::fabs(eta_cor(i, j)) > dt * eta_cor_bound(i, j)) eta_cor(i, j) = fsign(dt * eta_cor_bound(i, j), eta_cor(i, j));
    }
  }
  for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) eta_src(i, j) = G.mask2dT(i, j) * (Instep * eta_cor(i, j));
  // :1627-1643 halo updates on the wide domain
  halo_wide(eta_PF, 0); halo_wide(eta_src, 0); halo_wide(bt_rem_u, 1); halo_wide(bt_rem_v, 2);
  halo_wide(BT_force_u, 1); halo_wide(BT_force_v, 2);
  if (add_uh0) { halo_wide(uhbt0, 1); halo_wide(vhbt0, 2); }
  halo_wide(Cor_ref_u, 1); halo_wide(Cor_ref_v, 2);

  // :1727-1795 filter weights
  double dt_filt;
  if (CS->dt_bt_filter >= 0.0) dt_filt = 0.5 * fmax2(0.0, fmin2(CS->dt_bt_filter, 2.0 * dt));
  else dt_filt = 0.5 * fmax2(0.0, dt * fmin2(-CS->dt_bt_filter, 2.0));
  const int nfilter = (int)std::ceil(dt_filt / dtbt);
  if (nstep + nfilter == 0) return 2;
  const int nt = nstep + nfilter;
  std::vector<double> wt_vel(nt), wt_eta(nt), wt_trans(nt + 1), wt_accel(nt + 1), wt_accel2(nt + 1);
  double sum_wt_vel = 0.0, sum_wt_eta = 0.0, sum_wt_accel = 0.0, sum_wt_trans = 0.0;
  for (int n = 1; n <= nt; ++n) {
    if ((n == nstep) || (dt_filt - std::abs(n - nstep) * dtbt >= 0.0)) { wt_vel[n - 1] = 1.0; wt_eta[n - 1] = 1.0; }
    else if (dtbt + dt_filt - std::abs(n - nstep) * dtbt > 0.0) { wt_vel[n - 1] = 1.0 + (dt_filt / dtbt) - std::abs(n - nstep); wt_eta[n - 1] = wt_vel[n - 1]; }
    else { wt_vel[n - 1] = 0.0; wt_eta[n - 1] = 0.0; }
    sum_wt_vel = sum_wt_vel + wt_vel[n - 1]; sum_wt_eta = sum_wt_eta + wt_eta[n - 1];
  }
  wt_trans[nt] = 0.0; wt_accel[nt] = 0.0;
  for (int n = nt; n >= 1; --n) {
    wt_trans[n - 1] = wt_trans[n] + wt_eta[n - 1];
    wt_accel[n - 1] = wt_accel[n] + wt_vel[n - 1];
    sum_wt_accel = sum_wt_accel + wt_accel[n - 1]; sum_wt_trans = sum_wt_trans + wt_trans[n - 1];
  }
  const double I_sum_wt_vel = 1.0 / sum_wt_vel, I_sum_wt_accel = 1.0 / sum_wt_accel;
  const double I_sum_wt_eta = 1.0 / sum_wt_eta, I_sum_wt_trans = 1.0 / sum_wt_trans;
  for (int n = 1; n <= nt; ++n) {
    wt_vel[n - 1] = wt_vel[n - 1] * I_sum_wt_vel;
    wt_accel2[n - 1] = wt_accel[n - 1] * I_sum_wt_accel;
    wt_trans[n - 1] = wt_trans[n - 1] * I_sum_wt_trans;
    wt_accel[n - 1] = wt_accel[n - 1] * I_sum_wt_accel;
    wt_eta[n - 1] = wt_eta[n - 1] * I_sum_wt_eta;
  }

  // March the barotropic solver through all of its time steps (:1803-1812)
  {
    mom6cu_bt_timeloop_args T = {};
    T.eta = eta.p; T.ubt = ubt.p; T.vbt = vbt.p; T.uhbt0 = uhbt0.p; T.vhbt0 = vhbt0.p; T.Datu = Datu.p; T.Datv = Datv.p;
    T.BTCL_u = bu_s.data(); T.BTCL_v = bv_s.data(); T.eta_src = eta_src.p; T.eta_PF = eta_PF.p;
    T.gtot_E = gtot_E.p; T.gtot_W = gtot_W.p; T.gtot_N = gtot_N.p; T.gtot_S = gtot_S.p;
    T.f_4_u = f4u_s.data(); T.f_4_v = f4v_s.data(); T.bt_rem_u = bt_rem_u.p; T.bt_rem_v = bt_rem_v.p;
    T.BT_force_u = BT_force_u.p; T.BT_force_v = BT_force_v.p; T.Cor_ref_u = Cor_ref_u.p; T.Cor_ref_v = Cor_ref_v.p;
    T.IareaT_OBCmask = CS->IareaT_OBCmask; T.IdxCu = CS->IdxCu; T.IdyCv = CS->IdyCv;
    T.u_accel_bt = u_accel_bt.p; T.v_accel_bt = v_accel_bt.p; T.eta_sum = eta_sum.p; T.eta_wtd = eta_wtd.p;
    T.ubtav = CS->ubtav; T.vbtav = CS->vbtav; T.uhbtav = A->uhbtav; T.vhbtav = A->vhbtav; T.ubt_wtd = ubt_wtd.p; T.vbt_wtd = vbt_wtd.p;
    T.wt_vel = wt_vel.data(); T.wt_eta = wt_eta.data(); T.wt_accel = wt_accel.data(); T.wt_trans = wt_trans.data(); T.wt_accel2 = wt_accel2.data();
    T.dtbt = dtbt; T.dgeo_de = dgeo_de; T.bebt = CS->bebt; T.vel_underflow = CS->vel_underflow;
    T.nstep = nstep; T.nfilter = nfilter; T.use_BT_cont = 1; T.find_etaav = find_etaav ? 1 : 0;
    T.BT_project_velocity = CS->BT_project_velocity; T.use_old_coriolis_bracket_bug = CS->use_old_coriolis_bracket_bug;
    T.use_wide_halos = CS->use_wide_halos; T.min_stencil = CS->min_stencil;
    int rc = oracle_btstep_timeloop(d, &T, nullptr, nullptr, nthreads);
    if (rc) return rc;
  }
  // :1814-1847
  if (find_etaav) { const V2 etaav = G.H(A->etaav); for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) etaav(i, j) = eta_sum(i, j) * 1.0; }
  for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) e_anom(i, j) = dgeo_de * (0.5 * (eta(i, j) + eta_in(i, j)) - eta_PF(i, j));
  for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) eta_out(i, j) = eta_wtd(i, j) * 1.0;
  if (find_etaav) halo_G(G.H(A->etaav), 0);
  halo_G(e_anom, 0);
  halo_G(CSubtav, 1); halo_G(CSvbtav, 2); halo_G(uhbtav, 1); halo_G(vhbtav, 2);
  // btstep_layer_accel :3432-3504
  const double accel_underflow = CS->vel_underflow * Idt;
  for (int k = 1; k <= nz; ++k) {
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
      accel_layer_u(I, j, k) = (u_accel_bt(I, j) - (((pbce(I + 1, j, k) - gtot_W(I + 1, j)) * e_anom(I + 1, j)) -
                                                    ((pbce(I, j, k) - gtot_E(I, j)) * e_anom(I, j))) * CS_IdxCu(I, j));
      if (std::fabs(accel_layer_u(I, j, k)) < accel_underflow) accel_layer_u(I, j, k) = 0.0;
    }
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
      accel_layer_v(i, J, k) = (v_accel_bt(i, J) - (((pbce(i, J + 1, k) - gtot_S(i, J + 1)) * e_anom(i, J + 1)) -
                                                    ((pbce(i, J, k) - gtot_N(i, J)) * e_anom(i, J))) * CS_IdyCv(i, J));
      if (std::fabs(accel_layer_v(i, J, k)) < accel_underflow) accel_layer_v(i, J, k) = 0.0;
    }
  }
  return 0;
}

// btcalc :4360-4605 (no OBCs)
extern "C" int oracle_btcalc(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV,
                             const mom6cu_btcalc_args* A, int nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, nz = G.ke;
  const bool have_huv = A->h_u && A->h_v;
  const int HARMONIC = 1, ARITHMETIC = 2, HYBRID = 3;
  const int hs = A->hvel_scheme;
  bool use_default = false;
  if (!(have_huv || hs == HARMONIC || hs == HYBRID || hs == ARITHMETIC)) {
    if (A->may_use_default) use_default = true;
    else return 2;  // FATAL: Inconsistent settings of optional arguments and hvel_scheme.
  }
  const V3 h = G.H3(A->h), frhatu = G.U3(A->frhatu), frhatv = G.V3_(A->frhatv);
  V3 h_u, h_v;
  if (have_huv) { h_u = G.U3(A->h_u); h_v = G.V3_(A->h_v); }
  const V2 bathyT = G.H(A->bathyT);
  const double h_neglect = GV->H_subroundoff, Z_to_H = GV->Z_to_H;
#pragma omp parallel for
  for (int j = js; j <= je; ++j) {
    std::vector<double> hatu_s((size_t)(ie - is + 2) * nz), e_s((size_t)(ie - is + 2) * (nz + 1)), tot(ie - is + 2, 0.0), Dsh(ie - is + 2);
    auto hatu = [&](int I, int k) -> double& { return hatu_s[(size_t)(k - 1) * (ie - is + 2) + (I - (is - 1))]; };
    auto e_u = [&](int I, int K) -> double& { return e_s[(size_t)(K - 1) * (ie - is + 2) + (I - (is - 1))]; };
    auto hatutot = [&](int I) -> double& { return tot[I - (is - 1)]; };
    if (have_huv) {
      for (int k = 1; k <= nz; ++k) for (int I = is - 1; I <= ie; ++I) { hatu(I, k) = h_u(I, j, k); hatutot(I) = hatutot(I) + hatu(I, k); }
    } else if (hs == ARITHMETIC) {
      for (int k = 1; k <= nz; ++k) for (int I = is - 1; I <= ie; ++I) { hatu(I, k) = 0.5 * (h(I + 1, j, k) + h(I, j, k)); hatutot(I) = hatutot(I) + hatu(I, k); }
    } else if (hs == HYBRID || use_default) {
      for (int I = is - 1; I <= ie; ++I) {
        e_u(I, nz + 1) = -0.5 * Z_to_H * (bathyT(I + 1, j) + bathyT(I, j));
        Dsh[I - (is - 1)] = -Z_to_H * fmin2(bathyT(I + 1, j), bathyT(I, j));
      }
      for (int k = nz; k >= 1; --k) for (int I = is - 1; I <= ie; ++I) {
        const double D_shallow_u = Dsh[I - (is - 1)];
        e_u(I, k) = e_u(I, k + 1) + 0.5 * (h(I + 1, j, k) + h(I, j, k));
        const double h_arith = 0.5 * (h(I + 1, j, k) + h(I, j, k));
        if (e_u(I, k + 1) >= D_shallow_u) hatu(I, k) = h_arith;
        else {
          const double h_harm = (h(I + 1, j, k) * h(I, j, k)) / (h_arith + h_neglect);
          if (e_u(I, k) <= D_shallow_u) hatu(I, k) = h_harm;
          else {
            const double wt_arith = (e_u(I, k) - D_shallow_u) / (h_arith + h_neglect);
            hatu(I, k) = wt_arith * h_arith + (1.0 - wt_arith) * h_harm;
          }
        }
        hatutot(I) = hatutot(I) + hatu(I, k);
      }
    } else if (hs == HARMONIC) {
      for (int k = 1; k <= nz; ++k) for (int I = is - 1; I <= ie; ++I) {
        hatu(I, k) = 2.0 * (h(I + 1, j, k) * h(I, j, k)) / ((h(I + 1, j, k) + h(I, j, k)) + h_neglect);
        hatutot(I) = hatutot(I) + hatu(I, k);
      }
    }
    for (int I = is - 1; I <= ie; ++I) {
      const double Ihatutot = G.mask2dCu(I, j) / (hatutot(I) + h_neglect);
      for (int k = 1; k <= nz; ++k) frhatu(I, j, k) = hatu(I, k) * Ihatutot;
    }
  }
#pragma omp parallel for
  for (int J = js - 1; J <= je; ++J) {
    std::vector<double> hatv_s((size_t)(ie - is + 1) * nz), e_s((size_t)(ie - is + 1) * (nz + 1)), tot(ie - is + 1, 0.0), Dsh(ie - is + 1);
    auto hatv = [&](int i, int k) -> double& { return hatv_s[(size_t)(k - 1) * (ie - is + 1) + (i - is)]; };
    auto e_v = [&](int i, int K) -> double& { return e_s[(size_t)(K - 1) * (ie - is + 1) + (i - is)]; };
    auto hatvtot = [&](int i) -> double& { return tot[i - is]; };
    if (have_huv) {
      for (int k = 1; k <= nz; ++k) for (int i = is; i <= ie; ++i) { hatv(i, k) = h_v(i, J, k); hatvtot(i) = hatvtot(i) + hatv(i, k); }
    } else if (hs == ARITHMETIC) {
      for (int k = 1; k <= nz; ++k) for (int i = is; i <= ie; ++i) { hatv(i, k) = 0.5 * (h(i, J + 1, k) + h(i, J, k)); hatvtot(i) = hatvtot(i) + hatv(i, k); }
    } else if (hs == HYBRID || use_default) {
      for (int i = is; i <= ie; ++i) {
        e_v(i, nz + 1) = -0.5 * Z_to_H * (bathyT(i, J + 1) + bathyT(i, J));
        Dsh[i - is] = -Z_to_H * fmin2(bathyT(i, J + 1), bathyT(i, J));
      }
      for (int k = nz; k >= 1; --k) for (int i = is; i <= ie; ++i) {
        const double D_shallow_v = Dsh[i - is];
        e_v(i, k) = e_v(i, k + 1) + 0.5 * (h(i, J + 1, k) + h(i, J, k));
        const double h_arith = 0.5 * (h(i, J + 1, k) + h(i, J, k));
        if (e_v(i, k + 1) >= D_shallow_v) hatv(i, k) = h_arith;
        else {
          const double h_harm = (h(i, J + 1, k) * h(i, J, k)) / (h_arith + h_neglect);
          if (e_v(i, k) <= D_shallow_v) hatv(i, k) = h_harm;
          else {
            const double wt_arith = (e_v(i, k) - D_shallow_v) / (h_arith + h_neglect);
            hatv(i, k) = wt_arith * h_arith + (1.0 - wt_arith) * h_harm;
          }
        }
        hatvtot(i) = hatvtot(i) + hatv(i, k);
      }
    } else if (hs == HARMONIC) {
      for (int k = 1; k <= nz; ++k) for (int i = is; i <= ie; ++i) {
        hatv(i, k) = 2.0 * (h(i, J + 1, k) * h(i, J, k)) / ((h(i, J + 1, k) + h(i, J, k)) + h_neglect);
        hatvtot(i) = hatvtot(i) + hatv(i, k);
      }
    }
    for (int i = is; i <= ie; ++i) {
      const double Ihatvtot = G.mask2dCv(i, J) / (hatvtot(i) + h_neglect);
      for (int k = 1; k <= nz; ++k) frhatv(i, J, k) = hatv(i, k) * Ihatvtot;
    }
  }
  return 0;
}

// bt_mass_source :5243-5296
extern "C" int oracle_bt_mass_source(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV, const double* hp,
                                     const double* etap, int set_cor, double* eta_corp) {
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, nz = G.ke;
  const V3 h = G.H3(hp);
  const V2 eta = G.H(etap), eta_cor = G.H(eta_corp);
  for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
    double eta_h = GV->Boussinesq ? h(i, j, 1) - G.bathyT(i, j) * GV->Z_to_H : h(i, j, 1);
    for (int k = 2; k <= nz; ++k) eta_h = eta_h + h(i, j, k);
    const double d_eta = eta_h - eta(i, j);
    if (set_cor) eta_cor(i, j) = d_eta; else eta_cor(i, j) = eta_cor(i, j) + d_eta;
  }
  return 0;
}

// set_dtbt :3509-3633 with BT_cont_to_face_areas :5107-5134 / find_face_areas :5146-5237 (halo = 0, Boussinesq, no SAL)
extern "C" int oracle_set_dtbt(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV, const mom6cu_unit_scale* US,
                               const mom6cu_set_dtbt_args* a, double* dtbt, double* dtbt_max_out) {
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, nz = G.ke;
  if (!(a->pbce || a->have_gtot_est)) return 2;  // FATAL :3565
  if (!GV->Boussinesq) return 3;
  const double add_SSH = a->SSH_add;
  A2 Datu = G.aU(), Datv = G.aV(), gtot_E = G.aH(), gtot_W = G.aH(), gtot_N = G.aH(), gtot_S = G.aH();
  const V2 bathyT = G.H(a->bathyT);
  if (a->BT_cont) {
    const mom6cu_bt_cont* B = a->BT_cont;
    const V2 EE = G.U(B->FA_u_EE), E0 = G.U(B->FA_u_E0), W0 = G.U(B->FA_u_W0), WW = G.U(B->FA_u_WW);
    const V2 NN = G.V(B->FA_v_NN), N0 = G.V(B->FA_v_N0), S0 = G.V(B->FA_v_S0), SS = G.V(B->FA_v_SS);
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) Datu(I, j) = max4(EE(I, j), E0(I, j), W0(I, j), WW(I, j));
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) Datv(i, J) = max4(NN(i, J), N0(i, J), S0(i, J), SS(i, J));
  } else if (a->Nonlinear_continuity && a->eta) {
    const V2 eta = G.H(a->eta);
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
      const double H1 = bathyT(I, j) * GV->Z_to_H + eta(I, j), H2 = bathyT(I + 1, j) * GV->Z_to_H + eta(I + 1, j);
      Datu(I, j) = 0.0; if ((H1 > 0.0) && (H2 > 0.0)) Datu(I, j) = G.dy_Cu(I, j) * (2.0 * H1 * H2) / (H1 + H2);
    }
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
      const double H1 = bathyT(i, J) * GV->Z_to_H + eta(i, J), H2 = bathyT(i, J + 1) * GV->Z_to_H + eta(i, J + 1);
      Datv(i, J) = 0.0; if ((H1 > 0.0) && (H2 > 0.0)) Datv(i, J) = G.dx_Cv(i, J) * (2.0 * H1 * H2) / (H1 + H2);
    }
  } else {
    const double Z_to_H = GV->Z_to_H;
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I)
      Datu(I, j) = G.dy_Cu(I, j) * Z_to_H * fmax2(fmax2(bathyT(I + 1, j), bathyT(I, j)) + (a->Z_ref + add_SSH), 0.0);
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i)
      Datv(i, J) = G.dx_Cv(i, J) * Z_to_H * fmax2(fmax2(bathyT(i, J + 1), bathyT(i, J)) + (a->Z_ref + add_SSH), 0.0);
  }
  const double det_de = 0.0;
  const double dgeo_de = 1.0 + fmax2(0.0, a->G_extra - det_de);
  if (a->pbce) {
    const V3 pbce = G.H3(a->pbce), frhatu = G.U3(a->frhatu), frhatv = G.V3_(a->frhatv);
    for (int k = 1; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      gtot_E(i, j) = gtot_E(i, j) + pbce(i, j, k) * frhatu(i, j, k);
      gtot_W(i, j) = gtot_W(i, j) + pbce(i, j, k) * frhatu(i - 1, j, k);
      gtot_N(i, j) = gtot_N(i, j) + pbce(i, j, k) * frhatv(i, j, k);
      gtot_S(i, j) = gtot_S(i, j) + pbce(i, j, k) * frhatv(i, j - 1, k);
    }
  } else {
    for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      gtot_E(i, j) = a->gtot_est; gtot_W(i, j) = a->gtot_est; gtot_N(i, j) = a->gtot_est; gtot_S(i, j) = a->gtot_est;
    }
  }
  double min_max_dt2 = 1.0e38 * (US->s_to_T * US->s_to_T);
  for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
    const double Idt_max2 = 0.5 * (1.0 + 2.0 * a->bebt) * (G.IareaT(i, j) *
        (((gtot_E(i, j) * Datu(i, j) * G.IdxCu(i, j)) + (gtot_W(i, j) * Datu(i - 1, j) * G.IdxCu(i - 1, j))) +
         ((gtot_N(i, j) * Datv(i, j) * G.IdyCv(i, j)) + (gtot_S(i, j) * Datv(i, j - 1) * G.IdyCv(i, j - 1)))) +
        ((G.Coriolis2Bu(i, j) + G.Coriolis2Bu(i - 1, j - 1)) + (G.Coriolis2Bu(i - 1, j) + G.Coriolis2Bu(i, j - 1))) *
            (a->BT_Coriolis_scale * a->BT_Coriolis_scale));
    if (Idt_max2 * min_max_dt2 > 1.0) min_max_dt2 = 1.0 / Idt_max2;
  }
  const double dtbt_max = std::sqrt(min_max_dt2 / dgeo_de);
  *dtbt = a->dtbt_fraction * dtbt_max;
  *dtbt_max_out = dtbt_max;
  return 0;
}
