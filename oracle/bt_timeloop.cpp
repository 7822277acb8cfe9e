// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of btstep_timeloop and its helpers,
// /root/reference/src/core/MOM_barotropic.F90:2175-2832, :2956-3149, :3209-3384, :4610-4631.
// Frozen options: no OBCs, no dynamic_psurf, no linear_wave_drag,
// no clip_velocity, INTEGRAL_BT_CONTINUITY=False, no evolving face areas, no diagnostics.
#include "oracle.h"
#include "farray.hpp"
#include <cstdlib>
#include <algorithm>
#include <omp.h>

using namespace orc;

namespace {

// local_BT_cont_u_type field order, MOM_barotropic.F90:367-390 (v: :393-416)
enum { FA_EE = 0, FA_E0, FA_W0, FA_WW, UBT_WW, UBT_EE, CRV_W, CRV_E, UH_WW, UH_EE };

// find_uhbt, MOM_barotropic.F90:4610-4631 (find_vhbt :4744-4765 is identical with
// N<->E, S<->W so the same routine serves both).
inline double find_uhbt(double u, const double* BTC) {
  double uhbt;
  if (u == 0.0) {
    uhbt = 0.0;
  } else if (u < BTC[UBT_EE]) {
    uhbt = (u - BTC[UBT_EE]) * BTC[FA_EE] + BTC[UH_EE];
  } else if (u < 0.0) {
    uhbt = u * (BTC[FA_E0] + BTC[CRV_E] * (u * u));
  } else if (u <= BTC[UBT_WW]) {
    uhbt = u * (BTC[FA_W0] + BTC[CRV_W] * (u * u));
  } else {  // (u > BTC%uBT_WW)
    uhbt = (u - BTC[UBT_WW]) * BTC[FA_WW] + BTC[UH_WW];
  }
  return uhbt;
}

struct Ctx {
  int is, ie, js, je, isdw, iedw, jsdw, jedw;
};

}  // namespace

extern "C" void oracle_fill_halo_2d(const mom6cu_domain* d, double* f, int stagger, int wide) {
  const int isd = wide ? d->isdw : d->isd, ied = wide ? d->iedw : d->ied;
  const int jsd = wide ? d->jsdw : d->jsd, jed = wide ? d->jedw : d->jed;
  const int su = (stagger == 1 || stagger == 3) ? 1 : 0;  // staggered in i
  const int sv = (stagger == 2 || stagger == 3) ? 1 : 0;  // staggered in j
  V2 A(f, isd - su, ied, jsd - sv, jed);
  const int ni = d->iec - d->isc + 1, nj = d->jec - d->jsc + 1;
  // Symmetric-memory semantics (config_src/infra/FMS2/MOM_domain_infra.F90:171-216 ->
  // mpp_update_domains): points of the (symmetric) computational domain, including the
  // shared edge I=isc-1 / J=jsc-1 of staggered fields, are never overwritten.
  if (d->cyclic_x) {
    _Pragma("omp parallel for")
    for (int j = jsd - sv; j <= jed; ++j) {
      for (int i = isd - su; i <= d->isc - 1 - su; ++i) A(i, j) = A(i + ni, j);
      for (int i = d->iec + 1; i <= ied; ++i) A(i, j) = A(i - ni, j);
    }
  }
  if (d->cyclic_y) {
    for (int i = isd - su; i <= ied; ++i) {
      for (int j = jsd - sv; j <= d->jsc - 1 - sv; ++j) A(i, j) = A(i, j + nj);
      for (int j = d->jec + 1; j <= jed; ++j) A(i, j) = A(i, j - nj);
    }
  }
}

extern "C" int oracle_btstep_timeloop(const mom6cu_domain* d, const mom6cu_bt_timeloop_args* a,
                                      oracle_halo_fn halo, void* user, int nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  const int is = d->isc, ie = d->iec, js = d->jsc, je = d->jec;
  const int isdw = d->isdw, iedw = d->iedw, jsdw = d->jsdw, jedw = d->jedw;
  const int isd = d->isd, ied = d->ied, jsd = d->jsd, jed = d->jed;
  const bool use_BT_cont = a->use_BT_cont != 0;
  const bool find_etaav = a->find_etaav != 0;
  const bool project = a->BT_project_velocity != 0;
  const double dtbt = a->dtbt, dgeo_de = a->dgeo_de;
  const int nstep = a->nstep, nfilter = a->nfilter;

  // wide arrays
  V2 eta(a->eta, isdw, iedw, jsdw, jedw), ubt(a->ubt, isdw - 1, iedw, jsdw, jedw),
      vbt(a->vbt, isdw, iedw, jsdw - 1, jedw);
  V2 uhbt0((double*)a->uhbt0, isdw - 1, iedw, jsdw, jedw), vhbt0((double*)a->vhbt0, isdw, iedw, jsdw - 1, jedw);
  V2 Datu((double*)a->Datu, isdw - 1, iedw, jsdw, jedw), Datv((double*)a->Datv, isdw, iedw, jsdw - 1, jedw);
  VM2 BTCL_u((double*)a->BTCL_u, 10, isdw - 1, iedw, jsdw, jedw), BTCL_v((double*)a->BTCL_v, 10, isdw, iedw, jsdw - 1, jedw);
  V2 eta_src((double*)a->eta_src, isdw, iedw, jsdw, jedw), eta_PF((double*)a->eta_PF, isdw, iedw, jsdw, jedw);
  V2 gtot_E((double*)a->gtot_E, isdw, iedw, jsdw, jedw), gtot_W((double*)a->gtot_W, isdw, iedw, jsdw, jedw);
  V2 gtot_N((double*)a->gtot_N, isdw, iedw, jsdw, jedw), gtot_S((double*)a->gtot_S, isdw, iedw, jsdw, jedw);
  VM2 f_4_u((double*)a->f_4_u, 4, isdw - 1, iedw, jsdw, jedw), f_4_v((double*)a->f_4_v, 4, isdw, iedw, jsdw - 1, jedw);
  V2 bt_rem_u((double*)a->bt_rem_u, isdw - 1, iedw, jsdw, jedw), bt_rem_v((double*)a->bt_rem_v, isdw, iedw, jsdw - 1, jedw);
  V2 BT_force_u((double*)a->BT_force_u, isdw - 1, iedw, jsdw, jedw), BT_force_v((double*)a->BT_force_v, isdw, iedw, jsdw - 1, jedw);
  V2 Cor_ref_u((double*)a->Cor_ref_u, isdw - 1, iedw, jsdw, jedw), Cor_ref_v((double*)a->Cor_ref_v, isdw, iedw, jsdw - 1, jedw);
  V2 IareaT_OBCmask((double*)a->IareaT_OBCmask, isdw, iedw, jsdw, jedw);
  V2 IdxCu((double*)a->IdxCu, isdw - 1, iedw, jsdw, jedw), IdyCv((double*)a->IdyCv, isdw, iedw, jsdw - 1, jedw);
  V2 u_accel_bt(a->u_accel_bt, isdw - 1, iedw, jsdw, jedw), v_accel_bt(a->v_accel_bt, isdw, iedw, jsdw - 1, jedw);
  V2 eta_sum(a->eta_sum, isdw, iedw, jsdw, jedw), eta_wtd(a->eta_wtd, isdw, iedw, jsdw, jedw);
  // G-sized outputs
  V2 ubtav(a->ubtav, isd - 1, ied, jsd, jed), vbtav(a->vbtav, isd, ied, jsd - 1, jed);
  V2 uhbtav(a->uhbtav, isd - 1, ied, jsd, jed), vhbtav(a->vhbtav, isd, ied, jsd - 1, jed);
  V2 ubt_wtd(a->ubt_wtd, isd - 1, ied, jsd, jed), vbt_wtd(a->vbt_wtd, isd, ied, jsd - 1, jed);

  // Local variables (:2354-2377)
  A2 uhbt(isdw - 1, iedw, jsdw, jedw), ubt_prev(isdw - 1, iedw, jsdw, jedw), ubt_trans(isdw - 1, iedw, jsdw, jedw),
      PFu(isdw - 1, iedw, jsdw, jedw), Cor_u(isdw - 1, iedw, jsdw, jedw);
  A2 vhbt(isdw, iedw, jsdw - 1, jedw), vbt_prev(isdw, iedw, jsdw - 1, jedw), vbt_trans(isdw, iedw, jsdw - 1, jedw),
      PFv(isdw, iedw, jsdw - 1, jedw), Cor_v(isdw, iedw, jsdw - 1, jedw);
  A2 eta_pred(isdw, iedw, jsdw, jedw);

  // :2413-2422  Figure out the fullest arrays that could be updated.
  int stencil = std::max(1, a->min_stencil);
  int num_cycles = 1;
  if (a->use_wide_halos) num_cycles = std::min((is - isdw) / stencil, (js - jsdw) / stencil);
  const int isvf = is - (num_cycles - 1) * stencil, ievf = ie + (num_cycles - 1) * stencil;
  const int jsvf = js - (num_cycles - 1) * stencil, jevf = je + (num_cycles - 1) * stencil;

  // :2431-2436
  double trans_wt1, trans_wt2;
  if (project) {
    const double be_proj = a->bebt;
    trans_wt1 = (1.0 + be_proj); trans_wt2 = -be_proj;
  } else {
    trans_wt1 = a->bebt; trans_wt2 = (1.0 - a->bebt);
  }

  // :2455-2486  Zero out the arrays for various time-averaged quantities.
  if (find_etaav) {
    _Pragma("omp parallel for")
    for (int j = jsvf - 1; j <= jevf + 1; ++j) for (int i = isvf - 1; i <= ievf + 1; ++i) {
      eta_sum(i, j) = 0.0; eta_wtd(i, j) = 0.0;
    }
  } else {
    _Pragma("omp parallel for")
    for (int j = jsvf - 1; j <= jevf + 1; ++j) for (int i = isvf - 1; i <= ievf + 1; ++i) eta_wtd(i, j) = 0.0;
  }
  _Pragma("omp parallel for")
  for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
    ubtav(I, j) = 0.0; uhbtav(I, j) = 0.0; ubt_wtd(I, j) = 0.0;
  }
  _Pragma("omp parallel for")
  for (int j = jsvf - 1; j <= jevf + 1; ++j) for (int I = isvf - 1; I <= ievf; ++I) ubt_trans(I, j) = 0.0;
  _Pragma("omp parallel for")
  for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
    vbtav(i, J) = 0.0; vhbtav(i, J) = 0.0; vbt_wtd(i, J) = 0.0;
  }
  _Pragma("omp parallel for")
  for (int J = jsvf - 1; J <= jevf; ++J) for (int i = isvf - 1; i <= ievf + 1; ++i) vbt_trans(i, J) = 0.0;

  // :2504-2827  The following loop contains all of the time steps.
  int isv = is, iev = ie, jsv = js, jev = je;
  for (int n = 1; n <= nstep + nfilter; ++n) {
    // :2509-2518 Update the range of valid points, by a halo update or by marching inward.
    if ((iev - stencil < ie) || (jev - stencil < je)) {
      if (halo) {
        halo(user, a->eta, a->ubt, a->vbt);
      } else {
        oracle_fill_halo_2d(d, a->eta, 0, 1);
        oracle_fill_halo_2d(d, a->ubt, 1, 1);
        oracle_fill_halo_2d(d, a->vbt, 2, 1);
      }
      isv = isvf; iev = ievf; jsv = jsvf; jev = jevf;
    } else {
      isv = isv + stencil; iev = iev - stencil;
      jsv = jsv + stencil; jev = jev - stencil;
    }

    // :2520-2526 Store the previous velocities for time-filtered transports.
    _Pragma("omp parallel for")
    for (int j = jsv; j <= jev; ++j) for (int I = isv - 2; I <= iev + 1; ++I) ubt_prev(I, j) = ubt(I, j);
    _Pragma("omp parallel for")
    for (int J = jsv - 2; J <= jev + 1; ++J) for (int i = isv; i <= iev; ++i) vbt_prev(i, J) = vbt(i, J);

    // :2545-2550 -> btloop_eta_predictor :3035-3058
    if (!project) {
      if (use_BT_cont) {
        _Pragma("omp parallel for")
        for (int j = jsv - 1; j <= jev + 1; ++j) for (int I = isv - 2; I <= iev + 1; ++I)
          uhbt(I, j) = find_uhbt(ubt(I, j), BTCL_u.at(I, j)) + uhbt0(I, j);
        _Pragma("omp parallel for")
        for (int J = jsv - 2; J <= jev + 1; ++J) for (int i = isv - 1; i <= iev + 1; ++i)
          vhbt(i, J) = find_uhbt(vbt(i, J), BTCL_v.at(i, J)) + vhbt0(i, J);
        _Pragma("omp parallel for")
        for (int j = jsv - 1; j <= jev + 1; ++j) for (int i = isv - 1; i <= iev + 1; ++i)
          eta_pred(i, j) = (eta(i, j) + eta_src(i, j)) + (dtbt * IareaT_OBCmask(i, j)) *
                           ((uhbt(i - 1, j) - uhbt(i, j)) + (vhbt(i, j - 1) - vhbt(i, j)));
      } else {
        _Pragma("omp parallel for")
        for (int j = jsv - 1; j <= jev + 1; ++j) for (int i = isv - 1; i <= iev + 1; ++i)
          eta_pred(i, j) = (eta(i, j) + eta_src(i, j)) + (dtbt * IareaT_OBCmask(i, j)) *
                           (((Datu(i - 1, j) * ubt(i - 1, j) + uhbt0(i - 1, j)) -
                             (Datu(i, j) * ubt(i, j) + uhbt0(i, j))) +
                            ((Datv(i, j - 1) * vbt(i, j - 1) + vhbt0(i, j - 1)) -
                             (Datv(i, j) * vbt(i, j) + vhbt0(i, j))));
      }
    }

    // :2561
    const bool v_first = (((n + d->first_direction) % 2) == 1);

    // :2563-2572 -> btloop_find_PF :3118-3147
    {
      const V2& eta_PF_BT = project ? eta : (const V2&)eta_pred;
      int is_v, ie_v, js_u, je_u;
      if (v_first) { is_v = isv - 1; ie_v = iev + 1; js_u = jsv; je_u = jev; }
      else { is_v = isv; ie_v = iev; js_u = jsv - 1; je_u = jev + 1; }
      _Pragma("omp parallel for")
      for (int j = js_u; j <= je_u; ++j) for (int I = isv - 1; I <= iev; ++I) {
        const int i = I;
        PFu(I, j) = (((eta_PF_BT(i, j) - eta_PF(i, j)) * gtot_E(i, j)) -
                     ((eta_PF_BT(i + 1, j) - eta_PF(i + 1, j)) * gtot_W(i + 1, j))) *
                    dgeo_de * IdxCu(I, j);
      }
      _Pragma("omp parallel for")
      for (int J = jsv - 1; J <= jev; ++J) for (int i = is_v; i <= ie_v; ++i) {
        const int j = J;
        PFv(i, J) = (((eta_PF_BT(i, j) - eta_PF(i, j)) * gtot_N(i, j)) -
                     ((eta_PF_BT(i, j + 1) - eta_PF(i, j + 1)) * gtot_S(i, j + 1))) *
                    dgeo_de * IdyCv(i, J);
      }
      const double wt_accel2_n = a->wt_accel2[n - 1];
      if (find_etaav && (std::fabs(wt_accel2_n) > 0.0)) {
        _Pragma("omp parallel for")
        for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i)
          eta_sum(i, j) = eta_sum(i, j) + wt_accel2_n * eta_PF_BT(i, j);
      }
    }

    const double wt_accel_n = a->wt_accel[n - 1];
    // btloop_update_v :3209-3303
    auto update_v = [&](int is_v, int ie_v, int Js_v, int Je_v, bool use_bracket_bug) {
      if (use_bracket_bug) {
        _Pragma("omp parallel for")
        for (int J = Js_v; J <= Je_v; ++J) for (int i = is_v; i <= ie_v; ++i) {
          const int I = i, j = J;
          Cor_v(i, J) = -1.0 * (((f_4_v(1, i, J) * ubt(I - 1, j)) + (f_4_v(2, i, J) * ubt(I, j))) +
                                ((f_4_v(4, i, J) * ubt(I, j + 1)) + (f_4_v(3, i, J) * ubt(I - 1, j + 1)))) -
                        Cor_ref_v(i, J);
        }
      } else {
        _Pragma("omp parallel for")
        for (int J = Js_v; J <= Je_v; ++J) for (int i = is_v; i <= ie_v; ++i) {
          const int I = i, j = J;
          Cor_v(i, J) = -1.0 * (((f_4_v(1, i, J) * ubt(I - 1, j)) + (f_4_v(4, i, J) * ubt(I, j + 1))) +
                                ((f_4_v(2, i, J) * ubt(I, j)) + (f_4_v(3, i, J) * ubt(I - 1, j + 1)))) -
                        Cor_ref_v(i, J);
        }
      }
      _Pragma("omp parallel for")
      for (int J = Js_v; J <= Je_v; ++J) for (int i = is_v; i <= ie_v; ++i) {
        vbt(i, J) = bt_rem_v(i, J) * (vbt(i, J) + dtbt * ((BT_force_v(i, J) + Cor_v(i, J)) + PFv(i, J)));
        if (std::fabs(vbt(i, J)) < a->vel_underflow) vbt(i, J) = 0.0;
      }
      _Pragma("omp parallel for")
      for (int J = Js_v; J <= Je_v; ++J) for (int i = is_v; i <= ie_v; ++i)
        v_accel_bt(i, J) = v_accel_bt(i, J) + wt_accel_n * (Cor_v(i, J) + PFv(i, J));
    };
    // btloop_update_u :3306-3384
    auto update_u = [&](int Is_u, int Ie_u, int js_u, int je_u) {
      _Pragma("omp parallel for")
      for (int j = js_u; j <= je_u; ++j) for (int I = Is_u; I <= Ie_u; ++I) {
        const int i = I, J = j;
        Cor_u(I, j) = (((f_4_u(4, I, j) * vbt(i + 1, J)) + (f_4_u(1, I, j) * vbt(i, J - 1))) +
                       ((f_4_u(3, I, j) * vbt(i, J)) + (f_4_u(2, I, j) * vbt(i + 1, J - 1)))) -
                      Cor_ref_u(I, j);
        ubt(I, j) = bt_rem_u(I, j) * (ubt(I, j) + dtbt * ((BT_force_u(I, j) + Cor_u(I, j)) + PFu(I, j)));
        if (std::fabs(ubt(I, j)) < a->vel_underflow) ubt(I, j) = 0.0;
      }
      _Pragma("omp parallel for")
      for (int j = js_u; j <= je_u; ++j) for (int I = Is_u; I <= Ie_u; ++I)
        u_accel_bt(I, j) = u_accel_bt(I, j) + wt_accel_n * (Cor_u(I, j) + PFu(I, j));
    };

    // :2580-2600
    if (v_first) {
      update_v(isv - 1, iev + 1, jsv - 1, jev, false);
      update_u(isv - 1, iev, jsv, jev);
    } else {
      update_u(isv - 1, iev, jsv - 1, jev + 1);
      update_v(isv, iev, jsv - 1, jev, a->use_old_coriolis_bracket_bug != 0);
    }

    // :2602-2647 Determine the transports based on the updated velocities.
    if (use_BT_cont) {
      _Pragma("omp parallel for")
      for (int j = jsv; j <= jev; ++j) for (int I = isv - 1; I <= iev; ++I) {
        ubt_trans(I, j) = trans_wt1 * ubt(I, j) + trans_wt2 * ubt_prev(I, j);
        uhbt(I, j) = find_uhbt(ubt_trans(I, j), BTCL_u.at(I, j)) + uhbt0(I, j);
      }
      _Pragma("omp parallel for")
      for (int J = jsv - 1; J <= jev; ++J) for (int i = isv; i <= iev; ++i) {
        vbt_trans(i, J) = trans_wt1 * vbt(i, J) + trans_wt2 * vbt_prev(i, J);
        vhbt(i, J) = find_uhbt(vbt_trans(i, J), BTCL_v.at(i, J)) + vhbt0(i, J);
      }
    } else {
      _Pragma("omp parallel for")
      for (int j = jsv; j <= jev; ++j) for (int I = isv - 1; I <= iev; ++I) {
        ubt_trans(I, j) = trans_wt1 * ubt(I, j) + trans_wt2 * ubt_prev(I, j);
        uhbt(I, j) = Datu(I, j) * ubt_trans(I, j) + uhbt0(I, j);
      }
      _Pragma("omp parallel for")
      for (int J = jsv - 1; J <= jev; ++J) for (int i = isv; i <= iev; ++i) {
        vbt_trans(i, J) = trans_wt1 * vbt(i, J) + trans_wt2 * vbt_prev(i, J);
        vhbt(i, J) = Datv(i, J) * vbt_trans(i, J) + vhbt0(i, J);
      }
    }

    // :2689-2703 Contribute to the running sums of the transports and velocities.
    const double wt_trans_n = a->wt_trans[n - 1], wt_vel_n = a->wt_vel[n - 1];
    _Pragma("omp parallel for")
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
      ubtav(I, j) = ubtav(I, j) + wt_trans_n * ubt_trans(I, j);
      uhbtav(I, j) = uhbtav(I, j) + wt_trans_n * uhbt(I, j);
      ubt_wtd(I, j) = ubt_wtd(I, j) + wt_vel_n * ubt(I, j);
    }
    _Pragma("omp parallel for")
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
      vbtav(i, J) = vbtav(i, J) + wt_trans_n * vbt_trans(i, J);
      vhbtav(i, J) = vhbtav(i, J) + wt_trans_n * vhbt(i, J);
      vbt_wtd(i, J) = vbt_wtd(i, J) + wt_vel_n * vbt(i, J);
    }

    // :2721-2727 Update eta in a corrector step using the barotropic continuity equation.
    const double wt_eta_n = a->wt_eta[n - 1];
    _Pragma("omp parallel for")
    for (int j = jsv; j <= jev; ++j) for (int i = isv; i <= iev; ++i) {
      eta(i, j) = (eta(i, j) + eta_src(i, j)) + (dtbt * IareaT_OBCmask(i, j)) *
                  ((uhbt(i - 1, j) - uhbt(i, j)) + (vhbt(i, j - 1) - vhbt(i, j)));
      eta_wtd(i, j) = eta_wtd(i, j) + eta(i, j) * wt_eta_n;
    }
  }  // end of do n=1,ntimestep
  return 0;
}
