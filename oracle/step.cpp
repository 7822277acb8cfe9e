// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of step_MOM_dyn_split_RK2, /root/reference/src/core/MOM_dynamics_split_RK2.F90:294-1205, for the frozen
// option set of include/mom6cu.h (no OBCs, waves, fpmix, dynamic surface pressure; BT_cont with h_u/h_v; calc_dtbt false;
// set_viscous_ML a no-op; no diagnostics): the order of the stage calls, the elementwise glue (:425-430, :565-573,
// :594-601, :681-691, :791-793, :800-804, :924-932, :961-975, :1021-1023, :1060-1072) and the group passes (single
// tile: the periodic wrap of oracle_fill_halo_2d out to the full halo).  The stages are the oracle's own.
// PARITY: PINNED BY A REFERENCE RUN -- the reference's own step_MOM_dyn_split_RK2 with every stage it calls, executed by oracle/f90run,
// agrees bit for bit on 5 configurations (tests/test_reference_f90.py, step/*).
#include "oracle.h"
#include "ogrid.hpp"
#include <vector>

using namespace orc;

namespace {
void fill3(const mom6cu_domain* d, double* f, int st, int nk) {
  const int su = (st == 1 || st == 3), sv = (st == 2 || st == 3);
  const size_t plane = (size_t)(d->ied - d->isd + 1 + su) * (d->jed - d->jsd + 1 + sv);
  for (int k = 0; k < nk; ++k) oracle_fill_halo_2d(d, f + plane * k, st, 0);
}
}  // namespace

extern "C" int oracle_step_dyn_split_rk2(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV, const mom6cu_unit_scale* US,
                                         const mom6cu_continuity_cs* cont_cs, const mom6cu_coriolisadv_cs* corad_cs,
                                         const mom6cu_hor_visc_cs* hv_cs, const mom6cu_pressureforce_cs* pgf_cs,
                                         const mom6cu_vertvisc_cs* vv_cs, mom6cu_dyn_split_rk2_cs* CS, const mom6cu_step_dyn_args* a,
                                         int nthreads) {
  const OGrid G(d, Gp);
  const int nz = G.ke, is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  if (CS->unsupported || !CS->BT_cont || !CS->BT_cont->h_u || !CS->BT_cont->h_v || !CS->barotropic) return 3;
  const double dt = a->dt;
  int rc;
  A3 up(G.isd - 1, G.ied, G.jsd, G.jed, nz), vp(G.isd, G.ied, G.jsd - 1, G.jed, nz), hp(G.isd, G.ied, G.jsd, G.jed, nz);
  A3 u_bc_accel(G.isd - 1, G.ied, G.jsd, G.jed, nz), v_bc_accel(G.isd, G.ied, G.jsd - 1, G.jed, nz);
  A3 uh_in(G.isd - 1, G.ied, G.jsd, G.jed, nz), vh_in(G.isd, G.ied, G.jsd - 1, G.jed, nz);
  A2 eta_pred = G.aH();
  A3 a_u(G.isd - 1, G.ied, G.jsd, G.jed, nz + 1), a_v(G.isd, G.ied, G.jsd - 1, G.jed, nz + 1), h_u(G.isd - 1, G.ied, G.jsd, G.jed, nz),
      h_v(G.isd, G.ied, G.jsd - 1, G.jed, nz);
  const V3 u = G.U3(a->u_inst), v = G.V3_(a->v_inst), h = G.H3(a->h), uh = G.U3(a->uh), vh = G.V3_(a->vh), uhtr = G.U3(a->uhtr),
           vhtr = G.V3_(a->vhtr);
  const V3 CAu = G.U3(CS->CAu), CAv = G.V3_(CS->CAv), CAu_pred = G.U3(CS->CAu_pred), CAv_pred = G.V3_(CS->CAv_pred), PFu = G.U3(CS->PFu),
           PFv = G.V3_(CS->PFv), diffu = G.U3(CS->diffu), diffv = G.V3_(CS->diffv), u_accel_bt = G.U3(CS->u_accel_bt),
           v_accel_bt = G.V3_(CS->v_accel_bt), u_av = G.U3(CS->u_av), v_av = G.V3_(CS->v_av), h_av = G.H3(CS->h_av);
  const V2 eta = G.H(CS->eta);
  // :425-430
  for (size_t n = 0; n < hp.size(); ++n) hp.p[n] = h.p[n];
  // PressureForce :503
  mom6cu_pressureforce_args pa = {a->h, a->T, a->S, CS->PFu, CS->PFv, a->p_surf, CS->pbce, CS->eta_PF};
  if ((rc = oracle_pressure_force(d, Gp, GV, pgf_cs, &pa, nthreads))) return 100 + rc;
  // :556
  if (!CS->CAu_pred_stored) {
    mom6cu_coradcalc_args ca = {CS->u_av, CS->v_av, CS->h_av, a->uh, a->vh, CS->CAu_pred, CS->CAv_pred, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if ((rc = oracle_coradcalc(d, Gp, GV, US, corad_cs, &ca, nthreads))) return 200 + rc;
  }
  // :565-573, :594-601
  for (int k = 1; k <= nz; ++k) {
    for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) u_bc_accel(I, j, k) = (CAu_pred(I, j, k) + PFu(I, j, k)) + diffu(I, j, k);
    for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) v_bc_accel(i, J, k) = (CAv_pred(i, J, k) + PFv(i, J, k)) + diffv(i, J, k);
  }
  for (int k = 1; k <= nz; ++k) {
    for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) up(I, j, k) = G.mask2dCu(I, j) * (u(I, j, k) + dt * u_bc_accel(I, j, k));
    for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) vp(i, J, k) = G.mask2dCv(i, J) * (v(i, J, k) + dt * v_bc_accel(i, J, k));
  }
  // vertvisc_coef :609, vertvisc_remnant :610
  mom6cu_vertvisc_coef_args vc = {up.p, vp.p, a->h, a->Kv_bbl_u, a->Kv_bbl_v, a->bbl_thick_u, a->bbl_thick_v, a->Kv_shear, a->Kv_shear_Bu, a->ustar, dt};
  if ((rc = oracle_vertvisc_coef(d, Gp, GV, US, vv_cs, &vc, a_u.p, a_v.p, h_u.p, h_v.p))) return 300 + rc;
  if ((rc = oracle_vertvisc_remnant(d, Gp, vv_cs, a->Ray_u, a->Ray_v, CS->visc_rem_u, CS->visc_rem_v, dt, a_u.p, a_v.p, h_u.p, h_v.p))) return 310 + rc;
  // :616-617 pass_eta, pass_visc_rem
  oracle_fill_halo_2d(d, CS->eta, 0, 0);
  fill3(d, CS->visc_rem_u, 1, nz); fill3(d, CS->visc_rem_v, 2, nz);
  // :629 bt_mass_source
  if ((rc = oracle_bt_mass_source(d, Gp, GV, a->h, CS->eta, 1, CS->barotropic->eta_cor))) return 400 + rc;
  // :646-651 continuity, btcalc
  mom6cu_continuity_args c1 = {a->u_inst, a->v_inst, a->h, hp.p, uh_in.p, vh_in.p, dt, nullptr, nullptr, nullptr, nullptr, CS->visc_rem_u,
                               CS->visc_rem_v, nullptr, nullptr, CS->BT_cont, nullptr, nullptr};
  if ((rc = oracle_continuity(d, Gp, GV, cont_cs, &c1, nthreads))) return 500 + rc;
  mom6cu_btcalc_args bc = {a->h, CS->BT_cont->h_u, CS->BT_cont->h_v, (double*)CS->barotropic->frhatu, (double*)CS->barotropic->frhatv,
                           Gp->bathyT, CS->hvel_scheme, 0};
  if ((rc = oracle_btcalc(d, Gp, GV, &bc, nthreads))) return 600 + rc;
  // :663-669 set_dtbt
  if (a->calc_dtbt) {
    mom6cu_set_dtbt_args sd = {};
    sd.pbce = CS->pbce; sd.frhatu = CS->barotropic->frhatu; sd.frhatv = CS->barotropic->frhatv; sd.bathyT = Gp->bathyT;
    sd.bebt = CS->barotropic->bebt; sd.G_extra = CS->barotropic->G_extra; sd.dtbt_fraction = CS->dtbt_fraction;
    sd.BT_Coriolis_scale = CS->BT_Coriolis_scale; sd.Z_ref = CS->Z_ref; sd.Nonlinear_continuity = CS->BT_Nonlinear_continuity;
    if (CS->dtbt_use_bt_cont) sd.BT_cont = CS->BT_cont; else sd.eta = CS->eta;
    if ((rc = oracle_set_dtbt(d, Gp, GV, US, &sd, &CS->barotropic->dtbt, &CS->dtbt_max))) return 650 + rc;
  }
  // :673 btstep (predictor)
  mom6cu_btstep_args b1 = {};
  b1.U_in = a->u_inst; b1.V_in = a->v_inst; b1.eta_in = CS->eta; b1.dt = dt; b1.bc_accel_u = u_bc_accel.p; b1.bc_accel_v = v_bc_accel.p;
  b1.taux = a->taux; b1.tauy = a->tauy; b1.pbce = CS->pbce; b1.eta_PF_in = CS->eta_PF; b1.U_Cor = CS->u_av; b1.V_Cor = CS->v_av;
  b1.accel_layer_u = CS->u_accel_bt; b1.accel_layer_v = CS->v_accel_bt; b1.eta_out = eta_pred.p; b1.uhbtav = CS->uhbt; b1.vhbtav = CS->vhbt;
  b1.visc_rem_u = CS->visc_rem_u; b1.visc_rem_v = CS->visc_rem_v; b1.BT_cont = CS->BT_cont;
  if (CS->split_bottom_stress) { b1.taux_bot = CS->taux_bot; b1.tauy_bot = CS->tauy_bot; }
  b1.uh0 = uh_in.p; b1.vh0 = vh_in.p; b1.u_uh0 = a->u_inst; b1.v_vh0 = a->v_inst;
  if ((rc = oracle_btstep(d, Gp, GV, CS->barotropic, &b1, nthreads))) return 700 + rc;
  // :681-691
  const double dt_pred = dt * CS->be;
  for (int k = 1; k <= nz; ++k) {
    for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i)
      vp(i, J, k) = G.mask2dCv(i, J) * (v(i, J, k) + dt_pred * (v_bc_accel(i, J, k) + v_accel_bt(i, J, k)));
    for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I)
      up(I, j, k) = G.mask2dCu(I, j) * (u(I, j, k) + dt_pred * (u_bc_accel(I, j, k) + u_accel_bt(I, j, k)));
  }
  // :738-768
  vc.dt = dt_pred;
  if ((rc = oracle_vertvisc_coef(d, Gp, GV, US, vv_cs, &vc, a_u.p, a_v.p, h_u.p, h_v.p))) return 800 + rc;
  mom6cu_vertvisc_args vs = {up.p, vp.p, a->h, a->taux, a->tauy, a->Ray_u, a->Ray_v, dt_pred, CS->taux_bot, CS->tauy_bot};
  if ((rc = oracle_vertvisc(d, Gp, GV, vv_cs, &vs, a_u.p, a_v.p, h_u.p, h_v.p))) return 810 + rc;
  if ((rc = oracle_vertvisc_remnant(d, Gp, vv_cs, a->Ray_u, a->Ray_v, CS->visc_rem_u, CS->visc_rem_v, CS->visc_rem_dt_bug ? dt_pred : dt, a_u.p, a_v.p,
                                    h_u.p, h_v.p))) return 820 + rc;
  fill3(d, CS->visc_rem_u, 1, nz); fill3(d, CS->visc_rem_v, 2, nz);
  fill3(d, up.p, 1, nz); fill3(d, vp.p, 2, nz);
  // :781 continuity
  mom6cu_continuity_args c2 = {up.p, vp.p, a->h, hp.p, a->uh, a->vh, dt, nullptr, nullptr, CS->uhbt, CS->vhbt, CS->visc_rem_u, CS->visc_rem_v,
                               CS->u_av, CS->v_av, CS->BT_cont, nullptr, nullptr};
  if ((rc = oracle_continuity(d, Gp, GV, cont_cs, &c2, nthreads))) return 900 + rc;
  // :785 pass_hp_uv
  fill3(d, hp.p, 0, nz); fill3(d, CS->u_av, 1, nz); fill3(d, CS->v_av, 2, nz); fill3(d, a->uh, 1, nz); fill3(d, a->vh, 2, nz);
  // :800-804
  for (int k = 1; k <= nz; ++k) for (int j = js - 2; j <= je + 2; ++j) for (int i = is - 2; i <= ie + 2; ++i) h_av(i, j, k) = 0.5 * (h(i, j, k) + hp(i, j, k));
  // :821 bt_mass_source(hp, eta_pred, .false.)
  if ((rc = oracle_bt_mass_source(d, Gp, GV, hp.p, eta_pred.p, 0, CS->barotropic->eta_cor))) return 1000 + rc;
  // :824-836
  if (CS->begw != 0.0) {
    for (int k = 1; k <= nz; ++k) for (int j = js - 1; j <= je + 1; ++j) for (int i = is - 1; i <= ie + 1; ++i)
      hp(i, j, k) = (1.0 - CS->begw) * h(i, j, k) + CS->begw * hp(i, j, k);
    mom6cu_pressureforce_args pb = {hp.p, a->T, a->S, CS->PFu, CS->PFv, a->p_surf, CS->pbce, CS->eta_PF};
    if ((rc = oracle_pressure_force(d, Gp, GV, pgf_cs, &pb, nthreads))) return 1100 + rc;
  }
  // :869 btcalc
  if ((rc = oracle_btcalc(d, Gp, GV, &bc, nthreads))) return 1200 + rc;
  // :886 horizontal_viscosity, :895 CorAdCalc
  mom6cu_hor_visc_args ha = {CS->u_av, CS->v_av, CS->h_av, a->uh, a->vh, CS->diffu, CS->diffv, CS->BT_cont->h_u, CS->BT_cont->h_v, dt};
  if ((rc = oracle_horizontal_viscosity(d, Gp, GV, hv_cs, &ha, nthreads))) return 1300 + rc;
  mom6cu_coradcalc_args cb = {CS->u_av, CS->v_av, CS->h_av, a->uh, a->vh, CS->CAu, CS->CAv, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if ((rc = oracle_coradcalc(d, Gp, GV, US, corad_cs, &cb, nthreads))) return 1400 + rc;
  // :901-908
  for (int k = 1; k <= nz; ++k) {
    for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) u_bc_accel(I, j, k) = (CAu(I, j, k) + PFu(I, j, k)) + diffu(I, j, k);
    for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) v_bc_accel(i, J, k) = (CAv(i, J, k) + PFv(i, J, k)) + diffv(i, J, k);
  }
  // :939 btstep (corrector)
  b1.U_Cor = CS->u_av; b1.V_Cor = CS->v_av; b1.uh0 = a->uh; b1.vh0 = a->vh; b1.u_uh0 = CS->u_av; b1.v_vh0 = CS->v_av; b1.etaav = a->eta_av;
  if ((rc = oracle_btstep(d, Gp, GV, CS->barotropic, &b1, nthreads))) return 1500 + rc;
  // :951, :961-975
  for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) eta(i, j) = eta_pred(i, j);
  for (int k = 1; k <= nz; ++k) {
    for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I)
      u(I, j, k) = G.mask2dCu(I, j) * (u(I, j, k) + dt * (u_bc_accel(I, j, k) + u_accel_bt(I, j, k)));
    for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i)
      v(i, J, k) = G.mask2dCv(i, J) * (v(i, J, k) + dt * (v_bc_accel(i, J, k) + v_accel_bt(i, J, k)));
  }
  // :1001-1016
  vc.u = a->u_inst; vc.v = a->v_inst; vc.dt = dt;
  if ((rc = oracle_vertvisc_coef(d, Gp, GV, US, vv_cs, &vc, a_u.p, a_v.p, h_u.p, h_v.p))) return 1600 + rc;
  mom6cu_vertvisc_args vt = {a->u_inst, a->v_inst, a->h, a->taux, a->tauy, a->Ray_u, a->Ray_v, dt, CS->taux_bot, CS->tauy_bot};
  if ((rc = oracle_vertvisc(d, Gp, GV, vv_cs, &vt, a_u.p, a_v.p, h_u.p, h_v.p))) return 1610 + rc;
  if ((rc = oracle_vertvisc_remnant(d, Gp, vv_cs, a->Ray_u, a->Ray_v, CS->visc_rem_u, CS->visc_rem_v, dt, a_u.p, a_v.p, h_u.p, h_v.p))) return 1620 + rc;
  // :1021-1023
  for (int k = 1; k <= nz; ++k) for (int j = js - 2; j <= je + 2; ++j) for (int i = is - 2; i <= ie + 2; ++i) h_av(i, j, k) = h(i, j, k);
  fill3(d, CS->visc_rem_u, 1, nz); fill3(d, CS->visc_rem_v, 2, nz);
  fill3(d, a->u_inst, 1, nz); fill3(d, a->v_inst, 2, nz);
  // :1043 continuity
  mom6cu_continuity_args c3 = {a->u_inst, a->v_inst, a->h, a->h, a->uh, a->vh, dt, nullptr, nullptr, CS->uhbt, CS->vhbt, CS->visc_rem_u,
                               CS->visc_rem_v, CS->u_av, CS->v_av, nullptr, nullptr, nullptr};
  if ((rc = oracle_continuity(d, Gp, GV, cont_cs, &c3, nthreads))) return 1700 + rc;
  // :1047 pass_h, :1054 pass_av_uvh
  fill3(d, a->h, 0, nz);
  fill3(d, CS->u_av, 1, nz); fill3(d, CS->v_av, 2, nz); fill3(d, a->uh, 1, nz); fill3(d, a->vh, 2, nz);
  // :1060-1062, :1067-1072
  for (int k = 1; k <= nz; ++k) for (int j = js - 2; j <= je + 2; ++j) for (int i = is - 2; i <= ie + 2; ++i) h_av(i, j, k) = 0.5 * (h_av(i, j, k) + h(i, j, k));
  for (int k = 1; k <= nz; ++k) {
    for (int j = js - 2; j <= je + 2; ++j) for (int I = Isq - 2; I <= Ieq + 2; ++I) uhtr(I, j, k) = uhtr(I, j, k) + uh(I, j, k) * dt;
    for (int J = Jsq - 2; J <= Jeq + 2; ++J) for (int i = is - 2; i <= ie + 2; ++i) vhtr(i, J, k) = vhtr(i, J, k) + vh(i, J, k) * dt;
  }
  // :1075-1083
  if (CS->store_CAu) {
    mom6cu_coradcalc_args cc = {CS->u_av, CS->v_av, CS->h_av, a->uh, a->vh, CS->CAu_pred, CS->CAv_pred, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if ((rc = oracle_coradcalc(d, Gp, GV, US, corad_cs, &cc, nthreads))) return 1800 + rc;
    CS->CAu_pred_stored = 1;
  } else CS->CAu_pred_stored = 0;
  return 0;
}
