// step_MOM_dyn_split_RK2 (/root/reference/src/core/MOM_dynamics_split_RK2.F90:294-1205) with every field resident on the
// device: the stage kernels (PressureForce, CorAdCalc, vertvisc*, continuity, btcalc / bt_mass_source / btstep,
// horizontal_viscosity) are launched back to back on the compute stream through their device-level entry points, the
// elementwise glue between them is three small kernels, and the group passes are in-place halo updates (periodic wrap on
// one tile, NCCL between tiles).  Nothing touches host memory between the first upload and the last download; arguments
// given as resident planes are used in place.
#include "ctx.h"
#include "common.cuh"
#include "stage.h"
#include <algorithm>

using m6::Geom;

namespace {

// u_bc_accel = (CA + PF) + diff  (:565-573, :901-908), optionally followed by up = mask * (u + dt * u_bc_accel) (:594-601)
struct AccelK {
  int is, ie, js, je;
  const double *maskCu, *maskCv, *CAu, *CAv, *PFu, *PFv, *diffu, *diffv, *u, *v;
  double *bcu, *bcv, *up, *vp;
  double dt;
  int predict;
};
__global__ void step_bc_accel_kernel(Geom G, AccelK P) {
  const int i = P.is - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = P.js - 1 + blockIdx.y, k = blockIdx.z;
  if (i > P.ie) return;
  const long long g = G.idx(i, j), o = (long long)k * G.plane + g;
  if (j >= P.js) {  // u-points (Isq:Ieq, js:je)
    const double a = (P.CAu[o] + P.PFu[o]) + P.diffu[o];
    P.bcu[o] = a;
    if (P.predict) P.up[o] = P.maskCu[g] * (P.u[o] + P.dt * a);
  }
  if (i >= P.is) {  // v-points (is:ie, Jsq:Jeq)
    const double a = (P.CAv[o] + P.PFv[o]) + P.diffv[o];
    P.bcv[o] = a;
    if (P.predict) P.vp[o] = P.maskCv[g] * (P.v[o] + P.dt * a);
  }
}

// xp = mask * (x + dt * (bc_accel + accel_bt))  (:681-691 into up/vp; :961-975 in place)
struct VelK {
  int is, ie, js, je;
  const double *maskCu, *maskCv, *u, *v, *bcu, *bcv, *abu, *abv;
  double *up, *vp;
  double dt;
};
__global__ void step_vel_kernel(Geom G, VelK P) {
  const int i = P.is - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = P.js - 1 + blockIdx.y, k = blockIdx.z;
  if (i > P.ie) return;
  const long long g = G.idx(i, j), o = (long long)k * G.plane + g;
  if (j >= P.js) P.up[o] = P.maskCu[g] * (P.u[o] + P.dt * (P.bcu[o] + P.abu[o]));
  if (i >= P.is) P.vp[o] = P.maskCv[g] * (P.v[o] + P.dt * (P.bcv[o] + P.abv[o]));
}

// h-point glue on (is-m:ie+m, js-m:je+m): mode 0 out = 0.5*(a + b) (:800-804, :1060-1062); 1 out = a (:1021-1023);
// 2 out = (1-w)*a + w*b (:825-827)
__global__ void step_hmix_kernel(Geom G, int ilo, int ihi, int jlo, int jhi, int mode, double w, const double* __restrict__ a,
                                 const double* __restrict__ b, double* __restrict__ out) {
  const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x, j = jlo + blockIdx.y, k = blockIdx.z;
  if (i > ihi) return;
  const long long o = (long long)k * G.plane + G.idx(i, j);
  if (mode == 0) out[o] = 0.5 * (a[o] + b[o]);
  else if (mode == 1) out[o] = a[o];
  else out[o] = (1.0 - w) * a[o] + w * b[o];
}

// uhtr = uhtr + uh*dt on (Isq-2:Ieq+2, js-2:je+2), vhtr likewise (:1067-1072); eta = eta_pred on (is:ie, js:je) (:951)
__global__ void step_accum_kernel(Geom G, int is, int ie, int js, int je, double dt, const double* __restrict__ uh, const double* __restrict__ vh,
                                  double* __restrict__ uhtr, double* __restrict__ vhtr) {
  const int i = is - 3 + blockIdx.x * blockDim.x + threadIdx.x, j = js - 3 + blockIdx.y, k = blockIdx.z;
  if (i > ie + 2) return;
  const long long o = (long long)k * G.plane + G.idx(i, j);
  if (j >= js - 2 && j <= je + 2) uhtr[o] = uhtr[o] + uh[o] * dt;                  // I = Isq-2 .. Ieq+2
  if (i >= is - 2 && j <= je + 2) vhtr[o] = vhtr[o] + vh[o] * dt;                  // J = Jsq-2 .. Jeq+2
}
__global__ void step_copy2_kernel(Geom G, int is, int ie, int js, int je, const double* __restrict__ src, double* __restrict__ dst) {
  const int i = is + blockIdx.x * blockDim.x + threadIdx.x, j = js + blockIdx.y;
  if (i > ie) return;
  const long long g = G.idx(i, j);
  dst[g] = src[g];
}

}  // namespace

// Device time of each stage inside the step (mom6cu_last_step_stage_ms): event pairs on the compute stream, read after finish().
struct StageTimer {
  mom6cu_ctx* c;
  int n = 0;
  int ids[40];
  explicit StageTimer(mom6cu_ctx* c_) : c(c_) {}
  void start(int id) {
    if (n >= 40) return;
    while ((int)c->stage_ev.size() < 2 * (n + 1)) { cudaEvent_t e = nullptr; cudaEventCreate(&e); c->stage_ev.push_back(e); }
    ids[n] = id;
    cudaEventRecord(c->stage_ev[2 * n], c->stream);
  }
  void stop() { if (n < 40) { cudaEventRecord(c->stage_ev[2 * n + 1], c->stream); ++n; } }
  void collect() {  // the stream has been synchronised
    for (int i = 0; i < 8; ++i) c->stage_ms[i] = 0.0;
    for (int i = 0; i < n; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, c->stage_ev[2 * i], c->stage_ev[2 * i + 1]) == cudaSuccess) c->stage_ms[ids[i]] += ms;
    }
  }
};
#define M6_TIMED(id, expr) do { ST_.start(id); rc = (expr); ST_.stop(); if (rc) return rc; } while (0)

extern "C" int mom6cu_step_dyn_split_rk2(mom6cu_ctx* c, mom6cu_dyn_split_rk2_cs* CS, const mom6cu_step_dyn_args* a) {
  if (!c || !CS || !a) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid || !c->have_vgrid || !c->have_cont_cs || !c->have_corad_cs || !c->have_hv_cs || !c->have_pgf_cs || !c->have_vv_cs)
    return c->fail(MOM6CU_ERR_BAD_ARG, "step_MOM_dyn_split_RK2: the grid and the continuity, CoriolisAdv, hor_visc, PressureForce and "
                                       "vertvisc control structures must be set first");
  if (CS->unsupported) return c->fail(MOM6CU_ERR_UNSUPPORTED, "step_MOM_dyn_split_RK2: the host configuration uses an option outside the frozen set");
  if (!CS->BT_cont || !CS->BT_cont->h_u || !CS->BT_cont->h_v || !CS->barotropic)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "step_MOM_dyn_split_RK2: BT_cont with h_u, h_v (BT_THICK_SCHEME=FROM_BT_CONT) is required");
  if (!a->u_inst || !a->v_inst || !a->h || !a->uh || !a->vh || !a->uhtr || !a->vhtr || !a->eta_av || !a->taux || !a->tauy)
    return c->fail(MOM6CU_ERR_BAD_ARG, "step_MOM_dyn_split_RK2: null required argument");
  const Geom& G = c->g;
  const mom6cu_domain& d = c->dom;
  const int nz = G.nk, is = d.isc, ie = d.iec, js = d.jsc, je = d.jec;
  const double dt = a->dt;
  Stager S(c, "step.");
  StageTimer ST_(c);
  int rc;
  // ---- arguments
  double *u, *v, *h, *uh, *vh, *uhtr, *vhtr, *eta_av;
  const double *T, *Sa, *p_surf;
  VvCoefDev VC = {};
  VvDev VS = {};
  // What the pressure force reads goes first, on the compute stream; the velocities, viscosities and stresses follow on the copy
  // stream while PressureForce runs (host-array callers only: a resident plane is used in place and costs nothing here).
  if ((rc = S.io3(a->h, ST_H, "h", &h)) || (rc = S.in3(a->T, ST_H, "T", &T)) || (rc = S.in3(a->S, ST_H, "S", &Sa)) ||
      (rc = S.in2(a->p_surf, ST_H, "p_surf", &p_surf)) || (rc = S.io3(a->uh, ST_U, "uh", &uh)) || (rc = S.io3(a->vh, ST_V, "vh", &vh)) ||
      (rc = S.io3(a->uhtr, ST_U, "uhtr", &uhtr)) || (rc = S.io3(a->vhtr, ST_V, "vhtr", &vhtr)) || (rc = S.io2(a->eta_av, ST_H, "eta_av", &eta_av)))
    return rc;
  if ((rc = S.defer_begin())) return rc;  // page-locked sources only; pageable ones stay on the compute stream
  rc = 0;
  if ((rc = S.io3(a->u_inst, ST_U, "u", &u)) || (rc = S.io3(a->v_inst, ST_V, "v", &v)) ||
      (rc = S.in2(a->Kv_bbl_u, ST_U, "kbu", &VC.Kv_bbl_u)) || (rc = S.in2(a->Kv_bbl_v, ST_V, "kbv", &VC.Kv_bbl_v)) ||
      (rc = S.in2(a->bbl_thick_u, ST_U, "btu", &VC.bbl_thick_u)) || (rc = S.in2(a->bbl_thick_v, ST_V, "btv", &VC.bbl_thick_v)) ||
      (rc = S.in(a->Kv_shear, ST_H, 0, nz + 1, "kvs", &VC.Kv_shear)) || (rc = S.in(a->Kv_shear_Bu, ST_Q, 0, nz + 1, "kvq", &VC.Kv_shear_Bu)) ||
      (rc = S.in2(a->ustar, ST_H, "ustar", &VC.ustar)) || (rc = S.in2(a->taux, ST_U, "taux", &VS.taux)) ||
      (rc = S.in2(a->tauy, ST_V, "tauy", &VS.tauy)) || (rc = S.in3(a->Ray_u, ST_U, "Ray_u", &VS.Ray_u)) ||
      (rc = S.in3(a->Ray_v, ST_V, "Ray_v", &VS.Ray_v))) {}
  {
    const int rc2 = S.defer_end();
    if (rc) return rc;
    if (rc2) return rc2;
  }
  // ---- the control structure's arrays
  double *CAu, *CAv, *CAu_pred, *CAv_pred, *PFu, *PFv, *diffu, *diffv, *vru, *vrv, *abu, *abv, *u_av, *v_av, *h_av, *pbce, *eta, *eta_PF,
      *uhbt, *vhbt, *taux_bot, *tauy_bot;
#define CS3(f, st, dst) if (!CS->f) return c->fail(MOM6CU_ERR_BAD_ARG, "step_MOM_dyn_split_RK2: CS%%" #f " is null"); if ((rc = S.io3(CS->f, st, "cs." #f, &dst))) return rc
#define CS2(f, st, dst) if (!CS->f) return c->fail(MOM6CU_ERR_BAD_ARG, "step_MOM_dyn_split_RK2: CS%%" #f " is null"); if ((rc = S.io2(CS->f, st, "cs." #f, &dst))) return rc
  CS3(CAu, ST_U, CAu); CS3(CAv, ST_V, CAv); CS3(CAu_pred, ST_U, CAu_pred); CS3(CAv_pred, ST_V, CAv_pred); CS3(PFu, ST_U, PFu); CS3(PFv, ST_V, PFv);
  CS3(diffu, ST_U, diffu); CS3(diffv, ST_V, diffv); CS3(visc_rem_u, ST_U, vru); CS3(visc_rem_v, ST_V, vrv);
  CS3(u_accel_bt, ST_U, abu); CS3(v_accel_bt, ST_V, abv); CS3(u_av, ST_U, u_av); CS3(v_av, ST_V, v_av); CS3(h_av, ST_H, h_av); CS3(pbce, ST_H, pbce);
  CS2(eta, ST_H, eta); CS2(eta_PF, ST_H, eta_PF); CS2(uhbt, ST_U, uhbt); CS2(vhbt, ST_V, vhbt); CS2(taux_bot, ST_U, taux_bot); CS2(tauy_bot, ST_V, tauy_bot);
#undef CS3
#undef CS2
  ContinuityDev CD = {};
  {
    const mom6cu_bt_cont* B = CS->BT_cont;
    double** dst2[12] = {&CD.FA_u_EE, &CD.FA_u_E0, &CD.FA_u_W0, &CD.FA_u_WW, &CD.uBT_WW, &CD.uBT_EE,
                         &CD.FA_v_NN, &CD.FA_v_N0, &CD.FA_v_S0, &CD.FA_v_SS, &CD.vBT_SS, &CD.vBT_NN};
    double* src2[12] = {B->FA_u_EE, B->FA_u_E0, B->FA_u_W0, B->FA_u_WW, B->uBT_WW, B->uBT_EE,
                        B->FA_v_NN, B->FA_v_N0, B->FA_v_S0, B->FA_v_SS, B->vBT_SS, B->vBT_NN};
    static const char* nm[12] = {"bc.FA_u_EE", "bc.FA_u_E0", "bc.FA_u_W0", "bc.FA_u_WW", "bc.uBT_WW", "bc.uBT_EE",
                                 "bc.FA_v_NN", "bc.FA_v_N0", "bc.FA_v_S0", "bc.FA_v_SS", "bc.vBT_SS", "bc.vBT_NN"};
    for (int m = 0; m < 12; ++m) {
      if (!src2[m]) return c->fail(MOM6CU_ERR_BAD_ARG, "step_MOM_dyn_split_RK2: BT_cont%%%s is not allocated", nm[m] + 3);
      if ((rc = S.io2(src2[m], m < 6 ? ST_U : ST_V, nm[m], dst2[m]))) return rc;
    }
    if ((rc = S.io3(B->h_u, ST_U, "bc.h_u", &CD.h_u)) || (rc = S.io3(B->h_v, ST_V, "bc.h_v", &CD.h_v))) return rc;
  }
  mom6cu_barotropic_cs BCS = *CS->barotropic;
  if ((rc = m6_stage_barotropic_cs(c, S, CS->barotropic, &BCS))) return rc;
  // frhatu / frhatv are written by btcalc inside the step: they must come back to a host-side CS
  if (!c->is_plane(CS->barotropic->frhatu)) S.outs.push_back({BCS.frhatu, (double*)CS->barotropic->frhatu, ST_U, 0, nz});
  if (!c->is_plane(CS->barotropic->frhatv)) S.outs.push_back({BCS.frhatv, (double*)CS->barotropic->frhatv, ST_V, 0, nz});
  // ---- work arrays (:341-374)
  double *up = c->plane3("step.up"), *vp = c->plane3("step.vp"), *hp = c->plane3("step.hp"), *bcu = c->plane3("step.bcu"),
         *bcv = c->plane3("step.bcv"), *uh_in = c->plane3("step.uh_in"), *vh_in = c->plane3("step.vh_in"), *eta_pred = c->plane2("step.eta_pred");
  if (!up || !vp || !hp || !bcu || !bcv || !uh_in || !vh_in || !eta_pred) return MOM6CU_ERR_CUDA;
  if ((rc = S.begin())) return rc;

  const size_t b3 = sizeof(double) * (size_t)G.plane * nz;
  const dim3 g3((ie - is + 2 + 127) / 128, je - js + 2, nz);
  auto halo = [&](std::initializer_list<double*> f, std::initializer_list<int> st, int nk) -> int {
    std::vector<double*> ff(f); std::vector<int> ss(st);
    return m6_halo_update(c, ff.data(), ss.data(), (int)ff.size(), 0, nk);
  };
  auto hmix = [&](int m, int mode, double w, const double* x, const double* y, double* out) {
    const dim3 g((ie - is + 1 + 2 * m + 127) / 128, je - js + 1 + 2 * m, nz);
    M6_LAUNCH(c, step_hmix_kernel, g, 128, 0, G, is - m, ie + m, js - m, je + m, mode, w, x, y, out);
  };
  // :425-430  up = vp = 0, hp = h
  M6_CUDA(c, cudaMemsetAsync(up, 0, b3, c->stream));
  M6_CUDA(c, cudaMemsetAsync(vp, 0, b3, c->stream));
  M6_CUDA(c, cudaMemcpyAsync(hp, h, b3, cudaMemcpyDeviceToDevice, c->stream));
  // :503 PressureForce
  PgfDev PD = {h, T, Sa, p_surf, PFu, PFv, pbce, eta_PF};
  M6_TIMED(MOM6CU_STAGE_PRESSURE_FORCE, m6_pressure_force_run(c, PD));
  if ((rc = S.wait_deferred())) return rc;  // u, v, visc%, forces% have arrived
  // :556 CorAdCalc (predictor accelerations, unless stored by the previous step)
  CorAdDev CA = {};
  CA.u = u_av; CA.v = v_av; CA.h = h_av; CA.uh = uh; CA.vh = vh;
  if (!CS->CAu_pred_stored) { CA.CAu = CAu_pred; CA.CAv = CAv_pred; M6_TIMED(MOM6CU_STAGE_CORADCALC, m6_coradcalc_run(c, CA)); }
  // :565-601
  AccelK AK = {is, ie, js, je, c->grid.mask2dCu, c->grid.mask2dCv, CAu_pred, CAv_pred, PFu, PFv, diffu, diffv, u, v, bcu, bcv, up, vp, dt, 1};
  M6_LAUNCH(c, step_bc_accel_kernel, g3, 128, 0, G, AK);
  // :609-610 vertvisc_coef, vertvisc_remnant
  VC.u = up; VC.v = vp; VC.h = h; VC.dt = dt;
  M6_TIMED(MOM6CU_STAGE_VERTVISC, (rc = m6_vertvisc_coef_run(c, VC)) ? rc : m6_vertvisc_remnant_run(c, VS.Ray_u, VS.Ray_v, vru, vrv, dt));
  // :616-617 pass_eta, pass_visc_rem
  if ((rc = halo({eta}, {ST_H}, 1)) || (rc = halo({vru, vrv}, {ST_U, ST_V}, nz))) return rc;
  // :629 bt_mass_source(h, eta, .true.)
  if ((rc = m6_bt_mass_source_run(c, h, eta, 1, BCS.eta_cor))) return rc;
  // :646-651 continuity (layer fluxes for the barotropic solver), btcalc
  ContinuityDev C1 = CD;
  C1.u = u; C1.v = v; C1.hin = h; C1.h = hp; C1.uh = uh_in; C1.vh = vh_in; C1.dt = dt; C1.visc_rem_u = vru; C1.visc_rem_v = vrv; C1.have_BT_cont = 1;
  M6_TIMED(MOM6CU_STAGE_CONTINUITY, m6_continuity_run(c, C1));
  M6_TIMED(MOM6CU_STAGE_BTCALC, m6_btcalc_run(c, h, CD.h_u, CD.h_v, c->grid.bathyT, CS->hvel_scheme, 0, (double*)BCS.frhatu, (double*)BCS.frhatv));
  // :663-669 set_dtbt
  if (a->calc_dtbt) {
    mom6cu_set_dtbt_args sd = {};
    mom6cu_bt_cont Bd = {};
    Bd.FA_u_EE = CD.FA_u_EE; Bd.FA_u_E0 = CD.FA_u_E0; Bd.FA_u_W0 = CD.FA_u_W0; Bd.FA_u_WW = CD.FA_u_WW;
    Bd.FA_v_NN = CD.FA_v_NN; Bd.FA_v_N0 = CD.FA_v_N0; Bd.FA_v_S0 = CD.FA_v_S0; Bd.FA_v_SS = CD.FA_v_SS;
    sd.pbce = pbce; sd.frhatu = BCS.frhatu; sd.frhatv = BCS.frhatv; sd.bathyT = c->grid.bathyT; sd.bebt = BCS.bebt; sd.G_extra = BCS.G_extra;
    sd.dtbt_fraction = CS->dtbt_fraction; sd.BT_Coriolis_scale = CS->BT_Coriolis_scale; sd.Z_ref = CS->Z_ref;
    sd.Nonlinear_continuity = CS->BT_Nonlinear_continuity;
    if (CS->dtbt_use_bt_cont) sd.BT_cont = &Bd; else sd.eta = eta;
    if ((rc = m6_set_dtbt_run(c, sd, &BCS.dtbt, &CS->dtbt_max))) return rc;
    CS->barotropic->dtbt = BCS.dtbt;
  }
  // :673 btstep (predictor)
  BtstepDev B1 = {};
  B1.dt = dt; B1.U_in = u; B1.V_in = v; B1.eta_in = eta; B1.bc_accel_u = bcu; B1.bc_accel_v = bcv; B1.taux = VS.taux; B1.tauy = VS.tauy;
  B1.pbce = pbce; B1.eta_PF_in = eta_PF; B1.U_Cor = u_av; B1.V_Cor = v_av; B1.visc_rem_u = vru; B1.visc_rem_v = vrv;
  if (CS->split_bottom_stress) { B1.taux_bot = taux_bot; B1.tauy_bot = tauy_bot; }
  B1.uh0 = uh_in; B1.vh0 = vh_in; B1.u_uh0 = u; B1.v_vh0 = v;
  B1.accel_layer_u = abu; B1.accel_layer_v = abv; B1.eta_out = eta_pred; B1.uhbtav = uhbt; B1.vhbtav = vhbt; B1.etaav = nullptr;
  B1.have_BT_cont = 1;
  B1.FA_u_EE = CD.FA_u_EE; B1.FA_u_E0 = CD.FA_u_E0; B1.FA_u_W0 = CD.FA_u_W0; B1.FA_u_WW = CD.FA_u_WW; B1.uBT_WW = CD.uBT_WW; B1.uBT_EE = CD.uBT_EE;
  B1.FA_v_NN = CD.FA_v_NN; B1.FA_v_N0 = CD.FA_v_N0; B1.FA_v_S0 = CD.FA_v_S0; B1.FA_v_SS = CD.FA_v_SS; B1.vBT_SS = CD.vBT_SS; B1.vBT_NN = CD.vBT_NN;
  M6_TIMED(MOM6CU_STAGE_BTSTEP, m6_btstep_run(c, BCS, B1));
  // :681-691
  const double dt_pred = dt * CS->be;
  VelK VK = {is, ie, js, je, c->grid.mask2dCu, c->grid.mask2dCv, u, v, bcu, bcv, abu, abv, up, vp, dt_pred};
  M6_LAUNCH(c, step_vel_kernel, g3, 128, 0, G, VK);
  // :738-768 vertvisc_coef, vertvisc, vertvisc_remnant
  VC.dt = dt_pred;
  VS.u = up; VS.v = vp; VS.h = h; VS.dt = dt_pred; VS.taux_bot = taux_bot; VS.tauy_bot = tauy_bot;
  M6_TIMED(MOM6CU_STAGE_VERTVISC, (rc = m6_vertvisc_coef_run(c, VC)) ? rc : (rc = m6_vertvisc_run(c, VS)) ? rc :
           m6_vertvisc_remnant_run(c, VS.Ray_u, VS.Ray_v, vru, vrv, CS->visc_rem_dt_bug ? dt_pred : dt));
  if ((rc = halo({vru, vrv, up, vp}, {ST_U, ST_V, ST_U, ST_V}, nz))) return rc;
  // :781 continuity
  ContinuityDev C2 = CD;
  C2.u = up; C2.v = vp; C2.hin = h; C2.h = hp; C2.uh = uh; C2.vh = vh; C2.dt = dt; C2.uhbt = uhbt; C2.vhbt = vhbt; C2.visc_rem_u = vru;
  C2.visc_rem_v = vrv; C2.u_cor = u_av; C2.v_cor = v_av; C2.have_BT_cont = 1;
  M6_TIMED(MOM6CU_STAGE_CONTINUITY, m6_continuity_run(c, C2));
  // :785 pass_hp_uv
  if ((rc = halo({hp, u_av, v_av, uh, vh}, {ST_H, ST_U, ST_V, ST_U, ST_V}, nz))) return rc;
  // :800-804
  hmix(2, 0, 0., h, hp, h_av);
  // :821 bt_mass_source(hp, eta_pred, .false.)
  if ((rc = m6_bt_mass_source_run(c, hp, eta_pred, 0, BCS.eta_cor))) return rc;
  // :824-836
  if (CS->begw != 0.0) {
    hmix(1, 2, CS->begw, h, hp, hp);
    PgfDev P2 = {hp, T, Sa, p_surf, PFu, PFv, pbce, eta_PF};
    M6_TIMED(MOM6CU_STAGE_PRESSURE_FORCE, m6_pressure_force_run(c, P2));
  }
  // :869 btcalc
  M6_TIMED(MOM6CU_STAGE_BTCALC, m6_btcalc_run(c, h, CD.h_u, CD.h_v, c->grid.bathyT, CS->hvel_scheme, 0, (double*)BCS.frhatu, (double*)BCS.frhatv));
  // :886 horizontal_viscosity, :895 CorAdCalc
  HorViscDev HV = {u_av, v_av, h_av, CD.h_u, CD.h_v, diffu, diffv};
  M6_TIMED(MOM6CU_STAGE_HOR_VISC, m6_hor_visc_run(c, HV));
  CA.CAu = CAu; CA.CAv = CAv;
  M6_TIMED(MOM6CU_STAGE_CORADCALC, m6_coradcalc_run(c, CA));
  // :901-908
  AK.CAu = CAu; AK.CAv = CAv; AK.predict = 0;
  M6_LAUNCH(c, step_bc_accel_kernel, g3, 128, 0, G, AK);
  // :939 btstep (corrector)
  B1.uh0 = uh; B1.vh0 = vh; B1.u_uh0 = u_av; B1.v_vh0 = v_av; B1.etaav = eta_av;
  M6_TIMED(MOM6CU_STAGE_BTSTEP, m6_btstep_run(c, BCS, B1));
  // :951, :961-975
  M6_LAUNCH(c, step_copy2_kernel, dim3((ie - is + 128) / 128, je - js + 1), 128, 0, G, is, ie, js, je, eta_pred, eta);
  VK.up = u; VK.vp = v; VK.dt = dt;
  M6_LAUNCH(c, step_vel_kernel, g3, 128, 0, G, VK);
  // :1001-1016
  VC.u = u; VC.v = v; VC.dt = dt;
  VS.u = u; VS.v = v; VS.dt = dt;
  M6_TIMED(MOM6CU_STAGE_VERTVISC, (rc = m6_vertvisc_coef_run(c, VC)) ? rc : (rc = m6_vertvisc_run(c, VS)) ? rc :
           m6_vertvisc_remnant_run(c, VS.Ray_u, VS.Ray_v, vru, vrv, dt));
  // :1021-1023
  hmix(2, 1, 0., h, h, h_av);
  if ((rc = halo({vru, vrv, u, v}, {ST_U, ST_V, ST_U, ST_V}, nz))) return rc;
  if ((rc = S.early(u)) || (rc = S.early(v))) return rc;  // u_inst, v_inst are final: copy them back under the last continuity call
  // :1043 continuity (in place in h)
  ContinuityDev C3 = {};
  C3.u = u; C3.v = v; C3.hin = h; C3.h = h; C3.uh = uh; C3.vh = vh; C3.dt = dt; C3.uhbt = uhbt; C3.vhbt = vhbt; C3.visc_rem_u = vru;
  C3.visc_rem_v = vrv; C3.u_cor = u_av; C3.v_cor = v_av;
  M6_TIMED(MOM6CU_STAGE_CONTINUITY, m6_continuity_run(c, C3));
  // :1047 pass_h, :1054 pass_av_uvh
  if ((rc = halo({h, u_av, v_av, uh, vh}, {ST_H, ST_U, ST_V, ST_U, ST_V}, nz))) return rc;
  // :1060-1062, :1067-1072
  hmix(2, 0, 0., h_av, h, h_av);
  M6_LAUNCH(c, step_accum_kernel, dim3((ie - is + 6 + 127) / 128, je - js + 6, nz), 128, 0, G, is, ie, js, je, dt, uh, vh, uhtr, vhtr);
  // :1075-1083
  if (CS->store_CAu) {
    CA.CAu = CAu_pred; CA.CAv = CAv_pred;
    M6_TIMED(MOM6CU_STAGE_CORADCALC, m6_coradcalc_run(c, CA));
    CS->CAu_pred_stored = 1;
  } else CS->CAu_pred_stored = 0;
  M6_CUDA(c, cudaGetLastError());
  rc = S.finish();
  ST_.collect();
  return rc;
}
