// Vertical friction: vertvisc_coef (/root/reference/src/parameterizations/vertical/MOM_vert_friction.F90:1357-2310 with
// find_coupling_coef :2314-2924), vertvisc :557-1226 and vertvisc_remnant :1229-1354, for the frozen option set declared in
// include/mom6cu.h.  One thread per velocity column.  vertvisc_coef is a single bottom-up sweep: the coupling coefficient
// of interface K needs only the velocity-point thicknesses of layers K-1 and K and the normalised height z_i(K), all of
// which the sweep has in registers when it passes layer K-1, so no column arrays are kept (the surface-intensified
// options, which integrate downward, re-read the sweep's dz_vel / dz_harm from two scratch fields).  The solvers are the
// reference's Schopf & Loughe sweep with c1(k) in a thread-local column.  Expression order is the reference's (bitwise
// with -fmad=false).
#include "ctx.h"
#include "common.cuh"
#include "stage.h"
#include <algorithm>
#include <cfloat>
#include <cmath>

using m6::Geom;
using m6::fmax2;
using m6::fmin2;

namespace {

constexpr int KMAX = 128;

struct CoefK {
  mom6cu_vertvisc_cs CS;
  int nz, i0, i1, j0, j1;
  double h_neglect, H_to_Z, Z_to_H, a_cpl_max;
  const double *mask, *bathyT, *CoriolisBu;
  const double *vel, *h;            // 3-D
  const double *kv_bbl, *bbl_thick; // 2-D at this velocity point
  const double *Kv_shear, *Kv_shear_Bu, *ustar;
  double *a_out, *h_out;            // nk+1 / nk levels
  double *dzvel, *dzharm;           // scratch (nk levels), only with the surface-intensified options
};

__device__ __forceinline__ double botfn6(double z2) { return 1. / (1. + 0.09 * z2 * z2 * z2 * z2 * z2 * z2); }

template <int DIR>  // 0: u-points (I,j), 1: v-points (i,J)
__global__ void __launch_bounds__(128, 6) vv_coef_kernel(Geom G, CoefK P) {
  const int i = P.i0 + blockIdx.x * blockDim.x + threadIdx.x, j = P.j0 + blockIdx.y;
  if (i > P.i1) return;
  const long long g = G.idx(i, j), pl = G.plane;
  if (!(P.mask[g] > 0.)) return;
  const long long sB = DIR ? G.pitch : 1;           // offset of the second thickness cell
  const long long sQ = DIR ? -1 : -(long long)G.pitch;  // offset of the second vertex: Bu(I,J-1) for u, Bu(I-1,J) for v
  const mom6cu_vertvisc_cs& CS = P.CS;
  const int nz = P.nz;
  const double h_neglect = P.h_neglect, dz_neglect = CS.dZ_subroundoff, hn = CS.dZ_subroundoff;
  double I_Hbbl = 1. / (CS.Hbbl + dz_neglect);
  const double I_valBL = (CS.harm_BL_val > 0.0) ? 1.0 / CS.harm_BL_val : 0.0;
  double kv_bbl = 0., bbl_thick = 0.;
  if (CS.bottomdraglaw) { kv_bbl = P.kv_bbl[g]; bbl_thick = P.bbl_thick[g] + dz_neglect; I_Hbbl = 1. / bbl_thick; }
  const double DA = P.bathyT[g], DB = P.bathyT[g + sB];
  const double Dmin = fmin2(DA, DB);
  const bool surf = (CS.Kvml_invZ2 > 0.) || CS.fixed_LOTW_ML || CS.apply_LOTW_floor;
  const bool same_units = (P.H_to_Z == 1.0) && (h_neglect == dz_neglect);
  double z_i_below = 0., zh = 0., zcolA = -DA, zcolB = -DB;  // z_i(k+1)
  double hv_below = 0.;                                       // dz_vel(k+1)
  // The loads of the level above are issued before this level's arithmetic and stores (the compiler cannot move them across
  // the stores itself: it has to assume h_out / a_out may alias the inputs), so two levels of a column are in flight.
  double hA_n, hB_n, vel_n, ksA_n = 0., ksB_n = 0., kqA_n = 0., kqB_n = 0.;
  {
    const long long o = (long long)(nz - 1) * pl + g;
    hA_n = __ldg(P.h + o); hB_n = __ldg(P.h + o + sB); vel_n = __ldg(P.vel + o);
  }
  for (int k = nz; k >= 1; --k) {
    const long long o = (long long)(k - 1) * pl + g;
    const double hA = hA_n, hB = hB_n;
    const double ksA = ksA_n, ksB = ksB_n, kqA = kqA_n, kqB = kqB_n;   // Kv_shear[_Bu] at the interface K = k+1
    const double vel_k = vel_n;
    if (k > 1) {
      const long long on = o - pl;
      hA_n = __ldg(P.h + on); hB_n = __ldg(P.h + on + sB); vel_n = __ldg(P.vel + on);
      if (P.Kv_shear) { ksA_n = __ldg(P.Kv_shear + o); ksB_n = __ldg(P.Kv_shear + o + sB); }             // interface K = k
      if (P.Kv_shear_Bu) { kqA_n = __ldg(P.Kv_shear_Bu + o + sQ); kqB_n = __ldg(P.Kv_shear_Bu + o); }
    }
    const double dzA = P.H_to_Z * hA, dzB = P.H_to_Z * hB;
    const double h_harm = 2. * hA * hB / (hA + hB + h_neglect);
    const double h_arith = 0.5 * (hB + hA);
    const double h_delta = hB - hA;
    // with H_to_Z = 1 and equal roundoff thicknesses this is the very expression of h_harm: one division less per layer
    const double dz_harm = same_units ? h_harm : 2. * dzA * dzB / (dzA + dzB + dz_neglect);
    const double dz_arith = 0.5 * (dzB + dzA);
    const double vel = vel_k;
    double hvel, dz_vel, z_i;
    if (CS.harmonic_visc) {
      hvel = h_harm; dz_vel = dz_harm;
      if (vel * h_delta < 0) {
        const double botfn = botfn6(z_i_below);
        hvel = (1. - botfn) * h_harm + botfn * h_arith;
        dz_vel = (1. - botfn) * dz_harm + botfn * dz_arith;
      }
      z_i = z_i_below + dz_harm * I_Hbbl;
    } else {
      zcolA = zcolA + dzA; zcolB = zcolB + dzB;
      zh = zh + dz_harm;
      const double z_clear = fmax2(zcolA, zcolB) + Dmin;
      z_i = fmax2(zh, z_clear) * I_Hbbl;
      hvel = h_arith; dz_vel = dz_arith;
      if (vel * h_delta > 0.) {
        if (zh * I_Hbbl < CS.harm_BL_val) { hvel = h_harm; dz_vel = dz_harm; }
        else {
          double z2_wt = 1.;
          if (zh * I_Hbbl < 2. * CS.harm_BL_val) z2_wt = fmax2(0., fmin2(1., zh * I_Hbbl * I_valBL - 1.));
          const double botfn = botfn6(z2_wt * (fmax2(zh, z_clear) * I_Hbbl));
          hvel = (1. - botfn) * h_arith + botfn * h_harm;
          dz_vel = (1. - botfn) * dz_arith + botfn * dz_harm;
        }
      }
    }
    P.h_out[o] = hvel + h_neglect;
    if (surf) { P.dzvel[o] = dz_vel; P.dzharm[o] = dz_harm; }
    // the coupling coefficient of the interface below this layer, K = k+1 (find_coupling_coef :2434-2560)
    if (k == nz) {
      double a;
      if (CS.bottomdraglaw) a = kv_bbl / ((fmin2(dz_vel * 0.5, bbl_thick) + hn) + 0.0 * kv_bbl);
      else if (fabs(CS.Kv_extra_bbl) > 0.0) a = (CS.Kv + CS.Kv_extra_bbl) / ((0.5 * dz_vel + hn) + 0.0 * (CS.Kv + CS.Kv_extra_bbl));
      else a = CS.Kv / ((0.5 * dz_vel + hn) + 0.0 * CS.Kv);
      P.a_out[(long long)nz * pl + g] = (surf ? a : fmin2(P.a_cpl_max, a));
    } else {
      const int K = k + 1;
      const long long oK = (long long)(K - 1) * pl;
      double Kv_tot = CS.Kv;  // the Kvml_invZ2 term is added in the downward pass
      double Kv_extra = 0.;
      bool has_extra = false;
      if (P.Kv_shear) { Kv_extra = 0.5 * (ksA + ksB); has_extra = true; }
      double Kv_bu = 0.;
      if (P.Kv_shear_Bu) Kv_bu = 0.5 * (kqA + kqB);
      if (!(CS.Kvml_invZ2 > 0.)) {
        if (has_extra) Kv_tot = Kv_tot + Kv_extra;
        if (P.Kv_shear_Bu) Kv_tot = Kv_tot + Kv_bu;
        double a;
        if (CS.bottomdraglaw) {
          const double botfn = botfn6(z_i_below);
          Kv_tot = Kv_tot + (kv_bbl - CS.Kv) * botfn;
          const double dhc = 0.5 * (hv_below + dz_vel);
          const double h_shear = (dhc > bbl_thick) ? ((1. - botfn) * dhc + botfn * bbl_thick) + hn : dhc + hn;
          a = Kv_tot / (h_shear + (0.0 * Kv_tot));
        } else if (fabs(CS.Kv_extra_bbl) > 0.0) {
          Kv_tot = Kv_tot + CS.Kv_extra_bbl * botfn6(z_i_below);
          a = Kv_tot / (0.5 * (hv_below + dz_vel + hn) + 0.0 * Kv_tot);
        } else a = Kv_tot / (0.5 * (hv_below + dz_vel + hn) + 0.0 * Kv_tot);
        P.a_out[oK + g] = (surf ? a : fmin2(P.a_cpl_max, a));
      } else {
        P.a_out[oK + g] = z_i_below;  // parked for the downward pass, which needs z_i(K)
      }
    }
    z_i_below = z_i; hv_below = dz_vel;
  }
  if (!surf) { P.a_out[g] = fmin2(P.a_cpl_max, 0.0); return; }

  // ---- downward pass: the 1997-vintage Kvml_invZ2 enhancement (:2419-2432) and the law-of-the-wall options (:2643-2923)
  if (CS.Kvml_invZ2 > 0.) {
    const double I_Hmix = 1. / (CS.Hmix + hn);
    double z_t = hn * I_Hmix;
    for (int K = 2; K <= nz; ++K) {
      const long long oK = (long long)(K - 1) * pl, ok = (long long)(K - 2) * pl;
      z_t = z_t + P.dzharm[ok + g] * I_Hmix;
      double Kv_tot = CS.Kv + CS.Kvml_invZ2 / ((z_t * z_t) * (1. + 0.09 * z_t * z_t * z_t * z_t * z_t * z_t));
      if (P.Kv_shear) Kv_tot = Kv_tot + 0.5 * (P.Kv_shear[oK + g] + P.Kv_shear[oK + g + sB]);
      if (P.Kv_shear_Bu) Kv_tot = Kv_tot + 0.5 * (P.Kv_shear_Bu[oK + g + sQ] + P.Kv_shear_Bu[oK + g]);
      const double z_iK = P.a_out[oK + g];
      const double hvK = P.dzvel[oK + g], hvKm = P.dzvel[ok + g];
      double a;
      if (CS.bottomdraglaw) {
        const double botfn = botfn6(z_iK);
        Kv_tot = Kv_tot + (kv_bbl - CS.Kv) * botfn;
        const double dhc = 0.5 * (hvK + hvKm);
        const double h_shear = (dhc > bbl_thick) ? ((1. - botfn) * dhc + botfn * bbl_thick) + hn : dhc + hn;
        a = Kv_tot / (h_shear + (0.0 * Kv_tot));
      } else if (fabs(CS.Kv_extra_bbl) > 0.0) {
        Kv_tot = Kv_tot + CS.Kv_extra_bbl * botfn6(z_iK);
        a = Kv_tot / (0.5 * (hvK + hvKm + hn) + 0.0 * Kv_tot);
      } else a = Kv_tot / (0.5 * (hvK + hvKm + hn) + 0.0 * Kv_tot);
      P.a_out[oK + g] = a;
    }
  }
  if (CS.fixed_LOTW_ML || CS.apply_LOTW_floor) {
    const double u_star = 0.5 * (P.ustar[g] + P.ustar[g + sB]);
    const double absf = 0.5 * (fabs(P.CoriolisBu[g + sQ]) + fabs(P.CoriolisBu[g]));
    double h_ml = 0.0;
    int nk_in_ml = 0;
    for (int k = 1; k <= nz; ++k) {
      if (h_ml < CS.Hmix) {
        nk_in_ml = k;
        const double hvk = P.dzvel[(long long)(k - 1) * pl + g];
        if (h_ml + hvk < CS.Hmix) h_ml = h_ml + hvk; else h_ml = CS.Hmix;
      } else break;
    }
    if (u_star <= 0.0) nk_in_ml = 0;
    double z_t = 0.0;
    for (int K = 2; K <= nk_in_ml; ++K) {
      const long long oK = (long long)(K - 1) * pl, ok = (long long)(K - 2) * pl;
      const double hvK = P.dzvel[oK + g], hvKm = P.dzvel[ok + g];
      z_t = z_t + hvKm;
      const double temp1 = (z_t * h_ml - z_t * z_t);
      double a = P.a_out[oK + g];
      if (CS.apply_LOTW_floor && CS.fixed_LOTW_ML) {
        const double ustar2_denom = (CS.vonKar * P.Z_to_H * (u_star * u_star)) / (absf * temp1 + (h_ml + hn) * u_star);
        const double visc_ml = temp1 * ustar2_denom;
        const double a_ml = visc_ml / (0.25 * (hvK + hvKm + hn));
        const double a_floor = (h_ml - z_t) * ustar2_denom;
        a = fmax2(fmax2(a, a_ml), a_floor);
      } else if (CS.apply_LOTW_floor) {
        const double ustar2_denom = (CS.vonKar * P.Z_to_H * (u_star * u_star)) / (absf * temp1 + (h_ml + hn) * u_star);
        a = fmax2(a, (h_ml - z_t) * ustar2_denom);
      } else {
        const double visc_ml = u_star * CS.vonKar * (P.Z_to_H * temp1 * u_star) / (absf * temp1 + (h_ml + hn) * u_star);
        const double a_ml = visc_ml / (0.25 * (hvK + hvKm + hn) + 0.5 * 0.0 * visc_ml);
        a = fmax2(a, a_ml);
      }
      P.a_out[oK + g] = a;
    }
  }
  P.a_out[g] = fmin2(P.a_cpl_max, 0.0);
  for (int K = 2; K <= nz + 1; ++K) { const long long oK = (long long)(K - 1) * pl + g; P.a_out[oK] = fmin2(P.a_cpl_max, P.a_out[oK]); }
}

struct SolveK {
  int nz, i0, i1, j0, j1;    // columns visited
  int js_stress, js_solve;   // first row with a surface stress / first row solved (the u solver starts at j = G%isc, :778)
  int direct_stress;
  double dt, dt_Rho0, h_neglect, Hmix, I_Hmix, H_to_RZ;
  const double *mask, *a, *hh, *Ray, *tau, *h;
  double* x;        // u / v (in/out) or the remnant (out)
  double* tau_bot;  // optional
  // vertvisc_limit_vel :2926-3120 (lim_on: some test is active; applied on rows >= js_lim)
  int lim_on, cfl_based, js_lim;
  double vel_underflow, CFL_trunc, maxvel, H_report;
  const double *face, *areaT, *IareaT;  // G%dy_Cu / dx_Cv; G%areaT, G%IareaT
  unsigned long long* ntrunc;
};

// vertvisc_limit_vel without truncation files (:3016-3037 for u, :3090-3111 for v) for one velocity: the CFL number of the face
// velocity in the cell it flows out of; a truncated velocity is the one of CFL 0.9*CFL_trunc.  hsum = h(i,j,k) + h(i+1,j,k) (or j+1).
__device__ __forceinline__ double vv_limit(const SolveK& P, double u, double dtf, double IA0, double IA1, double A0, double A1,
                                           bool have_h, double hsum) {
  double un = u;
  bool trunc = false;
  if (fabs(u) < P.vel_underflow) un = 0.0;
  else if (P.cfl_based) {
    if (P.CFL_trunc > 0.0) {
      if ((u * dtf) * IA1 < -P.CFL_trunc) { un = (-0.9 * P.CFL_trunc) * (A1 / dtf); trunc = true; }
      else if ((u * dtf) * IA0 > P.CFL_trunc) { un = (0.9 * P.CFL_trunc) * (A0 / dtf); trunc = true; }
    }
  } else if (P.maxvel > 0.0 && fabs(u) > P.maxvel) { un = copysign(0.9 * P.maxvel, u); trunc = true; }
  if (trunc && have_h && (hsum > P.H_report)) atomicAdd(P.ntrunc, 1ULL);
  return un;
}

template <int DIR, bool REM>
__global__ void __launch_bounds__(128) vv_solve_kernel(Geom G, SolveK P) {
  const int i = P.i0 + blockIdx.x * blockDim.x + threadIdx.x, j = P.j0 + blockIdx.y;
  if (i > P.i1) return;
  const long long g = G.idx(i, j), pl = G.plane;
  const int nz = P.nz;
  const double mask = P.mask[g], dt = P.dt;
  double surface_stress = 0.0;
  bool tau_done = false, lim_done = false;
  if (!REM && j >= P.js_stress) {
    if (P.direct_stress) {  // :705-718
      if (mask > 0.) {
        const long long sB = DIR ? G.pitch : 1;
        double zDS = 0.0;
        const double stress = P.dt_Rho0 * P.tau[g];
        for (int k = 1; k <= nz; ++k) {
          const long long o = (long long)(k - 1) * pl + g;
          const double h_a = 0.5 * (P.h[o] + P.h[o + sB]) + P.h_neglect;
          double hfr = 1.0; if ((zDS + h_a) > P.Hmix) hfr = (P.Hmix - zDS) / h_a;
          P.x[o] = P.x[o] + P.I_Hmix * hfr * stress;
          zDS = zDS + h_a; if (zDS >= P.Hmix) break;
        }
      }
    } else surface_stress = P.dt_Rho0 * (mask * P.tau[g]);
  }
  if (mask > 0. && j >= P.js_solve) {
    double c1[KMAX + 1];
    // a, hh, Ray are read-only here; x is read and written at different levels.  The next level's operands are loaded before
    // this level's store (which the compiler must otherwise treat as a possible alias), so two levels are in flight.
    double a_k1 = __ldg(P.a + pl + g);                                   // a(K=2)
    double b_denom_1 = __ldg(P.hh + g) + dt * ((P.Ray ? __ldg(P.Ray + g) : 0.) + __ldg(P.a + g));
    double b1 = 1.0 / (b_denom_1 + dt * a_k1);
    double d1 = b_denom_1 * b1;
    double xm = REM ? b1 * __ldg(P.hh + g) : b1 * (__ldg(P.hh + g) * P.x[g] + surface_stress);
    double hk_n = 0., ray_n = 0., x_n = 0., a_n = 0.;
    if (nz >= 2) {
      hk_n = __ldg(P.hh + pl + g); if (P.Ray) ray_n = __ldg(P.Ray + pl + g);
      if (!REM) x_n = P.x[pl + g];
      a_n = __ldg(P.a + 2 * pl + g);                                       // a(K=3)
    }
    P.x[g] = xm;
    for (int k = 2; k <= nz; ++k) {
      const long long o = (long long)(k - 1) * pl + g;
      const double aK = a_k1, hk = hk_n, ray = ray_n, xk = x_n, a_below = a_n;
      if (k < nz) {
        hk_n = __ldg(P.hh + o + pl); if (P.Ray) ray_n = __ldg(P.Ray + o + pl);
        if (!REM) x_n = P.x[o + pl];
        a_n = __ldg(P.a + o + 2 * pl);
      }
      a_k1 = a_below;
      c1[k] = dt * aK * b1;
      b_denom_1 = hk + dt * ((P.Ray ? ray : 0.) + aK * d1);
      b1 = 1.0 / (b_denom_1 + dt * a_below);
      d1 = b_denom_1 * b1;
      xm = REM ? (hk + dt * aK * xm) * b1 : (hk * xk + dt * aK * xm) * b1;
      P.x[o] = xm;
    }
    // Without Rayleigh drag the bottom stress needs only the bottom velocity, so the truncation of vertvisc_limit_vel (which the
    // reference applies after the stress) can be folded into the back-substitution: the recurrence carries the untruncated
    // velocity in a register and the truncated one is what is stored.  No extra pass over the column.
    const bool fuse = !REM && P.lim_on && !P.Ray && j >= P.js_lim;
    if (fuse) {
      lim_done = true;
      const long long sB = DIR ? G.pitch : 1;
      const double dtf = dt * P.face[g], IA0 = P.IareaT[g], IA1 = P.IareaT[g + sB], A0 = P.areaT[g], A1 = P.areaT[g + sB];
      const long long ob = (long long)(nz - 1) * pl + g;
      if (P.tau_bot && j >= P.js_stress) { P.tau_bot[g] = P.H_to_RZ * (xm * P.a[(long long)nz * pl + g]); tau_done = true; }
      {
        const double un = vv_limit(P, xm, dtf, IA0, IA1, A0, A1, P.h != nullptr, P.h ? P.h[ob] + P.h[ob + sB] : 0.0);
        if (__double_as_longlong(un) != __double_as_longlong(xm)) P.x[ob] = un;
      }
      if (nz >= 2) {
        double xk_n = P.x[(long long)(nz - 2) * pl + g];
        for (int k = nz - 1; k >= 1; --k) {
          const long long o = (long long)(k - 1) * pl + g;
          const double xk = xk_n;
          if (k > 1) xk_n = P.x[o - pl];
          xm = xk + c1[k + 1] * xm;
          P.x[o] = vv_limit(P, xm, dtf, IA0, IA1, A0, A1, P.h != nullptr, P.h ? P.h[o] + P.h[o + sB] : 0.0);
        }
      }
    } else if (nz >= 2) {
      double xk_n = P.x[(long long)(nz - 2) * pl + g];
      for (int k = nz - 1; k >= 1; --k) {
        const long long o = (long long)(k - 1) * pl + g;
        const double xk = xk_n;
        if (k > 1) xk_n = P.x[o - pl];
        xm = xk + c1[k + 1] * xm;
        P.x[o] = xm;
      }
    }
  }
  if (!REM && P.tau_bot && j >= P.js_stress && !tau_done) {  // :903-912
    double t = P.H_to_RZ * (P.x[(long long)(nz - 1) * pl + g] * P.a[(long long)nz * pl + g]);
    if (P.Ray) for (int k = 1; k <= nz; ++k) { const long long o = (long long)(k - 1) * pl + g; t = t + P.H_to_RZ * (P.Ray[o] * P.x[o]); }
    P.tau_bot[g] = t;
  }
  if (!REM && P.lim_on && j >= P.js_lim && !lim_done) {
    // the separate pass (land columns, columns with Rayleigh drag): the stress above used the untruncated velocities
    const long long sB = DIR ? G.pitch : 1;
    const double dtf = dt * P.face[g], IA0 = P.IareaT[g], IA1 = P.IareaT[g + sB], A0 = P.areaT[g], A1 = P.areaT[g + sB];
    for (int k = 1; k <= nz; ++k) {
      const long long o = (long long)(k - 1) * pl + g;
      const double u = P.x[o];
      const double un = vv_limit(P, u, dtf, IA0, IA1, A0, A1, P.h != nullptr, P.h ? P.h[o] + P.h[o + sB] : 0.0);
      if (__double_as_longlong(un) != __double_as_longlong(u)) P.x[o] = un;
    }
  }
}

int need_cs(mom6cu_ctx* c, const char* who) {
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "%s: mom6cu_set_grid / mom6cu_set_vgrid have not been called", who);
  if (!c->have_vv_cs) return c->fail(MOM6CU_ERR_BAD_ARG, "MOM_vert_friction(%s): Module must be initialized before it is used.", who);
  if (c->g.nk > KMAX) return c->fail(MOM6CU_ERR_UNSUPPORTED, "%s: %d layers exceed the %d-layer column capacity", who, c->g.nk, KMAX);
  return 0;
}

struct Coefs { double *a_u, *a_v, *h_u, *h_v; };
int coef_planes(mom6cu_ctx* c, Coefs* K) {
  K->a_u = c->plane3k("vv.a_u", c->g.nk + 1); K->a_v = c->plane3k("vv.a_v", c->g.nk + 1);
  K->h_u = c->plane3("vv.h_u"); K->h_v = c->plane3("vv.h_v");
  return (K->a_u && K->a_v && K->h_u && K->h_v) ? 0 : MOM6CU_ERR_CUDA;
}

}  // namespace

extern "C" int mom6cu_set_cs_vertvisc(mom6cu_ctx* c, const mom6cu_vertvisc_cs* CS) {
  if (!c || !CS) return MOM6CU_ERR_BAD_ARG;
  if (CS->unsupported) return c->fail(MOM6CU_ERR_UNSUPPORTED, "vertvisc: the host configuration uses an option outside the frozen set (GL90, ice shelves, OBCs, Stokes mixing)");
  if (CS->dynamic_viscous_ML) return c->fail(MOM6CU_ERR_UNSUPPORTED, "vertvisc: DYNAMIC_VISCOUS_ML is not implemented");
  if (CS->nkml > 0) return c->fail(MOM6CU_ERR_UNSUPPORTED, "vertvisc: a bulk mixed layer (GV%%nkml = %d) is not implemented", CS->nkml);
  if (CS->answer_date < 20190101) return c->fail(MOM6CU_ERR_UNSUPPORTED, "vertvisc: answer_date %d < 20190101 is not implemented", CS->answer_date);
  c->vv_cs = *CS;
  c->have_vv_cs = true;
  return 0;
}

int m6_vertvisc_coef_run(mom6cu_ctx* c, const VvCoefDev& D) {
  int rc;
  if ((rc = need_cs(c, "coef"))) return rc;
  if (!c->vgrid.Boussinesq) return c->fail(MOM6CU_ERR_UNSUPPORTED, "vertvisc_coef: non-Boussinesq thickness_to_dz / find_ustar are not implemented");
  const mom6cu_vertvisc_cs& CS = c->vv_cs;
  const bool lotw = CS.fixed_LOTW_ML || CS.apply_LOTW_floor, surf = lotw || CS.Kvml_invZ2 > 0.;
  if (CS.bottomdraglaw && (!D.Kv_bbl_u || !D.Kv_bbl_v || !D.bbl_thick_u || !D.bbl_thick_v))
    return c->fail(MOM6CU_ERR_BAD_ARG, "vertvisc_coef: BOTTOMDRAGLAW needs visc%%Kv_bbl_[uv] and visc%%bbl_thick_[uv]");
  if (lotw && !D.ustar) return c->fail(MOM6CU_ERR_BAD_ARG, "vertvisc_coef: the law-of-the-wall options need forces%%ustar");
  const Geom& G = c->g;
  Coefs K;
  if ((rc = coef_planes(c, &K))) return rc;
  double *sc1 = nullptr, *sc2 = nullptr;
  if (surf && (!(sc1 = c->plane3("vv.dzvel")) || !(sc2 = c->plane3("vv.dzharm")))) return MOM6CU_ERR_CUDA;
  const mom6cu_domain& d = c->dom;
  CoefK U = {}, V = {};
  U.CS = CS; U.nz = G.nk; U.h_neglect = c->vgrid.H_subroundoff; U.H_to_Z = c->vgrid.H_to_Z; U.Z_to_H = c->vgrid.Z_to_H;
  U.a_cpl_max = 1.0e37 * c->vgrid.m_to_H * c->US.T_to_s;
  U.bathyT = c->grid.bathyT; U.CoriolisBu = c->grid.CoriolisBu; U.h = D.h; U.Kv_shear = D.Kv_shear; U.Kv_shear_Bu = D.Kv_shear_Bu; U.ustar = D.ustar;
  U.dzvel = sc1; U.dzharm = sc2;
  V = U;
  U.mask = c->grid.mask2dCu; U.vel = D.u; U.kv_bbl = D.Kv_bbl_u; U.bbl_thick = D.bbl_thick_u; U.a_out = K.a_u; U.h_out = K.h_u;
  U.i0 = d.isc - 1; U.i1 = d.iec; U.j0 = d.jsc; U.j1 = d.jec;
  V.mask = c->grid.mask2dCv; V.vel = D.v; V.kv_bbl = D.Kv_bbl_v; V.bbl_thick = D.bbl_thick_v; V.a_out = K.a_v; V.h_out = K.h_v;
  V.i0 = d.isc; V.i1 = d.iec; V.j0 = d.jsc - 1; V.j1 = d.jec;
  M6_LAUNCH(c, vv_coef_kernel<0>, dim3((U.i1 - U.i0 + 128) / 128, U.j1 - U.j0 + 1), 128, 0, G, U);
  M6_LAUNCH(c, vv_coef_kernel<1>, dim3((V.i1 - V.i0 + 128) / 128, V.j1 - V.j0 + 1), 128, 0, G, V);
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

extern "C" int mom6cu_vertvisc_coef(mom6cu_ctx* c, const mom6cu_vertvisc_coef_args* a) {
  if (!c || !a) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = need_cs(c, "coef"))) return rc;
  if (!a->u || !a->v || !a->h) return c->fail(MOM6CU_ERR_BAD_ARG, "vertvisc_coef: null required argument");
  const Geom& G = c->g;
  Stager S(c, "vvc.");
  VvCoefDev D = {};
  D.dt = a->dt;
  if ((rc = S.in3(a->u, ST_U, "u", &D.u)) || (rc = S.in3(a->v, ST_V, "v", &D.v)) || (rc = S.in3(a->h, ST_H, "h", &D.h)) ||
      (rc = S.in2(a->Kv_bbl_u, ST_U, "kbu", &D.Kv_bbl_u)) || (rc = S.in2(a->Kv_bbl_v, ST_V, "kbv", &D.Kv_bbl_v)) ||
      (rc = S.in2(a->bbl_thick_u, ST_U, "btu", &D.bbl_thick_u)) || (rc = S.in2(a->bbl_thick_v, ST_V, "btv", &D.bbl_thick_v)) ||
      (rc = S.in(a->Kv_shear, ST_H, 0, G.nk + 1, "kvs", &D.Kv_shear)) || (rc = S.in(a->Kv_shear_Bu, ST_Q, 0, G.nk + 1, "kvq", &D.Kv_shear_Bu)) ||
      (rc = S.in2(a->ustar, ST_H, "ustar", &D.ustar))) return rc;
  if ((rc = S.begin())) return rc;
  if ((rc = m6_vertvisc_coef_run(c, D))) return rc;
  return S.finish();
}

extern "C" int mom6cu_vertvisc_get_coef(mom6cu_ctx* c, double* a_u, double* a_v, double* h_u, double* h_v) {
  if (!c) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  Coefs K;
  int rc;
  if ((rc = coef_planes(c, &K))) return rc;
  const int nk = c->g.nk;
  if (a_u && (rc = m6_down(c, K.a_u, ST_U, 0, nk + 1, a_u))) return rc;
  if (a_v && (rc = m6_down(c, K.a_v, ST_V, 0, nk + 1, a_v))) return rc;
  if (h_u && (rc = m6_down(c, K.h_u, ST_U, 0, nk, h_u))) return rc;
  if (h_v && (rc = m6_down(c, K.h_v, ST_V, 0, nk, h_v))) return rc;
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int m6_vertvisc_run(mom6cu_ctx* c, const VvDev& D) {
  int rc;
  if ((rc = need_cs(c, "visc"))) return rc;
  const mom6cu_vertvisc_cs& CS = c->vv_cs;
  if (CS.direct_stress && !D.h) return c->fail(MOM6CU_ERR_BAD_ARG, "vertvisc: DIRECT_STRESS needs h");
  const Geom& G = c->g;
  Coefs K;
  if ((rc = coef_planes(c, &K))) return rc;
  const mom6cu_domain& d = c->dom;
  SolveK U = {};
  U.nz = G.nk; U.direct_stress = CS.direct_stress; U.dt = D.dt; U.dt_Rho0 = D.dt / c->vgrid.H_to_RZ; U.h_neglect = c->vgrid.H_subroundoff;
  U.Hmix = CS.Hmix_stress; U.I_Hmix = CS.direct_stress ? 1.0 / CS.Hmix_stress : 0.0; U.H_to_RZ = c->vgrid.H_to_RZ; U.h = D.h;
  U.vel_underflow = CS.vel_underflow; U.CFL_trunc = CS.CFL_trunc; U.maxvel = CS.maxvel; U.cfl_based = CS.CFL_based_trunc;
  U.lim_on = (CS.vel_underflow > 0.0) || (CS.CFL_based_trunc ? CS.CFL_trunc > 0.0 : CS.maxvel > 0.0);
  U.H_report = 6.0 * c->vgrid.Angstrom_H; U.areaT = c->grid.areaT; U.IareaT = c->grid.IareaT;
  U.ntrunc = reinterpret_cast<unsigned long long*>(c->buf("vv.ntrunc", 1));
  if (!U.ntrunc) return MOM6CU_ERR_CUDA;
  SolveK V = U;
  U.mask = c->grid.mask2dCu; U.a = K.a_u; U.hh = K.h_u; U.Ray = D.Ray_u; U.tau = D.taux; U.x = D.u; U.tau_bot = D.taux_bot;
  U.i0 = d.isc - 1; U.i1 = d.iec; U.j0 = std::min(d.jsc, d.isc); U.j1 = d.jec; U.js_stress = d.jsc; U.js_solve = d.isc;  // `do j=G%isc,G%jec` (:778)
  U.js_lim = d.jsc; U.face = c->grid.dy_Cu;
  V.mask = c->grid.mask2dCv; V.a = K.a_v; V.hh = K.h_v; V.Ray = D.Ray_v; V.tau = D.tauy; V.x = D.v; V.tau_bot = D.tauy_bot;
  V.i0 = d.isc; V.i1 = d.iec; V.j0 = d.jsc - 1; V.j1 = d.jec; V.js_stress = V.j0; V.js_solve = V.j0;
  V.js_lim = V.j0; V.face = c->grid.dx_Cv;
  M6_LAUNCH(c, (vv_solve_kernel<0, false>), dim3((U.i1 - U.i0 + 128) / 128, U.j1 - U.j0 + 1), 128, 0, G, U);
  M6_LAUNCH(c, (vv_solve_kernel<1, false>), dim3((V.i1 - V.i0 + 128) / 128, V.j1 - V.j0 + 1), 128, 0, G, V);
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

extern "C" long long mom6cu_vertvisc_ntrunc(mom6cu_ctx* c) {
  if (!c) return -1;
  const double* p = c->buf("vv.ntrunc", 1);
  unsigned long long n = 0;
  if (!p || cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess ||
      cudaMemcpy(&n, p, sizeof(n), cudaMemcpyDeviceToHost) != cudaSuccess) {
    c->fail(MOM6CU_ERR_CUDA, "vertvisc_ntrunc: the truncation counter could not be read");
    return -1;
  }
  return (long long)n;
}

extern "C" int mom6cu_vertvisc(mom6cu_ctx* c, const mom6cu_vertvisc_args* a) {
  if (!c || !a) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = need_cs(c, "visc"))) return rc;
  if (!a->u || !a->v || !a->taux || !a->tauy) return c->fail(MOM6CU_ERR_BAD_ARG, "vertvisc: null required argument");
  Stager S(c, "vv.");
  VvDev D = {};
  D.dt = a->dt;
  if ((rc = S.io3(a->u, ST_U, "u", &D.u)) || (rc = S.io3(a->v, ST_V, "v", &D.v)) || (rc = S.in3(a->h, ST_H, "h", &D.h)) ||
      (rc = S.in2(a->taux, ST_U, "taux", &D.taux)) || (rc = S.in2(a->tauy, ST_V, "tauy", &D.tauy)) ||
      (rc = S.in3(a->Ray_u, ST_U, "Ray_u", &D.Ray_u)) || (rc = S.in3(a->Ray_v, ST_V, "Ray_v", &D.Ray_v))) return rc;
  if (a->taux_bot && (rc = S.io2(a->taux_bot, ST_U, "taux_bot", &D.taux_bot))) return rc;
  if (a->tauy_bot && (rc = S.io2(a->tauy_bot, ST_V, "tauy_bot", &D.tauy_bot))) return rc;
  if ((rc = S.begin())) return rc;
  if ((rc = m6_vertvisc_run(c, D))) return rc;
  return S.finish();
}

int m6_vertvisc_remnant_run(mom6cu_ctx* c, const double* Ray_u, const double* Ray_v, double* visc_rem_u, double* visc_rem_v, double dt) {
  int rc;
  if ((rc = need_cs(c, "remant"))) return rc;
  const Geom& G = c->g;
  Coefs K;
  if ((rc = coef_planes(c, &K))) return rc;
  const mom6cu_domain& d = c->dom;
  SolveK U = {};
  U.nz = G.nk; U.dt = dt;
  SolveK V = U;
  U.mask = c->grid.mask2dCu; U.a = K.a_u; U.hh = K.h_u; U.Ray = Ray_u; U.x = visc_rem_u;
  U.i0 = d.isc - 1; U.i1 = d.iec; U.j0 = d.jsc; U.j1 = d.jec; U.js_stress = U.j0; U.js_solve = U.j0;
  V.mask = c->grid.mask2dCv; V.a = K.a_v; V.hh = K.h_v; V.Ray = Ray_v; V.x = visc_rem_v;
  V.i0 = d.isc; V.i1 = d.iec; V.j0 = d.jsc - 1; V.j1 = d.jec; V.js_stress = V.j0; V.js_solve = V.j0;
  M6_LAUNCH(c, (vv_solve_kernel<0, true>), dim3((U.i1 - U.i0 + 128) / 128, U.j1 - U.j0 + 1), 128, 0, G, U);
  M6_LAUNCH(c, (vv_solve_kernel<1, true>), dim3((V.i1 - V.i0 + 128) / 128, V.j1 - V.j0 + 1), 128, 0, G, V);
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

extern "C" int mom6cu_vertvisc_remnant(mom6cu_ctx* c, const double* Ray_u, const double* Ray_v, double* visc_rem_u, double* visc_rem_v, double dt) {
  if (!c || !visc_rem_u || !visc_rem_v) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = need_cs(c, "remant"))) return rc;
  Stager S(c, "vvr.");
  double *d_ru_out, *d_rv_out;
  const double *d_ru, *d_rv;
  if ((rc = S.io3(visc_rem_u, ST_U, "vru", &d_ru_out)) || (rc = S.io3(visc_rem_v, ST_V, "vrv", &d_rv_out)) ||
      (rc = S.in3(Ray_u, ST_U, "Ray_u", &d_ru)) || (rc = S.in3(Ray_v, ST_V, "Ray_v", &d_rv))) return rc;
  if ((rc = S.begin())) return rc;
  if ((rc = m6_vertvisc_remnant_run(c, d_ru, d_rv, d_ru_out, d_rv_out, dt))) return rc;
  return S.finish();
}
