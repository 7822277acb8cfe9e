// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of /root/reference/src/parameterizations/vertical/MOM_vert_friction.F90: vertvisc :557-1226,
// vertvisc_remnant :1229-1354, vertvisc_coef :1357-2310, find_coupling_coef :2314-2924, for the frozen option set of
// include/mom6cu.h (no OBCs, shelves, GL90, Stokes; answer_date >= 20190101; Boussinesq, so thickness_to_dz gives
// dz = GV%H_to_Z*h [MOM_interface_heights.F90:939] and find_ustar returns forces%ustar [MOM_forcing_type.F90:1264-1268]).
// The reference's row bound `do j=G%isc,G%jec` in the u solver (:778 etc.) is kept as written.
// PARITY: PINNED BY A REFERENCE RUN -- the reference's own MOM_vert_friction.F90, executed by oracle/f90run, agrees bit for bit on 11
// option sets + 4 cases that truncate velocities (vertvisc_limit_vel :2926-3120) (tests/test_reference_f90.py, vertvisc_family/*).
#include "oracle.h"
#include "ogrid.hpp"
#include <cfloat>
#include <cmath>
#include <vector>
#include <algorithm>

using namespace orc;

namespace {

inline double max3(double a, double b, double c) { return fmax2(fmax2(a, b), c); }

// One velocity column of vertvisc_coef + find_coupling_coef.  d = 0: u-point (I,j) between cells (i,j) and (i+1,j);
// d = 1: v-point (i,J) between (i,j) and (i,j+1).  All arrays 1-based in k.
struct ColIn {
  int nz;
  const double *hA, *hB;    // h of the two cells (1..nz)
  const double* vel;        // u or v of the column
  double DA, DB;            // bathyT of the two cells
  double kv_bbl, bbl_thick_in;
  const double *KvsA, *KvsB;  // Kv_shear of the two cells (interfaces 1..nz+1) or null
  const double *KbuA, *KbuB;  // Kv_shear_Bu of the two vertices or null
  double ustarA, ustarB, absf;
};

void coef_column(const mom6cu_vertvisc_cs* CS, const mom6cu_vgrid* GV, double a_cpl_max, const ColIn& c, double* a_out /*1..nz+1*/,
                 double* h_out /*1..nz*/) {
  const int nz = c.nz;
  const double h_neglect = GV->H_subroundoff, dz_neglect = CS->dZ_subroundoff;
  std::vector<double> hvel(nz + 2), dz_vel(nz + 2), dz_harm(nz + 2), z_i(nz + 3), a_cpl(nz + 3, 0.0);
  double I_Hbbl = 1. / (CS->Hbbl + dz_neglect);
  const double I_valBL = (CS->harm_BL_val > 0.0) ? 1.0 / CS->harm_BL_val : 0.0;
  double kv_bbl = 0., bbl_thick = 0.;
  if (CS->bottomdraglaw) { kv_bbl = c.kv_bbl; bbl_thick = c.bbl_thick_in + dz_neglect; I_Hbbl = 1. / bbl_thick; }
  const double Dmin = fmin2(c.DA, c.DB);
  z_i[nz + 1] = 0.;
  double zh = 0., zcolA = -c.DA, zcolB = -c.DB;
  for (int k = nz; k >= 1; --k) {
    const double hA = c.hA[k], hB = c.hB[k];
    const double dzA = GV->H_to_Z * hA, dzB = GV->H_to_Z * hB;
    const double h_harm = 2. * hA * hB / (hA + hB + h_neglect);
    const double h_arith = 0.5 * (hB + hA);
    const double h_delta = hB - hA;
    dz_harm[k] = 2. * dzA * dzB / (dzA + dzB + dz_neglect);
    const double dz_arith = 0.5 * (dzB + dzA);
    if (CS->harmonic_visc) {
      hvel[k] = h_harm; dz_vel[k] = dz_harm[k];
      if (c.vel[k] * h_delta < 0) {
        const double z2 = z_i[k + 1];
        const double botfn = 1. / (1. + 0.09 * z2 * z2 * z2 * z2 * z2 * z2);
        hvel[k] = (1. - botfn) * h_harm + botfn * h_arith;
        dz_vel[k] = (1. - botfn) * dz_harm[k] + botfn * dz_arith;
      }
      z_i[k] = z_i[k + 1] + dz_harm[k] * I_Hbbl;
    } else {
      zcolA = zcolA + dzA; zcolB = zcolB + dzB;
      zh = zh + dz_harm[k];
      const double z_clear = fmax2(zcolA, zcolB) + Dmin;
      z_i[k] = fmax2(zh, z_clear) * I_Hbbl;
      hvel[k] = h_arith; dz_vel[k] = dz_arith;
      if (c.vel[k] * h_delta > 0.) {
        if (zh * I_Hbbl < CS->harm_BL_val) { hvel[k] = h_harm; dz_vel[k] = dz_harm[k]; }
        else {
          double z2_wt = 1.;
          if (zh * I_Hbbl < 2. * CS->harm_BL_val) z2_wt = fmax2(0., fmin2(1., zh * I_Hbbl * I_valBL - 1.));
          const double z2 = z2_wt * (fmax2(zh, z_clear) * I_Hbbl);
          const double botfn = 1. / (1. + 0.09 * z2 * z2 * z2 * z2 * z2 * z2);
          hvel[k] = (1. - botfn) * h_arith + botfn * h_harm;
          dz_vel[k] = (1. - botfn) * dz_arith + botfn * dz_harm[k];
        }
      }
    }
  }
  // find_coupling_coef(a_cpl, dz_vel, do_i, dz_harm, bbl_thick, kv_bbl, z_i, h_ml, ...) :2314
  const double* hv = dz_vel.data();   // the routine's "hvel" argument is dz_vel
  const double hn = CS->dZ_subroundoff;  // its h_neglect is GV%dZ_subroundoff (:2391)
  const double I_amax = 0.0;
  double z_t = 0., I_Hmix = 0.;
  if (CS->Kvml_invZ2 > 0.) { I_Hmix = 1. / (CS->Hmix + hn); z_t = hn * I_Hmix; }
  for (int K = 2; K <= nz; ++K) {
    double Kv_tot = CS->Kv;
    if (CS->Kvml_invZ2 > 0.) {
      z_t = z_t + dz_harm[K - 1] * I_Hmix;
      Kv_tot = CS->Kv + CS->Kvml_invZ2 / ((z_t * z_t) * (1. + 0.09 * z_t * z_t * z_t * z_t * z_t * z_t));
    }
    if (c.KvsA) { const double Kv_add = 0.5 * (c.KvsA[K] + c.KvsB[K]); Kv_tot = Kv_tot + Kv_add; }
    if (c.KbuA) Kv_tot = Kv_tot + 0.5 * (c.KbuA[K] + c.KbuB[K]);
    if (CS->bottomdraglaw) {
      const double z2 = z_i[K];
      const double botfn = 1. / (1. + 0.09 * z2 * z2 * z2 * z2 * z2 * z2);
      Kv_tot = Kv_tot + (kv_bbl - CS->Kv) * botfn;
      const double dhc = 0.5 * (hv[K] + hv[K - 1]);
      double h_shear;
      if (dhc > bbl_thick) h_shear = ((1. - botfn) * dhc + botfn * bbl_thick) + hn;
      else h_shear = dhc + hn;
      a_cpl[K] = Kv_tot / (h_shear + (I_amax * Kv_tot));
    } else if (std::fabs(CS->Kv_extra_bbl) > 0.0) {
      const double z2 = z_i[K];
      const double botfn = 1. / (1. + 0.09 * z2 * z2 * z2 * z2 * z2 * z2);
      Kv_tot = Kv_tot + CS->Kv_extra_bbl * botfn;
      const double h_shear = 0.5 * (hv[K] + hv[K - 1] + hn);
      a_cpl[K] = Kv_tot / (h_shear + I_amax * Kv_tot);
    } else {
      const double h_shear = 0.5 * (hv[K] + hv[K - 1] + hn);
      a_cpl[K] = Kv_tot / (h_shear + I_amax * Kv_tot);
    }
  }
  if (CS->bottomdraglaw) {
    const double dhc = hv[nz] * 0.5;
    a_cpl[nz + 1] = kv_bbl / ((fmin2(dhc, bbl_thick) + hn) + I_amax * kv_bbl);
  } else if (std::fabs(CS->Kv_extra_bbl) > 0.0) {
    a_cpl[nz + 1] = (CS->Kv + CS->Kv_extra_bbl) / ((0.5 * hv[nz] + hn) + I_amax * (CS->Kv + CS->Kv_extra_bbl));
  } else a_cpl[nz + 1] = CS->Kv / ((0.5 * hv[nz] + hn) + I_amax * CS->Kv);
  if (CS->fixed_LOTW_ML || CS->apply_LOTW_floor) {  // :2643 with dynamic_viscous_ML = .false., GV%nkml = 0
    const double u_star = 0.5 * (c.ustarA + c.ustarB);
    double h_ml = 0.0;
    int nk_in_ml = 0;
    for (int k = 1; k <= nz; ++k) {
      if (h_ml < CS->Hmix) {
        nk_in_ml = k;
        if (h_ml + hv[k] < CS->Hmix) h_ml = h_ml + hv[k];
        else h_ml = CS->Hmix;
      }
    }
    if (u_star <= 0.0) nk_in_ml = 0;
    z_t = 0.0;
    for (int K = 2; K <= nk_in_ml; ++K) {
      z_t = z_t + hv[K - 1];
      const double temp1 = (z_t * h_ml - z_t * z_t);
      if (CS->apply_LOTW_floor && CS->fixed_LOTW_ML) {
        const double ustar2_denom = (CS->vonKar * GV->Z_to_H * (u_star * u_star)) / (c.absf * temp1 + (h_ml + hn) * u_star);
        const double visc_ml = temp1 * ustar2_denom;
        const double a_ml = visc_ml / (0.25 * (hv[K] + hv[K - 1] + hn));
        const double a_floor = (h_ml - z_t) * ustar2_denom;
        a_cpl[K] = max3(a_cpl[K], a_ml, a_floor);
      } else if (CS->apply_LOTW_floor) {
        const double ustar2_denom = (CS->vonKar * GV->Z_to_H * (u_star * u_star)) / (c.absf * temp1 + (h_ml + hn) * u_star);
        a_cpl[K] = fmax2(a_cpl[K], (h_ml - z_t) * ustar2_denom);
      } else {
        const double visc_ml = u_star * CS->vonKar * (GV->Z_to_H * temp1 * u_star) / (c.absf * temp1 + (h_ml + hn) * u_star);
        const double a_ml = visc_ml / (0.25 * (hv[K] + hv[K - 1] + hn) + 0.5 * I_amax * visc_ml);
        a_cpl[K] = fmax2(a_cpl[K], a_ml);
      }
    }
  }
  for (int K = 1; K <= nz + 1; ++K) a_out[K] = fmin2(a_cpl_max, a_cpl[K]);
  for (int k = 1; k <= nz; ++k) h_out[k] = hvel[k] + h_neglect;
}

}  // namespace

extern "C" int oracle_vertvisc_coef(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV, const mom6cu_unit_scale* US,
                                    const mom6cu_vertvisc_cs* CS, const mom6cu_vertvisc_coef_args* a, double* a_up, double* a_vp,
                                    double* h_up, double* h_vp) {
  const OGrid G(d, Gp);
  const int nz = G.ke, is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  if (CS->unsupported || CS->dynamic_viscous_ML || CS->nkml > 0 || CS->answer_date < 20190101) return 3;
  const V3 u = G.U3(a->u), v = G.V3_(a->v), h = G.H3(a->h);
  const V3 a_u = G.U3(a_up, nz + 1), a_v = G.V3_(a_vp, nz + 1), h_u = G.U3(h_up), h_v = G.V3_(h_vp);
  const V3 Kvs = a->Kv_shear ? G.H3(a->Kv_shear, nz + 1) : V3(), Kbu = a->Kv_shear_Bu ? G.Q3(a->Kv_shear_Bu, nz + 1) : V3();
  const V2 kbu = a->Kv_bbl_u ? G.U(a->Kv_bbl_u) : V2(), kbv = a->Kv_bbl_v ? G.V(a->Kv_bbl_v) : V2();
  const V2 btu = a->bbl_thick_u ? G.U(a->bbl_thick_u) : V2(), btv = a->bbl_thick_v ? G.V(a->bbl_thick_v) : V2();
  const V2 ust = a->ustar ? G.H(a->ustar) : V2();
  const bool lotw = CS->fixed_LOTW_ML || CS->apply_LOTW_floor;
  if (CS->bottomdraglaw && (!kbu.p || !kbv.p || !btu.p || !btv.p)) return 2;
  if (lotw && !ust.p) return 2;
  const double a_cpl_max = 1.0e37 * GV->m_to_H * US->T_to_s;
  std::vector<double> hA(nz + 2), hB(nz + 2), vel(nz + 2), kA(nz + 3), kB(nz + 3), qA(nz + 3), qB(nz + 3), ao(nz + 3), ho(nz + 2);
  for (int dir = 0; dir < 2; ++dir) {
    const int i0 = dir ? is : Isq, i1 = dir ? ie : Ieq, j0 = dir ? Jsq : js, j1 = dir ? Jeq : je;
    for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) {
      const bool do_i = (dir ? G.mask2dCv(i, j) : G.mask2dCu(i, j)) > 0.;
      if (!do_i) continue;
      const int ib = dir ? i : i + 1, jb = dir ? j + 1 : j;
      ColIn c;
      c.nz = nz;
      for (int k = 1; k <= nz; ++k) { hA[k] = h(i, j, k); hB[k] = h(ib, jb, k); vel[k] = dir ? v(i, j, k) : u(i, j, k); }
      c.hA = hA.data(); c.hB = hB.data(); c.vel = vel.data();
      c.DA = G.bathyT(i, j); c.DB = G.bathyT(ib, jb);
      c.kv_bbl = CS->bottomdraglaw ? (dir ? kbv(i, j) : kbu(i, j)) : 0.;
      c.bbl_thick_in = CS->bottomdraglaw ? (dir ? btv(i, j) : btu(i, j)) : 0.;
      c.KvsA = c.KvsB = c.KbuA = c.KbuB = nullptr;
      if (Kvs.p) { for (int K = 1; K <= nz + 1; ++K) { kA[K] = Kvs(i, j, K); kB[K] = Kvs(ib, jb, K); } c.KvsA = kA.data(); c.KvsB = kB.data(); }
      if (Kbu.p) {  // u: Bu(I,J-1), Bu(I,J); v: Bu(I-1,J), Bu(I,J)
        for (int K = 1; K <= nz + 1; ++K) { qA[K] = dir ? Kbu(i - 1, j, K) : Kbu(i, j - 1, K); qB[K] = Kbu(i, j, K); }
        c.KbuA = qA.data(); c.KbuB = qB.data();
      }
      c.ustarA = c.ustarB = c.absf = 0.;
      if (lotw) {
        c.ustarA = ust(i, j); c.ustarB = ust(ib, jb);
        c.absf = dir ? 0.5 * (std::fabs(G.CoriolisBu(i - 1, j)) + std::fabs(G.CoriolisBu(i, j)))
                     : 0.5 * (std::fabs(G.CoriolisBu(i, j - 1)) + std::fabs(G.CoriolisBu(i, j)));
      }
      coef_column(CS, GV, a_cpl_max, c, ao.data(), ho.data());
      const V3& A = dir ? a_v : a_u; const V3& H = dir ? h_v : h_u;
      for (int K = 1; K <= nz + 1; ++K) A(i, j, K) = ao[K];
      for (int k = 1; k <= nz; ++k) H(i, j, k) = ho[k];
    }
  }
  return 0;
}

namespace {
// the Schopf & Loughe tridiagonal sweep shared by vertvisc (:772-803, :1010-1045) and vertvisc_remnant (:1270-1300):
// x(k) is u (rem = false) or the remnant (rem = true, right-hand side h_u itself)
void solve_column(int nz, double dt, const double* a, const double* hh, const double* Ray, double surface_stress, bool rem, double* x) {
  std::vector<double> c1(nz + 2);
  double b_denom_1 = hh[1] + dt * ((Ray ? Ray[1] : 0.) + a[1]);
  double b1 = 1.0 / (b_denom_1 + dt * a[2]);
  double d1 = b_denom_1 * b1;
  x[1] = rem ? b1 * hh[1] : b1 * (hh[1] * x[1] + surface_stress);
  for (int k = 2; k <= nz; ++k) {
    c1[k] = dt * a[k] * b1;
    b_denom_1 = hh[k] + dt * ((Ray ? Ray[k] : 0.) + a[k] * d1);
    b1 = 1.0 / (b_denom_1 + dt * a[k + 1]);
    d1 = b_denom_1 * b1;
    x[k] = rem ? (hh[k] + dt * a[k] * x[k - 1]) * b1 : (hh[k] * x[k] + dt * a[k] * x[k - 1]) * b1;
  }
  for (int k = nz - 1; k >= 1; --k) x[k] = x[k] + c1[k + 1] * x[k + 1];
}
}  // namespace

static long long g_ntrunc = 0;

extern "C" int oracle_vertvisc(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV, const mom6cu_vertvisc_cs* CS,
                               const mom6cu_vertvisc_args* a, const double* a_up, const double* a_vp, const double* h_up, const double* h_vp) {
  const OGrid G(d, Gp);
  const int nz = G.ke, is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  if (CS->unsupported) return 3;
  const V3 u = G.U3(a->u), v = G.V3_(a->v);
  const V3 h = a->h ? G.H3(a->h) : V3();
  const V3 a_u = G.U3(a_up, nz + 1), a_v = G.V3_(a_vp, nz + 1), h_u = G.U3(h_up), h_v = G.V3_(h_vp);
  const V3 Ru = a->Ray_u ? G.U3(a->Ray_u) : V3(), Rv = a->Ray_v ? G.V3_(a->Ray_v) : V3();
  const V2 taux = G.U(a->taux), tauy = G.V(a->tauy);
  const double dt = a->dt, dt_Rho0 = dt / GV->H_to_RZ, h_neglect = GV->H_subroundoff;
  double Hmix = 0., I_Hmix = 0.;
  if (CS->direct_stress) { if (!h.p) return 2; Hmix = CS->Hmix_stress; I_Hmix = 1.0 / Hmix; }
  std::vector<double> aa(nz + 3), hh(nz + 2), rr(nz + 2), x(nz + 2);
  for (int dir = 0; dir < 2; ++dir) {
    const int i0 = dir ? is : Isq, i1 = dir ? ie : Ieq, j0 = dir ? Jsq : js, j1 = dir ? Jeq : je;
    const V3& vel = dir ? v : u; const V3& A = dir ? a_v : a_u; const V3& H = dir ? h_v : h_u; const V3& R = dir ? Rv : Ru;
    for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) {
      const double mask = dir ? G.mask2dCv(i, j) : G.mask2dCu(i, j);
      const double tau = dir ? tauy(i, j) : taux(i, j);
      double surface_stress;
      if (CS->direct_stress) {
        surface_stress = 0.0;
        if (mask > 0.) {
          double zDS = 0.0;
          const double stress = dt_Rho0 * tau;
          const int ib = dir ? i : i + 1, jb = dir ? j + 1 : j;
          for (int k = 1; k <= nz; ++k) {
            const double h_a = 0.5 * (h(i, j, k) + h(ib, jb, k)) + h_neglect;
            double hfr = 1.0; if ((zDS + h_a) > Hmix) hfr = (Hmix - zDS) / h_a;
            vel(i, j, k) = vel(i, j, k) + I_Hmix * hfr * stress;
            zDS = zDS + h_a; if (zDS >= Hmix) break;
          }
        }
      } else surface_stress = dt_Rho0 * (mask * tau);
      // the u solver runs over j = G%isc..G%jec (:778), the v solver over J = Jsq..Jeq
      const bool in_rows = dir ? true : (j >= is);
      if (mask > 0. && in_rows) {
        for (int k = 1; k <= nz; ++k) { hh[k] = H(i, j, k); rr[k] = R.p ? R(i, j, k) : 0.; x[k] = vel(i, j, k); }
        for (int K = 1; K <= nz + 1; ++K) aa[K] = A(i, j, K);
        solve_column(nz, dt, aa.data(), hh.data(), R.p ? rr.data() : nullptr, surface_stress, false, x.data());
        for (int k = 1; k <= nz; ++k) vel(i, j, k) = x[k];
      }
    }
    // the u solver's row range also reaches rows isc..jsc-1 when isc < jsc
    if (!dir && is < js) {
      for (int j = is; j < js; ++j) for (int i = i0; i <= i1; ++i) if (G.mask2dCu(i, j) > 0.) {
        for (int k = 1; k <= nz; ++k) { hh[k] = H(i, j, k); rr[k] = R.p ? R(i, j, k) : 0.; x[k] = vel(i, j, k); }
        for (int K = 1; K <= nz + 1; ++K) aa[K] = A(i, j, K);
        // surface_stress(I,j) is not set on these rows in the reference (uninitialised); 0 is used here
        solve_column(nz, dt, aa.data(), hh.data(), R.p ? rr.data() : nullptr, 0.0, false, x.data());
        for (int k = 1; k <= nz; ++k) vel(i, j, k) = x[k];
      }
    }
    double* tb = dir ? a->tauy_bot : a->taux_bot;
    if (tb) {
      const V2 T = dir ? G.V(tb) : G.U(tb);
      for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) {
        T(i, j) = GV->H_to_RZ * (vel(i, j, nz) * A(i, j, nz + 1));
        if (R.p) for (int k = 1; k <= nz; ++k) T(i, j) = T(i, j) + GV->H_to_RZ * (R(i, j, k) * vel(i, j, k));
      }
    }
  }
  // vertvisc_limit_vel :2926-3120 without truncation files (the reporting branch truncates the same velocities)
  const bool lim_on = (CS->vel_underflow > 0.0) || (CS->CFL_based_trunc ? CS->CFL_trunc > 0.0 : CS->maxvel > 0.0);
  if (lim_on) {
    const double maxvel = CS->maxvel, truncvel = 0.9 * maxvel, H_report = 6.0 * GV->Angstrom_H;
    for (int k = 1; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) {
      bool trunc = false;
      if (std::fabs(u(I, j, k)) < CS->vel_underflow) u(I, j, k) = 0.0;
      else if (CS->CFL_based_trunc) {
        if (CS->CFL_trunc > 0.0) {
          if ((u(I, j, k) * (dt * G.dy_Cu(I, j))) * G.IareaT(I + 1, j) < -CS->CFL_trunc) {
            u(I, j, k) = (-0.9 * CS->CFL_trunc) * (G.areaT(I + 1, j) / (dt * G.dy_Cu(I, j))); trunc = true;
          } else if ((u(I, j, k) * (dt * G.dy_Cu(I, j))) * G.IareaT(I, j) > CS->CFL_trunc) {
            u(I, j, k) = (0.9 * CS->CFL_trunc) * (G.areaT(I, j) / (dt * G.dy_Cu(I, j))); trunc = true;
          }
        }
      } else if (maxvel > 0.0 && std::fabs(u(I, j, k)) > maxvel) { u(I, j, k) = std::copysign(truncvel, u(I, j, k)); trunc = true; }
      if (trunc && h.p && (h(I, j, k) + h(I + 1, j, k) > H_report)) ++g_ntrunc;
    }
    for (int k = 1; k <= nz; ++k) for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) {
      bool trunc = false;
      if (std::fabs(v(i, J, k)) < CS->vel_underflow) v(i, J, k) = 0.0;
      else if (CS->CFL_based_trunc) {
        if (CS->CFL_trunc > 0.0) {
          if ((v(i, J, k) * (dt * G.dx_Cv(i, J))) * G.IareaT(i, J + 1) < -CS->CFL_trunc) {
            v(i, J, k) = (-0.9 * CS->CFL_trunc) * (G.areaT(i, J + 1) / (dt * G.dx_Cv(i, J))); trunc = true;
          } else if ((v(i, J, k) * (dt * G.dx_Cv(i, J))) * G.IareaT(i, J) > CS->CFL_trunc) {
            v(i, J, k) = (0.9 * CS->CFL_trunc) * (G.areaT(i, J) / (dt * G.dx_Cv(i, J))); trunc = true;
          }
        }
      } else if (maxvel > 0.0 && std::fabs(v(i, J, k)) > maxvel) { v(i, J, k) = std::copysign(truncvel, v(i, J, k)); trunc = true; }
      if (trunc && h.p && (h(i, J, k) + h(i, J + 1, k) > H_report)) ++g_ntrunc;
    }
  }
  return 0;
}

// CS%ntrunc: truncations counted by oracle_vertvisc since the last reset
extern "C" long long oracle_vertvisc_ntrunc(int reset) {
  const long long n = g_ntrunc;
  if (reset) g_ntrunc = 0;
  return n;
}

extern "C" int oracle_vertvisc_remnant(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vertvisc_cs* CS, const double* Ray_u,
                                       const double* Ray_v, double* vru, double* vrv, double dt, const double* a_up, const double* a_vp,
                                       const double* h_up, const double* h_vp) {
  const OGrid G(d, Gp);
  const int nz = G.ke, is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  const V3 a_u = G.U3(a_up, nz + 1), a_v = G.V3_(a_vp, nz + 1), h_u = G.U3(h_up), h_v = G.V3_(h_vp);
  const V3 Ru = Ray_u ? G.U3(Ray_u) : V3(), Rv = Ray_v ? G.V3_(Ray_v) : V3();
  const V3 ru = G.U3(vru), rv = G.V3_(vrv);
  std::vector<double> aa(nz + 3), hh(nz + 2), rr(nz + 2), x(nz + 2);
  for (int dir = 0; dir < 2; ++dir) {
    const int i0 = dir ? is : Isq, i1 = dir ? ie : Ieq, j0 = dir ? Jsq : js, j1 = dir ? Jeq : je;
    const V3& A = dir ? a_v : a_u; const V3& H = dir ? h_v : h_u; const V3& R = dir ? Rv : Ru; const V3& X = dir ? rv : ru;
    for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) if ((dir ? G.mask2dCv(i, j) : G.mask2dCu(i, j)) > 0.) {
      for (int k = 1; k <= nz; ++k) { hh[k] = H(i, j, k); rr[k] = R.p ? R(i, j, k) : 0.; }
      for (int K = 1; K <= nz + 1; ++K) aa[K] = A(i, j, K);
      solve_column(nz, dt, aa.data(), hh.data(), R.p ? rr.data() : nullptr, 0.0, true, x.data());
      for (int k = 1; k <= nz; ++k) X(i, j, k) = x[k];
    }
  }
  return 0;
}
