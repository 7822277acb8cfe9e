// One-column ALE remapping (remapping_core_h, /root/reference/src/ALE/MOM_remapping.F90:234-335, with the OM4-era
// reconstruction functions PCM/PLM/PPM_H4/PPM_IH4 and answer_date >= 20190101), written for one GPU thread per column:
// every loop has a trip count that depends on (n0, n1) only, so the lanes of a warp run the sub-cell walk in lockstep
// and the thread-local work arrays (interleaved by lane in local memory) are read and written with coalesced accesses.
// The expression order follows the reference line by line; with -fmad=false the results are bitwise those of the
// CPU code.  Also compiles as plain C++ (tests/harness) so the column logic can be checked without a GPU.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define M6R_HD __host__ __device__ __forceinline__
#else
#define M6R_HD inline
#endif

namespace m6remap {

M6R_HD double rmin(double a, double b) { return (a < b) ? a : b; }
M6R_HD double rmax(double a, double b) { return (a > b) ? a : b; }
M6R_HD double rsign(double a, double b) { return copysign(a, b); }
M6R_HD double rmin3(double a, double b, double c) { return rmin(rmin(a, b), c); }
M6R_HD double rmax3(double a, double b, double c) { return rmax(rmax(a, b), c); }

enum { SCHEME_PCM = 0, SCHEME_PLM = 2, SCHEME_PPM_H4 = 4, SCHEME_PPM_IH4 = 5 };
enum { INT_PCM = 0, INT_PLM = 1, INT_PPM = 3 };

struct Params {  // mom6cu_remapping_cs, resolved
  int scheme, boundary_extrapolation, force_bounds_in_subcell, force_bounds_in_target, om4;
  double h_neglect, h_neglect_edge;
};

// A strided view of one column: element k (1-based) is p[(k-1)*sk]
struct Col {
  const double* p; long sk;
  M6R_HD double operator()(int k) const { return p[(long)(k - 1) * sk]; }
};
struct ColOut {
  double* p; long sk;
  M6R_HD double& operator()(int k) const { return p[(long)(k - 1) * sk]; }
};

// The sub-cell grid of one (source, target) column pair: intersect_src_tgt_grids, MOM_remapping.F90:642-798.
// Indices fit a byte pair (n0 + n1 + 1 <= 2*KCAP + 1 <= 65535); all arrays 1-based.
template <int KCAP>
struct SubGrid {
  double h0[KCAP + 1], h1[KCAP + 1], h_sub[2 * KCAP + 2], h0_eff[KCAP + 1];
  uint16_t isub_src[2 * KCAP + 3], isrc_start[KCAP + 1], isrc_end[KCAP + 1], isrc_max[KCAP + 1], itgt_start[KCAP + 1], itgt_end[KCAP + 1];
  int n0, n1;
};

template <int KCAP>
M6R_HD void intersect(SubGrid<KCAP>& S) {
  const int n0 = S.n0, n1 = S.n1;
  double h0_supply = S.h0[1], h1_supply = S.h1[1];
  bool src_has_volume = true, tgt_has_volume = true;
  int i0 = 1, i1 = 1, i_start0 = 1, i_start1 = 1, i_max = 1;
  double dh_max = 0., dh0_eff = 0.;
  S.h_sub[1] = 0.; S.isrc_start[1] = 1; S.isrc_end[1] = 1; S.isrc_max[1] = 1; S.isub_src[1] = 1;
  for (int i_sub = 2; i_sub <= n0 + n1 + 1; ++i_sub) {
    const double dh = rmin(h0_supply, h1_supply);
    dh0_eff = dh0_eff + rmin(dh, h0_supply);
    S.isub_src[i_sub] = (uint16_t)i0;
    double hs = dh;
    if (dh >= dh_max) { i_max = i_sub; dh_max = dh; }
    const bool src_step = (h0_supply <= h1_supply && src_has_volume);
    const bool tgt_step = !src_step && (h0_supply >= h1_supply && tgt_has_volume);
    const bool src_tail = !src_step && !tgt_step && src_has_volume;
    const bool tgt_tail = !src_step && !tgt_step && !src_tail && tgt_has_volume;
    if (src_step || src_tail) {
      if (src_step) h1_supply = h1_supply - dh; else hs = h0_supply;
      S.isrc_start[i0] = (uint16_t)i_start0; S.isrc_end[i0] = (uint16_t)i_sub; i_start0 = i_sub + 1;
      S.isrc_max[i0] = (uint16_t)i_max; i_max = i_sub + 1; dh_max = 0.;
      S.h0_eff[i0] = dh0_eff;
      if (i0 < n0) { i0 = i0 + 1; h0_supply = S.h0[i0]; dh0_eff = 0.; }
      else { h0_supply = 0.; src_has_volume = false; }
    } else if (tgt_step || tgt_tail) {
      if (tgt_step) h0_supply = h0_supply - dh; else hs = h1_supply;
      S.itgt_start[i1] = (uint16_t)i_start1; S.itgt_end[i1] = (uint16_t)i_sub; i_start1 = i_sub + 1;
      if (i1 < n1) { i1 = i1 + 1; h1_supply = S.h1[i1]; }
      else { h1_supply = 0.; tgt_has_volume = false; }
    }
    S.h_sub[i_sub] = hs;
  }
}

// The reconstruction of one source column: edge values ppoly_E(:,1:2) and, for PLM, the slope coefficient.
template <int KCAP>
struct Recon {
  double u[KCAP + 1], E1[KCAP + 1], E2[KCAP + 1], c2[KCAP + 1];
};

// PLM_slope_wa, PLM_functions.F90:17-65
M6R_HD double PLM_slope_wa(double h_l, double h_c, double h_r, double h_neglect, double u_l, double u_c, double u_r) {
  const double sigma_r = u_r - u_c, sigma_l = u_c - u_l;
  const double sigma_c = 2.0 * (u_r - u_l) * (h_c / (h_l + 2.0 * h_c + h_r + h_neglect));
  const double u_min = rmin3(u_l, u_c, u_r), u_max = rmax3(u_l, u_c, u_r);
  double s = 0.0;
  if ((sigma_l * sigma_r) > 0.0) s = rsign(rmin(fabs(sigma_c), 2. * rmin(u_c - u_min, u_max - u_c)), sigma_c);
  if (u_c - 0.5 * fabs(s) < u_min || u_c + 0.5 * fabs(s) > u_max) s = s * (1. - DBL_EPSILON);
  if (fabs(s) < 1.E-140) s = 0.;
  return s;
}
// PLM_monotonized_slope :119-155
M6R_HD double PLM_monotonized_slope(double u_l, double u_c, double u_r, double s_l, double s_c, double s_r) {
  const double almost_two = 2. * (1. - DBL_EPSILON);
  const double e_r = u_l + 0.5 * s_l, e_l = u_r - 0.5 * s_r;
  double slp = fabs(s_c);
  double edge = u_c - 0.5 * s_c;
  if ((edge - e_r) * (u_c - edge) < 0.) { edge = 0.5 * (edge + e_r); slp = rmin(slp, fabs(edge - u_c) * almost_two); }
  edge = u_c + 0.5 * s_c;
  if ((edge - u_c) * (e_l - edge) < 0.) { edge = 0.5 * (edge + e_l); slp = rmin(slp, fabs(edge - u_c) * almost_two); }
  return rsign(slp, s_c);
}
// PLM_extrapolate_slope :160-184
M6R_HD double PLM_extrapolate_slope(double h_l, double h_c, double h_neglect, double u_l, double u_c) {
  const double hl = h_l + h_neglect, hc = h_c + h_neglect;
  const double left_edge = (u_l * hc + u_c * hl) / (hl + hc);
  return 2.0 * (u_c - left_edge);
}

// PLM_reconstruction :188-262 (+ PLM_boundary_extrapolation :267-300).  slp / mslp of the three cells around k are
// recomputed from the column instead of being stored.
template <int KCAP>
M6R_HD void PLM_reconstruction(int N, const double* h, Recon<KCAP>& R, double h_neglect, bool extrapolate) {
  const double almost_one = 1. - DBL_EPSILON;
  const double* u = R.u;
  // sliding window of slp(k-1), slp(k), slp(k+1), slp(k+2) and mslp(k), mslp(k+1)
  auto slp_at = [&](int k) -> double {
    return (k >= 2 && k <= N - 1) ? PLM_slope_wa(h[k - 1], h[k], h[k + 1], h_neglect, u[k - 1], u[k], u[k + 1]) : 0.;
  };
  R.E1[1] = u[1]; R.E2[1] = u[1]; R.c2[1] = 0.;
  double s_m = 0., s_c = slp_at(2), s_p = slp_at(3);  // slp(k-1), slp(k), slp(k+1) for k = 2
  double m_c = (N >= 3) ? PLM_monotonized_slope(u[1], u[2], u[3], s_m, s_c, s_p) : 0.;
  for (int k = 2; k <= N - 1; ++k) {
    const double s_pp = slp_at(k + 2);
    const double m_p = (k + 1 <= N - 1) ? PLM_monotonized_slope(u[k], u[k + 1], u[k + 2], s_c, s_p, s_pp) : 0.;
    const double slope = m_c;
    const double u_l = u[k] - 0.5 * slope, u_r = u[k] + 0.5 * slope;
    double c2 = (u_r - u_l);
    const double edge = c2 + u_l;
    const double e_r = u[k + 1] - 0.5 * rsign(m_p, s_p);
    if ((edge - u[k]) * (e_r - edge) < 0.) c2 = c2 * almost_one;
    R.E1[k] = u_l; R.E2[k] = u_r; R.c2[k] = c2;
    s_m = s_c; s_c = s_p; s_p = s_pp; m_c = m_p;
  }
  R.E1[N] = u[N]; R.E2[N] = u[N]; R.c2[N] = 0.;
  if (extrapolate) {
    double slope = -PLM_extrapolate_slope(h[2], h[1], h_neglect, u[2], u[1]);
    R.E1[1] = u[1] - 0.5 * slope; R.E2[1] = u[1] + 0.5 * slope; R.c2[1] = R.E2[1] - R.E1[1];
    slope = PLM_extrapolate_slope(h[N - 1], h[N], h_neglect, u[N - 1], u[N]);
    R.E1[N] = u[N] - 0.5 * slope; R.E2[N] = u[N] + 0.5 * slope; R.c2[N] = R.E2[N] - R.E1[N];
  }
}

// end_value_h4, regrid_edge_values.F90:633-760; dz, u, Csys 1-based of length 4
M6R_HD void end_value_h4(const double* dz, const double* u, double* Csys) {
  const double min_frac = 1.0e-6;
  double h1 = dz[1], h2 = dz[2], h3 = dz[3], h4 = dz[4];
  if ((h2 + h3) < min_frac * h1) h3 = min_frac * h1 - h2;
  if ((h3 + h4) < min_frac * h1) h4 = min_frac * h1 - h3;
  const double h12 = h1 + h2, h23 = h2 + h3, h34 = h3 + h4;
  const double h123 = h12 + h3, h234 = h2 + h34, h1234 = h12 + h34;
  const double I_denB3 = 1.0 / (h123 * h12 * h23);
  const double I_h12 = (h123 * h23) * I_denB3, I_h23 = (h12 * h123) * I_denB3, I_h123 = (h12 * h23) * I_denB3;
  const double I_denom = 1.0 / (h1234 * (h234 * h34));
  const double I_h234 = (h1234 * h34) * I_denom, I_h1234 = (h234 * h34) * I_denom;
  const double W11 = -h1 * (I_h1234 + I_h123 + I_h12);
  const double W21 = h1 * h12 * (I_h234 * I_h1234 + I_h23 * (I_h234 + I_h123));
  const double W31 = -h1 * h12 * h123 * I_denom;
  const double W12 = 2.0 * (I_h12 * (1.0 + (h1 + h12) * (I_h1234 + I_h123)) + h1 * I_h1234 * I_h123);
  const double W22 = -2.0 * ((h1 * h12 * I_h1234) * (I_h23 * (I_h234 + I_h123)) + (h1 + h12) * (I_h1234 * I_h234 + I_h23 * (I_h234 + I_h123)));
  const double W32 = 2.0 * ((h1 + h12) * h123 + h1 * h12) * I_denom;
  const double W13 = -3.0 * I_h12 * I_h123 * (1.0 + I_h1234 * ((h1 + h12) + h123));
  const double W23 = 3.0 * I_h23 * (I_h123 + I_h1234 * ((h1 + h12) + h123) * (I_h123 + I_h234));
  const double W33 = -3.0 * ((h1 + h12) + h123) * I_denom;
  const double W14 = 4.0 * I_h1234 * I_h123 * I_h12;
  const double W24 = -4.0 * I_h1234 * (I_h23 * (I_h123 + I_h234));
  const double W34 = 4.0 * I_denom;
  const double d21 = u[2] - u[1], d32 = u[3] - u[2], d43 = u[4] - u[3];
  Csys[1] = ((u[1] + (W11 * d21)) + (W21 * d32)) + (W31 * d43);
  Csys[2] = ((W12 * d21) + (W22 * d32)) + (W32 * d43);
  Csys[3] = ((W13 * d21) + (W23 * d32)) + (W33 * d43);
  Csys[4] = ((W14 * d21) + (W24 * d32)) + (W34 * d43);
}

// the two one-sided end values shared by the explicit and implicit h4 schemes
template <int KCAP>
M6R_HD void end_values(int N, const double* h, const double* u, double h_neglect, double* top /*[2]*/, double* bot /*[2]*/) {
  double dz[5], ut[5], C[5];
  for (int i = 1; i <= 4; ++i) { dz[i] = rmax(h_neglect, h[i]); ut[i] = u[i]; }
  end_value_h4(dz, ut, C);
  top[0] = C[1];
  top[1] = C[1] + dz[1] * (C[2] + dz[1] * (C[3] + dz[1] * C[4]));
  for (int i = 1; i <= 4; ++i) { dz[i] = rmax(h_neglect, h[N + 1 - i]); ut[i] = u[N + 1 - i]; }
  end_value_h4(dz, ut, C);
  bot[0] = C[1];
  bot[1] = C[1] + dz[1] * (C[2] + dz[1] * (C[3] + dz[1] * C[4]));
}

// edge_values_explicit_h4, regrid_edge_values.F90:213-330
template <int KCAP>
M6R_HD void edge_values_explicit_h4(int N, const double* h, Recon<KCAP>& R, double h_neglect) {
  const double hMinFrac = 1.e-5;
  const double* u = R.u;
  for (int i = 3; i <= N - 1; ++i) {
    double h0 = h[i - 2], h1 = h[i - 1], h2 = h[i], h3 = h[i + 1];
    if (h0 + h1 == 0.0 || h1 + h2 == 0.0 || h2 + h3 == 0.0) {
      const double h_min = hMinFrac * rmax(h_neglect, (h0 + h1) + (h2 + h3));
      h0 = rmax(h_min, h[i - 2]); h1 = rmax(h_min, h[i - 1]); h2 = rmax(h_min, h[i]); h3 = rmax(h_min, h[i + 1]);
    }
    const double I_h12 = 1.0 / (h1 + h2);
    const double I_den_et2 = 1.0 / (((h0 + h1) + h2) * (h0 + h1)), I_h012 = (h0 + h1) * I_den_et2;
    const double I_den_et3 = 1.0 / ((h1 + (h2 + h3)) * (h2 + h3)), I_h123 = (h2 + h3) * I_den_et3;
    const double et1 = (1.0 + (h1 * I_h012 + (h0 + h1) * I_h123)) * I_h12 * (h2 * (h2 + h3)) * u[i - 1] +
                       (1.0 + (h2 * I_h123 + (h2 + h3) * I_h012)) * I_h12 * (h1 * (h0 + h1)) * u[i];
    const double et2 = (h1 * (h2 * (h2 + h3)) * I_den_et2) * (u[i - 1] - u[i - 2]);
    const double et3 = (h2 * (h1 * (h0 + h1)) * I_den_et3) * (u[i] - u[i + 1]);
    const double e = (et1 + (et2 + et3)) / ((h0 + h1) + (h2 + h3));
    R.E1[i] = e; R.E2[i - 1] = e;
  }
  double top[2], bot[2];
  end_values<KCAP>(N, h, u, h_neglect, top, bot);
  R.E1[1] = top[0]; R.E2[1] = top[1]; R.E1[2] = top[1];
  R.E2[N] = bot[0]; R.E1[N] = bot[1]; R.E2[N - 1] = bot[1];
}

// edge_values_implicit_h4 :473-630 with solve_diag_dominant_tridiag (regrid_solvers.F90:246-280) inlined: the forward
// sweep stores X in E1(1:N) (X(N+1) in xN1) and c1 in E2 (overwritten by the back substitution's results).
template <int KCAP>
M6R_HD void edge_values_implicit_h4(int N, const double* h, Recon<KCAP>& R, double h_neglect) {
  const double* u = R.u;
  double top[2], bot[2];
  end_values<KCAP>(N, h, u, h_neglect, top, bot);
  // row 1: Al = 0, Ac = 1, Au = 0, R = top[0]
  double I_pivot = 1.0 / (1.0 + 0.0);
  double d1 = 1.0 * I_pivot;
  double* X = R.E1;   // X(1:N)
  double* c1 = R.c2;  // c1(1:N), scratch (c2 is only meaningful for PLM)
  c1[1] = 0.0 * I_pivot;
  X[1] = top[0] * I_pivot;
  for (int k = 2; k <= N; ++k) {  // rows 2..N of the (N+1)-system are built from cells i = k-1, k
    const int i = k - 1;
    double h0 = rmax(h[i], h_neglect), h1 = rmax(h[i + 1], h_neglect);
    if (fabs(h0) < 1.0e-12 * fabs(h1)) h0 = 1.0e-12 * h1;
    if (fabs(h1) < 1.0e-12 * fabs(h0)) h1 = 1.0e-12 * h0;
    const double I_h2 = 1.0 / ((h0 + h1) * (h0 + h1));
    const double alpha = (h1 * h1) * I_h2, beta = (h0 * h0) * I_h2, abmix = (h0 * h1) * I_h2;
    const double a = 2.0 * alpha * (alpha + 2.0 * beta + 3.0 * abmix);
    const double b = 2.0 * beta * (beta + 2.0 * alpha + 3.0 * abmix);
    const double Ac = 2.0 * abmix, Al = alpha, Au = beta, Rk = a * u[i] + b * u[i + 1];
    const double denom_t1 = Ac + d1 * Al;
    I_pivot = 1.0 / (denom_t1 + Au);
    d1 = denom_t1 * I_pivot;
    c1[k] = Au * I_pivot;
    X[k] = (Rk - Al * X[k - 1]) * I_pivot;
  }
  // last row N+1: Al = 0, Ac = 1, R = bot[0]
  I_pivot = 1.0 / (1.0 + d1 * 0.0);
  double xn = (bot[0] - 0.0 * X[N]) * I_pivot;  // X(N+1)
  R.E2[N] = xn;
  for (int k = N; k >= 1; --k) {
    xn = X[k] - c1[k] * xn;
    X[k] = xn;                       // E1(k) = X(k)
    if (k >= 2) R.E2[k - 1] = xn;    // E2(k-1) = X(k)
  }
}

// PPM_reconstruction, PPM_functions.F90:21-50: bound_edge_values (regrid_edge_values.F90:39-105),
// check_discontinuous_edge_values (:143-165), PPM_limiter_standard (PPM_functions.F90:55-120) as three sweeps;
// then PPM_boundary_extrapolation (:155-300) from the coefficients of cells 2 and N-1.
template <int KCAP>
M6R_HD void PPM_reconstruction(int N, const double* h, Recon<KCAP>& R, double h_neglect, bool extrapolate) {
  const double* u = R.u;
  for (int k = 1; k <= N; ++k) {
    const int km1 = (k - 1 > 1) ? k - 1 : 1, kp1 = (k + 1 < N) ? k + 1 : N;
    double slope_x_h = 0.0;
    if (((h[km1] + h[kp1]) + 2.0 * h[k]) > 0.0) {
      const double sigma_l = (u[k] - u[km1]);
      const double sigma_c = (u[kp1] - u[km1]) * (h[k] / ((h[km1] + h[kp1]) + 2.0 * h[k]));
      const double sigma_r = (u[kp1] - u[k]);
      if ((sigma_l * sigma_r) > 0.0) slope_x_h = rsign(rmin3(fabs(sigma_l), fabs(sigma_c), fabs(sigma_r)), sigma_c);
    }
    double e1 = R.E1[k], e2 = R.E2[k];
    if ((u[km1] - e1) * (e1 - u[k]) < 0.0) e1 = u[k] - rsign(rmin(fabs(slope_x_h), fabs(e1 - u[k])), slope_x_h);
    if ((u[kp1] - e2) * (e2 - u[k]) < 0.0) e2 = u[k] + rsign(rmin(fabs(slope_x_h), fabs(e2 - u[k])), slope_x_h);
    R.E1[k] = rmax(rmin(e1, rmax(u[km1], u[k])), rmin(u[km1], u[k]));
    R.E2[k] = rmax(rmin(e2, rmax(u[kp1], u[k])), rmin(u[kp1], u[k]));
  }
  for (int k = 1; k <= N - 1; ++k) {
    if ((R.E1[k + 1] - R.E2[k]) * (u[k + 1] - u[k]) < 0.0) {
      double u0_avg = 0.5 * (R.E2[k] + R.E1[k + 1]);
      u0_avg = rmax(rmin(u0_avg, rmax(u[k], u[k + 1])), rmin(u[k], u[k + 1]));
      R.E2[k] = u0_avg; R.E1[k + 1] = u0_avg;
    }
  }
  for (int k = 2; k <= N - 1; ++k) {
    const double u_l = u[k - 1], u_c = u[k], u_r = u[k + 1];
    double edge_l = R.E1[k], edge_r = R.E2[k];
    if ((u_r - u_c) * (u_c - u_l) <= 0.0) { edge_l = u_c; edge_r = u_c; }
    else {
      const double expr1 = 3.0 * (edge_r - edge_l) * ((u_c - edge_l) + (u_c - edge_r));
      const double expr2 = (edge_r - edge_l) * (edge_r - edge_l);
      if (expr1 > expr2) {
        edge_l = u_c + 2.0 * (u_c - edge_r);
        edge_l = rmax(rmin(edge_l, rmax(u_l, u_c)), rmin(u_l, u_c));
      } else if (expr1 < -expr2) {
        edge_r = u_c + 2.0 * (u_c - edge_l);
        edge_r = rmax(rmin(edge_r, rmax(u_r, u_c)), rmin(u_r, u_c));
      }
    }
    if (fabs(edge_r - edge_l) < rmax(1.e-60, DBL_EPSILON * fabs(u_c))) { edge_l = u_c; edge_r = u_c; }
    R.E1[k] = edge_l; R.E2[k] = edge_r;
  }
  R.E1[1] = u[1]; R.E2[1] = u[1]; R.E1[N] = u[N]; R.E2[N] = u[N];
  if (!extrapolate) return;
  {  // top boundary cell from the parabola of cell 2
    const int i0 = 1, i1 = 2;
    const double h0 = h[i0], h1 = h[i1], u0 = u[i0], u1 = u[i1];
    const double b = 4.0 * (u[i1] - R.E1[i1]) + 2.0 * (u[i1] - R.E2[i1]);  // ppoly_coef(i1,2)
    double u1_r = b * ((h0 + h_neglect) / (h1 + h_neglect));
    const double slope = 2.0 * (u1 - u0);
    if (fabs(u1_r) > fabs(slope)) u1_r = slope;
    double u0_r = R.E1[i1];
    double u0_l = 3.0 * u0 + 0.5 * u1_r - 2.0 * u0_r;
    const double exp1 = (u0_r - u0_l) * (u0 - 0.5 * (u0_l + u0_r));
    const double exp2 = (u0_r - u0_l) * (u0_r - u0_l) / 6.0;
    if (exp1 > exp2) u0_l = 3.0 * u0 - 2.0 * u0_r;
    if (exp1 < -exp2) u0_r = 3.0 * u0 - 2.0 * u0_l;
    R.E1[i0] = u0_l; R.E2[i0] = u0_r;
  }
  {  // bottom boundary cell from the parabola of cell N-1
    const int i0 = N - 1, i1 = N;
    const double h0 = h[i0], h1 = h[i1], u0 = u[i0], u1 = u[i1];
    // ppoly_coef(i0,2:3); PPM needs n0 >= 4 (build_reconstructions), so cell N-1 is an interior cell
    const double b = 4.0 * (u[i0] - R.E1[i0]) + 2.0 * (u[i0] - R.E2[i0]);
    const double c = 3.0 * ((R.E2[i0] - u[i0]) + (R.E1[i0] - u[i0]));
    double u1_l = (b + 2 * c);
    u1_l = u1_l * ((h1 + h_neglect) / (h0 + h_neglect));
    const double slope = 2.0 * (u1 - u0);
    if (fabs(u1_l) > fabs(slope)) u1_l = slope;
    double u0_l = R.E2[i0];
    double u0_r = 3.0 * u1 - 0.5 * u1_l - 2.0 * u0_l;
    const double exp1 = (u0_r - u0_l) * (u1 - 0.5 * (u0_l + u0_r));
    const double exp2 = (u0_r - u0_l) * (u0_r - u0_l) / 6.0;
    if (exp1 > exp2) u0_l = 3.0 * u1 - 2.0 * u0_r;
    if (exp1 < -exp2) u0_r = 3.0 * u1 - 2.0 * u0_l;
    R.E1[i1] = u0_l; R.E2[i1] = u0_r;
  }
}

// build_reconstructions_1d, MOM_remapping.F90:410-550; returns the integration method
template <int KCAP>
M6R_HD int build_reconstructions(const Params& P, int n0, const double* h0, Recon<KCAP>& R) {
  int scheme = P.scheme;
  if (n0 <= 1) scheme = SCHEME_PCM;
  else if (n0 <= 3) scheme = (scheme < SCHEME_PLM) ? scheme : (int)SCHEME_PLM;
  else if (n0 <= 4) scheme = (scheme < SCHEME_PPM_H4) ? scheme : (int)SCHEME_PPM_H4;
  if (scheme == SCHEME_PCM) {
    for (int k = 1; k <= n0; ++k) { R.E1[k] = R.u[k]; R.E2[k] = R.u[k]; }
    return INT_PCM;
  }
  if (scheme == SCHEME_PLM) {
    PLM_reconstruction<KCAP>(n0, h0, R, P.h_neglect, P.boundary_extrapolation != 0);
    return INT_PLM;
  }
  if (scheme == SCHEME_PPM_H4) edge_values_explicit_h4<KCAP>(n0, h0, R, P.h_neglect_edge);
  else edge_values_implicit_h4<KCAP>(n0, h0, R, P.h_neglect_edge);
  PPM_reconstruction<KCAP>(n0, h0, R, P.h_neglect, P.boundary_extrapolation != 0);
  return INT_PPM;
}

// average_value_ppoly, MOM_remapping.F90:1391-1490
template <int KCAP>
M6R_HD double average_value(const Recon<KCAP>& R, int method, int i0, double xa, double xb) {
  const double a_L = R.E1[i0], a_R = R.E2[i0], u_c = R.u[i0];
  if (xb > xa) {
    if (method == INT_PCM) return u_c;
    if (method == INT_PLM) return (a_L + R.c2[i0] * 0.5 * (xb + xa));
    const double mx = 0.5 * (xa + xb);
    const double a_c = 0.5 * ((u_c - a_L) + (u_c - a_R));
    if (mx < 0.5) {
      const double xa2b2ab = (xa * xa + xb * xb) + xa * xb;
      return a_L + ((a_R - a_L) * mx + a_c * (3. * (xb + xa) - 2. * xa2b2ab));
    }
    const double Ya = 1. - xa, Yb = 1. - xb, my = 0.5 * (Ya + Yb);
    const double Ya2b2ab = (Ya * Ya + Yb * Yb) + Ya * Yb;
    return a_R + ((a_L - a_R) * my + a_c * (3. * (Yb + Ya) - 2. * Ya2b2ab));
  }
  if (method == INT_PCM) return a_L;  // ppoly0_coefs(i0,1) = u0(i0) = E1 for PCM
  const double Ya = 1. - xa;
  if (method == INT_PLM) return (xa < 0.5) ? a_L + xa * (a_R - a_L) : a_R + Ya * (a_L - a_R);
  const double a_c = 3. * ((u_c - a_L) + (u_c - a_R));
  return (xa < 0.5) ? a_L + xa * ((a_R - a_L) + a_c * Ya) : a_R + Ya * ((a_L - a_R) + a_c * xa);
}

template <int KCAP>
struct SubVals { double u_sub[2 * KCAP + 2], uh_sub[2 * KCAP + 2]; };

// remap_src_to_sub_grid_om4 :845-958 / remap_src_to_sub_grid :962-1099, then remap_sub_to_tgt_grid_om4 :1103-1163.
// u1 is written through `out`; returns nothing (the error estimates of the reference are diagnostics only).
template <int KCAP, class Out>
M6R_HD void remap_via_sub_cells(const Params& P, const SubGrid<KCAP>& S, const Recon<KCAP>& R, int method, SubVals<KCAP>& V, Out out,
                                double conc_underflow) {
  const int n0 = S.n0, n1 = S.n1, ns = n0 + n1 + 1;
  const bool om4 = P.om4 != 0, fb = P.force_bounds_in_subcell != 0;
  double xa = 0., xb = 0., dh0_eff = 0.;
  if (om4) { V.uh_sub[1] = 0.; V.u_sub[1] = R.E1[1]; }
  const int first = om4 ? 2 : 1, last = om4 ? n0 + n1 : ns;
  for (int i_sub = first; i_sub <= last; ++i_sub) {
    const double dh = S.h_sub[i_sub];
    const int i0 = S.isub_src[i_sub];
    dh0_eff = dh0_eff + dh;
    const double hden = om4 ? S.h0_eff[i0] : S.h0[i0];
    double us;
    if (hden > 0.) {
      xb = dh0_eff / hden;
      xb = rmin(1., xb);
      us = average_value<KCAP>(R, method, i0, xa, xb);
    } else { xb = 1.; us = R.u[i0]; }
    if (fb) {
      us = rmax(us, rmin(R.E1[i0], R.E2[i0]));
      us = rmin(us, rmax(R.E1[i0], R.E2[i0]));
    }
    V.u_sub[i_sub] = us;
    V.uh_sub[i_sub] = dh * us;
    if (i_sub < ns) {
      if (S.isub_src[i_sub + 1] != i0) { dh0_eff = 0.; xa = 0.; }
      else xa = xb;
    }
  }
  if (om4) { V.u_sub[ns] = R.E2[n0]; V.uh_sub[ns] = R.E2[n0] * S.h_sub[ns]; }
  // adjust_thickest_subcell (every source cell up to the last one with volume)
  int i0_last_thick_cell = 0;
  for (int i0 = 1; i0 <= n0; ++i0) if (S.h0[i0] > 0.) i0_last_thick_cell = i0;
  for (int i0 = 1; i0 <= i0_last_thick_cell; ++i0) {
    const int i_max = S.isrc_max[i0];
    if (S.h_sub[i_max] > 0.) {
      double duh = 0.;
      for (int i_sub = S.isrc_start[i0]; i_sub <= S.isrc_end[i0]; ++i_sub) if (i_sub != i_max) duh = duh + V.uh_sub[i_sub];
      V.uh_sub[i_max] = R.u[i0] * S.h0[i0] - duh;
    }
  }
  // sub-cells to target cells
  const bool ft = P.force_bounds_in_target != 0;
  for (int i1 = 1; i1 <= n1; ++i1) {
    double r;
    const int s = S.itgt_start[i1], e = S.itgt_end[i1];
    if (S.h1[i1] > 0.) {
      double duh = 0., dh = 0.;
      double u1min = V.u_sub[s], u1max = V.u_sub[s];
      for (int i_sub = s; i_sub <= e; ++i_sub) {
        u1min = rmin(u1min, V.u_sub[i_sub]); u1max = rmax(u1max, V.u_sub[i_sub]);
        dh = dh + S.h_sub[i_sub];
        duh = duh + V.uh_sub[i_sub];
      }
      r = duh / dh;
      if (ft) r = rmax(u1min, rmin(u1max, r));
    } else r = V.u_sub[s];
    if (conc_underflow > 0.0 && fabs(r) < conc_underflow) r = 0.0;
    out(i1) = r;
  }
}

}  // namespace m6remap
