// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of the ALE vertical remapping core, /root/reference/src/ALE/:
//   MOM_remapping.F90: remapping_core_h :234-335, build_reconstructions_1d :410-550, intersect_src_tgt_grids :642-798,
//     remap_src_to_sub_grid_om4 :845-958, remap_src_to_sub_grid :962-1099, remap_sub_to_tgt_grid_om4 :1103-1163,
//     average_value_ppoly :1391-1490
//   PCM_functions.F90 :21-45, PLM_functions.F90 (PLM_slope_wa :17, PLM_monotonized_slope :119, PLM_extrapolate_slope :160,
//     PLM_reconstruction :188, PLM_boundary_extrapolation :267), PPM_functions.F90 (PPM_reconstruction :21,
//     PPM_limiter_standard :55, PPM_boundary_extrapolation :155), regrid_edge_values.F90 (bound_edge_values :39,
//     check_discontinuous_edge_values :143, edge_values_explicit_h4 :213, edge_values_implicit_h4 :473, end_value_h4 :633),
//     regrid_solvers.F90 solve_diag_dominant_tridiag :246
//   MOM_ALE.F90: ALE_remap_tracers :760-879 (column loop), ALE_remap_set_h_vel :882-925, ALE_remap_velocities :1089+.
// Only the answer_date >= 20190101 expressions are restated (the reference's default and what its unit tests pin).
// PARITY: PINNED by the reference's own known-answer vectors (remapping_unit_tests, MOM_remapping.F90:2072+), see
// tests/test_oracle_remap_kat.py.
// All 1-D arrays use the Fortran 1-based indices (element 0 unused).
#include "oracle.h"
#include "ogrid.hpp"
#include <cmath>
#include <cfloat>
#include <vector>
#include <omp.h>

using namespace orc;

namespace {

const double hMinFrac = 1.e-5;  // regrid_edge_values.F90:25
typedef std::vector<double> vd;
typedef std::vector<int> vi;
// ppoly_E(k,m) m=1,2 ; coefs(k,m) m=1..3  (k = 1..N)
struct Poly { int N; vd E1, E2, c1, c2, c3; explicit Poly(int n) : N(n), E1(n + 2, 0.), E2(n + 2, 0.), c1(n + 2, 0.), c2(n + 2, 0.), c3(n + 2, 0.) {} };

inline double min3(double a, double b, double c) { return fmin2(fmin2(a, b), c); }
inline double max3(double a, double b, double c) { return fmax2(fmax2(a, b), c); }

// PLM_slope_wa, PLM_functions.F90:17-65
double PLM_slope_wa(double h_l, double h_c, double h_r, double h_neglect, double u_l, double u_c, double u_r) {
  const double sigma_r = u_r - u_c, sigma_l = u_c - u_l;
  const double sigma_c = 2.0 * (u_r - u_l) * (h_c / (h_l + 2.0 * h_c + h_r + h_neglect));
  const double u_min = min3(u_l, u_c, u_r), u_max = max3(u_l, u_c, u_r);
  double s;
  if ((sigma_l * sigma_r) > 0.0) s = fsign(fmin2(std::fabs(sigma_c), 2. * fmin2(u_c - u_min, u_max - u_c)), sigma_c);
  else s = 0.0;
  if (u_c - 0.5 * std::fabs(s) < u_min || u_c + 0.5 * std::fabs(s) > u_max) s = s * (1. - DBL_EPSILON);
  if (std::fabs(s) < 1.E-140) s = 0.;
  return s;
}
// PLM_monotonized_slope :119-155
double PLM_monotonized_slope(double u_l, double u_c, double u_r, double s_l, double s_c, double s_r) {
  const double almost_two = 2. * (1. - DBL_EPSILON);
  const double e_r = u_l + 0.5 * s_l, e_l = u_r - 0.5 * s_r;
  double slp = std::fabs(s_c);
  double edge = u_c - 0.5 * s_c;
  if ((edge - e_r) * (u_c - edge) < 0.) { edge = 0.5 * (edge + e_r); slp = fmin2(slp, std::fabs(edge - u_c) * almost_two); }
  edge = u_c + 0.5 * s_c;
  if ((edge - u_c) * (e_l - edge) < 0.) { edge = 0.5 * (edge + e_l); slp = fmin2(slp, std::fabs(edge - u_c) * almost_two); }
  return fsign(slp, s_c);
}
// PLM_extrapolate_slope :160-184
double PLM_extrapolate_slope(double h_l, double h_c, double h_neglect, double u_l, double u_c) {
  const double hl = h_l + h_neglect, hc = h_c + h_neglect;
  const double left_edge = (u_l * hc + u_c * hl) / (hl + hc);
  return 2.0 * (u_c - left_edge);
}
// PCM_reconstruction, PCM_functions.F90:21-45
void PCM_reconstruction(int N, const double* u, Poly& P) {
  for (int k = 1; k <= N; ++k) { P.c1[k] = u[k]; P.E1[k] = u[k]; P.E2[k] = u[k]; }
}
// PLM_reconstruction :188-262
void PLM_reconstruction(int N, const double* h, const double* u, Poly& P, double h_neglect) {
  const double almost_one = 1. - DBL_EPSILON;
  vd slp(N + 2, 0.), mslp(N + 2, 0.);
  for (int k = 2; k <= N - 1; ++k) slp[k] = PLM_slope_wa(h[k - 1], h[k], h[k + 1], h_neglect, u[k - 1], u[k], u[k + 1]);
  slp[1] = 0.; slp[N] = 0.;
  for (int k = 2; k <= N - 1; ++k) mslp[k] = PLM_monotonized_slope(u[k - 1], u[k], u[k + 1], slp[k - 1], slp[k], slp[k + 1]);
  mslp[1] = 0.; mslp[N] = 0.;
  P.E1[1] = u[1]; P.E2[1] = u[1]; P.c1[1] = u[1]; P.c2[1] = 0.;
  for (int k = 2; k <= N - 1; ++k) {
    const double slope = mslp[k];
    const double u_l = u[k] - 0.5 * slope, u_r = u[k] + 0.5 * slope;
    P.E1[k] = u_l; P.E2[k] = u_r; P.c1[k] = u_l; P.c2[k] = (u_r - u_l);
    const double edge = P.c2[k] + P.c1[k];
    const double e_r = u[k + 1] - 0.5 * fsign(mslp[k + 1], slp[k + 1]);
    if ((edge - u[k]) * (e_r - edge) < 0.) P.c2[k] = P.c2[k] * almost_one;
  }
  P.E1[N] = u[N]; P.E2[N] = u[N]; P.c1[N] = u[N]; P.c2[N] = 0.;
}
// PLM_boundary_extrapolation :267-300
void PLM_boundary_extrapolation(int N, const double* h, const double* u, Poly& P, double h_neglect) {
  double slope = -PLM_extrapolate_slope(h[2], h[1], h_neglect, u[2], u[1]);
  P.E1[1] = u[1] - 0.5 * slope; P.E2[1] = u[1] + 0.5 * slope;
  P.c1[1] = P.E1[1]; P.c2[1] = P.E2[1] - P.E1[1];
  slope = PLM_extrapolate_slope(h[N - 1], h[N], h_neglect, u[N - 1], u[N]);
  P.E1[N] = u[N] - 0.5 * slope; P.E2[N] = u[N] + 0.5 * slope;
  P.c1[N] = P.E1[N]; P.c2[N] = P.E2[N] - P.E1[N];
}
// bound_edge_values, regrid_edge_values.F90:39-105 (answer_date >= 20190101 branch)
void bound_edge_values(int N, const double* h, const double* u, Poly& P, double h_neglect) {
  for (int k = 1; k <= N; ++k) {
    const int km1 = std::max(1, k - 1), kp1 = std::min(k + 1, N);
    double slope_x_h = 0.0;
    if (((h[km1] + h[kp1]) + 2.0 * h[k]) > 0.0) {
      const double sigma_l = (u[k] - u[km1]);
      const double sigma_c = (u[kp1] - u[km1]) * (h[k] / ((h[km1] + h[kp1]) + 2.0 * h[k]));
      const double sigma_r = (u[kp1] - u[k]);
      if ((sigma_l * sigma_r) > 0.0) slope_x_h = fsign(min3(std::fabs(sigma_l), std::fabs(sigma_c), std::fabs(sigma_r)), sigma_c);
    }
    if ((u[km1] - P.E1[k]) * (P.E1[k] - u[k]) < 0.0) P.E1[k] = u[k] - fsign(fmin2(std::fabs(slope_x_h), std::fabs(P.E1[k] - u[k])), slope_x_h);
    if ((u[kp1] - P.E2[k]) * (P.E2[k] - u[k]) < 0.0) P.E2[k] = u[k] + fsign(fmin2(std::fabs(slope_x_h), std::fabs(P.E2[k] - u[k])), slope_x_h);
    P.E1[k] = fmax2(fmin2(P.E1[k], fmax2(u[km1], u[k])), fmin2(u[km1], u[k]));
    P.E2[k] = fmax2(fmin2(P.E2[k], fmax2(u[kp1], u[k])), fmin2(u[kp1], u[k]));
  }
}
// check_discontinuous_edge_values :143-165
void check_discontinuous_edge_values(int N, const double* u, Poly& P) {
  for (int k = 1; k <= N - 1; ++k) {
    if ((P.E1[k + 1] - P.E2[k]) * (u[k + 1] - u[k]) < 0.0) {
      double u0_avg = 0.5 * (P.E2[k] + P.E1[k + 1]);
      u0_avg = fmax2(fmin2(u0_avg, fmax2(u[k], u[k + 1])), fmin2(u[k], u[k + 1]));
      P.E2[k] = u0_avg; P.E1[k + 1] = u0_avg;
    }
  }
}
// end_value_h4 :633-760: dz, u, Csys are 1-based length-4 arrays
void end_value_h4(const double* dz, const double* u, double* Csys) {
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
  double Wt[4][5];  // Wt[n][m] = Wt(n,m)
  Wt[1][1] = -h1 * (I_h1234 + I_h123 + I_h12);
  Wt[2][1] = h1 * h12 * (I_h234 * I_h1234 + I_h23 * (I_h234 + I_h123));
  Wt[3][1] = -h1 * h12 * h123 * I_denom;
  Wt[1][2] = 2.0 * (I_h12 * (1.0 + (h1 + h12) * (I_h1234 + I_h123)) + h1 * I_h1234 * I_h123);
  Wt[2][2] = -2.0 * ((h1 * h12 * I_h1234) * (I_h23 * (I_h234 + I_h123)) + (h1 + h12) * (I_h1234 * I_h234 + I_h23 * (I_h234 + I_h123)));
  Wt[3][2] = 2.0 * ((h1 + h12) * h123 + h1 * h12) * I_denom;
  Wt[1][3] = -3.0 * I_h12 * I_h123 * (1.0 + I_h1234 * ((h1 + h12) + h123));
  Wt[2][3] = 3.0 * I_h23 * (I_h123 + I_h1234 * ((h1 + h12) + h123) * (I_h123 + I_h234));
  Wt[3][3] = -3.0 * ((h1 + h12) + h123) * I_denom;
  Wt[1][4] = 4.0 * I_h1234 * I_h123 * I_h12;
  Wt[2][4] = -4.0 * I_h1234 * (I_h23 * (I_h123 + I_h234));
  Wt[3][4] = 4.0 * I_denom;
  Csys[1] = ((u[1] + (Wt[1][1] * (u[2] - u[1]))) + (Wt[2][1] * (u[3] - u[2]))) + (Wt[3][1] * (u[4] - u[3]));
  Csys[2] = ((Wt[1][2] * (u[2] - u[1])) + (Wt[2][2] * (u[3] - u[2]))) + (Wt[3][2] * (u[4] - u[3]));
  Csys[3] = ((Wt[1][3] * (u[2] - u[1])) + (Wt[2][3] * (u[3] - u[2]))) + (Wt[3][3] * (u[4] - u[3]));
  Csys[4] = ((Wt[1][4] * (u[2] - u[1])) + (Wt[2][4] * (u[3] - u[2]))) + (Wt[3][4] * (u[4] - u[3]));
}
// edge_values_explicit_h4 :213-330 (answer_date >= 20190101)
void edge_values_explicit_h4(int N, const double* h, const double* u, Poly& P, double h_neglect) {
  for (int i = 3; i <= N - 1; ++i) {
    double h0 = h[i - 2], h1 = h[i - 1], h2 = h[i], h3 = h[i + 1];
    if (h0 + h1 == 0.0 || h1 + h2 == 0.0 || h2 + h3 == 0.0) {
      const double h_min = hMinFrac * fmax2(h_neglect, (h0 + h1) + (h2 + h3));
      h0 = fmax2(h_min, h[i - 2]); h1 = fmax2(h_min, h[i - 1]); h2 = fmax2(h_min, h[i]); h3 = fmax2(h_min, h[i + 1]);
    }
    const double I_h12 = 1.0 / (h1 + h2);
    const double I_den_et2 = 1.0 / (((h0 + h1) + h2) * (h0 + h1)), I_h012 = (h0 + h1) * I_den_et2;
    const double I_den_et3 = 1.0 / ((h1 + (h2 + h3)) * (h2 + h3)), I_h123 = (h2 + h3) * I_den_et3;
    const double et1 = (1.0 + (h1 * I_h012 + (h0 + h1) * I_h123)) * I_h12 * (h2 * (h2 + h3)) * u[i - 1] +
                       (1.0 + (h2 * I_h123 + (h2 + h3) * I_h012)) * I_h12 * (h1 * (h0 + h1)) * u[i];
    const double et2 = (h1 * (h2 * (h2 + h3)) * I_den_et2) * (u[i - 1] - u[i - 2]);
    const double et3 = (h2 * (h1 * (h0 + h1)) * I_den_et3) * (u[i] - u[i + 1]);
    P.E1[i] = (et1 + (et2 + et3)) / ((h0 + h1) + (h2 + h3));
    P.E2[i - 1] = P.E1[i];
  }
  double dz[5], ut[5], C[5];
  for (int i = 1; i <= 4; ++i) { dz[i] = fmax2(h_neglect, h[i]); ut[i] = u[i]; }
  end_value_h4(dz, ut, C);
  P.E1[1] = C[1];
  P.E2[1] = C[1] + dz[1] * (C[2] + dz[1] * (C[3] + dz[1] * C[4]));
  P.E1[2] = P.E2[1];
  for (int i = 1; i <= 4; ++i) { dz[i] = fmax2(h_neglect, h[N + 1 - i]); ut[i] = u[N + 1 - i]; }
  end_value_h4(dz, ut, C);
  P.E2[N] = C[1];
  P.E1[N] = C[1] + dz[1] * (C[2] + dz[1] * (C[3] + dz[1] * C[4]));
  P.E2[N - 1] = P.E1[N];
}
// solve_diag_dominant_tridiag, regrid_solvers.F90:246-280 (1-based arrays of length N)
void solve_diag_dominant_tridiag(const double* Al, const double* Ac, const double* Au, const double* R, double* X, int N) {
  vd c1(N + 2);
  double I_pivot = 1.0 / (Ac[1] + Au[1]);
  double d1 = Ac[1] * I_pivot;
  c1[1] = Au[1] * I_pivot;
  X[1] = R[1] * I_pivot;
  for (int k = 2; k <= N - 1; ++k) {
    const double denom_t1 = Ac[k] + d1 * Al[k];
    I_pivot = 1.0 / (denom_t1 + Au[k]);
    d1 = denom_t1 * I_pivot;
    c1[k] = Au[k] * I_pivot;
    X[k] = (R[k] - Al[k] * X[k - 1]) * I_pivot;
  }
  I_pivot = 1.0 / (Ac[N] + d1 * Al[N]);
  X[N] = (R[N] - Al[N] * X[N - 1]) * I_pivot;
  for (int k = N - 1; k >= 1; --k) X[k] = X[k] - c1[k] * X[k + 1];
}
// edge_values_implicit_h4 :473-630 (answer_date >= 20190101)
void edge_values_implicit_h4(int N, const double* h, const double* u, Poly& P, double h_neglect) {
  vd tri_l(N + 3, 0.), tri_c(N + 3, 0.), tri_u(N + 3, 0.), tri_b(N + 3, 0.), tri_x(N + 3, 0.);
  for (int i = 1; i <= N - 1; ++i) {
    double h0 = fmax2(h[i], h_neglect), h1 = fmax2(h[i + 1], h_neglect);
    if (std::fabs(h0) < 1.0e-12 * std::fabs(h1)) h0 = 1.0e-12 * h1;
    if (std::fabs(h1) < 1.0e-12 * std::fabs(h0)) h1 = 1.0e-12 * h0;
    const double I_h2 = 1.0 / ((h0 + h1) * (h0 + h1));
    const double alpha = (h1 * h1) * I_h2, beta = (h0 * h0) * I_h2, abmix = (h0 * h1) * I_h2;
    const double a = 2.0 * alpha * (alpha + 2.0 * beta + 3.0 * abmix);
    const double b = 2.0 * beta * (beta + 2.0 * alpha + 3.0 * abmix);
    tri_c[i + 1] = 2.0 * abmix;
    tri_l[i + 1] = alpha; tri_u[i + 1] = beta;
    tri_b[i + 1] = a * u[i] + b * u[i + 1];
  }
  double dz[5], ut[5], C[5];
  for (int i = 1; i <= 4; ++i) { dz[i] = fmax2(h_neglect, h[i]); ut[i] = u[i]; }
  end_value_h4(dz, ut, C);
  tri_b[1] = C[1]; tri_c[1] = 1.0; tri_u[1] = 0.0;
  for (int i = 1; i <= 4; ++i) { dz[i] = fmax2(h_neglect, h[N + 1 - i]); ut[i] = u[N + 1 - i]; }
  end_value_h4(dz, ut, C);
  tri_b[N + 1] = C[1]; tri_c[N + 1] = 1.0; tri_l[N + 1] = 0.0;
  solve_diag_dominant_tridiag(tri_l.data(), tri_c.data(), tri_u.data(), tri_b.data(), tri_x.data(), N + 1);
  P.E1[1] = tri_x[1];
  for (int i = 2; i <= N; ++i) { P.E1[i] = tri_x[i]; P.E2[i - 1] = tri_x[i]; }
  P.E2[N] = tri_x[N + 1];
}
// PPM_limiter_standard, PPM_functions.F90:55-120
void PPM_limiter_standard(int N, const double* h, const double* u, Poly& P, double h_neglect) {
  bound_edge_values(N, h, u, P, h_neglect);
  check_discontinuous_edge_values(N, u, P);
  for (int k = 2; k <= N - 1; ++k) {
    const double u_l = u[k - 1], u_c = u[k], u_r = u[k + 1];
    double edge_l = P.E1[k], edge_r = P.E2[k];
    if ((u_r - u_c) * (u_c - u_l) <= 0.0) { edge_l = u_c; edge_r = u_c; }
    else {
      const double expr1 = 3.0 * (edge_r - edge_l) * ((u_c - edge_l) + (u_c - edge_r));
      const double expr2 = (edge_r - edge_l) * (edge_r - edge_l);
      if (expr1 > expr2) {
        edge_l = u_c + 2.0 * (u_c - edge_r);
        edge_l = fmax2(fmin2(edge_l, fmax2(u_l, u_c)), fmin2(u_l, u_c));
      } else if (expr1 < -expr2) {
        edge_r = u_c + 2.0 * (u_c - edge_l);
        edge_r = fmax2(fmin2(edge_r, fmax2(u_r, u_c)), fmin2(u_r, u_c));
      }
    }
    if (std::fabs(edge_r - edge_l) < fmax2(1.e-60, DBL_EPSILON * std::fabs(u_c))) { edge_l = u_c; edge_r = u_c; }
    P.E1[k] = edge_l; P.E2[k] = edge_r;
  }
  P.E1[1] = u[1]; P.E2[1] = u[1]; P.E1[N] = u[N]; P.E2[N] = u[N];
}
// PPM_reconstruction :21-50
void PPM_reconstruction(int N, const double* h, const double* u, Poly& P, double h_neglect) {
  PPM_limiter_standard(N, h, u, P, h_neglect);
  for (int k = 1; k <= N; ++k) {
    const double edge_l = P.E1[k], edge_r = P.E2[k];
    P.c1[k] = edge_l;
    P.c2[k] = 4.0 * (u[k] - edge_l) + 2.0 * (u[k] - edge_r);
    P.c3[k] = 3.0 * ((edge_r - u[k]) + (edge_l - u[k]));
  }
}
// PPM_boundary_extrapolation :155-300
void PPM_boundary_extrapolation(int N, const double* h, const double* u, Poly& P, double h_neglect) {
  int i0 = 1, i1 = 2;
  double h0 = h[i0], h1 = h[i1], u0 = u[i0], u1 = u[i1];
  double b = P.c2[i1];
  double u1_r = b * ((h0 + h_neglect) / (h1 + h_neglect));
  double slope = 2.0 * (u1 - u0);
  if (std::fabs(u1_r) > std::fabs(slope)) u1_r = slope;
  double u0_r = P.E1[i1];
  double u0_l = 3.0 * u0 + 0.5 * u1_r - 2.0 * u0_r;
  double exp1 = (u0_r - u0_l) * (u0 - 0.5 * (u0_l + u0_r));
  double exp2 = (u0_r - u0_l) * (u0_r - u0_l) / 6.0;
  if (exp1 > exp2) u0_l = 3.0 * u0 - 2.0 * u0_r;
  if (exp1 < -exp2) u0_r = 3.0 * u0 - 2.0 * u0_l;
  P.E1[i0] = u0_l; P.E2[i0] = u0_r;
  P.c1[i0] = u0_l; P.c2[i0] = 6.0 * u0 - 4.0 * u0_l - 2.0 * u0_r; P.c3[i0] = 3.0 * (u0_r + u0_l - 2.0 * u0);
  i0 = N - 1; i1 = N;
  h0 = h[i0]; h1 = h[i1]; u0 = u[i0]; u1 = u[i1];
  b = P.c2[i0];
  const double c = P.c3[i0];
  double u1_l = (b + 2 * c);
  u1_l = u1_l * ((h1 + h_neglect) / (h0 + h_neglect));
  slope = 2.0 * (u1 - u0);
  if (std::fabs(u1_l) > std::fabs(slope)) u1_l = slope;
  u0_l = P.E2[i0];
  u0_r = 3.0 * u1 - 0.5 * u1_l - 2.0 * u0_l;
  exp1 = (u0_r - u0_l) * (u1 - 0.5 * (u0_l + u0_r));
  exp2 = (u0_r - u0_l) * (u0_r - u0_l) / 6.0;
  if (exp1 > exp2) u0_l = 3.0 * u1 - 2.0 * u0_r;
  if (exp1 < -exp2) u0_r = 3.0 * u1 - 2.0 * u0_l;
  P.E1[i1] = u0_l; P.E2[i1] = u0_r;
  P.c1[i1] = u0_l; P.c2[i1] = 6.0 * u1 - 4.0 * u0_l - 2.0 * u0_r; P.c3[i1] = 3.0 * (u0_r + u0_l - 2.0 * u1);
}

}  // namespace

// ALE_PLM_edge_values (MOM_ALE.F90:1518-1576) for one column: Q_t, Q_b of a PLM reconstruction; arrays are 1-based (element 0 unused).
// Used by the pressure-force restatement (TS_PLM_edge_values :1495), oracle/pgf.cpp.
void orc::ale_plm_edge_values_column(int nk, const double* h, const double* Q, bool bdry_extrap, double h_neglect, double* Q_t, double* Q_b) {
  vd slp(nk + 2, 0.);
  slp[1] = 0.;
  for (int k = 2; k <= nk - 1; ++k) slp[k] = PLM_slope_wa(h[k - 1], h[k], h[k + 1], h_neglect, Q[k - 1], Q[k], Q[k + 1]);
  slp[nk] = 0.;
  for (int k = 2; k <= nk - 1; ++k) {
    const double mslp = PLM_monotonized_slope(Q[k - 1], Q[k], Q[k + 1], slp[k - 1], slp[k], slp[k + 1]);
    Q_t[k] = Q[k] - 0.5 * mslp;
    Q_b[k] = Q[k] + 0.5 * mslp;
  }
  if (bdry_extrap) {
    double mslp = -PLM_extrapolate_slope(h[2], h[1], h_neglect, Q[2], Q[1]);
    Q_t[1] = Q[1] - 0.5 * mslp; Q_b[1] = Q[1] + 0.5 * mslp;
    mslp = PLM_extrapolate_slope(h[nk - 1], h[nk], h_neglect, Q[nk - 1], Q[nk]);
    Q_t[nk] = Q[nk] - 0.5 * mslp; Q_b[nk] = Q[nk] + 0.5 * mslp;
  } else {
    Q_t[1] = Q[1]; Q_b[1] = Q[1]; Q_t[nk] = Q[nk]; Q_b[nk] = Q[nk];
  }
}

// One field of TS_PPM_edge_values (MOM_ALE.F90:1620-1660; answer_date >= 20190101): edge_values_implicit_h4 + PPM_reconstruction
// (+ PPM_boundary_extrapolation), returning ppol_E(:,1), ppol_E(:,2).  1-based arrays.
void orc::ale_ppm_edge_values_column(int nk, const double* h, const double* Q, bool bdry_extrap, double h_neglect, double h_neglect_edge,
                                     double* Q_t, double* Q_b) {
  Poly P(nk);
  edge_values_implicit_h4(nk, h, Q, P, h_neglect_edge);
  PPM_reconstruction(nk, h, Q, P, h_neglect);
  if (bdry_extrap) PPM_boundary_extrapolation(nk, h, Q, P, h_neglect);
  for (int k = 1; k <= nk; ++k) { Q_t[k] = P.E1[k]; Q_b[k] = P.E2[k]; }
}

namespace {

enum { INTEGRATION_PCM = 0, INTEGRATION_PLM = 1, INTEGRATION_PPM = 3 };

// build_reconstructions_1d :410-550
int build_reconstructions_1d(const mom6cu_remapping_cs* CS, int n0, const double* h0, const double* u0, Poly& P) {
  const double h_neglect = CS->h_neglect, h_neg_edge = CS->h_neglect_edge;
  int scheme = CS->remapping_scheme;
  if (n0 <= 1) scheme = MOM6CU_REMAPPING_PCM;
  else if (n0 <= 3) scheme = std::min(scheme, (int)MOM6CU_REMAPPING_PLM);
  else if (n0 <= 4 && scheme != 10) scheme = std::min(scheme, (int)MOM6CU_REMAPPING_PPM_H4);
  switch (scheme) {
    case MOM6CU_REMAPPING_PCM: PCM_reconstruction(n0, u0, P); return INTEGRATION_PCM;
    case MOM6CU_REMAPPING_PLM:
      PLM_reconstruction(n0, h0, u0, P, h_neglect);
      if (CS->boundary_extrapolation) PLM_boundary_extrapolation(n0, h0, u0, P, h_neglect);
      return INTEGRATION_PLM;
    case MOM6CU_REMAPPING_PPM_H4:
      edge_values_explicit_h4(n0, h0, u0, P, h_neg_edge);
      PPM_reconstruction(n0, h0, u0, P, h_neglect);
      if (CS->boundary_extrapolation) PPM_boundary_extrapolation(n0, h0, u0, P, h_neglect);
      return INTEGRATION_PPM;
    case MOM6CU_REMAPPING_PPM_IH4:
      edge_values_implicit_h4(n0, h0, u0, P, h_neg_edge);
      PPM_reconstruction(n0, h0, u0, P, h_neglect);
      if (CS->boundary_extrapolation) PPM_boundary_extrapolation(n0, h0, u0, P, h_neglect);
      return INTEGRATION_PPM;
    default: return -999;
  }
}

// average_value_ppoly :1391-1490
double average_value_ppoly(const double* u0, const Poly& P, int method, int i0, double xa, double xb) {
  double u_ave = 0.;
  if (xb > xa) {
    if (method == INTEGRATION_PCM) u_ave = u0[i0];
    else if (method == INTEGRATION_PLM) u_ave = (P.c1[i0] + P.c2[i0] * 0.5 * (xb + xa));
    else {
      const double mx = 0.5 * (xa + xb);
      const double a_L = P.E1[i0], a_R = P.E2[i0], u_c = u0[i0];
      const double a_c = 0.5 * ((u_c - a_L) + (u_c - a_R));
      if (mx < 0.5) {
        const double xa2b2ab = (xa * xa + xb * xb) + xa * xb;
        u_ave = a_L + ((a_R - a_L) * mx + a_c * (3. * (xb + xa) - 2. * xa2b2ab));
      } else {
        const double Ya = 1. - xa, Yb = 1. - xb, my = 0.5 * (Ya + Yb);
        const double Ya2b2ab = (Ya * Ya + Yb * Yb) + Ya * Yb;
        u_ave = a_R + ((a_L - a_R) * my + a_c * (3. * (Yb + Ya) - 2. * Ya2b2ab));
      }
    }
  } else {
    if (method == INTEGRATION_PCM) u_ave = P.c1[i0];
    else if (method == INTEGRATION_PLM) {
      const double a_L = P.E1[i0], a_R = P.E2[i0], Ya = 1. - xa;
      if (xa < 0.5) u_ave = a_L + xa * (a_R - a_L); else u_ave = a_R + Ya * (a_L - a_R);
    } else {
      const double a_L = P.E1[i0], a_R = P.E2[i0], u_c = u0[i0];
      const double a_c = 3. * ((u_c - a_L) + (u_c - a_R)), Ya = 1. - xa;
      if (xa < 0.5) u_ave = a_L + xa * ((a_R - a_L) + a_c * Ya); else u_ave = a_R + Ya * ((a_L - a_R) + a_c * xa);
    }
  }
  return u_ave;
}

struct Sub { vd h_sub, uh_sub, u_sub, h0_eff; vi isub_src, isrc_start, isrc_end, isrc_max, itgt_start, itgt_end;
  Sub(int n0, int n1) : h_sub(n0 + n1 + 3, 0.), uh_sub(n0 + n1 + 3, 0.), u_sub(n0 + n1 + 3, 0.), h0_eff(n0 + 2, 0.), isub_src(n0 + n1 + 3, 0),
                        isrc_start(n0 + 2, 0), isrc_end(n0 + 2, 0), isrc_max(n0 + 2, 0), itgt_start(n1 + 2, 0), itgt_end(n1 + 2, 0) {} };

// intersect_src_tgt_grids :642-798
void intersect_src_tgt_grids(int n0, const double* h0, int n1, const double* h1, Sub& S) {
  double h0_supply = h0[1], h1_supply = h1[1];
  bool src_has_volume = true, tgt_has_volume = true;
  int i0 = 1, i1 = 1, i_start0 = 1, i_start1 = 1, i_max = 1;
  double dh_max = 0., dh0_eff = 0.;
  S.h_sub[1] = 0.; S.isrc_start[1] = 1; S.isrc_end[1] = 1; S.isrc_max[1] = 1; S.isub_src[1] = 1;
  for (int i_sub = 2; i_sub <= n0 + n1 + 1; ++i_sub) {
    const double dh = fmin2(h0_supply, h1_supply);
    dh0_eff = dh0_eff + fmin2(dh, h0_supply);
    S.isub_src[i_sub] = i0;
    S.h_sub[i_sub] = dh;
    if (dh >= dh_max) { i_max = i_sub; dh_max = dh; }
    if (h0_supply <= h1_supply && src_has_volume) {
      h1_supply = h1_supply - dh;
      S.isrc_start[i0] = i_start0; S.isrc_end[i0] = i_sub; i_start0 = i_sub + 1;
      S.isrc_max[i0] = i_max; i_max = i_sub + 1; dh_max = 0.;
      S.h0_eff[i0] = dh0_eff;
      if (i0 < n0) { i0 = i0 + 1; h0_supply = h0[i0]; dh0_eff = 0.; }
      else { h0_supply = 0.; src_has_volume = false; }
    } else if (h0_supply >= h1_supply && tgt_has_volume) {
      h0_supply = h0_supply - dh;
      S.itgt_start[i1] = i_start1; S.itgt_end[i1] = i_sub; i_start1 = i_sub + 1;
      if (i1 < n1) { i1 = i1 + 1; h1_supply = h1[i1]; }
      else { h1_supply = 0.; tgt_has_volume = false; }
    } else if (src_has_volume) {
      S.h_sub[i_sub] = h0_supply;
      S.isrc_start[i0] = i_start0; S.isrc_end[i0] = i_sub; i_start0 = i_sub + 1;
      S.isrc_max[i0] = i_max; i_max = i_sub + 1; dh_max = 0.;
      S.h0_eff[i0] = dh0_eff;
      if (i0 < n0) { i0 = i0 + 1; h0_supply = h0[i0]; dh0_eff = 0.; }
      else { h0_supply = 0.; src_has_volume = false; }
    } else if (tgt_has_volume) {
      S.h_sub[i_sub] = h1_supply;
      S.itgt_start[i1] = i_start1; S.itgt_end[i1] = i_sub; i_start1 = i_sub + 1;
      if (i1 < n1) { i1 = i1 + 1; h1_supply = h1[i1]; }
      else { h1_supply = 0.; tgt_has_volume = false; }
    }
  }
}

// remap_src_to_sub_grid_om4 :845-958 (om4 = true) / remap_src_to_sub_grid :962-1099 (om4 = false)
void remap_src_to_sub_grid(bool om4, int n0, const double* h0, const double* u0, const Poly& P, int n1, Sub& S, int method,
                           bool force_bounds_in_subcell, double& u02_err) {
  vd u0_min(n0 + 2), u0_max(n0 + 2);
  int i0_last_thick_cell = 0;
  for (int i0 = 1; i0 <= n0; ++i0) {
    u0_min[i0] = fmin2(P.E1[i0], P.E2[i0]); u0_max[i0] = fmax2(P.E1[i0], P.E2[i0]);
    if (h0[i0] > 0.) i0_last_thick_cell = i0;
  }
  double xa = 0., xb, dh0_eff = 0.;
  u02_err = 0.;
  auto one = [&](int i_sub, bool reset) {
    const double dh = S.h_sub[i_sub];
    const int i0 = S.isub_src[i_sub];
    dh0_eff = dh0_eff + dh;
    const double hden = om4 ? S.h0_eff[i0] : h0[i0];
    if (hden > 0.) {
      xb = dh0_eff / hden;
      xb = fmin2(1., xb);
      S.u_sub[i_sub] = average_value_ppoly(u0, P, method, i0, xa, xb);
    } else { xb = 1.; S.u_sub[i_sub] = u0[i0]; }
    if (force_bounds_in_subcell) {
      const double u_orig = S.u_sub[i_sub];
      S.u_sub[i_sub] = fmax2(S.u_sub[i_sub], u0_min[i0]);
      S.u_sub[i_sub] = fmin2(S.u_sub[i_sub], u0_max[i0]);
      u02_err = u02_err + dh * std::fabs(S.u_sub[i_sub] - u_orig);
    }
    S.uh_sub[i_sub] = dh * S.u_sub[i_sub];
    if (reset) {
      if (S.isub_src[i_sub + 1] != i0) { dh0_eff = 0.; xa = 0.; }
      else xa = xb;
    }
  };
  if (om4) {
    S.uh_sub[1] = 0.; S.u_sub[1] = P.E1[1];
    for (int i_sub = 2; i_sub <= n0 + n1; ++i_sub) one(i_sub, true);
    S.u_sub[n0 + n1 + 1] = P.E2[n0];
    S.uh_sub[n0 + n1 + 1] = P.E2[n0] * S.h_sub[n0 + n1 + 1];
  } else {
    for (int i_sub = 1; i_sub <= n0 + n1; ++i_sub) one(i_sub, true);
    one(n0 + n1 + 1, false);
  }
  for (int i0 = 1; i0 <= i0_last_thick_cell; ++i0) {  // adjust_thickest_subcell
    const int i_max = S.isrc_max[i0];
    const double dh_max = S.h_sub[i_max];
    if (dh_max > 0.) {
      double duh = 0.;
      for (int i_sub = S.isrc_start[i0]; i_sub <= S.isrc_end[i0]; ++i_sub) if (i_sub != i_max) duh = duh + S.uh_sub[i_sub];
      S.uh_sub[i_max] = u0[i0] * h0[i0] - duh;
      u02_err = u02_err + max3(std::fabs(S.uh_sub[i_max]), std::fabs(u0[i0] * h0[i0]), std::fabs(duh));
    }
  }
}

// remap_sub_to_tgt_grid_om4 :1103-1163
void remap_sub_to_tgt_grid_om4(int n1, const double* h1, const Sub& S, bool force_bounds_in_target, double* u1, double& uh_err) {
  double u1min = 0., u1max = 0.;
  uh_err = 0.;
  for (int i1 = 1; i1 <= n1; ++i1) {
    if (h1[i1] > 0.) {
      double duh = 0., dh = 0.;
      int i_sub = S.itgt_start[i1];
      if (force_bounds_in_target) { u1min = S.u_sub[i_sub]; u1max = S.u_sub[i_sub]; }
      for (i_sub = S.itgt_start[i1]; i_sub <= S.itgt_end[i1]; ++i_sub) {
        if (force_bounds_in_target) { u1min = fmin2(u1min, S.u_sub[i_sub]); u1max = fmax2(u1max, S.u_sub[i_sub]); }
        dh = dh + S.h_sub[i_sub];
        duh = duh + S.uh_sub[i_sub];
        uh_err = uh_err + fmax2(std::fabs(duh), std::fabs(S.uh_sub[i_sub])) * DBL_EPSILON;
      }
      u1[i1] = duh / dh;
      uh_err = uh_err + std::fabs(duh) * DBL_EPSILON;
      if (force_bounds_in_target) {
        const double u_orig = u1[i1];
        u1[i1] = fmax2(u1min, fmin2(u1max, u1[i1]));
        uh_err = uh_err + dh * std::fabs(u1[i1] - u_orig);
      }
    } else u1[i1] = S.u_sub[S.itgt_start[i1]];
  }
}

// remapping_core_h :234-335 (OM4-era reconstruction functions branch); h0,u0,h1,u1 1-based
void remapping_core_h(const mom6cu_remapping_cs* CS, int n0, const double* h0, const double* u0, int n1, const double* h1, double* u1,
                      double* net_err) {
  Sub S(n0, n1);
  intersect_src_tgt_grids(n0, h0, n1, h1, S);
  Poly P(n0);
  const int iMethod = build_reconstructions_1d(CS, n0, h0, u0, P);
  double u02_err, uh_err;
  remap_src_to_sub_grid(CS->om4_remap_via_sub_cells != 0, n0, h0, u0, P, n1, S, iMethod, CS->force_bounds_in_subcell != 0, u02_err);
  remap_sub_to_tgt_grid_om4(n1, h1, S, CS->force_bounds_in_target != 0, u1, uh_err);
  if (net_err) *net_err = uh_err + u02_err;
}

}  // namespace

extern "C" {

// ---- entry points for the known-answer tests (0-based C arrays in, copied to the 1-based work arrays)
int oracle_remapping_core_h(const mom6cu_remapping_cs* CS, int n0, const double* h0, const double* u0, int n1, const double* h1,
                            double* u1, double* net_err) {
  if (CS->answer_date < 20190101) return 3;   // only the answer_date >= 20190101 expressions are restated
  vd H0(n0 + 2), U0(n0 + 2), H1(n1 + 2), U1(n1 + 2);
  for (int k = 0; k < n0; ++k) { H0[k + 1] = h0[k]; U0[k + 1] = u0[k]; }
  for (int k = 0; k < n1; ++k) H1[k + 1] = h1[k];
  remapping_core_h(CS, n0, H0.data(), U0.data(), n1, H1.data(), U1.data(), net_err);
  for (int k = 0; k < n1; ++k) u1[k] = U1[k + 1];
  return 0;
}

int oracle_remap_intersect(int n0, const double* h0, int n1, const double* h1, double* h_sub, double* h0_eff, int* isrc_start,
                           int* isrc_end, int* isrc_max, int* itgt_start, int* itgt_end, int* isub_src) {
  vd H0(n0 + 2), H1(n1 + 2);
  for (int k = 0; k < n0; ++k) H0[k + 1] = h0[k];
  for (int k = 0; k < n1; ++k) H1[k + 1] = h1[k];
  Sub S(n0, n1);
  intersect_src_tgt_grids(n0, H0.data(), n1, H1.data(), S);
  for (int k = 0; k < n0 + n1 + 1; ++k) { h_sub[k] = S.h_sub[k + 1]; isub_src[k] = S.isub_src[k + 1]; }
  for (int k = 0; k < n0; ++k) { h0_eff[k] = S.h0_eff[k + 1]; isrc_start[k] = S.isrc_start[k + 1]; isrc_end[k] = S.isrc_end[k + 1]; isrc_max[k] = S.isrc_max[k + 1]; }
  for (int k = 0; k < n1; ++k) { itgt_start[k] = S.itgt_start[k + 1]; itgt_end[k] = S.itgt_end[k + 1]; }
  return 0;
}

// which: 0 PCM_reconstruction, 1 PLM_reconstruction, 2 PLM_reconstruction + PLM_boundary_extrapolation,
//        3 edge_values_explicit_h4, 4 PPM_reconstruction (edge values in E are input), 5 edge_values_implicit_h4.
// E is (2,N) and coefs (3,N), row-major.
int oracle_remap_reconstruct(int which, int N, const double* h, const double* u, double h_neglect, double* E, double* coefs) {
  vd H(N + 2), U(N + 2);
  for (int k = 0; k < N; ++k) { H[k + 1] = h[k]; U[k + 1] = u[k]; }
  Poly P(N);
  for (int k = 0; k < N; ++k) { P.E1[k + 1] = E[k]; P.E2[k + 1] = E[N + k]; }
  switch (which) {
    case 0: PCM_reconstruction(N, U.data(), P); break;
    case 1: PLM_reconstruction(N, H.data(), U.data(), P, h_neglect); break;
    case 2: PLM_reconstruction(N, H.data(), U.data(), P, h_neglect); PLM_boundary_extrapolation(N, H.data(), U.data(), P, h_neglect); break;
    case 3: edge_values_explicit_h4(N, H.data(), U.data(), P, h_neglect); break;
    case 4: PPM_reconstruction(N, H.data(), U.data(), P, h_neglect); break;
    case 5: edge_values_implicit_h4(N, H.data(), U.data(), P, h_neglect); break;
    default: return 2;
  }
  for (int k = 0; k < N; ++k) { E[k] = P.E1[k + 1]; E[N + k] = P.E2[k + 1]; coefs[k] = P.c1[k + 1]; coefs[N + k] = P.c2[k + 1]; coefs[2 * N + k] = P.c3[k + 1]; }
  return 0;
}

// PLM reconstruction + boundary extrapolation, then remap_src_to_sub_grid[_om4] and remap_sub_to_tgt_grid_om4 (KAT tests 3-5)
int oracle_remap_plm_sub(int om4, int n0, const double* h0, const double* u0, int n1, const double* h1, double h_neglect, double* u_sub,
                         double* u1) {
  vd H0(n0 + 2), U0(n0 + 2), H1(n1 + 2), U1(n1 + 2);
  for (int k = 0; k < n0; ++k) { H0[k + 1] = h0[k]; U0[k + 1] = u0[k]; }
  for (int k = 0; k < n1; ++k) H1[k + 1] = h1[k];
  Sub S(n0, n1);
  intersect_src_tgt_grids(n0, H0.data(), n1, H1.data(), S);
  Poly P(n0);
  PLM_reconstruction(n0, H0.data(), U0.data(), P, h_neglect);
  PLM_boundary_extrapolation(n0, H0.data(), U0.data(), P, h_neglect);
  double e1, e2;
  remap_src_to_sub_grid(om4 != 0, n0, H0.data(), U0.data(), P, n1, S, INTEGRATION_PLM, false, e1);
  remap_sub_to_tgt_grid_om4(n1, H1.data(), S, false, U1.data(), e2);
  for (int k = 0; k < n0 + n1 + 1; ++k) u_sub[k] = S.u_sub[k + 1];
  for (int k = 0; k < n1; ++k) u1[k] = U1[k + 1];
  return 0;
}

// ALE_remap_tracers column loop, MOM_ALE.F90:806-826: remap one h-point field from h_old to h_new where mask2dT > 0
int oracle_ale_remap_scalar(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_remapping_cs* CS, const double* h_old,
                            const double* h_new, double* field, double conc_underflow, int nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  const OGrid G(d, Gp);
  const int nz = G.ke;
  const V3 ho = G.H3(h_old), hn = G.H3(h_new), t = G.H3(field);
#pragma omp parallel for
  for (int j = G.jsc; j <= G.jec; ++j) {
    vd h1(nz + 2), h2(nz + 2), u0(nz + 2), col(nz + 2);
    for (int i = G.isc; i <= G.iec; ++i) if (G.mask2dT(i, j) > 0.) {
      for (int k = 1; k <= nz; ++k) { h1[k] = ho(i, j, k); h2[k] = hn(i, j, k); u0[k] = t(i, j, k); }
      remapping_core_h(CS, nz, h1.data(), u0.data(), nz, h2.data(), col.data(), nullptr);
      if (conc_underflow > 0.0) for (int k = 1; k <= nz; ++k) if (std::fabs(col[k]) < conc_underflow) col[k] = 0.0;
      for (int k = 1; k <= nz; ++k) t(i, j, k) = col[k];
    }
  }
  return 0;
}

// ALE_remap_set_h_vel, MOM_ALE.F90:882-925 (no partial cells, no OBCs)
int oracle_ale_remap_set_h_vel(const mom6cu_domain* d, const mom6cu_grid* Gp, const double* h_new, double* h_up, double* h_vp) {
  const OGrid G(d, Gp);
  const V3 hn = G.H3(h_new), h_u = G.U3(h_up), h_v = G.V3_(h_vp);
  for (int k = 1; k <= G.ke; ++k) {
    for (int j = G.jsc; j <= G.jec; ++j) for (int I = G.IscB; I <= G.IecB; ++I) if (G.mask2dCu(I, j) > 0.)
      h_u(I, j, k) = 0.5 * (hn(I, j, k) + hn(I + 1, j, k));
    for (int J = G.JscB; J <= G.JecB; ++J) for (int i = G.isc; i <= G.iec; ++i) if (G.mask2dCv(i, J) > 0.)
      h_v(i, J, k) = 0.5 * (hn(i, J, k) + hn(i, J + 1, k));
  }
  return 0;
}

// ALE_remap_velocities, MOM_ALE.F90:1089-1300 (no KE-conserving correction, no near-bottom masking, no diagnostics)
int oracle_ale_remap_velocities(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_remapping_cs* CS, const double* h_old_u,
                                const double* h_old_v, const double* h_new_u, const double* h_new_v, double* up, double* vp, int nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  const OGrid G(d, Gp);
  const int nz = G.ke;
  const V3 hou = G.U3(h_old_u), hov = G.V3_(h_old_v), hnu = G.U3(h_new_u), hnv = G.V3_(h_new_v), u = G.U3(up), v = G.V3_(vp);
#pragma omp parallel for
  for (int j = G.jsc; j <= G.jec; ++j) {
    vd h1(nz + 2), h2(nz + 2), u0(nz + 2), col(nz + 2);
    for (int I = G.IscB; I <= G.IecB; ++I) if (G.mask2dCu(I, j) > 0.) {
      for (int k = 1; k <= nz; ++k) { h1[k] = hou(I, j, k); h2[k] = hnu(I, j, k); u0[k] = u(I, j, k); }
      remapping_core_h(CS, nz, h1.data(), u0.data(), nz, h2.data(), col.data(), nullptr);
      for (int k = 1; k <= nz; ++k) u(I, j, k) = col[k];
    }
  }
#pragma omp parallel for
  for (int J = G.JscB; J <= G.JecB; ++J) {
    vd h1(nz + 2), h2(nz + 2), u0(nz + 2), col(nz + 2);
    for (int i = G.isc; i <= G.iec; ++i) if (G.mask2dCv(i, J) > 0.) {
      for (int k = 1; k <= nz; ++k) { h1[k] = hov(i, J, k); h2[k] = hnv(i, J, k); u0[k] = v(i, J, k); }
      remapping_core_h(CS, nz, h1.data(), u0.data(), nz, h2.data(), col.data(), nullptr);
      for (int k = 1; k <= nz; ++k) v(i, J, k) = col[k];
    }
  }
  return 0;
}

// remap_dyn_split_RK2_aux_vars, MOM_dynamics_split_RK2.F90:1302-1331 (CS%remap_aux true; single tile halo fill)
int oracle_remap_dyn_split_rk2_aux_vars(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_remapping_cs* rCS,
                                        const mom6cu_dyn_split_rk2_cs* CS, const double* h_old_u, const double* h_old_v,
                                        const double* h_new_u, const double* h_new_v, int nthreads) {
  const OGrid G(d, Gp);
  auto fill = [&](double* f, int st) {
    const int su = (st == 1), sv = (st == 2);
    const size_t plane = (size_t)(d->ied - d->isd + 1 + su) * (d->jed - d->jsd + 1 + sv);
    for (int k = 0; k < G.ke; ++k) oracle_fill_halo_2d(d, f + plane * k, st, 0);
  };
  if (CS->store_CAu) {
    oracle_ale_remap_velocities(d, Gp, rCS, h_old_u, h_old_v, h_new_u, h_new_v, CS->u_av, CS->v_av, nthreads);
    oracle_ale_remap_velocities(d, Gp, rCS, h_old_u, h_old_v, h_new_u, h_new_v, CS->CAu_pred, CS->CAv_pred, nthreads);
    fill(CS->u_av, 1); fill(CS->v_av, 2); fill(CS->CAu_pred, 1); fill(CS->CAv_pred, 2);
  }
  return oracle_ale_remap_velocities(d, Gp, rCS, h_old_u, h_old_v, h_new_u, h_new_v, CS->diffu, CS->diffv, nthreads);
}

}  // extern "C"

// The column edge-value routines the pressure force uses (ALE_PLM_edge_values / one field of TS_PPM_edge_values), for the known-answer tests
// (tests/test_oracle_remap_kat.py): scheme 1 = PLM, 2 = PPM (implicit h4 edge values).  0-based arrays of length nk.
extern "C" int oracle_ale_edge_values(int scheme, int nk, const double* h, const double* Q, int bdry_extrap, double h_neglect, double* Q_t,
                                      double* Q_b) {
  if (nk < 2 || (scheme == 2 && nk < 4)) return 2;
  std::vector<double> hh(nk + 2), qq(nk + 2), qt(nk + 2), qb(nk + 2);
  for (int k = 1; k <= nk; ++k) { hh[k] = h[k - 1]; qq[k] = Q[k - 1]; }
  if (scheme == 1) orc::ale_plm_edge_values_column(nk, hh.data(), qq.data(), bdry_extrap != 0, h_neglect, qt.data(), qb.data());
  else orc::ale_ppm_edge_values_column(nk, hh.data(), qq.data(), bdry_extrap != 0, h_neglect, h_neglect, qt.data(), qb.data());
  for (int k = 1; k <= nk; ++k) { Q_t[k - 1] = qt[k]; Q_b[k - 1] = qb[k]; }
  return 0;
}
