// Streaming form of the one-column ALE remapping of remap_column.cuh (remapping_core_h,
// /root/reference/src/ALE/MOM_remapping.F90:234-335) for the PCM / PLM / PPM_H4 reconstructions: the same arithmetic in
// the same order, but organised around the source cell that is being consumed, so that a thread keeps only scalars and
// one output column instead of ~9 KB of column arrays (whose local-memory traffic made the array form run at 11x its
// algorithmic bytes).
//
//  * The sub-cell sequence of intersect_src_tgt_grids (:642-798) is a merge of the two interface lists; its state is six
//    scalars, so it is *replayed* instead of stored.  For each source cell the merge is run three times from the cell's
//    checkpoint: (A) to find h0_eff, the thickest sub-cell and whether it has volume; (B) to evaluate the sub-cell
//    averages and the k-ordered sum that adjust_thickest_subcell needs (:919-937 / :1060-1078); (C) to feed the target
//    accumulators of remap_sub_to_tgt_grid_om4 (:1103-1163) in sub-cell order with the adjusted value in place.
//  * The reconstruction of a source cell (edge_values_explicit_h4 + bound_edge_values + check_discontinuous_edge_values
//    + PPM_limiter_standard [+ PPM_boundary_extrapolation], or PLM_reconstruction [+ PLM_boundary_extrapolation]) is a
//    function of a 5- to 6-cell window of (h, u), so it is evaluated when the walk enters the cell.
// Every value is produced by the same expression as in remap_column.cuh; tests/test_remap_column_host.py compares both
// with the oracle bit for bit.  PPM_IH4 needs a column-wide tridiagonal solve and stays on the array form.
#pragma once
#include "remap_column.cuh"

namespace m6remap {

// The merge state of intersect_src_tgt_grids before sub-cell i_sub is formed
struct Merge {
  double h0s, h1s;     // h0_supply, h1_supply
  int i0, i1, i_sub;   // current source / target cell, index of the NEXT sub-cell to form
  bool src, tgt;       // src_has_volume, tgt_has_volume
};
struct Step { double dh, hs; int i0; bool src_retire, tgt_retire; };

// one pass of the loop body of intersect_src_tgt_grids (:691-790); H0 / H1 give h0(k) / h1(k)
template <class FH0, class FH1>
M6R_HD Step merge_step(Merge& M, int n0, int n1, const FH0& H0, const FH1& H1) {
  Step s;
  const double dh = rmin(M.h0s, M.h1s);
  s.dh = dh; s.hs = dh; s.i0 = M.i0; s.src_retire = false; s.tgt_retire = false;
  const bool src_step = (M.h0s <= M.h1s && M.src);
  const bool tgt_step = !src_step && (M.h0s >= M.h1s && M.tgt);
  const bool src_tail = !src_step && !tgt_step && M.src;
  const bool tgt_tail = !src_step && !tgt_step && !src_tail && M.tgt;
  if (src_step || src_tail) {
    if (src_step) M.h1s = M.h1s - dh; else s.hs = M.h0s;
    s.src_retire = true;
    if (M.i0 < n0) { M.i0 = M.i0 + 1; M.h0s = H0(M.i0); }
    else { M.h0s = 0.; M.src = false; }
  } else if (tgt_step || tgt_tail) {
    if (tgt_step) M.h0s = M.h0s - dh; else s.hs = M.h1s;
    s.tgt_retire = true;
    if (M.i1 < n1) { M.i1 = M.i1 + 1; M.h1s = H1(M.i1); }
    else { M.h1s = 0.; M.tgt = false; }
  }
  M.i_sub = M.i_sub + 1;
  return s;
}

// ---- reconstructions of one cell from the column accessors --------------------------------------------------------
struct CellRec { double u, E1, E2, c2; };

// interface value e(i) between cells i-1 and i, 3 <= i <= N-1: edge_values_explicit_h4, regrid_edge_values.F90:240-270
template <class FH, class FU>
M6R_HD double edge_h4(int i, const FH& H, const FU& U, double h_neglect) {
  const double hMinFrac = 1.e-5;
  double h0 = H(i - 2), h1 = H(i - 1), h2 = H(i), h3 = H(i + 1);
  if (h0 + h1 == 0.0 || h1 + h2 == 0.0 || h2 + h3 == 0.0) {
    const double h_min = hMinFrac * rmax(h_neglect, (h0 + h1) + (h2 + h3));
    h0 = rmax(h_min, H(i - 2)); h1 = rmax(h_min, H(i - 1)); h2 = rmax(h_min, H(i)); h3 = rmax(h_min, H(i + 1));
  }
  const double I_h12 = 1.0 / (h1 + h2);
  const double I_den_et2 = 1.0 / (((h0 + h1) + h2) * (h0 + h1)), I_h012 = (h0 + h1) * I_den_et2;
  const double I_den_et3 = 1.0 / ((h1 + (h2 + h3)) * (h2 + h3)), I_h123 = (h2 + h3) * I_den_et3;
  const double et1 = (1.0 + (h1 * I_h012 + (h0 + h1) * I_h123)) * I_h12 * (h2 * (h2 + h3)) * U(i - 1) +
                     (1.0 + (h2 * I_h123 + (h2 + h3) * I_h012)) * I_h12 * (h1 * (h0 + h1)) * U(i);
  const double et2 = (h1 * (h2 * (h2 + h3)) * I_den_et2) * (U(i - 1) - U(i - 2));
  const double et3 = (h2 * (h1 * (h0 + h1)) * I_den_et3) * (U(i) - U(i + 1));
  return (et1 + (et2 + et3)) / ((h0 + h1) + (h2 + h3));
}

struct Ends { double top0, top1, bot0, bot1; };  // E1(1), E2(1)=E1(2);  E2(N), E1(N)=E2(N-1)

template <class FH, class FU>
M6R_HD Ends end_values_acc(int N, const FH& H, const FU& U, double h_neglect) {
  double dz[5], ut[5], C[5];
  Ends e;
  for (int i = 1; i <= 4; ++i) { dz[i] = rmax(h_neglect, H(i)); ut[i] = U(i); }
  end_value_h4(dz, ut, C);
  e.top0 = C[1];
  e.top1 = C[1] + dz[1] * (C[2] + dz[1] * (C[3] + dz[1] * C[4]));
  for (int i = 1; i <= 4; ++i) { dz[i] = rmax(h_neglect, H(N + 1 - i)); ut[i] = U(N + 1 - i); }
  end_value_h4(dz, ut, C);
  e.bot0 = C[1];
  e.bot1 = C[1] + dz[1] * (C[2] + dz[1] * (C[3] + dz[1] * C[4]));
  return e;
}

// raw interface value e(i), 1 <= i <= N+1 (N >= 4)
template <class FH, class FU>
M6R_HD double raw_edge(int i, int N, const Ends& e, const FH& H, const FU& U, double h_neglect_edge) {
  if (i == 1) return e.top0;
  if (i == 2) return e.top1;
  if (i == N + 1) return e.bot0;
  if (i == N) return e.bot1;
  return edge_h4(i, H, U, h_neglect_edge);
}

// bound_edge_values (regrid_edge_values.F90:39-105) for cell k: bounded (E1, E2) from the raw ones
template <class FH, class FU>
M6R_HD void bound_cell(int k, int N, double e1, double e2, const FH& H, const FU& U, double& E1, double& E2) {
  const int km1 = (k - 1 > 1) ? k - 1 : 1, kp1 = (k + 1 < N) ? k + 1 : N;
  const double um = U(km1), uc = U(k), up = U(kp1), hm = H(km1), hc = H(k), hp = H(kp1);
  double slope_x_h = 0.0;
  if (((hm + hp) + 2.0 * hc) > 0.0) {
    const double sigma_l = (uc - um);
    const double sigma_c = (up - um) * (hc / ((hm + hp) + 2.0 * hc));
    const double sigma_r = (up - uc);
    if ((sigma_l * sigma_r) > 0.0) slope_x_h = rsign(rmin3(fabs(sigma_l), fabs(sigma_c), fabs(sigma_r)), sigma_c);
  }
  if ((um - e1) * (e1 - uc) < 0.0) e1 = uc - rsign(rmin(fabs(slope_x_h), fabs(e1 - uc)), slope_x_h);
  if ((up - e2) * (e2 - uc) < 0.0) e2 = uc + rsign(rmin(fabs(slope_x_h), fabs(e2 - uc)), slope_x_h);
  E1 = rmax(rmin(e1, rmax(um, uc)), rmin(um, uc));
  E2 = rmax(rmin(e2, rmax(up, uc)), rmin(up, uc));
}

// edges of cell k after bound_edge_values and check_discontinuous_edge_values (:143-165), 1 <= k <= N
template <class FH, class FU>
M6R_HD void bounded_continuous_cell(int k, int N, const Ends& e, const FH& H, const FU& U, double h_neglect_edge, double& E1, double& E2) {
  const double eL = raw_edge(k, N, e, H, U, h_neglect_edge), eR = raw_edge(k + 1, N, e, H, U, h_neglect_edge);
  double b1, b2;
  bound_cell(k, N, eL, eR, H, U, b1, b2);
  E1 = b1; E2 = b2;
  if (k >= 2) {  // pair (k-1, k): E2b(k-1) against E1b(k)
    const double eLL = raw_edge(k - 1, N, e, H, U, h_neglect_edge);
    double m1, m2;
    bound_cell(k - 1, N, eLL, eL, H, U, m1, m2);
    if ((b1 - m2) * (U(k) - U(k - 1)) < 0.0) {
      double avg = 0.5 * (m2 + b1);
      avg = rmax(rmin(avg, rmax(U(k - 1), U(k))), rmin(U(k - 1), U(k)));
      E1 = avg;
    }
  }
  if (k <= N - 1) {  // pair (k, k+1): E2b(k) against E1b(k+1)
    const double eRR = raw_edge(k + 2, N, e, H, U, h_neglect_edge);
    double p1, p2;
    bound_cell(k + 1, N, eR, eRR, H, U, p1, p2);
    if ((p1 - b2) * (U(k + 1) - U(k)) < 0.0) {
      double avg = 0.5 * (b2 + p1);
      avg = rmax(rmin(avg, rmax(U(k), U(k + 1))), rmin(U(k), U(k + 1)));
      E2 = avg;
    }
  }
}

// final PPM edges of an interior cell 2 <= k <= N-1: PPM_limiter_standard, PPM_functions.F90:75-112
template <class FH, class FU>
M6R_HD void ppm_interior_cell(int k, int N, const Ends& e, const FH& H, const FU& U, double h_neglect_edge, double& E1, double& E2) {
  double edge_l, edge_r;
  bounded_continuous_cell(k, N, e, H, U, h_neglect_edge, edge_l, edge_r);
  const double u_l = U(k - 1), u_c = U(k), u_r = U(k + 1);
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
  E1 = edge_l; E2 = edge_r;
}

// slp(k) and mslp(k) of PLM_reconstruction (PLM_functions.F90:214-232); 0 outside 2..N-1
template <class FH, class FU>
M6R_HD double plm_slp(int k, int N, const FH& H, const FU& U, double h_neglect) {
  return (k >= 2 && k <= N - 1) ? PLM_slope_wa(H(k - 1), H(k), H(k + 1), h_neglect, U(k - 1), U(k), U(k + 1)) : 0.;
}
template <class FH, class FU>
M6R_HD double plm_mslp(int k, int N, const FH& H, const FU& U, double h_neglect) {
  if (!(k >= 2 && k <= N - 1)) return 0.;
  return PLM_monotonized_slope(U(k - 1), U(k), U(k + 1), plm_slp(k - 1, N, H, U, h_neglect), plm_slp(k, N, H, U, h_neglect),
                               plm_slp(k + 1, N, H, U, h_neglect));
}

// The PPM_H4 reconstruction visited in cell order: the bounded edge pairs of cells k-1, k, k+1 and the raw interface
// value e(k+2) are carried from one cell to the next, so each interface value and each bound_edge_values evaluation is
// computed once per column instead of once per neighbour (same expressions, same results as ppm_interior_cell).
struct PpmWin {
  int k;                               // the interior cell the window is centred on (2 <= k <= N-1), 0 = not started
  double Bm1, Bm2, Bc1, Bc2, Bp1, Bp2;  // bounded (E1, E2) of cells k-1, k, k+1
  double e_next;                        // raw e(k+2)
  double f1, f2;                        // final edges of cell k
};

template <class FH, class FU>
M6R_HD void ppm_win_finish(PpmWin& W, const FU& U) {
  const int k = W.k;
  double edge_l = W.Bc1, edge_r = W.Bc2;
  if ((W.Bc1 - W.Bm2) * (U(k) - U(k - 1)) < 0.0) {
    double avg = 0.5 * (W.Bm2 + W.Bc1);
    avg = rmax(rmin(avg, rmax(U(k - 1), U(k))), rmin(U(k - 1), U(k)));
    edge_l = avg;
  }
  if ((W.Bp1 - W.Bc2) * (U(k + 1) - U(k)) < 0.0) {
    double avg = 0.5 * (W.Bc2 + W.Bp1);
    avg = rmax(rmin(avg, rmax(U(k), U(k + 1))), rmin(U(k), U(k + 1)));
    edge_r = avg;
  }
  const double u_l = U(k - 1), u_c = U(k), u_r = U(k + 1);
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
  W.f1 = edge_l; W.f2 = edge_r;
}

// centre the window on interior cell k: k = 2 starts it, k = W.k + 1 slides it
template <class FH, class FU>
M6R_HD void ppm_win_goto(PpmWin& W, int k, int N, const Ends& e, const FH& H, const FU& U, double h_neglect_edge) {
  if (W.k == k) return;
  if (W.k + 1 == k && W.k >= 2) {
    W.Bm1 = W.Bc1; W.Bm2 = W.Bc2; W.Bc1 = W.Bp1; W.Bc2 = W.Bp2;
    const double e_new = raw_edge(k + 2, N, e, H, U, h_neglect_edge);
    bound_cell(k + 1, N, W.e_next, e_new, H, U, W.Bp1, W.Bp2);
    W.e_next = e_new;
  } else {
    const double e0 = raw_edge(k - 1, N, e, H, U, h_neglect_edge), e1 = raw_edge(k, N, e, H, U, h_neglect_edge),
                 e2 = raw_edge(k + 1, N, e, H, U, h_neglect_edge), e3 = raw_edge(k + 2, N, e, H, U, h_neglect_edge);
    bound_cell(k - 1, N, e0, e1, H, U, W.Bm1, W.Bm2);
    bound_cell(k, N, e1, e2, H, U, W.Bc1, W.Bc2);
    bound_cell(k + 1, N, e2, e3, H, U, W.Bp1, W.Bp2);
    W.e_next = e3;
  }
  W.k = k;
  ppm_win_finish<FH, FU>(W, U);
}

// the reconstruction of source cell k as build_reconstructions_1d (MOM_remapping.F90:410-550) leaves it
template <class FH, class FU>
M6R_HD CellRec recon_cell(const Params& P, int scheme, int k, int N, const Ends& e, const FH& H, const FU& U, PpmWin& W) {
  CellRec r;
  r.u = U(k); r.c2 = 0.;
  if (scheme == SCHEME_PCM) { r.E1 = r.u; r.E2 = r.u; return r; }
  if (scheme == SCHEME_PLM) {
    const double almost_one = 1. - DBL_EPSILON;
    if (k >= 2 && k <= N - 1) {
      const double slope = plm_mslp(k, N, H, U, P.h_neglect);
      const double u_l = r.u - 0.5 * slope, u_r = r.u + 0.5 * slope;
      double c2 = (u_r - u_l);
      const double edge = c2 + u_l;
      const double e_r = U(k + 1) - 0.5 * rsign(plm_mslp(k + 1, N, H, U, P.h_neglect), plm_slp(k + 1, N, H, U, P.h_neglect));
      if ((edge - r.u) * (e_r - edge) < 0.) c2 = c2 * almost_one;
      r.E1 = u_l; r.E2 = u_r; r.c2 = c2;
    } else if (!P.boundary_extrapolation) { r.E1 = r.u; r.E2 = r.u; r.c2 = 0.; }
    else if (k == 1) {
      const double slope = -PLM_extrapolate_slope(H(2), H(1), P.h_neglect, U(2), U(1));
      r.E1 = r.u - 0.5 * slope; r.E2 = r.u + 0.5 * slope; r.c2 = r.E2 - r.E1;
    } else {
      const double slope = PLM_extrapolate_slope(H(N - 1), H(N), P.h_neglect, U(N - 1), U(N));
      r.E1 = r.u - 0.5 * slope; r.E2 = r.u + 0.5 * slope; r.c2 = r.E2 - r.E1;
    }
    return r;
  }
  // PPM_H4 (N >= 4)
  if (k >= 2 && k <= N - 1) { ppm_win_goto(W, k, N, e, H, U, P.h_neglect_edge); r.E1 = W.f1; r.E2 = W.f2; return r; }
  if (!P.boundary_extrapolation) { r.E1 = r.u; r.E2 = r.u; return r; }
  if (k == 1) {  // PPM_boundary_extrapolation, top (PPM_functions.F90:196-240)
    ppm_win_goto(W, 2, N, e, H, U, P.h_neglect_edge);
    const double E1n = W.f1, E2n = W.f2;
    const double h0 = H(1), h1 = H(2), u0 = U(1), u1 = U(2);
    const double b = 4.0 * (u1 - E1n) + 2.0 * (u1 - E2n);
    double u1_r = b * ((h0 + P.h_neglect) / (h1 + P.h_neglect));
    const double slope = 2.0 * (u1 - u0);
    if (fabs(u1_r) > fabs(slope)) u1_r = slope;
    double u0_r = E1n;
    double u0_l = 3.0 * u0 + 0.5 * u1_r - 2.0 * u0_r;
    const double exp1 = (u0_r - u0_l) * (u0 - 0.5 * (u0_l + u0_r));
    const double exp2 = (u0_r - u0_l) * (u0_r - u0_l) / 6.0;
    if (exp1 > exp2) u0_l = 3.0 * u0 - 2.0 * u0_r;
    if (exp1 < -exp2) u0_r = 3.0 * u0 - 2.0 * u0_l;
    r.E1 = u0_l; r.E2 = u0_r;
    return r;
  }
  {  // bottom (:242-296)
    ppm_win_goto(W, N - 1, N, e, H, U, P.h_neglect_edge);  // already there when the cells are visited in order
    const double E1n = W.f1, E2n = W.f2;
    const double h0 = H(N - 1), h1 = H(N), u0 = U(N - 1), u1 = U(N);
    const double b = 4.0 * (u0 - E1n) + 2.0 * (u0 - E2n);
    const double c = 3.0 * ((E2n - u0) + (E1n - u0));
    double u1_l = (b + 2 * c);
    u1_l = u1_l * ((h1 + P.h_neglect) / (h0 + P.h_neglect));
    const double slope = 2.0 * (u1 - u0);
    if (fabs(u1_l) > fabs(slope)) u1_l = slope;
    double u0_l = E2n;
    double u0_r = 3.0 * u1 - 0.5 * u1_l - 2.0 * u0_l;
    const double exp1 = (u0_r - u0_l) * (u1 - 0.5 * (u0_l + u0_r));
    const double exp2 = (u0_r - u0_l) * (u0_r - u0_l) / 6.0;
    if (exp1 > exp2) u0_l = 3.0 * u1 - 2.0 * u0_r;
    if (exp1 < -exp2) u0_r = 3.0 * u1 - 2.0 * u0_l;
    r.E1 = u0_l; r.E2 = u0_r;
    return r;
  }
}

// average_value_ppoly (MOM_remapping.F90:1391-1490) on a CellRec
M6R_HD double average_rec(const CellRec& R, int method, double xa, double xb) {
  const double a_L = R.E1, a_R = R.E2, u_c = R.u;
  if (xb > xa) {
    if (method == INT_PCM) return u_c;
    if (method == INT_PLM) return (a_L + R.c2 * 0.5 * (xb + xa));
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
  if (method == INT_PCM) return a_L;
  const double Ya = 1. - xa;
  if (method == INT_PLM) return (xa < 0.5) ? a_L + xa * (a_R - a_L) : a_R + Ya * (a_L - a_R);
  const double a_c = 3. * ((u_c - a_L) + (u_c - a_R));
  return (xa < 0.5) ? a_L + xa * ((a_R - a_L) + a_c * Ya) : a_R + Ya * ((a_L - a_R) + a_c * xa);
}

// the position of the walk inside a source cell (remap_src_to_sub_grid's xa, dh0_eff)
struct Walk { double xa, dh0; };

// the value of one sub-cell (:875-905 / :1000-1043)
M6R_HD double sub_value(const Params& P, const CellRec& R, int method, double hden, double dh, Walk& W, bool same_cell_next) {
  W.dh0 = W.dh0 + dh;
  double xb, us;
  if (hden > 0.) {
    xb = W.dh0 / hden;
    xb = rmin(1., xb);
    us = average_rec(R, method, W.xa, xb);
  } else { xb = 1.; us = R.u; }
  if (P.force_bounds_in_subcell) {
    us = rmax(us, rmin(R.E1, R.E2));
    us = rmin(us, rmax(R.E1, R.E2));
  }
  if (same_cell_next) W.xa = xb; else { W.dh0 = 0.; W.xa = 0.; }
  return us;
}

// the accumulators of the target cell being assembled (remap_sub_to_tgt_grid_om4 :1117-1160)
struct Tgt { double duh, dh, umin, umax, ufirst; bool open; };

template <class Out>
M6R_HD void tgt_feed(const Params& P, Tgt& T, double us, double uh, double hs, bool retire, int i1, double h1, Out& out, double underflow) {
  if (!T.open) { T.open = true; T.duh = 0.; T.dh = 0.; T.umin = us; T.umax = us; T.ufirst = us; }
  T.umin = rmin(T.umin, us); T.umax = rmax(T.umax, us);
  T.dh = T.dh + hs;
  T.duh = T.duh + uh;
  if (retire) {
    double r;
    if (h1 > 0.) {
      r = T.duh / T.dh;
      if (P.force_bounds_in_target) r = rmax(T.umin, rmin(T.umax, r));
    } else r = T.ufirst;
    if (underflow > 0.0 && fabs(r) < underflow) r = 0.0;
    out(i1) = r;
    T.open = false;
  }
}

// remapping_core_h for one column and one field, streaming.  H0, U0, H1 are accessors k -> value (1-based); out(i1) is
// assigned once per target cell, in increasing i1, possibly before all of U0 has been read (it must not alias U0).
template <class FH0, class FU0, class FH1, class Out>
M6R_HD void remap_stream(const Params& P, int n0, int n1, const FH0& H0, const FU0& U0, const FH1& H1, Out out, double underflow) {
  int scheme = P.scheme;
  if (n0 <= 1) scheme = SCHEME_PCM;
  else if (n0 <= 3) scheme = (scheme < SCHEME_PLM) ? scheme : (int)SCHEME_PLM;
  else if (n0 <= 4) scheme = (scheme < SCHEME_PPM_H4) ? scheme : (int)SCHEME_PPM_H4;
  const int method = (scheme == SCHEME_PCM) ? INT_PCM : (scheme == SCHEME_PLM) ? INT_PLM : INT_PPM;
  Ends ends = {0., 0., 0., 0.};
  if (scheme == SCHEME_PPM_H4) ends = end_values_acc(n0, H0, U0, P.h_neglect_edge);
  const bool om4 = P.om4 != 0;
  const int ns = n0 + n1 + 1;
  int last_thick = 0;
  for (int k = 1; k <= n0; ++k) if (H0(k) > 0.) last_thick = k;

  Merge M = {H0(1), H1(1), 1, 1, 2, true, true};  // sub-cell 1 is the zero-thickness one at the top of source cell 1
  Tgt T = {0., 0., 0., 0., 0., false};
  Walk W = {0., 0.};
  PpmWin PW = {0, 0., 0., 0., 0., 0., 0., 0., 0., 0.};
  CellRec R = recon_cell(P, scheme, 1, n0, ends, H0, U0, PW);
  double hden_last = 0.;
  for (int i0 = 1; i0 <= n0; ++i0) {
    if (i0 > 1) R = recon_cell(P, scheme, i0, n0, ends, H0, U0, PW);
    const double h0c = H0(i0);
    // ---- (A) look ahead over this cell's sub-cells: h0_eff, the thickest sub-cell, the last index
    Merge A = M;
    double h0_eff = 0., dh_max = 0., hs_max = 0.;
    int i_max = (i0 == 1) ? 1 : A.i_sub;  // sub-cell 1 (h_sub = 0) is replaced as soon as the first real sub-cell is formed
    for (;;) {
      const int is = A.i_sub;
      const double h0s_before = A.h0s;
      const Step s = merge_step(A, n0, n1, H0, H1);
      h0_eff = h0_eff + rmin(s.dh, h0s_before);
      if (s.dh >= dh_max) { i_max = is; dh_max = s.dh; hs_max = s.hs; }
      if (s.src_retire) break;
    }
    const double hden = om4 ? h0_eff : h0c;
    hden_last = hden;
    const bool adjust = (i0 <= last_thick) && (hs_max > 0.);
    // ---- (B) the sum of the other sub-cells' uh, in sub-cell order
    double uh_adj = 0.;
    if (adjust) {
      Merge B = M; Walk WB = {0., 0.};
      double duh = 0.;
      if (i0 == 1) {  // sub-cell 1
        double uh1;
        if (om4) uh1 = 0.;
        else { const double us1 = sub_value(P, R, method, hden, 0., WB, true); uh1 = 0. * us1; }
        if (i_max != 1) duh = duh + uh1;
      }
      for (;;) {
        const int is = B.i_sub;
        const Step s = merge_step(B, n0, n1, H0, H1);
        double us;
        if (om4 && is == ns) us = R.E2;  // the OM4 variant takes the bottom edge value for the last sub-cell (:912-913)
        else us = sub_value(P, R, method, hden, s.hs, WB, !s.src_retire);
        if (is != i_max) duh = duh + s.hs * us;
        if (s.src_retire) break;
      }
      uh_adj = R.u * h0c - duh;
    }
    // ---- (C) feed the targets
    if (i0 == 1) {
      double us1, uh1;
      if (om4) { us1 = R.E1; uh1 = 0.; }
      else { us1 = sub_value(P, R, method, hden, 0., W, true); uh1 = 0. * us1; }
      if (adjust && i_max == 1) uh1 = uh_adj;
      tgt_feed(P, T, us1, uh1, 0., false, M.i1, H1(M.i1), out, underflow);
    }
    for (;;) {
      const int is = M.i_sub, i1 = M.i1;
      const Step s = merge_step(M, n0, n1, H0, H1);
      double us, uh;
      if (om4 && is == ns) { us = R.E2; uh = R.E2 * s.hs; }
      else { us = sub_value(P, R, method, hden, s.hs, W, !(s.src_retire && i0 < n0)); uh = s.hs * us; }
      if (adjust && is == i_max) uh = uh_adj;
      tgt_feed(P, T, us, uh, s.hs, s.tgt_retire, i1, s.tgt_retire ? H1(i1) : 0., out, underflow);
      if (s.src_retire) break;
    }
  }
  // ---- the sub-cells below the last source cell (isub_src = n0: the walk continues inside cell n0, no adjustment)
  while (M.i_sub <= ns) {
    const int is = M.i_sub, i1 = M.i1;
    const Step s = merge_step(M, n0, n1, H0, H1);
    double us, uh;
    if (om4 && is == ns) { us = R.E2; uh = R.E2 * s.hs; }
    else { us = sub_value(P, R, method, hden_last, s.hs, W, true); uh = s.hs * us; }
    tgt_feed(P, T, us, uh, s.hs, s.tgt_retire, i1, s.tgt_retire ? H1(i1) : 0., out, underflow);
  }
}

}  // namespace m6remap
