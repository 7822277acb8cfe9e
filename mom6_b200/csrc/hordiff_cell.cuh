// The per-face and per-cell work of tracer_hordiff's along-surface path (src/tracer/MOM_tracer_hor_diff.F90:203-340, :354-366,
// :537-604) as host/device code on the unified plane layout (common.cuh): the kernels of hordiff.cu call it with one thread per face
// / cell, tests/harness/hordiff_host.cpp compiles the same functions with g++ and loops over the tile, so the code the GPU threads
// run is checked bit for bit against the oracle without a GPU (tests/test_tracer_hordiff.py).
// The reference's 2-D scratch Coef_x, Coef_y, Ihdxdy (:553-567) is evaluated where a cell needs it (the same expressions, so the same
// bits); the Jacobi update T = T + dTr of a layer (:570-597) writes a second copy of the field.
#pragma once
#include <math.h>
#if defined(__CUDACC__)
#define M6D_HD __host__ __device__ __forceinline__
#else
#define M6D_HD inline
#endif

namespace m6hd {

M6D_HD double fmx(double a, double b) { return (a > b) ? a : b; }
M6D_HD double fmn(double a, double b) { return (a < b) ? a : b; }

struct Par {
  double dt, Idt, h_neglect, KhTr, KhTr_min, KhTr_max, pass_coeff, pass_min, max_diff_CFL;
  int use_VarMix, Resoln_scaled;
  int use_Eady, use_MEKE;  // KhTr_Slope_Cff > 0 with VarMix; allocated(MEKE%Kh) with VarMix
  double Slope_Cff, KhTr_fac;
};

// khdt_x(I,j) | khdt_y(i,J) :204-327.  g: plane offset of the face and of its western / southern cell, sd: offset to the other cell;
// lenC = G%dy_Cu | G%dx_Cv, IdC = G%IdxCu | G%IdyCv.
M6D_HD double khdt_face(const Par& P, const long long g, const long long sd, const double* lenC, const double* IdC, const double* areaT,
                        const double* Res_fn_h, const double* Rd_dx_h, const double* L2, const double* SN, const double* MEKE_Kh) {
  double khdt;
  if (P.use_VarMix) {
    double Kh_loc = P.KhTr;
    if (P.use_Eady) Kh_loc = Kh_loc + P.Slope_Cff * L2[g] * SN[g];                          // :208 | :227
    if (P.use_MEKE) Kh_loc = Kh_loc + P.KhTr_fac * sqrt(MEKE_Kh[g] * MEKE_Kh[g + sd]);     // :209-210 | :228-229
    if (P.KhTr_max > 0.) Kh_loc = fmn(Kh_loc, P.KhTr_max);
    if (P.Resoln_scaled) Kh_loc = Kh_loc * 0.5 * (Res_fn_h[g] + Res_fn_h[g + sd]);
    double Kh = fmx(Kh_loc, P.KhTr_min);
    if (P.pass_coeff > 0.) {
      const double Rd_dx = 0.5 * (Rd_dx_h[g] + Rd_dx_h[g + sd]);
      Kh_loc = Kh * fmx(P.pass_min, P.pass_coeff * Rd_dx);
      if (P.KhTr_max > 0.) Kh_loc = fmn(Kh_loc, P.KhTr_max);
      Kh = fmx(Kh_loc, P.KhTr_min);
    }
    khdt = P.dt * (Kh * (lenC[g] * IdC[g]));
  } else {
    khdt = P.dt * (P.KhTr * (lenC[g] * IdC[g]));
  }
  if (P.max_diff_CFL > 0.0) {
    const double khdt_max = 0.125 * P.max_diff_CFL * fmn(areaT[g], areaT[g + sd]);
    khdt = fmn(khdt, khdt_max);
  }
  return khdt;
}

// CFL(i,j) :357-358
M6D_HD double cfl_cell(const long long g, const long long pitch, const double* khdt_x, const double* khdt_y, const double* IareaT) {
  return 2.0 * ((khdt_x[g - 1] + khdt_x[g]) + (khdt_y[g - pitch] + khdt_y[g])) * IareaT[g];
}

// Coef_x(I,j,1) | Coef_y(i,J,1) of one layer :553-561 (hk points at the layer's plane)
M6D_HD double coef_face(const double scale, const double h_neglect, const double khdt, const double h0, const double h1) {
  return ((scale * khdt) * 2.0 * (h0 * h1)) / (h0 + h1 + h_neglect);
}

// T(i,j,k) + dTr(i,j) :570-597 for the cell at offset gk = g + k*plane
M6D_HD double diffuse_cell(const Par& P, const double scale, const long long g, const long long gk, const long long pitch, const double* h,
                           const double* T, const double* khdt_x, const double* khdt_y, const double* IareaT) {
  const double hc = h[gk], hw = h[gk - 1], he = h[gk + 1], hs = h[gk - pitch], hn = h[gk + pitch];
  const double cW = coef_face(scale, P.h_neglect, khdt_x[g - 1], hw, hc), cE = coef_face(scale, P.h_neglect, khdt_x[g], hc, he);
  const double cS = coef_face(scale, P.h_neglect, khdt_y[g - pitch], hs, hc), cN = coef_face(scale, P.h_neglect, khdt_y[g], hc, hn);
  const double Ihdxdy = IareaT[g] / (hc + P.h_neglect);
  const double Tc = T[gk];
  const double dTr = Ihdxdy * (((cW * (T[gk - 1] - Tc)) - (cE * (Tc - T[gk + 1]))) + ((cS * (T[gk - pitch] - Tc)) - (cN * (Tc - T[gk + pitch]))));
  return Tc + dTr;
}

// Reg%Tr(m)%df_x | df_y increment :576-583 for the face at offset gk
M6D_HD double dflux_face(const Par& P, const double scale, const long long g, const long long gk, const long long sd, const double* h,
                         const double* T, const double* khdt) {
  const double c = coef_face(scale, P.h_neglect, khdt[g], h[gk], h[gk + sd]);
  return c * (T[gk] - T[gk + sd]) * P.Idt;
}

}  // namespace m6hd
