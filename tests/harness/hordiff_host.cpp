// TEST HARNESS ONLY: compiles the host/device code of mom6_b200/csrc/hordiff_cell.cuh as plain C++ and runs it over a tile in the
// library's unified plane layout (tests/test_tracer_hordiff.py).  Not part of the product: nothing in mom6_b200/ loads this.
#include "../../mom6_b200/csrc/hordiff_cell.cuh"

static m6hd::Par par_of(const double* p) {
  m6hd::Par P;
  P.dt = p[0]; P.Idt = p[1]; P.h_neglect = p[2]; P.KhTr = p[3]; P.KhTr_min = p[4]; P.KhTr_max = p[5]; P.pass_coeff = p[6]; P.pass_min = p[7];
  P.max_diff_CFL = p[8]; P.use_VarMix = (int)p[9]; P.Resoln_scaled = (int)p[10];
  P.use_Eady = (int)p[11]; P.use_MEKE = (int)p[12]; P.Slope_Cff = p[13]; P.KhTr_fac = p[14];
  return P;
}
// box = {is, ie, js, je, i0, j0}; every field a plane of rows x pitch doubles with idx(i,j) = (j - j0)*pitch + (i - i0)
extern "C" double hd_host_khdt(const double* par, const int* box, long long pitch, const double* dy_Cu, const double* IdxCu, const double* dx_Cv,
                               const double* IdyCv, const double* areaT, const double* IareaT, const double* Res_fn_h, const double* Rd_dx_h,
                               double* khdt_x, double* khdt_y, const double* L2u, const double* SN_u, const double* L2v, const double* SN_v,
                               const double* MEKE_Kh) {
  const m6hd::Par P = par_of(par);
  const int is = box[0], ie = box[1], js = box[2], je = box[3], i0 = box[4], j0 = box[5];
  auto idx = [&](int i, int j) { return (long long)(j - j0) * pitch + (i - i0); };
  for (int j = js; j <= je; ++j) for (int i = is - 1; i <= ie; ++i) khdt_x[idx(i, j)] = m6hd::khdt_face(P, idx(i, j), 1, dy_Cu, IdxCu, areaT, Res_fn_h, Rd_dx_h, L2u, SN_u, MEKE_Kh);
  for (int j = js - 1; j <= je; ++j) for (int i = is; i <= ie; ++i) khdt_y[idx(i, j)] = m6hd::khdt_face(P, idx(i, j), pitch, dx_Cv, IdyCv, areaT, Res_fn_h, Rd_dx_h, L2v, SN_v, MEKE_Kh);
  double max_CFL = 0.0;
  for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
    double c = m6hd::cfl_cell(idx(i, j), pitch, khdt_x, khdt_y, IareaT);
    if (!(c > 0.0)) c = 0.0;
    max_CFL = (c > max_CFL) ? c : max_CFL;
  }
  return max_CFL;
}
// one Jacobi sweep of one tracer: df_x / df_y (may be null) are incremented from T, then Tnew = T + dTr on the computational domain
extern "C" void hd_host_sweep(const double* par, const int* box, long long pitch, long long plane, int nk, double scale, double underflow,
                              const double* h, const double* T, const double* khdt_x, const double* khdt_y, const double* IareaT, double* Tnew,
                              double* df_x, double* df_y) {
  const m6hd::Par P = par_of(par);
  const int is = box[0], ie = box[1], js = box[2], je = box[3], i0 = box[4], j0 = box[5];
  auto idx = [&](int i, int j) { return (long long)(j - j0) * pitch + (i - i0); };
  for (int k = 0; k < nk; ++k) {
    const long long ok = (long long)k * plane;
    if (df_x) for (int j = js; j <= je; ++j) for (int i = is - 1; i <= ie; ++i)
      df_x[idx(i, j) + ok] = df_x[idx(i, j) + ok] + m6hd::dflux_face(P, scale, idx(i, j), idx(i, j) + ok, 1, h, T, khdt_x);
    if (df_y) for (int j = js - 1; j <= je; ++j) for (int i = is; i <= ie; ++i)
      df_y[idx(i, j) + ok] = df_y[idx(i, j) + ok] + m6hd::dflux_face(P, scale, idx(i, j), idx(i, j) + ok, pitch, h, T, khdt_y);
    for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      double t = m6hd::diffuse_cell(P, scale, idx(i, j), idx(i, j) + ok, pitch, h, T, khdt_x, khdt_y, IareaT);
      if (underflow > 0.0 && fabs(t) < underflow) t = 0.0;
      Tnew[idx(i, j) + ok] = t;
    }
  }
}
