// TEST HARNESS ONLY: compiles the host/device code of mom6_b200/csrc/mle_mu.cuh and mle_column.cuh as plain C++ and runs it over a
// tile in the library's unified plane layout (tests/test_mle.py).  Not part of the product: nothing in mom6_b200/ loads this.
#include "../../mom6_b200/csrc/mle_column.cuh"
extern "C" double mle_host_mu(double sigma, double dh) { return m6mle::mu(sigma, dh); }
extern "C" double mle_host_density(int form, double r0, double dT, double dS, double dp, double T, double S, double p) {
  const m6mle::Eos E = {form, r0, dT, dS, dp};
  return m6mle::density(E, T, S, p);
}
// par: the m6mle::Par members in declaration order, ints as doubles (nk, dt ... tail_dh, aFac1..bFac2, filt1, filt2, res_upscale,
// eos form, Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp).  Every field is a plane of `rows` x `pitch` doubles with the SAME offset
// idx(i,j) = (j - j0)*pitch + (i - i0) whatever its staggering; 3-D fields are nk consecutive planes.  box = {is, ie, js, je, i0, j0}.
extern "C" void mle_host_run(const double* par, const int* box, long long pitch, long long plane, double* h, double* uhtr, double* vhtr,
                             const double* T, const double* S, const double* ustar, const double* h_MLD, const double* Rd_dx_h,
                             double* MLD_filtered, double* MLD_filtered_slow, const double* areaT, const double* IareaT,
                             const double* CoriolisBu, const double* mask2dCu, const double* mask2dCv, const double* dxCu,
                             const double* dyCu, const double* dxCv, const double* dyCv, const double* IdxCu, const double* IdyCv,
                             double* scratch /* 4 planes + 2*nk planes */) {
  m6mle::Par P;
  int n = 0;
  P.nk = (int)par[n++];
  P.dt = par[n++]; P.Z_to_H = par[n++]; P.Angstrom_H = par[n++]; P.h_neglect = par[n++]; P.g_Rho0 = par[n++]; P.I4dt = par[n++];
  P.h_min = par[n++]; P.vonKar_x_pi2 = par[n++]; P.ustar_min = par[n++]; P.coef = par[n++]; P.coef2 = par[n++]; P.front_length = par[n++];
  P.stretch = par[n++]; P.tail_dh = par[n++]; P.aFac1 = par[n++]; P.bFac1 = par[n++]; P.aFac2 = par[n++]; P.bFac2 = par[n++];
  P.filt1 = (int)par[n++]; P.filt2 = (int)par[n++]; P.res_upscale = (int)par[n++];
  P.eos.form = (int)par[n++]; P.eos.Rho_T0_S0 = par[n++]; P.eos.dRho_dT = par[n++]; P.eos.dRho_dS = par[n++]; P.eos.dRho_dp = par[n++];
  P.detect = (int)par[n++]; P.density_diff = par[n++];
  const int is = box[0], ie = box[1], js = box[2], je = box[3], i0 = box[4], j0 = box[5];
  auto idx = [&](int i, int j) { return (long long)(j - j0) * pitch + (i - i0); };
  double *hf = scratch, *hs = scratch + plane, *rf = scratch + 2 * plane, *rs = scratch + 3 * plane;
  double *uhml = scratch + 4 * plane, *vhml = uhml + (long long)P.nk * plane;
  for (int j = js - 1; j <= je + 1; ++j)
    for (int i = is - 1; i <= ie + 1; ++i) m6mle::column(P, idx(i, j), plane, h, T, S, h_MLD, MLD_filtered, MLD_filtered_slow, hf, hs, rf, rs);
  for (int j = js; j <= je; ++j)
    for (int i = is - 1; i <= ie; ++i)
      m6mle::face(P, idx(i, j), 1, pitch, plane, h, areaT, ustar, Rd_dx_h, hf, hs, rf, rs, CoriolisBu, mask2dCu, dxCu, dyCu, dyCu, IdxCu, uhml, uhtr);
  for (int j = js - 1; j <= je; ++j)
    for (int i = is; i <= ie; ++i)
      m6mle::face(P, idx(i, j), pitch, 1, plane, h, areaT, ustar, Rd_dx_h, hf, hs, rf, rs, CoriolisBu, mask2dCv, dxCv, dyCv, dxCv, IdyCv, vhml, vhtr);
  for (int k = 0; k < P.nk; ++k)
    for (int j = js; j <= je; ++j)
      for (int i = is; i <= ie; ++i) m6mle::update(P, idx(i, j), idx(i, j) + (long long)k * plane, pitch, uhml, vhml, IareaT, h);
}
