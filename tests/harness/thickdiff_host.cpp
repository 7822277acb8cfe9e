// TEST HARNESS ONLY: compiles the host/device code of mom6_b200/csrc/thickdiff_column.cuh as plain C++ and runs it over a tile in the
// library's unified plane layout (tests/test_thickness_diffuse.py).  Not part of the product: nothing in mom6_b200/ loads this.
#include "../../mom6_b200/csrc/thickdiff_column.cuh"
// par: nk, eos_form, Resoln_scaled, have_p_surf (as doubles), then the double members of m6td::Par in declaration order.
// box = {is, ie, js, je, i0, j0}; every field a plane of rows x pitch doubles with idx(i,j) = (j - j0)*pitch + (i - i0).
// scratch: 3 fields of nk+1 planes (e, pres, rsum), 6 of nk planes (h_frac, Tf, Sf, c1, uhD, vhD), then 5 of nk+1 planes for face_ext.
extern "C" void td_host_run(const double* par, const int* box, long long pitch, long long plane, double* h, double* uhtr, double* vhtr,
                            const double* T, const double* S, const double* p_surf, const double* Res_fn_u, const double* Res_fn_v, double* uhGM,
                            double* vhGM, const double* areaT, const double* IareaT, const double* bathyT, const double* IdxCu, const double* IdyCu,
                            const double* dy_Cu, const double* IdxCv, const double* IdyCv, const double* dx_Cv, double* scratch,
                            const double* mask2dCu, const double* mask2dCv, const double* slope_x, const double* slope_y, const double* cg1,
                            const double* MEKE_Kh) {
  m6td::Par P;
  int n = 0;
  P.nk = (int)par[n++]; P.eos_form = (int)par[n++]; P.Resoln_scaled = (int)par[n++]; P.have_p_surf = (int)par[n++];
  P.dt = par[n++]; P.I4dt = par[n++]; P.Angstrom_H = par[n++]; P.h_neglect = par[n++]; P.h_neglect2 = par[n++]; P.dz_neglect = par[n++];
  P.H_to_Z = par[n++]; P.Z_to_H = par[n++]; P.g_H_to_RZ = par[n++]; P.Z_to_L = par[n++];
  P.Khth = par[n++]; P.Khth_Min = par[n++]; P.Khth_Max = par[n++]; P.max_Khth_CFL = par[n++]; P.I_slope_max2 = par[n++]; P.kap_dt_x2 = par[n++]; P.h0 = par[n++];
  P.dRho_dT = par[n++]; P.dRho_dS = par[n++];
  P.stored_slopes = (int)par[n++]; P.FGNV = (int)par[n++]; P.use_MEKE_Kh = (int)par[n++];
  P.G_rho0 = par[n++]; P.dz_neglect2 = par[n++]; P.N2_floor = par[n++]; P.FGNV_scale = par[n++]; P.KhTh_fac = par[n++];
  const bool ext = P.stored_slopes || P.FGNV || P.use_MEKE_Kh;
  const int is = box[0], ie = box[1], js = box[2], je = box[3], i0 = box[4], j0 = box[5];
  auto idx = [&](int i, int j) { return (long long)(j - j0) * pitch + (i - i0); };
  const long long n1 = (long long)(P.nk + 1) * plane, n0 = (long long)P.nk * plane;
  double *e = scratch, *pres = e + n1, *rsum = pres + n1, *hfr = rsum + n1, *Tf = hfr + n0, *Sf = Tf + n0, *c1 = Sf + n0, *uhD = c1 + n0, *vhD = uhD + n0;
  double *sfn_s = vhD + n0, *ratio_s = sfn_s + n1, *hN2_s = ratio_s + n1, *c2_s = hN2_s + n1, *c1_s = c2_s + n1;  // face_ext scratch (5 x (nk+1) planes)
  for (int j = js - 1; j <= je + 1; ++j)
    for (int i = is - 1; i <= ie + 1; ++i) m6td::column(P, idx(i, j), plane, h, T, S, p_surf, areaT, bathyT, e, pres, rsum, hfr, Tf, Sf, c1);
  for (int j = js; j <= je; ++j)
    for (int i = is - 1; i <= ie; ++i)
      if (ext) m6td::face_ext(P, idx(i, j), 1, plane, h, e, pres, rsum, hfr, Tf, Sf, areaT, IdxCu, dy_Cu, IdxCu, IdyCu, Res_fn_u, mask2dCu, slope_x, cg1,
                              MEKE_Kh, sfn_s, ratio_s, hN2_s, c2_s, c1_s, uhD, uhtr, uhGM);
      else m6td::face(P, idx(i, j), 1, plane, h, e, pres, rsum, hfr, Tf, Sf, areaT, IdxCu, dy_Cu, IdxCu, IdyCu, Res_fn_u, uhD, uhtr, uhGM);
  for (int j = js - 1; j <= je; ++j)
    for (int i = is; i <= ie; ++i)
      if (ext) m6td::face_ext(P, idx(i, j), pitch, plane, h, e, pres, rsum, hfr, Tf, Sf, areaT, IdyCv, dx_Cv, IdxCv, IdyCv, Res_fn_v, mask2dCv, slope_y,
                              cg1, MEKE_Kh, sfn_s, ratio_s, hN2_s, c2_s, c1_s, vhD, vhtr, vhGM);
      else m6td::face(P, idx(i, j), pitch, plane, h, e, pres, rsum, hfr, Tf, Sf, areaT, IdyCv, dx_Cv, IdxCv, IdyCv, Res_fn_v, vhD, vhtr, vhGM);
  for (int k = 0; k < P.nk; ++k)
    for (int j = js; j <= je; ++j)
      for (int i = is; i <= ie; ++i) m6td::update(P, idx(i, j), idx(i, j) + (long long)k * plane, pitch, uhD, vhD, IareaT, h);
}
