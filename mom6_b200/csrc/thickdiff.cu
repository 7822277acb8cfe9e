// thickness_diffuse -> thickness_diffuse_full (src/parameterizations/lateral/MOM_thickness_diffuse.F90:134-1670; MOM.F90:1388, the
// isopycnal-height (GM) diffusion that changes h, uhtr, vhtr between the dynamics step and the tracer step) on the device.
//  * td_column_kernel: one thread per column of (isc-1:iec+1, jsc-1:jec+1): interface heights (find_eta), running sums of the volume
//    available to a face, interface pressures, and the vertically smoothed T, S of vert_fill_TS (a tridiagonal solve per column).
//  * td_face_kernel<DIR>: one thread per velocity face, one bottom-up sweep: density derivatives at the interface, the neutral slope,
//    the unlimited streamfunction and the transport limited by the available volumes; the layer-1 transport closes the column.
//  * td_update_kernel: h -= dt * IareaT * div(uhD, vhD), floored at Angstrom_H.
// The column / face / update code is host/device code in thickdiff_column.cuh (checked against the oracle on the host).
#include "ctx.h"
#include "thickdiff_column.cuh"
#include <cmath>

using m6::Geom;

namespace {

using TdP = m6td::Par;
struct TdBox { int is, ie, js, je; };

__global__ void __launch_bounds__(128) td_column_kernel(const Geom G, const TdP P, const TdBox B, const double* __restrict__ h,
                                                        const double* __restrict__ T_in, const double* __restrict__ S_in,
                                                        const double* __restrict__ p_surf, const double* __restrict__ areaT,
                                                        const double* __restrict__ bathyT, double* __restrict__ e, double* __restrict__ pres,
                                                        double* __restrict__ rsum, double* __restrict__ h_frac, double* __restrict__ Tf,
                                                        double* __restrict__ Sf, double* __restrict__ c1) {
  const int i = (B.is - 1) + blockIdx.x * blockDim.x + threadIdx.x, j = (B.js - 1) + blockIdx.y;
  if (i > B.ie + 1 || j > B.je + 1) return;
  m6td::column(P, G.idx(i, j), G.plane, h, T_in, S_in, p_surf, areaT, bathyT, e, pres, rsum, h_frac, Tf, Sf, c1);
}

template <int DIR>
__global__ void __launch_bounds__(128) td_face_kernel(const Geom G, const TdP P, const TdBox B, const double* __restrict__ h,
                                                      const double* __restrict__ e, const double* __restrict__ pres,
                                                      const double* __restrict__ rsum, const double* __restrict__ h_frac,
                                                      const double* __restrict__ T, const double* __restrict__ S,
                                                      const double* __restrict__ areaT, const double* __restrict__ IdC,
                                                      const double* __restrict__ lenC, const double* __restrict__ IdxC,
                                                      const double* __restrict__ IdyC, const double* __restrict__ Res_fn, double* __restrict__ hD,
                                                      double* __restrict__ htr, double* __restrict__ hGM) {
  const int i = (DIR == 0 ? B.is - 1 : B.is) + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (DIR == 0 ? B.js : B.js - 1) + blockIdx.y;
  if (i > B.ie || j > B.je) return;
  m6td::face(P, G.idx(i, j), (DIR == 0) ? 1 : G.pitch, G.plane, h, e, pres, rsum, h_frac, T, S, areaT, IdC, lenC, IdxC, IdyC, Res_fn, hD, htr, hGM);
}

// the same face with stored slopes / the FGNV streamfunction / the MEKE diffusivity (m6td::face_ext): a kernel of its own, so that the
// default selection keeps its register budget
template <int DIR>
__global__ void __launch_bounds__(128) td_face_ext_kernel(const Geom G, const TdP P, const TdBox B, const double* __restrict__ h,
                                                          const double* __restrict__ e, const double* __restrict__ pres,
                                                          const double* __restrict__ rsum, const double* __restrict__ h_frac,
                                                          const double* __restrict__ T, const double* __restrict__ S,
                                                          const double* __restrict__ areaT, const double* __restrict__ IdC,
                                                          const double* __restrict__ lenC, const double* __restrict__ IdxC,
                                                          const double* __restrict__ IdyC, const double* __restrict__ Res_fn,
                                                          const double* __restrict__ maskC, const double* __restrict__ slope,
                                                          const double* __restrict__ cg1, const double* __restrict__ MEKE_Kh, double* __restrict__ sfn_s,
                                                          double* __restrict__ ratio_s, double* __restrict__ hN2_s, double* __restrict__ c2_s,
                                                          double* __restrict__ c1_s, double* __restrict__ hD, double* __restrict__ htr,
                                                          double* __restrict__ hGM) {
  const int i = (DIR == 0 ? B.is - 1 : B.is) + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (DIR == 0 ? B.js : B.js - 1) + blockIdx.y;
  if (i > B.ie || j > B.je) return;
  m6td::face_ext(P, G.idx(i, j), (DIR == 0) ? 1 : G.pitch, G.plane, h, e, pres, rsum, h_frac, T, S, areaT, IdC, lenC, IdxC, IdyC, Res_fn, maskC, slope, cg1,
                 MEKE_Kh, sfn_s, ratio_s, hN2_s, c2_s, c1_s, hD, htr, hGM);
}

__global__ void td_update_kernel(const Geom G, const TdP P, const TdBox B, const double* __restrict__ uhD, const double* __restrict__ vhD,
                                 const double* __restrict__ IareaT, double* __restrict__ h) {
  const int i = B.is + blockIdx.x * blockDim.x + threadIdx.x, j = B.js + blockIdx.y;
  if (i > B.ie || j > B.je) return;
  const long long g = G.idx(i, j);
  m6td::update(P, g, g + (long long)blockIdx.z * G.plane, G.pitch, uhD, vhD, IareaT, h);
}

}  // namespace

extern "C" int mom6cu_thickness_diffuse(mom6cu_ctx* c, const mom6cu_thickness_diffuse_cs* CS, const mom6cu_thickness_diffuse_args* a) {
  if (!c || !CS || !a) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "thickness_diffuse: mom6cu_set_grid / mom6cu_set_vgrid have not been called");
  const mom6cu_vgrid& GV = c->vgrid;
  if (CS->read_khth || CS->detangle_interfaces || CS->interface_Kh || CS->use_stanley_gm || CS->use_GME_thickness_diffuse || CS->find_work ||
      CS->Depth_scaled_KhTh || CS->use_Visbeck || CS->use_QG_Leith_GM || CS->khth_struct)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "thickness_diffuse: KHTH (+ MEKE%%Kh) with the CFL / min / max / resolution-function limits, slopes from T and S or "
                                           "stored, and the limited or FGNV streamfunction are implemented (no Visbeck / QG Leith, vertical structure, depth "
                                           "scaling, detangling, KH_ETA, Stanley, GM work, 2-d KHTH from a file)");
  if (CS->use_FGNV_streamfn && !a->cg1) return c->fail(MOM6CU_ERR_BAD_ARG, "cg1 must be associated when using FGNV streamfunction.");
  if ((CS->use_stored_slopes && (!a->slope_x || !a->slope_y)) || (CS->use_MEKE_Kh && !a->MEKE_Kh))
    return c->fail(MOM6CU_ERR_BAD_ARG, "thickness_diffuse: VarMix%%slope_x / slope_y or MEKE%%Kh is not allocated");
  if (!GV.Boussinesq) return c->fail(MOM6CU_ERR_UNSUPPORTED, "thickness_diffuse: only the Boussinesq branch is implemented");
  if (CS->EOS_form != MOM6CU_EOS_LINEAR && CS->EOS_form != MOM6CU_EOS_WRIGHT)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "thickness_diffuse: an equation of state (LINEAR or WRIGHT) is required (the constant-density-layer branch is not implemented)");
  if (!CS->thickness_diffuse || !(CS->Khth > 0.0 || CS->use_variable_mixing)) return 0;  // :196-198
  if (!a->h || !a->uhtr || !a->vhtr || !a->T || !a->S) return c->fail(MOM6CU_ERR_BAD_ARG, "thickness_diffuse: null required argument");
  const bool Resoln_scaled = CS->use_variable_mixing && CS->Resoln_scaled_KhTh;
  if (Resoln_scaled && (!a->Res_fn_u || !a->Res_fn_v)) return c->fail(MOM6CU_ERR_BAD_ARG, "thickness_diffuse: VarMix%%Res_fn_u / Res_fn_v are not allocated");
  const Geom& G = c->g;
  const mom6cu_domain& d = c->dom;
  const GridDev& Gd = c->grid;
  const int nz = G.nk;
  if (nz < 2 || !(a->dt > 0.0)) return c->fail(MOM6CU_ERR_BAD_ARG, "thickness_diffuse: needs at least two layers and dt > 0");
  if (std::min(std::min(d.isc - d.isd, d.ied - d.iec), std::min(d.jsc - d.jsd, d.jed - d.jec)) < 1)
    return c->fail(MOM6CU_ERR_BAD_ARG, "find_eta called with an overly large halo_size.");
  Stager S(c, "td.");
  int rc;
  double *d_h, *d_uhtr, *d_vhtr, *d_ugm = nullptr, *d_vgm = nullptr;
  const double *d_T, *d_S, *d_ps = nullptr, *d_ru = nullptr, *d_rv = nullptr;
  if ((rc = S.io3(a->h, ST_H, "h", &d_h)) || (rc = S.io3(a->uhtr, ST_U, "uhtr", &d_uhtr)) || (rc = S.io3(a->vhtr, ST_V, "vhtr", &d_vhtr)) ||
      (rc = S.in3(a->T, ST_H, "T", &d_T)) || (rc = S.in3(a->S, ST_H, "S", &d_S)) || (rc = S.in2(a->p_surf, ST_H, "p_surf", &d_ps)) ||
      (rc = S.in2(Resoln_scaled ? a->Res_fn_u : nullptr, ST_U, "Res_fn_u", &d_ru)) || (rc = S.in2(Resoln_scaled ? a->Res_fn_v : nullptr, ST_V, "Res_fn_v", &d_rv)))
    return rc;
  const bool ext = CS->use_stored_slopes || CS->use_FGNV_streamfn || CS->use_MEKE_Kh;
  const double *d_sx = nullptr, *d_sy = nullptr, *d_cg = nullptr, *d_kh = nullptr;
  if (CS->use_stored_slopes && ((rc = S.in(a->slope_x, ST_U, 0, nz + 1, "slope_x", &d_sx)) || (rc = S.in(a->slope_y, ST_V, 0, nz + 1, "slope_y", &d_sy)))) return rc;
  if (CS->use_FGNV_streamfn && (rc = S.in2(a->cg1, ST_H, "cg1", &d_cg))) return rc;
  if (CS->use_MEKE_Kh && (rc = S.in2(a->MEKE_Kh, ST_H, "MEKE_Kh", &d_kh))) return rc;
  if (a->uhGM && (rc = S.io3(a->uhGM, ST_U, "uhGM", &d_ugm))) return rc;
  if (a->vhGM && (rc = S.io3(a->vhGM, ST_V, "vhGM", &d_vgm))) return rc;
  double *e = c->plane3k("td.e", nz + 1), *pres = c->plane3k("td.pres", nz + 1), *rsum = c->plane3k("td.rsum", nz + 1);
  double *hfr = c->plane3("td.h_frac"), *Tf = c->plane3("td.Tf"), *Sf = c->plane3("td.Sf"), *c1 = c->plane3("td.c1");
  double *uhD = c->plane3("td.uhD"), *vhD = c->plane3("td.vhD");
  if (!e || !pres || !rsum || !hfr || !Tf || !Sf || !c1 || !uhD || !vhD) return MOM6CU_ERR_CUDA;
  double *xs = nullptr, *xr = nullptr, *xn = nullptr, *xc2 = nullptr, *xc1 = nullptr;  // face_ext scratch
  if (ext) {
    xs = c->plane3k("td.x_sfn", nz + 1); xr = c->plane3k("td.x_ratio", nz + 1); xn = c->plane3k("td.x_hN2", nz + 1);
    xc2 = c->plane3k("td.x_c2", nz + 1); xc1 = c->plane3k("td.x_c1", nz + 1);
    if (!xs || !xr || !xn || !xc2 || !xc1) return MOM6CU_ERR_CUDA;
  }
  TdP P = {};
  P.nk = nz; P.eos_form = CS->EOS_form; P.Resoln_scaled = Resoln_scaled ? 1 : 0; P.have_p_surf = a->p_surf ? 1 : 0;
  P.dt = a->dt; P.I4dt = 0.25 / a->dt; P.Angstrom_H = GV.Angstrom_H; P.h_neglect = GV.H_subroundoff; P.h_neglect2 = GV.H_subroundoff * GV.H_subroundoff;
  P.dz_neglect = CS->dZ_subroundoff; P.H_to_Z = GV.H_to_Z; P.Z_to_H = GV.Z_to_H; P.g_H_to_RZ = GV.g_Earth * GV.H_to_RZ; P.Z_to_L = c->US.Z_to_L;
  P.Khth = CS->Khth; P.Khth_Min = CS->Khth_Min; P.Khth_Max = CS->Khth_Max; P.max_Khth_CFL = CS->max_Khth_CFL;
  P.I_slope_max2 = 1.0 / (CS->slope_max * CS->slope_max);
  const double kappa_dt = CS->kappa_smooth * a->dt;
  P.kap_dt_x2 = (2.0 * kappa_dt) * (c->US.Z_to_m * GV.m_to_H);
  P.h0 = 1.0e-16 * sqrt(0.5 * P.kap_dt_x2);  // larger_h_denom = .true.  (MOM_isopycnal_slopes.F90:656-659)
  P.dRho_dT = CS->dRho_dT; P.dRho_dS = CS->dRho_dS;
  P.stored_slopes = CS->use_stored_slopes ? 1 : 0; P.FGNV = CS->use_FGNV_streamfn ? 1 : 0; P.use_MEKE_Kh = CS->use_MEKE_Kh ? 1 : 0;
  P.G_rho0 = GV.g_Earth / GV.Rho0; P.dz_neglect2 = CS->dZ_subroundoff * CS->dZ_subroundoff; P.N2_floor = CS->N2_floor; P.FGNV_scale = CS->FGNV_scale;
  P.KhTh_fac = CS->MEKE_KhTh_fac;
  const TdBox B = {d.isc, d.iec, d.jsc, d.jec};
  if ((rc = S.begin())) return rc;
  const int ni = d.iec - d.isc + 1, nj = d.jec - d.jsc + 1;
  M6_LAUNCH(c, td_column_kernel, dim3((ni + 2 + 127) / 128, nj + 2), 128, 0, G, P, B, d_h, d_T, d_S, d_ps, Gd.areaT, Gd.bathyT, e, pres, rsum, hfr, Tf, Sf, c1);
  if (!ext) {
    M6_LAUNCH(c, td_face_kernel<0>, dim3((ni + 1 + 127) / 128, nj), 128, 0, G, P, B, d_h, e, pres, rsum, hfr, Tf, Sf, Gd.areaT, Gd.IdxCu, Gd.dy_Cu, Gd.IdxCu,
              Gd.IdyCu, d_ru, uhD, d_uhtr, d_ugm);
    M6_LAUNCH(c, td_face_kernel<1>, dim3((ni + 127) / 128, nj + 1), 128, 0, G, P, B, d_h, e, pres, rsum, hfr, Tf, Sf, Gd.areaT, Gd.IdyCv, Gd.dx_Cv, Gd.IdxCv,
              Gd.IdyCv, d_rv, vhD, d_vhtr, d_vgm);
  } else {
    M6_LAUNCH(c, td_face_ext_kernel<0>, dim3((ni + 1 + 127) / 128, nj), 128, 0, G, P, B, d_h, e, pres, rsum, hfr, Tf, Sf, Gd.areaT, Gd.IdxCu, Gd.dy_Cu,
              Gd.IdxCu, Gd.IdyCu, d_ru, Gd.mask2dCu, d_sx, d_cg, d_kh, xs, xr, xn, xc2, xc1, uhD, d_uhtr, d_ugm);
    M6_LAUNCH(c, td_face_ext_kernel<1>, dim3((ni + 127) / 128, nj + 1), 128, 0, G, P, B, d_h, e, pres, rsum, hfr, Tf, Sf, Gd.areaT, Gd.IdyCv, Gd.dx_Cv,
              Gd.IdxCv, Gd.IdyCv, d_rv, Gd.mask2dCv, d_sy, d_cg, d_kh, xs, xr, xn, xc2, xc1, vhD, d_vhtr, d_vgm);
  }
  M6_LAUNCH(c, td_update_kernel, dim3((ni + 127) / 128, nj, nz), 128, 0, G, P, B, uhD, vhD, Gd.IareaT, d_h);
  M6_CUDA(c, cudaGetLastError());
  return S.finish();
}
