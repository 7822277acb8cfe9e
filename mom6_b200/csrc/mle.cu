// mixedlayer_restrat -> mixedlayer_restrat_OM4 (src/parameterizations/lateral/MOM_mixed_layer_restrat.F90:149-714): the
// Fox-Kemper et al. (2008) mixed-layer-eddy restratification in general coordinates, which adjusts h, uhtr and vhtr between
// the dynamics step and the tracer advection (MOM.F90:1422) -- on the device so that the state stays resident across it.
//
//  * mle_column_kernel: one thread per column of (isc-1:iec+1, jsc-1:jec+1).  MLD time filters (:316-346) and a top-down sweep
//    that accumulates the mixed-layer thickness and the thickness-weighted surface-referenced density for the fast and the slow
//    filtered depths (:384-410), stopping at the base of the deeper of the two.
//  * mle_face_kernel<DIR>: one thread per velocity face.  The overturning amplitudes uDml / uDml_slow (:464-495), then three
//    top-down passes over the mixed layer -- limit uDml by the available volumes, limit uDml_slow, form the transports (:500-535).
//  * mle_update_kernel: h -= dt * IareaT * div(uhml, vhml), floored at Angstrom_H / 2 (:623-627).
// The column / face / update code is host/device code in mle_column.cuh (checked against the oracle on the host, tests/test_mle.py);
// its header explains what is evaluated differently from the reference's loops and why the bits are the same.
#include "ctx.h"
#include "mle_column.cuh"

using m6::Geom;

namespace {

using MleP = m6mle::Par;
struct MleBox { int is, ie, js, je; };

__global__ void __launch_bounds__(128) mle_column_kernel(const Geom G, const MleP P, const MleBox B, const double* __restrict__ h,
                                                         const double* __restrict__ T, const double* __restrict__ S,
                                                         const double* __restrict__ h_MLD, double* __restrict__ MLD_filtered,
                                                         double* __restrict__ MLD_filtered_slow, double* __restrict__ htot_fast,
                                                         double* __restrict__ htot_slow, double* __restrict__ Rml_av_fast,
                                                         double* __restrict__ Rml_av_slow) {
  const int i = (B.is - 1) + blockIdx.x * blockDim.x + threadIdx.x, j = (B.js - 1) + blockIdx.y;
  if (i > B.ie + 1 || j > B.je + 1) return;
  m6mle::column(P, G.idx(i, j), G.plane, h, T, S, h_MLD, MLD_filtered, MLD_filtered_slow, htot_fast, htot_slow, Rml_av_fast, Rml_av_slow);
}

// DIR 0: u faces (I = isc-1..iec, j = jsc..jec), neighbour cell i+1;  DIR 1: v faces (i = isc..iec, J = jsc-1..jec), neighbour j+1
template <int DIR>
__global__ void __launch_bounds__(128) mle_face_kernel(const Geom G, const MleP P, const MleBox B, const double* __restrict__ h,
                                                       const double* __restrict__ areaT, const double* __restrict__ ustar,
                                                       const double* __restrict__ Rd_dx_h, const double* __restrict__ htot_fast,
                                                       const double* __restrict__ htot_slow, const double* __restrict__ Rml_av_fast,
                                                       const double* __restrict__ Rml_av_slow, const double* __restrict__ CoriolisBu,
                                                       const double* __restrict__ maskC, const double* __restrict__ dxC,
                                                       const double* __restrict__ dyC, const double* __restrict__ Idn /* IdxCu | IdyCv */,
                                                       double* __restrict__ hml, double* __restrict__ htr) {
  const int i = (DIR == 0 ? B.is - 1 : B.is) + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (DIR == 0 ? B.js : B.js - 1) + blockIdx.y;
  if (i > B.ie || j > B.je) return;
  m6mle::face(P, G.idx(i, j), (DIR == 0) ? 1 : G.pitch, (DIR == 0) ? G.pitch : 1, G.plane, h, areaT, ustar, Rd_dx_h, htot_fast, htot_slow,
              Rml_av_fast, Rml_av_slow, CoriolisBu, maskC, dxC, dyC, (DIR == 0) ? dyC : dxC, Idn, hml, htr);
}

__global__ void mle_update_kernel(const Geom G, const MleP P, const MleBox B, const double* __restrict__ uhml, const double* __restrict__ vhml,
                                  const double* __restrict__ IareaT, double* __restrict__ h) {
  const int i = B.is + blockIdx.x * blockDim.x + threadIdx.x, j = B.js + blockIdx.y;
  if (i > B.ie || j > B.je) return;
  const long long g = G.idx(i, j);
  m6mle::update(P, g, g + (long long)blockIdx.z * G.plane, G.pitch, uhml, vhml, IareaT, h);
}

__global__ void mle_mu_kernel(const int n, const double* __restrict__ sigma, const double* __restrict__ dh, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = m6mle::mu(sigma[e], dh[e]);
}

}  // namespace

extern "C" int mom6cu_mle_mu(mom6cu_ctx* c, int n, const double* sigma, const double* dh, double* out) {
  if (!c || n < 1 || !sigma || !dh || !out) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  double* b = c->buf("mle.mu", (size_t)3 * n);
  if (!b) return MOM6CU_ERR_CUDA;
  M6_CUDA(c, cudaMemcpyAsync(b, sigma, sizeof(double) * n, cudaMemcpyDefault, c->stream));
  M6_CUDA(c, cudaMemcpyAsync(b + n, dh, sizeof(double) * n, cudaMemcpyDefault, c->stream));
  M6_LAUNCH(c, mle_mu_kernel, (n + 127) / 128, 128, 0, n, b, b + n, b + 2 * n);
  M6_CUDA(c, cudaGetLastError());
  M6_CUDA(c, cudaMemcpyAsync(out, b + 2 * n, sizeof(double) * n, cudaMemcpyDefault, c->stream));
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int mom6cu_mixedlayer_restrat(mom6cu_ctx* c, mom6cu_mle_cs* CS, double* h, double* uhtr, double* vhtr, const double* T,
                                         const double* S, const double* ustar, double dt, const double* h_MLD, const double* Rd_dx_h) {
  if (!c || !CS || !h || !uhtr || !vhtr || !ustar) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "mixedlayer_restrat: mom6cu_set_grid / mom6cu_set_vgrid have not been called");
  const mom6cu_vgrid& GV = c->vgrid;
  if (!GV.Boussinesq) return c->fail(MOM6CU_ERR_UNSUPPORTED, "mixedlayer_restrat_OM4: only the Boussinesq branch (:375-412) is implemented");
  if (CS->use_Bodner || CS->use_Stanley_ML || CS->fl_from_file)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "mixedlayer_restrat: the Bodner et al. (2023) variant, the Stanley SGS variance and a front length from a file are not implemented");
  if ((CS->EOS_form != MOM6CU_EOS_LINEAR && CS->EOS_form != MOM6CU_EOS_WRIGHT) || !T || !S)
    return c->fail(MOM6CU_ERR_BAD_ARG, "mixedlayer_restrat_OM4: An equation of state must be used with this module.");
  if (CS->front_length > 0. && !Rd_dx_h) return c->fail(MOM6CU_ERR_BAD_ARG, "mixedlayer_restrat_OM4: The resolution argument, Rd/dx, was not associated.");
  const bool detect = CS->MLE_density_diff > 0.;  // detect_mld :298-299 takes precedence over the boundary-layer depth
  if (!detect && (!CS->MLE_use_PBL_MLD || !h_MLD)) return c->fail(MOM6CU_ERR_BAD_ARG, "mixedlayer_restrat_OM4: No MLD to use for MLE parameterization.");
  if (CS->MLE_tail_dh != 0.0)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "mixedlayer_restrat_OM4: MLE_TAIL_DH /= 0 makes mu a real power, which is not bit-reproducible across math libraries");
  if ((CS->MLE_MLD_decay_time > 0. && !CS->MLD_filtered) || (CS->MLE_MLD_decay_time2 > 0. && !CS->MLD_filtered_slow))
    return c->fail(MOM6CU_ERR_BAD_ARG, "mixedlayer_restrat: CS%%MLD_filtered / MLD_filtered_slow is null");
  const Geom& G = c->g;
  const mom6cu_domain& d = c->dom;
  const int nz = G.nk;
  Stager St(c, "mle.");
  int rc;
  double *d_h, *d_uhtr, *d_vhtr, *d_f1 = nullptr, *d_f2 = nullptr;
  const double *d_T, *d_S, *d_us, *d_ml, *d_rd = nullptr;
  if ((rc = St.io3(h, ST_H, "h", &d_h)) || (rc = St.io3(uhtr, ST_U, "uhtr", &d_uhtr)) || (rc = St.io3(vhtr, ST_V, "vhtr", &d_vhtr)) ||
      (rc = St.in3(T, ST_H, "T", &d_T)) || (rc = St.in3(S, ST_H, "S", &d_S)) || (rc = St.in2(ustar, ST_H, "ustar", &d_us)) ||
      (rc = St.in2(detect ? nullptr : h_MLD, ST_H, "h_MLD", &d_ml)) || (rc = St.in2(Rd_dx_h, ST_H, "Rd_dx_h", &d_rd))) return rc;
  if (CS->MLD_filtered && (rc = St.io2(CS->MLD_filtered, ST_H, "MLD_filtered", &d_f1))) return rc;
  if (CS->MLD_filtered_slow && (rc = St.io2(CS->MLD_filtered_slow, ST_H, "MLD_filtered_slow", &d_f2))) return rc;
  double *d_uhml = c->plane3("mle.uhml"), *d_vhml = c->plane3("mle.vhml");
  double *d_hf = c->plane2("mle.htot_fast"), *d_hs = c->plane2("mle.htot_slow"), *d_rf = c->plane2("mle.Rml_av_fast"), *d_rs = c->plane2("mle.Rml_av_slow");
  if (!d_uhml || !d_vhml || !d_hf || !d_hs || !d_rf || !d_rs) return MOM6CU_ERR_CUDA;
  MleP P = {};
  P.nk = nz;
  const MleBox B = {d.isc, d.iec, d.jsc, d.jec};
  P.dt = dt; P.Z_to_H = GV.Z_to_H; P.Angstrom_H = GV.Angstrom_H; P.h_neglect = GV.H_subroundoff;
  P.g_Rho0 = GV.H_to_Z * GV.g_Earth / GV.Rho0;
  P.I4dt = 0.25 / dt; P.h_min = 0.5 * GV.Angstrom_H; P.vonKar_x_pi2 = CS->vonKar * 9.8696; P.ustar_min = CS->ustar_min;
  P.coef = CS->ml_restrat_coef; P.coef2 = CS->ml_restrat_coef2; P.front_length = CS->front_length; P.stretch = CS->MLE_MLD_stretch;
  P.tail_dh = CS->MLE_tail_dh;
  P.filt1 = CS->MLE_MLD_decay_time > 0.; P.filt2 = CS->MLE_MLD_decay_time2 > 0.;
  if (P.filt1) { P.aFac1 = CS->MLE_MLD_decay_time / (dt + CS->MLE_MLD_decay_time); P.bFac1 = dt / (dt + CS->MLE_MLD_decay_time); }
  if (P.filt2) { P.aFac2 = CS->MLE_MLD_decay_time2 / (dt + CS->MLE_MLD_decay_time2); P.bFac2 = dt / (dt + CS->MLE_MLD_decay_time2); }
  P.res_upscale = CS->front_length > 0.;
  P.eos = {CS->EOS_form, CS->Rho_T0_S0, CS->dRho_dT, CS->dRho_dS, CS->dRho_dp};
  P.detect = detect ? 1 : 0; P.density_diff = CS->MLE_density_diff;
  if ((rc = St.begin())) return rc;
  const GridDev& Gd = c->grid;
  {
    const dim3 grid((d.iec - d.isc + 3 + 127) / 128, d.jec - d.jsc + 3);
    M6_LAUNCH(c, mle_column_kernel, grid, 128, 0, G, P, B, d_h, d_T, d_S, d_ml, d_f1, d_f2, d_hf, d_hs, d_rf, d_rs);
  }
  {
    const dim3 gu((d.iec - d.isc + 2 + 127) / 128, d.jec - d.jsc + 1), gv((d.iec - d.isc + 1 + 127) / 128, d.jec - d.jsc + 2);
    M6_LAUNCH(c, mle_face_kernel<0>, gu, 128, 0, G, P, B, d_h, Gd.areaT, d_us, d_rd, d_hf, d_hs, d_rf, d_rs, Gd.CoriolisBu, Gd.mask2dCu, Gd.dxCu, Gd.dyCu,
              Gd.IdxCu, d_uhml, d_uhtr);
    M6_LAUNCH(c, mle_face_kernel<1>, gv, 128, 0, G, P, B, d_h, Gd.areaT, d_us, d_rd, d_hf, d_hs, d_rf, d_rs, Gd.CoriolisBu, Gd.mask2dCv, Gd.dxCv, Gd.dyCv,
              Gd.IdyCv, d_vhml, d_vhtr);
  }
  {
    const dim3 grid((d.iec - d.isc + 1 + 127) / 128, d.jec - d.jsc + 1, nz);
    M6_LAUNCH(c, mle_update_kernel, grid, 128, 0, G, P, B, d_uhml, d_vhml, Gd.IareaT, d_h);
  }
  M6_CUDA(c, cudaGetLastError());
  return St.finish();
}
