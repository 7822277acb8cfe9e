// mixedlayer_restrat -> mixedlayer_restrat_OM4 (src/parameterizations/lateral/MOM_mixed_layer_restrat.F90:149-714): the
// Fox-Kemper et al. (2008) mixed-layer-eddy restratification in general coordinates, which adjusts h, uhtr and vhtr between
// the dynamics step and the tracer advection (MOM.F90:1422) -- on the device so that the state stays resident across it.
//
//  * mle_column_kernel: one thread per column of (isc-1:iec+1, jsc-1:jec+1).  MLD time filters (:316-346), the volume
//    available to each face h_avail (:382), and a single top-down sweep that accumulates the mixed-layer thickness and the
//    thickness-weighted surface-referenced density for the fast and the slow filtered depths (:384-410).  The reference's
//    row-level "keep_going" exit only skips work no column of the row needs; here each thread stops on its own.
//  * mle_face_kernel<DIR>: one thread per velocity face.  The overturning amplitudes uDml / uDml_slow (:464-495), then three
//    top-down passes over the column -- limit uDml by the available volumes, limit uDml_slow, form the transports (:500-535) --
//    recomputing the profile a(k), b(k) = mu(z_top) - mu(z_bottom) in each pass instead of storing it (the sequence of
//    operations, hence every bit, is the same).
//  * mle_update_kernel: h -= dt * IareaT * div(uhml, vhml), floored at Angstrom_H / 2 (:623-627).
#include "ctx.h"
#include "mle_mu.cuh"

using m6::Geom;

namespace {

struct MleP {
  int nk, is, ie, js, je;
  double dt, Z_to_H, Angstrom_H, h_neglect, g_Rho0, I4dt, h_min, vonKar_x_pi2, ustar_min, coef, coef2, front_length, stretch, tail_dh;
  double aFac1, bFac1, aFac2, bFac2;
  int filt1, filt2, res_upscale;
  m6mle::Eos eos;
};

__global__ void __launch_bounds__(128) mle_column_kernel(const Geom G, const MleP P, const double* __restrict__ h, const double* __restrict__ T,
                                                         const double* __restrict__ S, const double* __restrict__ h_MLD,
                                                         const double* __restrict__ areaT, double* __restrict__ MLD_filtered,
                                                         double* __restrict__ MLD_filtered_slow, double* __restrict__ h_avail,
                                                         double* __restrict__ htot_fast, double* __restrict__ htot_slow,
                                                         double* __restrict__ Rml_av_fast, double* __restrict__ Rml_av_slow) {
  const int i = (P.is - 1) + blockIdx.x * blockDim.x + threadIdx.x, j = (P.js - 1) + blockIdx.y;
  if (i > P.ie + 1 || j > P.je + 1) return;
  const long long g = G.idx(i, j);
  double MLD_fast = P.stretch * h_MLD[g];
  if (P.filt1) {
    const double f = m6mle::fmx(MLD_fast, P.bFac1 * MLD_fast + P.aFac1 * MLD_filtered[g]);
    MLD_filtered[g] = f;
    MLD_fast = f;
  }
  double MLD_slow = MLD_fast;
  if (P.filt2) {
    const double f = m6mle::fmx(MLD_fast, P.bFac2 * MLD_fast + P.aFac2 * MLD_filtered_slow[g]);
    MLD_filtered_slow[g] = f;
    MLD_slow = f;
  }
  const double aT = areaT[g];
  double hf = 0.0, hs = 0.0, rf = 0.0, rs = 0.0;
  for (int k = 0; k < P.nk; ++k) {
    const long long gk = g + (long long)k * G.plane;
    const double hk = h[gk];
    h_avail[gk] = m6mle::fmx(P.I4dt * aT * (hk - P.Angstrom_H), 0.0);
    const bool nf = hf < MLD_fast, ns = hs < MLD_slow;
    if (nf || ns) {
      const double rho = m6mle::density(P.eos, T[gk], S[gk], 0.0);
      if (nf) { const double dh = m6mle::fmn(hk, MLD_fast - hf); rf = rf + dh * rho; hf = hf + dh; }
      if (ns) { const double dh = m6mle::fmn(hk, MLD_slow - hs); rs = rs + dh * rho; hs = hs + dh; }
    }
  }
  htot_fast[g] = hf; htot_slow[g] = hs;
  Rml_av_fast[g] = -(P.g_Rho0 * rf) / (hf + P.h_neglect);
  Rml_av_slow[g] = -(P.g_Rho0 * rs) / (hs + P.h_neglect);
}

// DIR 0: u faces (I = isc-1..iec, j = jsc..jec), neighbour cell i+1;  DIR 1: v faces (i = isc..iec, J = jsc-1..jec), neighbour j+1
template <int DIR>
__global__ void __launch_bounds__(128) mle_face_kernel(const Geom G, const MleP P, const double* __restrict__ h, const double* __restrict__ h_avail,
                                                       const double* __restrict__ ustar, const double* __restrict__ Rd_dx_h,
                                                       const double* __restrict__ htot_fast, const double* __restrict__ htot_slow,
                                                       const double* __restrict__ Rml_av_fast, const double* __restrict__ Rml_av_slow,
                                                       const double* __restrict__ CoriolisBu, const double* __restrict__ maskC,
                                                       const double* __restrict__ dxC, const double* __restrict__ dyC,
                                                       const double* __restrict__ Idn /* IdxCu | IdyCv */, double* __restrict__ hml,
                                                       double* __restrict__ htr) {
  const int i = (DIR == 0 ? P.is - 1 : P.is) + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (DIR == 0 ? P.js : P.js - 1) + blockIdx.y;
  if (i > P.ie || j > P.je) return;
  const long long g = G.idx(i, j), sd = (DIR == 0) ? 1 : G.pitch, pl = G.plane;
  const double u_star = m6mle::fmx(P.ustar_min, 0.5 * (P.Z_to_H * ustar[g] + P.Z_to_H * ustar[g + sd]));
  // |f| at the face: q points (I,J-1),(I,J) for u; (I-1,J),(I,J) for v
  const double absf = 0.5 * (fabs(CoriolisBu[g - ((DIR == 0) ? G.pitch : 1)]) + fabs(CoriolisBu[g]));
  double res_scaling_fac = 0.0;
  if (P.res_upscale) {
    const double lfront = 0.5 * (P.front_length + P.front_length);
    double I_LFront = 0.0; if (lfront != 0.0) I_LFront = 1.0 / lfront;
    const double dx = dxC[g], dy = dyC[g];
    res_scaling_fac = (sqrt(0.5 * ((dx * dx) + (dy * dy))) * I_LFront) * m6mle::fmn(1., 0.5 * (Rd_dx_h[g] + Rd_dx_h[g + sd]));
  }
  const double len = (DIR == 0) ? dyC[g] : dxC[g];  // G%dyCu | G%dxCv
  const double hfs = htot_fast[g] + htot_fast[g + sd], hss = htot_slow[g] + htot_slow[g + sd];
  double h_vel = 0.5 * (hfs + P.h_neglect);
  double mom_mixrate = P.vonKar_x_pi2 * (u_star * u_star) / (absf * (h_vel * h_vel) + 4.0 * (h_vel + P.h_neglect) * u_star);
  double timescale = 0.0625 * (absf + 2.0 * mom_mixrate) / ((absf * absf) + (mom_mixrate * mom_mixrate));
  timescale = timescale * P.coef;
  if (P.res_upscale) timescale = timescale * res_scaling_fac;
  double Dml = timescale * maskC[g] * len * Idn[g] * (Rml_av_fast[g + sd] - Rml_av_fast[g]) * (h_vel * h_vel);
  h_vel = 0.5 * (hss + P.h_neglect);
  mom_mixrate = P.vonKar_x_pi2 * (u_star * u_star) / (absf * (h_vel * h_vel) + 4.0 * (h_vel + P.h_neglect) * u_star);
  timescale = 0.0625 * (absf + 2.0 * mom_mixrate) / ((absf * absf) + (mom_mixrate * mom_mixrate));
  timescale = timescale * P.coef2;
  if (P.res_upscale) timescale = timescale * res_scaling_fac;
  double Dml_slow = timescale * maskC[g] * len * Idn[g] * (Rml_av_slow[g + sd] - Rml_av_slow[g]) * (h_vel * h_vel);
  if (Dml + Dml_slow == 0.) {
    for (int k = 0; k < P.nk; ++k) hml[g + (long long)k * pl] = 0.0;
    return;
  }
  const double IhTot = 2.0 / (hfs + P.h_neglect), IhTot_slow = 2.0 / (hss + P.h_neglect);
  const double tail = P.tail_dh;
  // pass 1: limit Dml by the volume available on the upwind side of each layer (:504-513)
  double zpa = 0.0;
  for (int k = 0; k < P.nk; ++k) {
    const long long gk = g + (long long)k * pl;
    const double hAtVel = 0.5 * (h[gk] + h[gk + sd]);
    double a = m6mle::mu(zpa, tail);
    zpa = zpa - (hAtVel * IhTot);
    a = a - m6mle::mu(zpa, tail);
    if (a * Dml > 0.0) { const double av = h_avail[gk]; if (a * Dml > av) Dml = av / a; }
    else if (a * Dml < 0.0) { const double av = h_avail[gk + sd]; if (-a * Dml > av) Dml = -av / a; }
  }
  // pass 2: limit Dml_slow by what Dml leaves (:514-526)
  zpa = 0.0;
  double zpb = 0.0;
  for (int k = 0; k < P.nk; ++k) {
    const long long gk = g + (long long)k * pl;
    const double hAtVel = 0.5 * (h[gk] + h[gk + sd]);
    double a = m6mle::mu(zpa, tail);
    zpa = zpa - (hAtVel * IhTot);
    a = a - m6mle::mu(zpa, tail);
    double b = m6mle::mu(zpb, tail);
    zpb = zpb - (hAtVel * IhTot_slow);
    b = b - m6mle::mu(zpb, tail);
    if (b * Dml_slow > 0.0) {
      const double room = h_avail[gk] - a * Dml;
      if (b * Dml_slow > room) Dml_slow = m6mle::fmx(0., room) / b;
    } else if (b * Dml_slow < 0.0) {
      const double room = h_avail[gk + sd] + a * Dml;
      if (-b * Dml_slow > room) Dml_slow = -m6mle::fmx(0., room) / b;
    }
  }
  // pass 3: the transports (:527-530)
  zpa = 0.0; zpb = 0.0;
  for (int k = 0; k < P.nk; ++k) {
    const long long gk = g + (long long)k * pl;
    const double hAtVel = 0.5 * (h[gk] + h[gk + sd]);
    double a = m6mle::mu(zpa, tail);
    zpa = zpa - (hAtVel * IhTot);
    a = a - m6mle::mu(zpa, tail);
    double b = m6mle::mu(zpb, tail);
    zpb = zpb - (hAtVel * IhTot_slow);
    b = b - m6mle::mu(zpb, tail);
    const double t = a * Dml + b * Dml_slow;
    hml[gk] = t;
    htr[gk] = htr[gk] + t * P.dt;
  }
}

__global__ void mle_update_kernel(const Geom G, const MleP P, const double* __restrict__ uhml, const double* __restrict__ vhml,
                                  const double* __restrict__ IareaT, double* __restrict__ h) {
  const int i = P.is + blockIdx.x * blockDim.x + threadIdx.x, j = P.js + blockIdx.y;
  if (i > P.ie || j > P.je) return;
  const long long g = G.idx(i, j), gk = g + (long long)blockIdx.z * G.plane;
  double hn = h[gk] - P.dt * IareaT[g] * ((uhml[gk] - uhml[gk - 1]) + (vhml[gk] - vhml[gk - G.pitch]));
  if (hn < P.h_min) hn = P.h_min;
  h[gk] = hn;
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
  if (CS->MLE_density_diff > 0.) return c->fail(MOM6CU_ERR_UNSUPPORTED, "mixedlayer_restrat_OM4: detect_mld (MLE_DENSITY_DIFF > 0) is not implemented");
  if (!CS->MLE_use_PBL_MLD || !h_MLD) return c->fail(MOM6CU_ERR_BAD_ARG, "mixedlayer_restrat_OM4: No MLD to use for MLE parameterization.");
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
      (rc = St.in2(h_MLD, ST_H, "h_MLD", &d_ml)) || (rc = St.in2(Rd_dx_h, ST_H, "Rd_dx_h", &d_rd))) return rc;
  if (CS->MLD_filtered && (rc = St.io2(CS->MLD_filtered, ST_H, "MLD_filtered", &d_f1))) return rc;
  if (CS->MLD_filtered_slow && (rc = St.io2(CS->MLD_filtered_slow, ST_H, "MLD_filtered_slow", &d_f2))) return rc;
  double *d_av = c->plane3("mle.h_avail"), *d_uhml = c->plane3("mle.uhml"), *d_vhml = c->plane3("mle.vhml");
  double *d_hf = c->plane2("mle.htot_fast"), *d_hs = c->plane2("mle.htot_slow"), *d_rf = c->plane2("mle.Rml_av_fast"), *d_rs = c->plane2("mle.Rml_av_slow");
  if (!d_av || !d_uhml || !d_vhml || !d_hf || !d_hs || !d_rf || !d_rs) return MOM6CU_ERR_CUDA;
  MleP P = {};
  P.nk = nz; P.is = d.isc; P.ie = d.iec; P.js = d.jsc; P.je = d.jec;
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
  if ((rc = St.begin())) return rc;
  const GridDev& Gd = c->grid;
  {
    const dim3 grid((d.iec - d.isc + 3 + 127) / 128, d.jec - d.jsc + 3);
    M6_LAUNCH(c, mle_column_kernel, grid, 128, 0, G, P, d_h, d_T, d_S, d_ml, Gd.areaT, d_f1, d_f2, d_av, d_hf, d_hs, d_rf, d_rs);
  }
  {
    const dim3 gu((d.iec - d.isc + 2 + 127) / 128, d.jec - d.jsc + 1), gv((d.iec - d.isc + 1 + 127) / 128, d.jec - d.jsc + 2);
    M6_LAUNCH(c, mle_face_kernel<0>, gu, 128, 0, G, P, d_h, d_av, d_us, d_rd, d_hf, d_hs, d_rf, d_rs, Gd.CoriolisBu, Gd.mask2dCu, Gd.dxCu, Gd.dyCu,
              Gd.IdxCu, d_uhml, d_uhtr);
    M6_LAUNCH(c, mle_face_kernel<1>, gv, 128, 0, G, P, d_h, d_av, d_us, d_rd, d_hf, d_hs, d_rf, d_rs, Gd.CoriolisBu, Gd.mask2dCv, Gd.dxCv, Gd.dyCv,
              Gd.IdyCv, d_vhml, d_vhtr);
  }
  {
    const dim3 grid((d.iec - d.isc + 1 + 127) / 128, d.jec - d.jsc + 1, nz);
    M6_LAUNCH(c, mle_update_kernel, grid, 128, 0, G, P, d_uhml, d_vhml, Gd.IareaT, d_h);
  }
  M6_CUDA(c, cudaGetLastError());
  return St.finish();
}
