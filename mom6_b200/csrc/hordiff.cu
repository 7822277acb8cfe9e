// tracer_hordiff, the along-surface path (src/tracer/MOM_tracer_hor_diff.F90:119-640; MOM.F90:1526, the lateral tracer diffusion that
// follows advect_tracer in step_MOM_tracer_dyn) on the device, so the tracers stay resident through the tracer step.
//  * hd_khdt_kernel<DIR>: khdt_x / khdt_y of every face (:204-327), one thread per face.
//  * hd_cfl_kernel: the diffusive CFL number of every cell and its maximum (:354-362) -- a maximum is exact in any order: warp
//    shuffles, then one 64-bit integer atomicMax per warp on the bit pattern (non-negative doubles order like integers).
//  * hd_diffuse_kernel: one Jacobi sweep of one tracer (:553-597), one thread per cell, blockIdx.z = layer; reads the tracer and
//    writes the updated copy (the reference's dTr scratch), with the conc_underflow flush folded in (:599-604); hd_copy_box_kernel
//    puts the computational domain back.
//  * hd_dflux_kernel<DIR>: the optional df_x / df_y diagnostics (:576-583).
// num_itts sweeps, each preceded by the halo update of the tracers (do_group_pass :541); max_across_PEs is a one-word NCCL all-reduce.
#include "ctx.h"
#include "hordiff_cell.cuh"
#include <cfloat>
#include <cmath>
#include <vector>

using m6::Geom;

namespace {

using HdP = m6hd::Par;
struct HdBox { int is, ie, js, je; };

template <int DIR>
__global__ void __launch_bounds__(128) hd_khdt_kernel(const Geom G, const HdP P, const HdBox B, const double* __restrict__ lenC,
                                                      const double* __restrict__ IdC, const double* __restrict__ areaT,
                                                      const double* __restrict__ Res_fn_h, const double* __restrict__ Rd_dx_h,
                                                      const double* __restrict__ L2, const double* __restrict__ SN,
                                                      const double* __restrict__ MEKE_Kh, double* __restrict__ khdt) {
  const int i = (DIR == 0 ? B.is - 1 : B.is) + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (DIR == 0 ? B.js : B.js - 1) + blockIdx.y;
  if (i > B.ie || j > B.je) return;
  const long long g = G.idx(i, j);
  khdt[g] = m6hd::khdt_face(P, g, (DIR == 0) ? 1 : G.pitch, lenC, IdC, areaT, Res_fn_h, Rd_dx_h, L2, SN, MEKE_Kh);
}

__global__ void __launch_bounds__(128) hd_cfl_kernel(const Geom G, const HdBox B, const double* __restrict__ khdt_x,
                                                     const double* __restrict__ khdt_y, const double* __restrict__ IareaT,
                                                     unsigned long long* __restrict__ max_bits) {
  const int i = B.is + blockIdx.x * blockDim.x + threadIdx.x, j = B.js + blockIdx.y;
  double c = 0.0;
  if (i <= B.ie && j <= B.je) {
    c = m6hd::cfl_cell(G.idx(i, j), G.pitch, khdt_x, khdt_y, IareaT);
    if (!(c > 0.0)) c = 0.0;  // `if (max_CFL < CFL) max_CFL = CFL` starting from 0: negatives and NaNs never win
  }
  unsigned long long b = (unsigned long long)__double_as_longlong(c);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, b, o); b = (x > b) ? x : b; }
  if ((threadIdx.x & 31) == 0 && b != 0ULL) atomicMax(max_bits, b);
}

__global__ void __launch_bounds__(128) hd_diffuse_kernel(const Geom G, const HdP P, const HdBox B, const double scale, const double underflow,
                                                         const double* __restrict__ h, const double* __restrict__ T,
                                                         const double* __restrict__ khdt_x, const double* __restrict__ khdt_y,
                                                         const double* __restrict__ IareaT, double* __restrict__ Tnew) {
  const int i = B.is + blockIdx.x * blockDim.x + threadIdx.x, j = B.js + blockIdx.y;
  if (i > B.ie || j > B.je) return;
  const long long g = G.idx(i, j), gk = g + (long long)blockIdx.z * G.plane;
  double t = m6hd::diffuse_cell(P, scale, g, gk, G.pitch, h, T, khdt_x, khdt_y, IareaT);
  if (underflow > 0.0 && fabs(t) < underflow) t = 0.0;
  Tnew[gk] = t;
}

template <int DIR>
__global__ void __launch_bounds__(128) hd_dflux_kernel(const Geom G, const HdP P, const HdBox B, const double scale, const double* __restrict__ h,
                                                       const double* __restrict__ T, const double* __restrict__ khdt, double* __restrict__ df) {
  const int i = (DIR == 0 ? B.is - 1 : B.is) + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (DIR == 0 ? B.js : B.js - 1) + blockIdx.y;
  if (i > B.ie || j > B.je) return;
  const long long g = G.idx(i, j), gk = g + (long long)blockIdx.z * G.plane;
  df[gk] = df[gk] + m6hd::dflux_face(P, scale, g, gk, (DIR == 0) ? 1 : G.pitch, h, T, khdt);
}

__global__ void hd_copy_box_kernel(const Geom G, const HdBox B, const double* __restrict__ src, double* __restrict__ dst) {
  const int i = B.is + blockIdx.x * blockDim.x + threadIdx.x, j = B.js + blockIdx.y;
  if (i > B.ie || j > B.je) return;
  const long long gk = G.idx(i, j) + (long long)blockIdx.z * G.plane;
  dst[gk] = src[gk];
}

__global__ void hd_zero_faces_kernel(const Geom G, const HdBox B, const int dir, double* __restrict__ df) {
  const int i = (dir == 0 ? B.is - 1 : B.is) + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (dir == 0 ? B.js : B.js - 1) + blockIdx.y;
  if (i > B.ie || j > B.je) return;
  df[G.idx(i, j) + (long long)blockIdx.z * G.plane] = 0.0;
}

}  // namespace

extern "C" int mom6cu_tracer_hordiff(mom6cu_ctx* c, const mom6cu_tracer_hor_diff_cs* CS, const mom6cu_tracer_hordiff_args* a) {
  if (!c || !CS || !a) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "tracer_hordiff: mom6cu_set_grid / mom6cu_set_vgrid have not been called");
  if (CS->use_neutral_diffusion || CS->use_hor_bnd_diffusion || CS->Diffuse_ML_interior)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "tracer_hordiff: neutral diffusion, horizontal boundary diffusion and DIFFUSE_ML_TO_INTERIOR are not implemented");
  const int ntr = a->ntr;
  c->last_iterations = 0;
  if (ntr < 0 || (ntr > 0 && (!a->tr || !a->h))) return c->fail(MOM6CU_ERR_BAD_ARG, "tracer_hordiff: null required argument");
  if (ntr == 0 || (CS->KhTr <= 0.0 && !CS->use_variable_mixing)) return 0;  // :153
  const bool use_VarMix = CS->use_variable_mixing != 0, Resoln_scaled = use_VarMix && CS->Resoln_scaled_KhTr;
  if (Resoln_scaled && !a->Res_fn_h) return c->fail(MOM6CU_ERR_BAD_ARG, "tracer_hordiff: VarMix%%Res_fn_h is not allocated");
  if (use_VarMix && CS->KhTr_passivity_coeff > 0. && !a->Rd_dx_h) return c->fail(MOM6CU_ERR_BAD_ARG, "tracer_hordiff: VarMix%%Rd_dx_h is not allocated");
  const bool use_Eady = use_VarMix && CS->KhTr_Slope_Cff > 0., use_MEKE = use_VarMix && CS->use_MEKE_Kh;
  if ((use_Eady && (!a->L2u || !a->SN_u || !a->L2v || !a->SN_v)) || (use_MEKE && !a->MEKE_Kh))
    return c->fail(MOM6CU_ERR_BAD_ARG, "tracer_hordiff: VarMix%%L2u / SN_u / L2v / SN_v or MEKE%%Kh is not allocated");
  if (!(a->dt > 0.0)) return c->fail(MOM6CU_ERR_BAD_ARG, "tracer_hordiff: dt must be positive");
  const mom6cu_domain& d = c->dom;
  const Geom& G = c->g;
  const GridDev& Gd = c->grid;
  const int nz = G.nk;
  if (std::min(std::min(d.isc - d.isd, d.ied - d.iec), std::min(d.jsc - d.jsd, d.jed - d.jec)) < 1)
    return c->fail(MOM6CU_ERR_BAD_ARG, "tracer_hordiff: needs a halo of at least one point");
  Stager S(c, "hd.");
  int rc;
  const double *d_h, *d_res = nullptr, *d_rd = nullptr;
  if ((rc = S.in3(a->h, ST_H, "h", &d_h)) || (rc = S.in2(a->Res_fn_h, ST_H, "Res_fn_h", &d_res)) || (rc = S.in2(a->Rd_dx_h, ST_H, "Rd_dx_h", &d_rd))) return rc;
  const double *d_l2u = nullptr, *d_snu = nullptr, *d_l2v = nullptr, *d_snv = nullptr, *d_kh = nullptr;
  if (use_Eady && ((rc = S.in2(a->L2u, ST_U, "L2u", &d_l2u)) || (rc = S.in2(a->SN_u, ST_U, "SN_u", &d_snu)) || (rc = S.in2(a->L2v, ST_V, "L2v", &d_l2v)) ||
                   (rc = S.in2(a->SN_v, ST_V, "SN_v", &d_snv)))) return rc;
  if (use_MEKE && (rc = S.in2(a->MEKE_Kh, ST_H, "MEKE_Kh", &d_kh))) return rc;
  std::vector<double*> TA(ntr), TB(ntr), DX(ntr, nullptr), DY(ntr, nullptr);
  for (int m = 0; m < ntr; ++m) {
    char nm[32];
    if (!a->tr[m]) return c->fail(MOM6CU_ERR_BAD_ARG, "tracer_hordiff: tracer %d is null", m);
    snprintf(nm, sizeof nm, "tr%d", m);
    if ((rc = S.io3(a->tr[m], ST_H, nm, &TA[m]))) return rc;
    snprintf(nm, sizeof nm, "hd.trB%d", m);
    if (!(TB[m] = c->plane3(nm))) return MOM6CU_ERR_CUDA;
    if (a->df_x && a->df_x[m]) { snprintf(nm, sizeof nm, "df_x%d", m); if ((rc = S.io3(a->df_x[m], ST_U, nm, &DX[m]))) return rc; }
    if (a->df_y && a->df_y[m]) { snprintf(nm, sizeof nm, "df_y%d", m); if ((rc = S.io3(a->df_y[m], ST_V, nm, &DY[m]))) return rc; }
  }
  double *khx = c->plane2("hd.khdt_x"), *khy = c->plane2("hd.khdt_y");
  unsigned long long* d_max = (unsigned long long*)c->buf("hd.maxCFL", 2);
  double* h_max = c->host_scratch("hd.maxCFL", 2);
  if (!khx || !khy || !d_max || !h_max) return MOM6CU_ERR_CUDA;
  HdP P = {};
  P.dt = a->dt; P.Idt = 1.0 / a->dt; P.h_neglect = c->vgrid.H_subroundoff; P.KhTr = CS->KhTr; P.KhTr_min = CS->KhTr_min; P.KhTr_max = CS->KhTr_max;
  P.pass_coeff = CS->KhTr_passivity_coeff; P.pass_min = CS->KhTr_passivity_min; P.max_diff_CFL = CS->max_diff_CFL;
  P.use_VarMix = use_VarMix ? 1 : 0; P.Resoln_scaled = Resoln_scaled ? 1 : 0;
  P.use_Eady = use_Eady ? 1 : 0; P.use_MEKE = use_MEKE ? 1 : 0; P.Slope_Cff = CS->KhTr_Slope_Cff; P.KhTr_fac = CS->MEKE_KhTr_fac;
  const HdBox B = {d.isc, d.iec, d.jsc, d.jec};
  if ((rc = S.begin())) return rc;
  const int ni = d.iec - d.isc + 1, nj = d.jec - d.jsc + 1;
  const dim3 gu((ni + 1 + 127) / 128, nj), gv((ni + 127) / 128, nj + 1), gh((ni + 127) / 128, nj);
  M6_LAUNCH(c, hd_khdt_kernel<0>, gu, 128, 0, G, P, B, Gd.dy_Cu, Gd.IdxCu, Gd.areaT, d_res, d_rd, d_l2u, d_snu, d_kh, khx);
  M6_LAUNCH(c, hd_khdt_kernel<1>, gv, 128, 0, G, P, B, Gd.dx_Cv, Gd.IdyCv, Gd.areaT, d_res, d_rd, d_l2v, d_snv, d_kh, khy);
  int num_itts = 1;
  double I_numitts = 1.0;
  if (CS->check_diffusive_CFL) {  // :354-366
    M6_CUDA(c, cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), c->stream));
    M6_LAUNCH(c, hd_cfl_kernel, gh, 128, 0, G, B, khx, khy, Gd.IareaT, d_max);
    M6_CUDA(c, cudaMemcpyAsync(h_max, d_max, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    M6_CUDA(c, cudaStreamSynchronize(c->stream));
    double max_CFL = h_max[0];
    if (c->nranks > 1 && (rc = m6_allreduce_max_doubles(c, &max_CFL, 1))) return rc;  // max_across_PEs :361
    num_itts = std::max(1, (int)std::ceil(max_CFL - 4.0 * DBL_EPSILON));
    I_numitts = 1.0 / ((double)num_itts);
  } else if (CS->max_diff_CFL > 0.0) {
    num_itts = std::max(1, (int)std::ceil(CS->max_diff_CFL - 4.0 * DBL_EPSILON));
    I_numitts = 1.0 / ((double)num_itts);
  }
  const dim3 gu3(gu.x, gu.y, nz), gv3(gv.x, gv.y, nz), gh3(gh.x, gh.y, nz);
  for (int m = 0; m < ntr; ++m) {  // :374-390
    if (DX[m]) M6_LAUNCH(c, hd_zero_faces_kernel, gu3, 128, 0, G, B, 0, DX[m]);
    if (DY[m]) M6_LAUNCH(c, hd_zero_faces_kernel, gv3, 128, 0, G, B, 1, DY[m]);
  }
  const double scale = I_numitts;
  for (int itt = 1; itt <= num_itts; ++itt) {
    {  // do_group_pass(CS%pass_t) :541
      std::vector<int> st(ntr, ST_H);
      for (int f0 = 0; f0 < ntr; f0 += 8)
        if ((rc = m6_halo_update(c, TA.data() + f0, st.data() + f0, std::min(8, ntr - f0), 0, nz))) return rc;
    }
    for (int m = 0; m < ntr; ++m) {
      if (DX[m]) M6_LAUNCH(c, hd_dflux_kernel<0>, gu3, 128, 0, G, P, B, scale, d_h, TA[m], khx, DX[m]);
      if (DY[m]) M6_LAUNCH(c, hd_dflux_kernel<1>, gv3, 128, 0, G, P, B, scale, d_h, TA[m], khy, DY[m]);
      const double uf = a->conc_underflow ? a->conc_underflow[m] : 0.0;
      M6_LAUNCH(c, hd_diffuse_kernel, gh3, 128, 0, G, P, B, scale, uf, d_h, TA[m], khx, khy, Gd.IareaT, TB[m]);
      // the computational domain goes back into the caller's field (the next sweep, the halo update and the staging copy work on it)
      M6_LAUNCH(c, hd_copy_box_kernel, gh3, 128, 0, G, B, TB[m], TA[m]);
    }
  }
  M6_CUDA(c, cudaGetLastError());
  c->last_iterations = num_itts;
  return S.finish();
}
