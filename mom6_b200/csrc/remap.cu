// ALE vertical remapping on the device: ALE_remap_tracers (/root/reference/src/ALE/MOM_ALE.F90:760-879),
// ALE_remap_set_h_vel (:882-925), ALE_remap_velocities (:1089-1300) and batched remapping_core_h
// (src/ALE/MOM_remapping.F90:234-335).  One thread per column (remap_column.cuh); the sub-cell grid of a column is
// built once and shared by every field remapped between the same pair of grids.
#include "ctx.h"
#include "common.cuh"
#include "remap_column.cuh"
#include "remap_stream.cuh"
#include <cstdlib>

using m6::Geom;
using namespace m6remap;

namespace {

constexpr int MAXF = 16;  // fields per launch
constexpr int REMAP_MINB_DEFAULT = 4;  // measured: 10.2 ms (4 CTAs/SM) vs 12.4 ms (1) per field at 1440x1080x75
struct Fields { double* p[MAXF]; double underflow[MAXF]; int n; };

template <int KCAP>
__device__ void remap_one_column(const Params& P, int n0, int n1, Col h0, Col h1, const Fields& F, long off_sc, long sk0, long sk1) {
  SubGrid<KCAP> S;
  Recon<KCAP> R;
  SubVals<KCAP> V;
  S.n0 = n0; S.n1 = n1;
  for (int k = 1; k <= n0; ++k) S.h0[k] = h0(k);
  for (int k = 1; k <= n1; ++k) S.h1[k] = h1(k);
  intersect<KCAP>(S);
  for (int f = 0; f < F.n; ++f) {
    double* col = F.p[f] + off_sc;
    for (int k = 1; k <= n0; ++k) R.u[k] = col[(long)(k - 1) * sk0];
    const int method = build_reconstructions<KCAP>(P, n0, S.h0, R);
    remap_via_sub_cells<KCAP>(P, S, R, method, V, ColOut{col, sk1}, F.underflow[f]);
  }
}

// columns of 3-D fields on the planes of G: (i, j) in [ilo, ihi] x [jlo, jhi] with mask > 0
template <int KCAP>
__global__ void __launch_bounds__(128) remap_planes_kernel(Geom G, Params P, int nk, int ilo, int ihi, int jlo, int jhi,
                                                          const double* __restrict__ mask, const double* __restrict__ h_old,
                                                          const double* __restrict__ h_new, Fields F) {
  const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x, j = jlo + blockIdx.y;
  if (i > ihi || j > jhi) return;
  const long g = G.idx(i, j);
  if (!(mask[g] > 0.)) return;
  const long pl = G.plane;
  remap_one_column<KCAP>(P, nk, nk, Col{h_old + g, pl}, Col{h_new + g, pl}, F, g, pl, pl);
}

// streaming form (remap_stream.cuh) for PCM / PLM / PPM_H4: one thread per (column, field); the only thread-local array is the
// output column, because a target cell may be completed before the source cells below it have been read
template <int KCAP, int MINB>
__global__ void __launch_bounds__(128, MINB) remap_planes_stream_kernel(Geom G, Params P, int nk, int ilo, int ihi, int jlo, int jhi,
                                                                 const double* __restrict__ mask, const double* __restrict__ h_old,
                                                                 const double* __restrict__ h_new, Fields F) {
  const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x, j = jlo + blockIdx.y, f = blockIdx.z;
  if (i > ihi || j > jhi) return;
  const long g = G.idx(i, j);
  if (!(mask[g] > 0.)) return;
  const long pl = G.plane;
  const double* h0 = h_old + g - pl;  // 1-based level index
  const double* h1 = h_new + g - pl;
  double* col = F.p[f] + g - pl;
  double u1[KCAP + 1];
  remap_stream(P, nk, nk, [=](int k) { return __ldg(h0 + (long)k * pl); }, [=](int k) { return col[(long)k * pl]; },
               [=](int k) { return __ldg(h1 + (long)k * pl); }, ColOut{u1 + 1, 1}, F.underflow[f]);
  for (int k = 1; k <= nk; ++k) col[(long)k * pl] = u1[k];
}

template <int KCAP>
__global__ void __launch_bounds__(128) remap_batch_stream_kernel(Params P, int ncol, int n0, int n1, const double* __restrict__ h0,
                                                                const double* __restrict__ u0, const double* __restrict__ h1,
                                                                double* __restrict__ u1) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  const double *H0 = h0 + (long)c * n0 - 1, *U0 = u0 + (long)c * n0 - 1, *H1 = h1 + (long)c * n1 - 1;
  remap_stream(P, n0, n1, [=](int k) { return H0[k]; }, [=](int k) { return U0[k]; }, [=](int k) { return H1[k]; },
               ColOut{u1 + (long)c * n1, 1}, 0.0);
}

bool use_stream(const Params& P) {
  static int force_array = -1;
  if (force_array < 0) { const char* e = getenv("MOM6CU_REMAP_ARRAY"); force_array = (e && atoi(e)) ? 1 : 0; }
  return !force_array && P.scheme != SCHEME_PPM_IH4;
}

// a batch of independent columns stored row-major (ncol, n): stride n between columns, 1 between levels
template <int KCAP>
__global__ void __launch_bounds__(128) remap_batch_kernel(Params P, int ncol, int n0, int n1, const double* __restrict__ h0,
                                                         const double* __restrict__ u0, const double* __restrict__ h1,
                                                         double* __restrict__ u1) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  SubGrid<KCAP> S;
  Recon<KCAP> R;
  SubVals<KCAP> V;
  S.n0 = n0; S.n1 = n1;
  for (int k = 1; k <= n0; ++k) { S.h0[k] = h0[(long)c * n0 + k - 1]; R.u[k] = u0[(long)c * n0 + k - 1]; }
  for (int k = 1; k <= n1; ++k) S.h1[k] = h1[(long)c * n1 + k - 1];
  intersect<KCAP>(S);
  const int method = build_reconstructions<KCAP>(P, n0, S.h0, R);
  remap_via_sub_cells<KCAP>(P, S, R, method, V, ColOut{u1 + (long)c * n1, 1}, 0.0);
}

// ALE_remap_set_h_vel, MOM_ALE.F90:882-925 (no OBCs, no partial-cell h_vel_mask)
__global__ void set_h_vel_kernel(Geom G, int is, int ie, int js, int je, const double* __restrict__ mCu, const double* __restrict__ mCv,
                                 const double* __restrict__ hn, double* __restrict__ h_u, double* __restrict__ h_v) {
  const int i = is - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = js - 1 + blockIdx.y, k = blockIdx.z;
  if (i > ie || j > je) return;
  const long g = G.idx(i, j), o = (long)k * G.plane + g;
  if (j >= js && mCu[g] > 0.) h_u[o] = 0.5 * (hn[o] + hn[o + 1]);
  if (i >= is && mCv[g] > 0.) h_v[o] = 0.5 * (hn[o] + hn[o + G.pitch]);
}

int check_cs(mom6cu_ctx* c, const mom6cu_remapping_cs* CS, int nmax, Params* P) {
  if (!CS) return c->fail(MOM6CU_ERR_BAD_ARG, "remapping: null CS");
  const int s = CS->remapping_scheme;
  if (s != MOM6CU_REMAPPING_PCM && s != MOM6CU_REMAPPING_PLM && s != MOM6CU_REMAPPING_PPM_H4 && s != MOM6CU_REMAPPING_PPM_IH4)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "remapping: scheme %d is not implemented (PCM, PLM, PPM_H4, PPM_IH4 are)", s);
  if (CS->answer_date < 20190101)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "remapping: answer_date %d < 20190101 is not implemented", CS->answer_date);
  if (nmax > 128) return c->fail(MOM6CU_ERR_UNSUPPORTED, "remapping: %d levels exceed the 128-level column capacity", nmax);
  if (nmax < 1) return c->fail(MOM6CU_ERR_BAD_ARG, "remapping: empty columns");
  *P = {s, CS->boundary_extrapolation, CS->force_bounds_in_subcell, CS->force_bounds_in_target, CS->om4_remap_via_sub_cells,
        CS->h_neglect, CS->h_neglect_edge};
  return 0;
}

int launch_planes(mom6cu_ctx* c, const Params& P, int ilo, int ihi, int jlo, int jhi, const double* mask, const double* h_old,
                  const double* h_new, const Fields& F) {
  const Geom& G = c->g;
  const dim3 grid((ihi - ilo + 1 + 127) / 128, jhi - jlo + 1), block(128);
  if (use_stream(P)) {
    const dim3 gs(grid.x, grid.y, F.n);
    static int minb = -1;  // CTAs/SM the kernel is compiled for: 3 (156 registers) or 4 (128, a few spills); MOM6CU_REMAP_MINB overrides
    if (minb < 0) { const char* e = getenv("MOM6CU_REMAP_MINB"); minb = e ? atoi(e) : REMAP_MINB_DEFAULT; }
#define M6_RS(K, B) M6_LAUNCH(c, (remap_planes_stream_kernel<K, B>), gs, block, 0, G, P, G.nk, ilo, ihi, jlo, jhi, mask, h_old, h_new, F)
    if (minb >= 4) { if (G.nk <= 40) M6_RS(40, 4); else if (G.nk <= 80) M6_RS(80, 4); else M6_RS(128, 4); }
    else { if (G.nk <= 40) M6_RS(40, 1); else if (G.nk <= 80) M6_RS(80, 1); else M6_RS(128, 1); }
#undef M6_RS
    M6_CUDA(c, cudaGetLastError());
    return 0;
  }
  if (G.nk <= 40) M6_LAUNCH(c, remap_planes_kernel<40>, grid, block, 0, G, P, G.nk, ilo, ihi, jlo, jhi, mask, h_old, h_new, F);
  else if (G.nk <= 80) M6_LAUNCH(c, remap_planes_kernel<80>, grid, block, 0, G, P, G.nk, ilo, ihi, jlo, jhi, mask, h_old, h_new, F);
  else M6_LAUNCH(c, remap_planes_kernel<128>, grid, block, 0, G, P, G.nk, ilo, ihi, jlo, jhi, mask, h_old, h_new, F);
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int mom6cu_ale_remap_tracers(mom6cu_ctx* c, const mom6cu_remapping_cs* CS, const double* h_old, const double* h_new, int ntr,
                                        double* const* tr, const double* conc_underflow) {
  if (!c || !h_old || !h_new || ntr < 0 || (ntr > 0 && !tr)) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid) return c->fail(MOM6CU_ERR_BAD_ARG, "ALE_remap_tracers: mom6cu_set_grid has not been called");
  Params P;
  int rc;
  if ((rc = check_cs(c, CS, c->g.nk, &P))) return rc;
  Stager S(c, "remap.");
  const double *d_ho, *d_hn;
  if ((rc = S.in3(h_old, ST_H, "h_old", &d_ho)) || (rc = S.in3(h_new, ST_H, "h_new", &d_hn))) return rc;
  std::vector<double*> d_tr(ntr);
  for (int m = 0; m < ntr; ++m) {
    if (!tr[m]) return c->fail(MOM6CU_ERR_BAD_ARG, "ALE_remap_tracers: tracer %d is null", m);
    char name[32]; snprintf(name, sizeof name, "tr%d", m);
    if ((rc = S.io3(tr[m], ST_H, name, &d_tr[m]))) return rc;
  }
  if ((rc = S.begin())) return rc;
  const mom6cu_domain& d = c->dom;
  for (int m0 = 0; m0 < ntr; m0 += MAXF) {
    Fields F = {};
    F.n = std::min(MAXF, ntr - m0);
    for (int m = 0; m < F.n; ++m) { F.p[m] = d_tr[m0 + m]; F.underflow[m] = conc_underflow ? conc_underflow[m0 + m] : 0.0; }
    if ((rc = launch_planes(c, P, d.isc, d.iec, d.jsc, d.jec, c->grid.mask2dT, d_ho, d_hn, F))) return rc;
  }
  return S.finish();
}

extern "C" int mom6cu_ale_remap_set_h_vel(mom6cu_ctx* c, const double* h_new, double* h_u, double* h_v) {
  if (!c || !h_new || !h_u || !h_v) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid) return c->fail(MOM6CU_ERR_BAD_ARG, "ALE_remap_set_h_vel: mom6cu_set_grid has not been called");
  Stager S(c, "remap.");
  const double* d_hn; double *d_hu, *d_hv;
  int rc;
  if ((rc = S.in3(h_new, ST_H, "h_new", &d_hn)) || (rc = S.io3(h_u, ST_U, "h_u", &d_hu)) || (rc = S.io3(h_v, ST_V, "h_v", &d_hv))) return rc;
  if ((rc = S.begin())) return rc;
  const mom6cu_domain& d = c->dom;
  const dim3 grid((d.iec - d.isc + 2 + 127) / 128, d.jec - d.jsc + 2, c->g.nk);
  M6_LAUNCH(c, set_h_vel_kernel, grid, 128, 0, c->g, d.isc, d.iec, d.jsc, d.jec, c->grid.mask2dCu, c->grid.mask2dCv, d_hn, d_hu, d_hv);
  return S.finish();
}

extern "C" int mom6cu_ale_remap_velocities(mom6cu_ctx* c, const mom6cu_remapping_cs* CS, const double* h_old_u, const double* h_old_v,
                                           const double* h_new_u, const double* h_new_v, double* u, double* v) {
  if (!c || !h_old_u || !h_old_v || !h_new_u || !h_new_v || !u || !v) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid) return c->fail(MOM6CU_ERR_BAD_ARG, "ALE_remap_velocities: mom6cu_set_grid has not been called");
  Params P;
  int rc;
  if ((rc = check_cs(c, CS, c->g.nk, &P))) return rc;
  Stager S(c, "remapv.");
  const double *d_hou, *d_hov, *d_hnu, *d_hnv; double *d_u, *d_v;
  if ((rc = S.in3(h_old_u, ST_U, "h_old_u", &d_hou)) || (rc = S.in3(h_old_v, ST_V, "h_old_v", &d_hov)) ||
      (rc = S.in3(h_new_u, ST_U, "h_new_u", &d_hnu)) || (rc = S.in3(h_new_v, ST_V, "h_new_v", &d_hnv)) ||
      (rc = S.io3(u, ST_U, "u", &d_u)) || (rc = S.io3(v, ST_V, "v", &d_v))) return rc;
  if ((rc = S.begin())) return rc;
  const mom6cu_domain& d = c->dom;
  Fields F = {};
  F.n = 1; F.p[0] = d_u;
  if ((rc = launch_planes(c, P, d.isc - 1, d.iec, d.jsc, d.jec, c->grid.mask2dCu, d_hou, d_hnu, F))) return rc;
  F.p[0] = d_v;
  if ((rc = launch_planes(c, P, d.isc, d.iec, d.jsc - 1, d.jec, c->grid.mask2dCv, d_hov, d_hnv, F))) return rc;
  return S.finish();
}

extern "C" int mom6cu_remapping_core_h(mom6cu_ctx* c, const mom6cu_remapping_cs* CS, int ncol, int n0, const double* h0, const double* u0,
                                       int n1, const double* h1, double* u1) {
  if (!c || ncol < 0 || !h0 || !u0 || !h1 || !u1) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  Params P;
  int rc;
  if ((rc = check_cs(c, CS, std::max(n0, n1), &P))) return rc;
  if (n0 < 1 || n1 < 1) return c->fail(MOM6CU_ERR_BAD_ARG, "remapping_core_h: empty columns");
  if (ncol == 0) return 0;
  const size_t s0 = (size_t)ncol * n0, s1 = (size_t)ncol * n1;
  cudaPointerAttributes at = {};
  const bool on_dev = (cudaPointerGetAttributes(&at, u1) == cudaSuccess && at.type == cudaMemoryTypeDevice);
  cudaGetLastError();
  const double *d_h0 = h0, *d_u0 = u0, *d_h1 = h1; double* d_u1 = u1;
  if (!on_dev) {
    double* b = c->buf("remap.batch", 2 * s0 + 2 * s1);
    if (!b) return MOM6CU_ERR_CUDA;
    M6_CUDA(c, cudaMemcpyAsync(b, h0, s0 * 8, cudaMemcpyHostToDevice, c->stream));
    M6_CUDA(c, cudaMemcpyAsync(b + s0, u0, s0 * 8, cudaMemcpyHostToDevice, c->stream));
    M6_CUDA(c, cudaMemcpyAsync(b + 2 * s0, h1, s1 * 8, cudaMemcpyHostToDevice, c->stream));
    d_h0 = b; d_u0 = b + s0; d_h1 = b + 2 * s0; d_u1 = b + 2 * s0 + s1;
  }
  M6_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  const dim3 grid((ncol + 127) / 128), block(128);
  const int nmax = std::max(n0, n1);
  if (use_stream(P)) M6_LAUNCH(c, remap_batch_stream_kernel<1>, grid, block, 0, P, ncol, n0, n1, d_h0, d_u0, d_h1, d_u1);
  else if (nmax <= 40) M6_LAUNCH(c, remap_batch_kernel<40>, grid, block, 0, P, ncol, n0, n1, d_h0, d_u0, d_h1, d_u1);
  else if (nmax <= 80) M6_LAUNCH(c, remap_batch_kernel<80>, grid, block, 0, P, ncol, n0, n1, d_h0, d_u0, d_h1, d_u1);
  else M6_LAUNCH(c, remap_batch_kernel<128>, grid, block, 0, P, ncol, n0, n1, d_h0, d_u0, d_h1, d_u1);
  M6_CUDA(c, cudaGetLastError());
  M6_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  if (!on_dev) M6_CUDA(c, cudaMemcpyAsync(u1, d_u1, s1 * 8, cudaMemcpyDeviceToHost, c->stream));
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  M6_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->last_ms = ms; c->total_ms = ms;
  return 0;
}

extern "C" int mom6cu_remap_dyn_split_rk2_aux_vars(mom6cu_ctx* c, const mom6cu_remapping_cs* remapCS, const mom6cu_dyn_split_rk2_cs* CS,
                                                   const double* h_old_u, const double* h_old_v, const double* h_new_u, const double* h_new_v) {
  if (!c || !CS || !h_old_u || !h_old_v || !h_new_u || !h_new_v) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid) return c->fail(MOM6CU_ERR_BAD_ARG, "remap_dyn_split_RK2_aux_vars: mom6cu_set_grid has not been called");
  if (!CS->diffu || !CS->diffv || (CS->store_CAu && (!CS->u_av || !CS->v_av || !CS->CAu_pred || !CS->CAv_pred)))
    return c->fail(MOM6CU_ERR_BAD_ARG, "remap_dyn_split_RK2_aux_vars: a control-structure array is null");
  Params P;
  int rc;
  if ((rc = check_cs(c, remapCS, c->g.nk, &P))) return rc;
  Stager S(c, "remapx.");
  const double *d_hou, *d_hov, *d_hnu, *d_hnv;
  double *uav = nullptr, *vav = nullptr, *cau = nullptr, *cav = nullptr, *du, *dv;
  if ((rc = S.in3(h_old_u, ST_U, "h_old_u", &d_hou)) || (rc = S.in3(h_old_v, ST_V, "h_old_v", &d_hov)) ||
      (rc = S.in3(h_new_u, ST_U, "h_new_u", &d_hnu)) || (rc = S.in3(h_new_v, ST_V, "h_new_v", &d_hnv)) ||
      (rc = S.io3(CS->diffu, ST_U, "diffu", &du)) || (rc = S.io3(CS->diffv, ST_V, "diffv", &dv))) return rc;
  if (CS->store_CAu && ((rc = S.io3(CS->u_av, ST_U, "u_av", &uav)) || (rc = S.io3(CS->v_av, ST_V, "v_av", &vav)) ||
                        (rc = S.io3(CS->CAu_pred, ST_U, "CAu_pred", &cau)) || (rc = S.io3(CS->CAv_pred, ST_V, "CAv_pred", &cav)))) return rc;
  if ((rc = S.begin())) return rc;
  const mom6cu_domain& d = c->dom;
  auto remap_pair = [&](double* u, double* v) -> int {
    Fields F = {};
    F.n = 1; F.p[0] = u;
    int r = launch_planes(c, P, d.isc - 1, d.iec, d.jsc, d.jec, c->grid.mask2dCu, d_hou, d_hnu, F);
    if (r) return r;
    F.p[0] = v;
    return launch_planes(c, P, d.isc, d.iec, d.jsc - 1, d.jec, c->grid.mask2dCv, d_hov, d_hnv, F);
  };
  if (CS->store_CAu) {
    if ((rc = remap_pair(uav, vav)) || (rc = remap_pair(cau, cav))) return rc;
    double* f[4] = {uav, vav, cau, cav};
    const int st[4] = {ST_U, ST_V, ST_U, ST_V};
    if ((rc = m6_halo_update(c, f, st, 4, 0, c->g.nk))) return rc;  // pass_vector :1322-1325
  }
  if ((rc = remap_pair(du, dv))) return rc;
  return S.finish();
}
