// ALE_regridding_and_remapping (src/core/MOM.F90:1751-1926): the thermodynamic-cadence pass that follows the dynamics steps,
// as ONE entry on resident fields -- halo update of T, S, h; ALE_update_regrid_weights; ALE_regrid; ALE_remap_tracers;
// ALE_remap_set_h_vel (old and new grid); ALE_remap_velocities; remap_dyn_split_RK2_aux_vars; remap_vertvisc_aux_vars
// (src/parameterizations/vertical/MOM_set_viscosity.F90:2849-2873 -> ALE_remap_interface_vals MOM_ALE.F90:1303-1339,
// ALE_remap_vertex_vals :1342-1382 -> interpolate_column MOM_remapping.F90:1247-1314); h = h_new.
// The stages are the kernels of regrid.cu / remap.cu called on planes (no host traffic between them); new here are the
// interface / vertex interpolation kernels (one thread per column, interp_column.cuh) and the final copy.
#include "ctx.h"
#include "interp_column.cuh"
#include <vector>

using m6::Geom;

namespace {

constexpr int KCAP = 128;  // layers per column the interpolation kernels hold in thread-local memory

// ALE_remap_interface_vals :1322-1337 (VERTEX = false) / ALE_remap_vertex_vals :1362-1381 (VERTEX = true)
template <bool VERTEX>
__global__ void __launch_bounds__(128) interp_vals_kernel(const Geom G, const int nk, const int ilo, const int ihi, const int jlo, const int jhi,
                                                          const double* __restrict__ mask2dT, const double* __restrict__ h_old,
                                                          const double* __restrict__ h_new, double* __restrict__ val) {
  const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x, j = jlo + blockIdx.y;
  if (i > ihi || j > jhi) return;
  const long long g = G.idx(i, j), pl = G.plane;
  double m00 = mask2dT[g], m11 = 0., m10 = 0., m01 = 0., I_mask_sum = 0.;
  if (VERTEX) {
    m11 = mask2dT[g + G.pitch + 1]; m10 = mask2dT[g + 1]; m01 = mask2dT[g + G.pitch];
    if (!((m00 + m11) + (m10 + m01) > 0.0)) return;
    I_mask_sum = 1.0 / ((m00 + m11) + (m10 + m01));
  } else if (!(m00 > 0.)) return;
  double vs[KCAP + 1];
  for (int k = 0; k <= nk; ++k) vs[k] = val[g + (long long)k * pl];
  auto thick = [&](const double* h, int k) -> double {  // k = 1..nk
    const double* p = h + g + (long long)(k - 1) * pl;
    if (!VERTEX) return p[0];
    return ((m00 * p[0] + m11 * p[G.pitch + 1]) + (m10 * p[1] + m01 * p[G.pitch])) * I_mask_sum;
  };
  m6interp::interpolate_column(nk, [&](int k) { return thick(h_old, k); }, [&](int k) { return vs[k - 1]; }, nk,
                               [&](int k) { return thick(h_new, k); }, [&](int k, double v) { val[g + (long long)(k - 1) * pl] = v; }, false);
}

// one column per thread, contiguous columns (the reference's unit-test vectors through the C ABI)
__global__ void interp_batch_kernel(const int ncol, const int nsrc, const int ndest, const double* __restrict__ h_src,
                                    const double* __restrict__ u_src, const double* __restrict__ h_dest, double* __restrict__ u_dest,
                                    const int mask_edges) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  const double *hs = h_src + (long long)c * nsrc - 1, *us = u_src + (long long)c * (nsrc + 1) - 1, *hd = h_dest + (long long)c * ndest - 1;
  double* ud = u_dest + (long long)c * (ndest + 1) - 1;
  m6interp::interpolate_column(nsrc, [=](int k) { return hs[k]; }, [=](int k) { return us[k]; }, ndest, [=](int k) { return hd[k]; },
                               [=](int k, double v) { ud[k] = v; }, mask_edges != 0);
}

// "h(i,j,k) = h_new(i,j,k)" on is-1..ie+1, js-1..je+1 (MOM.F90:1875-1878)
__global__ void copy_window_kernel(const Geom G, const int ilo, const int ihi, const int jlo, const int jhi, const double* __restrict__ src,
                                   double* __restrict__ dst) {
  const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x, j = jlo + blockIdx.y;
  if (i > ihi || j > jhi) return;
  const long long g = G.idx(i, j) + (long long)blockIdx.z * G.plane;
  dst[g] = src[g];
}

int run_interp(mom6cu_ctx* c, bool vertex, const double* h_old, const double* h_new, double* val) {
  const mom6cu_domain& d = c->dom;
  if (c->g.nk > KCAP) return c->fail(MOM6CU_ERR_UNSUPPORTED, "ALE_remap_interface_vals: %d layers exceed the %d-layer column capacity", c->g.nk, KCAP);
  const int ilo = d.isc - (vertex ? 1 : 0), jlo = d.jsc - (vertex ? 1 : 0);
  const dim3 grid((d.iec - ilo + 1 + 127) / 128, d.jec - jlo + 1);
  if (vertex) M6_LAUNCH(c, interp_vals_kernel<true>, grid, 128, 0, c->g, c->g.nk, ilo, d.iec, jlo, d.jec, c->grid.mask2dT, h_old, h_new, val);
  else M6_LAUNCH(c, interp_vals_kernel<false>, grid, 128, 0, c->g, c->g.nk, ilo, d.iec, jlo, d.jec, c->grid.mask2dT, h_old, h_new, val);
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

}  // namespace

static int remap_vals_entry(mom6cu_ctx* c, bool vertex, const double* h_old, const double* h_new, double* val) {
  if (!c || !h_old || !h_new || !val) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid) return c->fail(MOM6CU_ERR_BAD_ARG, "ALE_remap_%s_vals: mom6cu_set_grid has not been called", vertex ? "vertex" : "interface");
  Stager S(c, vertex ? "ivx." : "ivl.");
  int rc;
  const double *d_ho, *d_hn; double* d_val;
  if ((rc = S.in3(h_old, ST_H, "h_old", &d_ho)) || (rc = S.in3(h_new, ST_H, "h_new", &d_hn)) ||
      (rc = S.io(val, vertex ? ST_Q : ST_H, 0, c->g.nk + 1, "val", &d_val))) return rc;
  if ((rc = S.begin()) || (rc = run_interp(c, vertex, d_ho, d_hn, d_val))) return rc;
  return S.finish();
}

extern "C" int mom6cu_ale_remap_interface_vals(mom6cu_ctx* c, const double* h_old, const double* h_new, double* int_val) {
  return remap_vals_entry(c, false, h_old, h_new, int_val);
}
extern "C" int mom6cu_ale_remap_vertex_vals(mom6cu_ctx* c, const double* h_old, const double* h_new, double* vert_val) {
  return remap_vals_entry(c, true, h_old, h_new, vert_val);
}

extern "C" int mom6cu_interpolate_column(mom6cu_ctx* c, int ncol, int nsrc, const double* h_src, const double* u_src, int ndest,
                                         const double* h_dest, double* u_dest, int mask_edges) {
  if (!c || ncol < 1 || nsrc < 1 || ndest < 1 || !h_src || !u_src || !h_dest || !u_dest) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  const size_t n_hs = (size_t)ncol * nsrc, n_us = (size_t)ncol * (nsrc + 1), n_hd = (size_t)ncol * ndest, n_ud = (size_t)ncol * (ndest + 1);
  double* b = c->buf("interp.batch", n_hs + n_us + n_hd + n_ud);
  if (!b) return MOM6CU_ERR_CUDA;
  double *d_hs = b, *d_us = d_hs + n_hs, *d_hd = d_us + n_us, *d_ud = d_hd + n_hd;
  M6_CUDA(c, cudaMemcpyAsync(d_hs, h_src, n_hs * 8, cudaMemcpyDefault, c->stream));
  M6_CUDA(c, cudaMemcpyAsync(d_us, u_src, n_us * 8, cudaMemcpyDefault, c->stream));
  M6_CUDA(c, cudaMemcpyAsync(d_hd, h_dest, n_hd * 8, cudaMemcpyDefault, c->stream));
  M6_LAUNCH(c, interp_batch_kernel, (ncol + 127) / 128, 128, 0, ncol, nsrc, ndest, d_hs, d_us, d_hd, d_ud, mask_edges);
  M6_CUDA(c, cudaGetLastError());
  M6_CUDA(c, cudaMemcpyAsync(u_dest, d_ud, n_ud * 8, cudaMemcpyDefault, c->stream));
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int mom6cu_ale_regridding_and_remapping(mom6cu_ctx* c, mom6cu_ale_cs* CS, const mom6cu_dyn_split_rk2_cs* dynCS,
                                                   const mom6cu_ale_args* a) {
  if (!c || !CS || !a || !a->u || !a->v || !a->h || a->ntr < 0 || (a->ntr > 0 && !a->tr)) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "ALE_regridding_and_remapping: mom6cu_set_grid / mom6cu_set_vgrid have not been called");
  if (CS->remap_uv_using_old_alg || CS->do_conv_adj || CS->use_hybgen_unmix)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "ALE_regridding_and_remapping: REMAP_UV_USING_OLD_ALG, convective adjustment and hybgen unmixing are not implemented");
  if (CS->remap_aux_vars && !dynCS) return c->fail(MOM6CU_ERR_BAD_ARG, "ALE_regridding_and_remapping: remap_aux_vars without the MOM_dyn_split_RK2_CS");
  if (a->iT >= a->ntr || a->iS >= a->ntr) return c->fail(MOM6CU_ERR_BAD_ARG, "ALE_regridding_and_remapping: tv%%T / tv%%S index outside the tracer registry");
  const Geom& G = c->g;
  const mom6cu_domain& d = c->dom;
  const int nz = G.nk;
  Stager S(c, "ale.");
  int rc;
  double *d_u, *d_v, *d_h, *d_Kd = nullptr, *d_Kv = nullptr, *d_KvB = nullptr;
  if ((rc = S.io3(a->u, ST_U, "u", &d_u)) || (rc = S.io3(a->v, ST_V, "v", &d_v)) || (rc = S.io3(a->h, ST_H, "h", &d_h))) return rc;
  std::vector<double*> d_tr(a->ntr);
  for (int m = 0; m < a->ntr; ++m) {
    if (!a->tr[m]) return c->fail(MOM6CU_ERR_BAD_ARG, "ALE_regridding_and_remapping: tracer %d is null", m);
    char name[32]; snprintf(name, sizeof name, "tr%d", m);
    if ((rc = S.io3(a->tr[m], ST_H, name, &d_tr[m]))) return rc;
  }
  if (CS->remap_aux_vars) {
    if (a->Kd_shear && (rc = S.io(a->Kd_shear, ST_H, 0, nz + 1, "Kd_shear", &d_Kd))) return rc;
    if (a->Kv_shear && (rc = S.io(a->Kv_shear, ST_H, 0, nz + 1, "Kv_shear", &d_Kv))) return rc;
    if (a->Kv_shear_Bu && (rc = S.io(a->Kv_shear_Bu, ST_Q, 0, nz + 1, "Kv_shear_Bu", &d_KvB))) return rc;
  }
  double* d_hn = c->plane3("ale.h_new");
  double* d_dz = c->plane3k("ale.dzRegrid", nz + 1);
  double *d_hou = c->plane3("ale.h_old_u"), *d_hov = c->plane3("ale.h_old_v"), *d_hnu = c->plane3("ale.h_new_u"), *d_hnv = c->plane3("ale.h_new_v");
  if (!d_hn || !d_dz || !d_hou || !d_hov || !d_hnu || !d_hnv) return MOM6CU_ERR_CUDA;
  M6_CUDA(c, cudaStreamSynchronize(c->stream));  // the uploads are done; the stage entries below time themselves
  double total_ms = 0.0;
  cudaEvent_t e0 = c->ev0;  // (the stage entries reuse ev0 / ev1; their times are summed instead)
  (void)e0;
  // ---- pass_T_S_h :1800-1806
  {
    double* f[3]; int st[3] = {ST_H, ST_H, ST_H}; int n = 0;
    if (a->iT >= 0) f[n++] = d_tr[a->iT];
    if (a->iS >= 0) f[n++] = d_tr[a->iS];
    f[n++] = d_h;
    if ((rc = m6_halo_update(c, f, st, n, 0, nz))) return rc;
  }
  // ---- ALE_update_regrid_weights (MOM_ALE.F90:1719-1733)
  {
    double w = 0.0;
    if (CS->regrid_time_scale > 0.0) w = CS->regrid_time_scale / (CS->regrid_time_scale + a->dtdia);
    CS->regridCS.old_grid_weight = w;
  }
  // ---- ALE_regrid :1831-1835, ALE_remap_tracers :1839
  if ((rc = mom6cu_ale_regrid(c, &CS->regridCS, d_h, d_hn, d_dz))) return rc;
  total_ms += c->last_ms;
  if (a->ntr > 0) {
    if ((rc = mom6cu_ale_remap_tracers(c, &CS->remapCS, d_h, d_hn, a->ntr, d_tr.data(), a->conc_underflow))) return rc;
    total_ms += c->last_ms;
  }
  // ---- thicknesses at velocity points on the old and the new grid :1842-1847, velocities :1850
  if ((rc = mom6cu_ale_remap_set_h_vel(c, d_h, d_hou, d_hov))) return rc;
  total_ms += c->last_ms;
  if ((rc = mom6cu_ale_remap_set_h_vel(c, d_hn, d_hnu, d_hnv))) return rc;
  total_ms += c->last_ms;
  if ((rc = mom6cu_ale_remap_velocities(c, &CS->vel_remapCS, d_hou, d_hov, d_hnu, d_hnv, d_u, d_v))) return rc;
  total_ms += c->last_ms;
  if (CS->remap_aux_vars) {  // :1855-1871
    if ((rc = mom6cu_remap_dyn_split_rk2_aux_vars(c, &CS->vel_remapCS, dynCS, d_hou, d_hov, d_hnu, d_hnv))) return rc;
    total_ms += c->last_ms;
    if ((rc = S.begin())) return rc;
    if (d_Kd && (rc = run_interp(c, false, d_h, d_hn, d_Kd))) return rc;
    if (d_Kv && (rc = run_interp(c, false, d_h, d_hn, d_Kv))) return rc;
    if (d_KvB && (rc = run_interp(c, true, d_h, d_hn, d_KvB))) return rc;
    if (d_Kv) {  // pass_var(CS%visc%Kv_shear, G%Domain, To_All+Omit_Corners, halo=1) :1870
      double* f[1] = {d_Kv}; int st[1] = {ST_H};
      if ((rc = m6_halo_update(c, f, st, 1, 0, nz + 1))) return rc;
    }
  } else if ((rc = S.begin())) return rc;
  // ---- replace the old grid with the new one :1875-1878
  {
    const dim3 grid((d.iec - d.isc + 3 + 127) / 128, d.jec - d.jsc + 3, nz);
    M6_LAUNCH(c, copy_window_kernel, grid, 128, 0, G, d.isc - 1, d.iec + 1, d.jsc - 1, d.jec + 1, d_hn, d_h);
    M6_CUDA(c, cudaGetLastError());
  }
  if ((rc = S.finish())) return rc;
  c->last_ms += total_ms; c->total_ms = c->last_ms;
  return 0;
}
