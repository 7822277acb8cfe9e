// ALE regridding, Z* coordinate: ALE_regrid (/root/reference/src/ALE/MOM_ALE.F90:518-554) -> regridding_main
// (src/ALE/MOM_regridding.F90:846-972) -> build_zstar_grid :1257-1367 [build_zstar_column coord_zlike.F90:63-144,
// filtered_grid_motion MOM_regridding.F90:1105-1252, adjust_interface_motion :1796-1857] and calc_h_new_by_dz :1008.
// One thread per column.  The only per-thread array is the column of new interface positions (reused for the
// interface displacements); the old interface heights are re-accumulated from the bottom in each sweep, which repeats the
// reference's own sum order (zOld(k) = zOld(k+1) + h(k)), so every value is bitwise the reference's.
#include "ctx.h"
#include "common.cuh"
#include <cfloat>
#include <vector>

using m6::Geom;
using m6::fmax2;
using m6::fmin2;

namespace {

constexpr int KMAX = 128;

struct RegridK {
  int nk, is, ie, js, je;
  double min_thickness, old_grid_weight, zs, zd, Z_ref, Z_to_H;
  const double *mask2dT, *bathyT, *res;  // res: coordinateResolution(nk) on the device
  const double* h; double* h_new; double* dz;
  int* err;
};

// filtered_grid_motion :1179-1232 for one interface
__device__ __forceinline__ double filtered_dz(double sgn, double z_new_k, double z_old_k, double z_old_1, double zs, double zd, double wtd,
                                              double Iwtd, double dzwt, double Idzwt, double dInt_zs_zd, double Aq) {
  const double dz_tgt = sgn * (z_new_k - z_old_k);
  const double zr1 = sgn * (z_old_k - z_old_1);
  if ((zr1 > zd) && (zr1 + wtd * dz_tgt > zd)) return sgn * wtd * dz_tgt;
  if ((zr1 < zs) && (zr1 + dz_tgt < zs)) return sgn * dz_tgt;
  double Int_zd, Int_zs;
  if (zr1 >= zd) { Int_zd = Iwtd * (zd - zr1); Int_zs = Int_zd - dInt_zs_zd; }
  else if (zr1 <= zs) { Int_zs = (zs - zr1); Int_zd = dInt_zs_zd + (zs - zr1); }
  else {
    Int_zd = (zd - zr1) * (Iwtd * (0.5 * (zd + zr1) - zs) + 0.5 * (zd - zr1)) * Idzwt;
    Int_zs = (zs - zr1) * (0.5 * Iwtd * ((zr1 - zs)) + (zd - 0.5 * (zr1 + zs))) * Idzwt;
  }
  if (dz_tgt >= Int_zd) return sgn * ((zd - zr1) + wtd * (dz_tgt - Int_zd));
  if (dz_tgt <= Int_zs) return sgn * ((zs - zr1) + (dz_tgt - Int_zs));
  double dz0, z0, F0;
  if (zr1 <= zs) { dz0 = zs - zr1; z0 = zs; F0 = dz_tgt - Int_zs; }
  else if (zr1 >= zd) { dz0 = zd - zr1; z0 = zd; F0 = dz_tgt - Int_zd; }
  else { dz0 = 0.0; z0 = zr1; F0 = dz_tgt; }
  const double Bq = (dzwt + 2.0 * Aq * (z0 - zs));
  return sgn * (dz0 + 2.0 * F0 * dzwt / (Bq + sqrt(Bq * Bq + 4.0 * Aq * F0 * dzwt)));
}

__global__ void __launch_bounds__(128) regrid_zstar_kernel(Geom G, RegridK P) {
  const int i = P.is - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = P.js - 1 + blockIdx.y;
  if (i > P.ie + 1) return;
  const long long g = G.idx(i, j), pl = G.plane;
  const int nk = P.nk;
  const double* __restrict__ h = P.h + g;
  double* __restrict__ dzI = P.dz + g;
  double* __restrict__ hn = P.h_new + g;
  if (P.mask2dT[g] == 0.) {  // :1299-1302 and calc_h_new_by_dz :1035
    for (int k = 0; k <= nk; ++k) dzI[k * pl] = 0.;
    for (int k = 0; k < nk; ++k) hn[k * pl] = h[k * pl];
    return;
  }
  double z[KMAX + 2];  // 1-based: zNew, then the interface displacements
  const double depth = fmax2((P.bathyT[g] + P.Z_ref) * P.Z_to_H, 0.0);  // regridding_main :918
  double total = 0.0, hmin = h[0];
  for (int k = 1; k <= nk; ++k) { const double hk = h[(k - 1) * pl]; total = total + hk; hmin = fmin2(hmin, hk); }
  if (hmin < 0.0 && i >= P.is && i <= P.ie && j >= P.js && j <= P.je) atomicCAS(P.err, 0, 14);  // :962-969
  // build_zstar_column
  const double min_thk = fmin2(P.min_thickness, total / (double)nk);
  const double eta = total - depth;
  const double stretching = total / (depth + 0.);
  z[1] = eta;
  for (int k = 1; k <= nk; ++k) z[k + 1] = z[k] - stretching * P.res[k - 1] * P.Z_to_H;
  z[nk + 1] = -depth;
  for (int k = nk; k >= 1; --k) if (z[k] < (z[k + 1] + min_thk)) z[k] = z[k + 1] + min_thk;
  // zOld(1)
  double zo = -depth;
  for (int k = nk; k >= 1; --k) zo = zo + h[(k - 1) * pl];
  const double z_old_1 = zo, z_old_b = -depth;
  // filtered_grid_motion
  const double prod = (z_old_b - z_old_1) * (z[nk + 1] - z[1]);
  if (prod < 0.0) atomicCAS(P.err, 0, 11);
  if (prod == 0.0) { for (int k = 1; k <= nk + 1; ++k) z[k] = 0.0; }
  else {
    const double sgn = ((z_old_b - z_old_1) + (z[nk + 1] - z[1]) > 0.0) ? 1.0 : -1.0;
    const double zs = P.zs, zd = P.zd, wtd = 1.0 - P.old_grid_weight, Iwtd = 1.0 / wtd, dzwt = (zd - zs);
    const double Idzwt = (fabs(zd - zs) > 0.0) ? 1.0 / (zd - zs) : 0.0;
    const double dInt_zs_zd = 0.5 * (1.0 + Iwtd) * (zd - zs), Aq = 0.5 * (Iwtd - 1.0);
    zo = -depth;
    for (int k = nk + 1; k >= 2; --k) {
      if (k <= nk) zo = zo + h[(k - 1) * pl];  // zOld(k) = zOld(k+1) + h(k)
      z[k] = filtered_dz(sgn, z[k], zo, z_old_1, zs, zd, wtd, Iwtd, dzwt, Idzwt, dInt_zs_zd, Aq);
    }
    z[1] = 0.0;
  }
  // adjust_interface_motion: the roundoff test (top down) ...
  const double eps = DBL_EPSILON;
  double h_err = 0.;
  for (int k = 1; k <= nk; ++k) {
    const double hk = h[(k - 1) * pl];
    h_err = h_err + fmax2(fmax2(hk, fabs(z[k])), fabs(z[k + 1])) * eps;
    if (hk + (z[k] - z[k + 1]) < -3.0 * h_err) atomicCAS(P.err, 0, 12);
  }
  // ... and the minimum-thickness adjustment (bottom up), fused with calc_h_new_by_dz
  dzI[(long long)nk * pl] = z[nk + 1];
  for (int k = nk; k >= 2; --k) {
    const double hk = h[(k - 1) * pl];
    double h_new = hk + (z[k] - z[k + 1]);
    if (h_new < P.min_thickness) z[k] = (z[k + 1] - hk) + P.min_thickness;
    h_new = hk + (z[k] - z[k + 1]);
    if (h_new < 0.) z[k] = (1. - eps) * (z[k + 1] - hk);
    h_new = hk + (z[k] - z[k + 1]);
    if (h_new < 0.) atomicCAS(P.err, 0, 13);
    dzI[(long long)(k - 1) * pl] = z[k];
    hn[(long long)(k - 1) * pl] = fmax2(0., hk + (z[k] - z[k + 1]));
  }
  dzI[0] = z[1];
  hn[0] = fmax2(0., h[0] + (z[1] - z[2]));
}

}  // namespace

extern "C" int mom6cu_ale_regrid(mom6cu_ctx* c, const mom6cu_regridding_cs* CS, const double* h, double* h_new, double* dzRegrid) {
  if (!c || !CS || !h || !h_new || !dzRegrid) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "ALE_regrid: mom6cu_set_grid / mom6cu_set_vgrid have not been called");
  const Geom& G = c->g;
  if (CS->regridding_scheme != MOM6CU_REGRIDDING_ZSTAR)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "ALE_regrid: only the Z* coordinate (REGRIDDING_ZSTAR = %d) is implemented, not %d", MOM6CU_REGRIDDING_ZSTAR,
                   CS->regridding_scheme);
  if (CS->nk != G.nk) return c->fail(MOM6CU_ERR_UNSUPPORTED, "ALE_regrid: CS%%nk = %d differs from GV%%ke = %d", CS->nk, G.nk);
  if (G.nk > KMAX) return c->fail(MOM6CU_ERR_UNSUPPORTED, "ALE_regrid: %d layers exceed the %d-layer column capacity", G.nk, KMAX);
  if (!CS->coordinateResolution) return c->fail(MOM6CU_ERR_BAD_ARG, "ALE_regrid: coordinateResolution is null");
  if (!c->vgrid.Boussinesq) return c->fail(MOM6CU_ERR_UNSUPPORTED, "ALE_regrid: the non-Boussinesq nominal depth (tv%%SpV_avg) is not implemented");
  Stager S(c, "regrid.");
  int rc;
  const double* d_h; double *d_hn, *d_dz;
  if ((rc = S.in3(h, ST_H, "h", &d_h)) || (rc = S.io3(h_new, ST_H, "h_new", &d_hn)) || (rc = S.io(dzRegrid, ST_H, 0, G.nk + 1, "dz", &d_dz))) return rc;
  double* d_res = c->buf("regrid.res", KMAX);
  int* d_err = (int*)c->buf("regrid.err", 2);
  int* h_err = (int*)c->host_scratch("regrid.err", 2);
  if (!d_res || !d_err || !h_err) return MOM6CU_ERR_CUDA;
  M6_CUDA(c, cudaMemcpyAsync(d_res, CS->coordinateResolution, sizeof(double) * G.nk, cudaMemcpyHostToDevice, c->stream));
  M6_CUDA(c, cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  if ((rc = S.begin())) return rc;
  // ALE_regrid :544 dzRegrid(:,:,:) = 0.0
  M6_CUDA(c, cudaMemsetAsync(d_dz, 0, sizeof(double) * (size_t)G.plane * (G.nk + 1), c->stream));
  const mom6cu_domain& d = c->dom;
  RegridK P = {G.nk, d.isc, d.iec, d.jsc, d.jec, CS->min_thickness, CS->old_grid_weight, CS->depth_of_time_filter_shallow,
               CS->depth_of_time_filter_deep, CS->Z_ref, c->US.Z_to_m * c->vgrid.m_to_H, c->grid.mask2dT, c->grid.bathyT, d_res, d_h, d_hn, d_dz, d_err};
  const dim3 grid((d.iec - d.isc + 3 + 127) / 128, d.jec - d.jsc + 3);
  M6_LAUNCH(c, regrid_zstar_kernel, grid, 128, 0, G, P);
  M6_CUDA(c, cudaGetLastError());
  M6_CUDA(c, cudaMemcpyAsync(h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if ((rc = S.finish())) return rc;
  switch (h_err[0]) {
    case 0: return 0;
    case 11: return c->fail(MOM6CU_ERR_BAD_ARG, "filtered_grid_motion: z_old and z_new use different sign conventions.");
    case 12: return c->fail(MOM6CU_ERR_BAD_ARG, "MOM_regridding: adjust_interface_motion() - implied h<0 is larger than roundoff!");
    case 13: return c->fail(MOM6CU_ERR_BAD_ARG, "MOM_regridding: adjust_interface_motion() - Repeated adjustment for roundoff h<0 failed!");
    default: return c->fail(MOM6CU_ERR_BAD_ARG, "regridding_main: negative thickness encountered.");
  }
}
