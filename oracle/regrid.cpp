// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of the Z* regridding path, /root/reference/src/ALE/: ALE_regrid MOM_ALE.F90:518-554 -> regridding_main
// MOM_regridding.F90:846-972 (Boussinesq branch :917-920, REGRIDDING_ZSTAR :925-927, negative-thickness check :962-969)
// -> build_zstar_grid :1257-1367 -> build_zstar_column coord_zlike.F90:63-144 (no rigid top), filtered_grid_motion
// MOM_regridding.F90:1105-1252, adjust_interface_motion :1796-1857; calc_h_new_by_dz :1008-1042.
// PARITY: PINNED BY A REFERENCE RUN -- ALE_regrid as called by the reference's own ALE_regridding_and_remapping (MOM.F90), executed by
// oracle/f90run, agrees bit for bit (tests/test_reference_f90.py, ale/*).
#include "oracle.h"
#include "ogrid.hpp"
#include <cfloat>
#include <cmath>
#include <vector>
#include <algorithm>

using namespace orc;

namespace {

// build_zstar_column, coord_zlike.F90:63-144 (z_rigid_top, eta_orig absent).  zInterface is 1-based (nk+1).
void build_zstar_column(const mom6cu_regridding_cs* CS, double depth, double total_thickness, double* zInterface, double z_scale) {
  const int nk = CS->nk;
  const double min_thickness = fmin2(CS->min_thickness, total_thickness / (double)nk);
  const double z0_top = 0.;
  const double eta = total_thickness - depth;
  const double stretching = total_thickness / (depth + z0_top);
  zInterface[1] = eta;
  for (int k = 1; k <= nk; ++k) {
    const double dh = stretching * CS->coordinateResolution[k - 1] * z_scale;
    zInterface[k + 1] = zInterface[k] - dh;
  }
  zInterface[nk + 1] = -depth;
  for (int k = nk; k >= 1; --k)
    if (zInterface[k] < (zInterface[k + 1] + min_thickness)) zInterface[k] = zInterface[k + 1] + min_thickness;
}

// filtered_grid_motion :1105-1252; returns 1 for the FATAL of :1146
int filtered_grid_motion(const mom6cu_regridding_cs* CS, int nk, const double* z_old, const double* z_new, double* dz_g) {
  const int cnk = CS->nk;
  double sgn;
  const double prod = (z_old[nk + 1] - z_old[1]) * (z_new[cnk + 1] - z_new[1]);
  if (prod < 0.0) return 1;
  else if (prod == 0.0) { for (int k = 1; k <= cnk + 1; ++k) dz_g[k] = 0.0; return 0; }
  else if ((z_old[nk + 1] - z_old[1]) + (z_new[cnk + 1] - z_new[1]) > 0.0) sgn = 1.0;
  else sgn = -1.0;
  const double zs = CS->depth_of_time_filter_shallow, zd = CS->depth_of_time_filter_deep;
  const double wtd = 1.0 - CS->old_grid_weight;
  const double Iwtd = 1.0 / wtd;
  const double dzwt = (zd - zs);
  double Idzwt = 0.0; if (std::fabs(zd - zs) > 0.0) Idzwt = 1.0 / (zd - zs);
  const double dInt_zs_zd = 0.5 * (1.0 + Iwtd) * (zd - zs);
  const double Aq = 0.5 * (Iwtd - 1.0);
  dz_g[1] = 0.0;
  double z_old_k = z_old[1];
  for (int k = 2; k <= cnk + 1; ++k) {
    if (k <= nk + 1) z_old_k = z_old[k];
    const double dz_tgt = sgn * (z_new[k] - z_old_k);
    const double zr1 = sgn * (z_old_k - z_old[1]);
    if ((zr1 > zd) && (zr1 + wtd * dz_tgt > zd)) dz_g[k] = sgn * wtd * dz_tgt;
    else if ((zr1 < zs) && (zr1 + dz_tgt < zs)) dz_g[k] = sgn * dz_tgt;
    else {
      double Int_zd, Int_zs;
      if (zr1 >= zd) { Int_zd = Iwtd * (zd - zr1); Int_zs = Int_zd - dInt_zs_zd; }
      else if (zr1 <= zs) { Int_zs = (zs - zr1); Int_zd = dInt_zs_zd + (zs - zr1); }
      else {
        Int_zd = (zd - zr1) * (Iwtd * (0.5 * (zd + zr1) - zs) + 0.5 * (zd - zr1)) * Idzwt;
        Int_zs = (zs - zr1) * (0.5 * Iwtd * ((zr1 - zs)) + (zd - 0.5 * (zr1 + zs))) * Idzwt;
      }
      if (dz_tgt >= Int_zd) dz_g[k] = sgn * ((zd - zr1) + wtd * (dz_tgt - Int_zd));
      else if (dz_tgt <= Int_zs) dz_g[k] = sgn * ((zs - zr1) + (dz_tgt - Int_zs));
      else {
        double dz0, z0, F0;
        if (zr1 <= zs) { dz0 = zs - zr1; z0 = zs; F0 = dz_tgt - Int_zs; }
        else if (zr1 >= zd) { dz0 = zd - zr1; z0 = zd; F0 = dz_tgt - Int_zd; }
        else { dz0 = 0.0; z0 = zr1; F0 = dz_tgt; }
        const double Bq = (dzwt + 2.0 * Aq * (z0 - zs));
        dz_g[k] = sgn * (dz0 + 2.0 * F0 * dzwt / (Bq + std::sqrt(Bq * Bq + 4.0 * Aq * F0 * dzwt)));
      }
    }
  }
  return 0;
}

// adjust_interface_motion :1796-1857 (CS%nk == nk); returns 1 / 2 for its two FATALs
int adjust_interface_motion(const mom6cu_regridding_cs* CS, int nk, const double* h_old, double* dz_int) {
  const double eps = DBL_EPSILON;
  double h_total = 0., h_err = 0.;
  for (int k = 1; k <= nk; ++k) {
    h_total = h_total + h_old[k];
    h_err = h_err + fmax2(fmax2(h_old[k], std::fabs(dz_int[k])), std::fabs(dz_int[k + 1])) * eps;
    const double h_new = h_old[k] + (dz_int[k] - dz_int[k + 1]);
    if (h_new < -3.0 * h_err) return 1;
  }
  for (int k = nk; k >= 2; --k) {
    double h_new = h_old[k] + (dz_int[k] - dz_int[k + 1]);
    if (h_new < CS->min_thickness) dz_int[k] = (dz_int[k + 1] - h_old[k]) + CS->min_thickness;
    h_new = h_old[k] + (dz_int[k] - dz_int[k + 1]);
    if (h_new < 0.) dz_int[k] = (1. - eps) * (dz_int[k + 1] - h_old[k]);
    h_new = h_old[k] + (dz_int[k] - dz_int[k + 1]);
    if (h_new < 0.) return 2;
  }
  return 0;
}

}  // namespace

// returns 0, or 10 + code of the first FATAL met (11 sign conventions, 12 implied h<0, 13 repeated adjustment failed, 14 negative h)
extern "C" int oracle_ale_regrid(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV, const mom6cu_unit_scale* US,
                                 const mom6cu_regridding_cs* CS, const double* hp, double* h_newp, double* dzp) {
  const OGrid G(d, Gp);
  const int nz = G.ke;
  if (CS->regridding_scheme != MOM6CU_REGRIDDING_ZSTAR || CS->nk != nz) return 3;
  const V3 h = G.H3(hp), h_new = G.H3(h_newp), dzI = G.H3(dzp, nz + 1);
  const double Z_to_H = US->Z_to_m * GV->m_to_H;
  dzI.fill(0.0);  // ALE_regrid :544
  std::vector<double> zOld(nz + 2), zNew(nz + 2), dz(nz + 2), hcol(nz + 2);
  int rc = 0;
  for (int j = G.jsc - 1; j <= G.jec + 1; ++j) for (int i = G.isc - 1; i <= G.iec + 1; ++i) {
    if (G.mask2dT(i, j) == 0.) { for (int k = 1; k <= nz + 1; ++k) dzI(i, j, k) = 0.; continue; }
    const double nominalDepth = fmax2((G.bathyT(i, j) + CS->Z_ref) * Z_to_H, 0.0);  // regridding_main :918
    double totalThickness = 0.0;
    for (int k = 1; k <= nz; ++k) totalThickness = totalThickness + h(i, j, k);
    zOld[nz + 1] = -nominalDepth;
    for (int k = nz; k >= 1; --k) zOld[k] = zOld[k + 1] + h(i, j, k);
    build_zstar_column(CS, nominalDepth, totalThickness, zNew.data(), Z_to_H);
    for (int k = 1; k <= nz + 1; ++k) dz[k] = dzI(i, j, k);
    if (filtered_grid_motion(CS, nz, zOld.data(), zNew.data(), dz.data()) && !rc) rc = 11;
    for (int k = 1; k <= nz; ++k) hcol[k] = h(i, j, k);
    const int r2 = adjust_interface_motion(CS, nz, hcol.data(), dz.data());
    if (r2 && !rc) rc = 11 + r2;
    for (int k = 1; k <= nz + 1; ++k) dzI(i, j, k) = dz[k];
  }
  // calc_h_new_by_dz :1008-1042
  for (int j = G.jsc - 1; j <= G.jec + 1; ++j) for (int i = G.isc - 1; i <= G.iec + 1; ++i) {
    if (G.mask2dT(i, j) > 0.) for (int k = 1; k <= nz; ++k) h_new(i, j, k) = fmax2(0., h(i, j, k) + (dzI(i, j, k) - dzI(i, j, k + 1)));
    else for (int k = 1; k <= nz; ++k) h_new(i, j, k) = h(i, j, k);
  }
  for (int j = G.jsc; j <= G.jec; ++j) for (int i = G.isc; i <= G.iec; ++i) if (G.mask2dT(i, j) > 0.) {
    double mn = h(i, j, 1);
    for (int k = 2; k <= nz; ++k) mn = fmin2(mn, h(i, j, k));
    if (mn < 0.0 && !rc) rc = 14;
  }
  return rc;
}
