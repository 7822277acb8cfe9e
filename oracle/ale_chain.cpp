// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of interpolate_column (/root/reference/src/ALE/MOM_remapping.F90:1247-1314), ALE_remap_interface_vals
// (src/ALE/MOM_ALE.F90:1303-1339), ALE_remap_vertex_vals (:1342-1382), remap_vertvisc_aux_vars
// (src/parameterizations/vertical/MOM_set_viscosity.F90:2849-2873), ALE_update_regrid_weights (MOM_ALE.F90:1719-1733) and
// of the order of operations of ALE_regridding_and_remapping (src/core/MOM.F90:1751-1926; no OBCs, ice shelves, particles,
// diagnostics), which calls the routines restated in regrid.cpp / remap.cpp.
// PARITY: interpolate_column is PINNED by the reference's unit-test vectors (MOM_remapping.F90:2648-2682), see
// tests/test_ale_chain.py; the chain itself is PINNED BY A REFERENCE RUN: ALE_regridding_and_remapping of the reference's own MOM.F90,
// executed by oracle/f90run, agrees bit for bit on 4 configurations (tests/test_reference_f90.py, ale/*).
#include "oracle.h"
#include "ogrid.hpp"
#include <vector>

using namespace orc;

// interpolate_column :1247-1314.  All arrays 0-based here (h_src[nsrc], u_src[nsrc+1], h_dest[ndest], u_dest[ndest+1]).
extern "C" void oracle_interpolate_column(int nsrc, const double* h_src, const double* u_src, int ndest, const double* h_dest,
                                          double* u_dest, int mask_edges) {
  std::vector<double> frac_pos(ndest + 2);
  std::vector<int> k_src(ndest + 2);
  int ks = 0;
  double dh = 0., x_dest = 0.;
  for (int k_dest = 1; k_dest <= ndest + 1; ++k_dest) {
    while (dh <= x_dest && ks < nsrc) {
      x_dest = x_dest - dh;
      ks = ks + 1;
      dh = h_src[ks - 1];
    }
    k_src[k_dest] = ks;
    if (dh > 0.) frac_pos[k_dest] = fmax2(0., fmin2(1., x_dest / dh));
    else frac_pos[k_dest] = 0.5;
    if (k_dest <= ndest) x_dest = x_dest + h_dest[k_dest - 1];
  }
  for (int k_dest = 1; k_dest <= ndest + 1; ++k_dest) {
    ks = k_src[k_dest];
    u_dest[k_dest - 1] = (1.0 - frac_pos[k_dest]) * u_src[ks - 1] + frac_pos[k_dest] * u_src[ks];
  }
  if (mask_edges) {
    for (int k_dest = 1; k_dest <= ndest; ++k_dest) { if (h_dest[k_dest - 1] > 0.) break; u_dest[k_dest - 1] = 0.0; }
    for (int k_dest = ndest; k_dest >= 1; --k_dest) { if (h_dest[k_dest - 1] > 0.) break; u_dest[k_dest] = 0.0; }
  }
}

// ALE_remap_interface_vals :1303-1339
extern "C" int oracle_ale_remap_interface_vals(const mom6cu_domain* d, const mom6cu_grid* Gp, const double* h_oldp, const double* h_newp,
                                               double* int_valp) {
  const OGrid G(d, Gp);
  const int nz = G.ke;
  const V3 h_old = G.H3(h_oldp), h_new = G.H3(h_newp), int_val = G.H3(int_valp, nz + 1);
  std::vector<double> val_src(nz + 1), val_tgt(nz + 1), h_src(nz), h_tgt(nz);
  for (int j = G.jsc; j <= G.jec; ++j) for (int i = G.isc; i <= G.iec; ++i) if (G.mask2dT(i, j) > 0.) {
    for (int k = 1; k <= nz; ++k) { h_src[k - 1] = h_old(i, j, k); h_tgt[k - 1] = h_new(i, j, k); }
    for (int K = 1; K <= nz + 1; ++K) val_src[K - 1] = int_val(i, j, K);
    oracle_interpolate_column(nz, h_src.data(), val_src.data(), nz, h_tgt.data(), val_tgt.data(), 0);
    for (int K = 1; K <= nz + 1; ++K) int_val(i, j, K) = val_tgt[K - 1];
  }
  return 0;
}

// ALE_remap_vertex_vals :1342-1382
extern "C" int oracle_ale_remap_vertex_vals(const mom6cu_domain* d, const mom6cu_grid* Gp, const double* h_oldp, const double* h_newp,
                                            double* vert_valp) {
  const OGrid G(d, Gp);
  const int nz = G.ke;
  const V3 h_old = G.H3(h_oldp), h_new = G.H3(h_newp), vert_val = G.Q3(vert_valp, nz + 1);
  std::vector<double> val_src(nz + 1), val_tgt(nz + 1), h_src(nz), h_tgt(nz);
  for (int J = G.JscB; J <= G.JecB; ++J) for (int I = G.IscB; I <= G.IecB; ++I) {
    const int i = I, j = J;
    if ((G.mask2dT(i, j) + G.mask2dT(i + 1, j + 1)) + (G.mask2dT(i + 1, j) + G.mask2dT(i, j + 1)) > 0.0) {
      const double I_mask_sum = 1.0 / ((G.mask2dT(i, j) + G.mask2dT(i + 1, j + 1)) + (G.mask2dT(i + 1, j) + G.mask2dT(i, j + 1)));
      for (int k = 1; k <= nz; ++k) {
        h_src[k - 1] = ((G.mask2dT(i, j) * h_old(i, j, k) + G.mask2dT(i + 1, j + 1) * h_old(i + 1, j + 1, k)) +
                        (G.mask2dT(i + 1, j) * h_old(i + 1, j, k) + G.mask2dT(i, j + 1) * h_old(i, j + 1, k))) * I_mask_sum;
        h_tgt[k - 1] = ((G.mask2dT(i, j) * h_new(i, j, k) + G.mask2dT(i + 1, j + 1) * h_new(i + 1, j + 1, k)) +
                        (G.mask2dT(i + 1, j) * h_new(i + 1, j, k) + G.mask2dT(i, j + 1) * h_new(i, j + 1, k))) * I_mask_sum;
      }
      for (int K = 1; K <= nz + 1; ++K) val_src[K - 1] = vert_val(I, J, K);
      oracle_interpolate_column(nz, h_src.data(), val_src.data(), nz, h_tgt.data(), val_tgt.data(), 0);
      for (int K = 1; K <= nz + 1; ++K) vert_val(I, J, K) = val_tgt[K - 1];
    }
  }
  return 0;
}

// ALE_regridding_and_remapping, MOM.F90:1751-1926 on one tile
extern "C" int oracle_ale_regridding_and_remapping(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV,
                                                   const mom6cu_unit_scale* US, mom6cu_ale_cs* CS, const mom6cu_dyn_split_rk2_cs* dynCS,
                                                   const mom6cu_ale_args* a, int nthreads) {
  if (CS->remap_uv_using_old_alg || CS->do_conv_adj || CS->use_hybgen_unmix) return 3;
  // remap.cpp restates the answer_date >= 20190101 expressions only: refuse the older ones, as the device does (csrc/remap.cu:120),
  // rather than answer with the newer arithmetic (the reference run of tests/refcases.py showed that they differ)
  if (CS->remapCS.answer_date < 20190101 || CS->vel_remapCS.answer_date < 20190101) return 3;
  const OGrid G(d, Gp);
  const int nz = G.ke;
  const size_t plH = (size_t)(d->ied - d->isd + 1) * (d->jed - d->jsd + 1);
  auto fillH = [&](double* f, int nk) { for (int k = 0; k < nk; ++k) oracle_fill_halo_2d(d, f + plH * k, 0, 0); };
  // pass_T_S_h :1800-1806
  if (a->iT >= 0) fillH(a->tr[a->iT], nz);
  if (a->iS >= 0) fillH(a->tr[a->iS], nz);
  fillH(a->h, nz);
  // ALE_update_regrid_weights :1825
  double w = 0.0;
  if (CS->regrid_time_scale > 0.0) w = CS->regrid_time_scale / (CS->regrid_time_scale + a->dtdia);
  CS->regridCS.old_grid_weight = w;
  A3 h_new(G.isd, G.ied, G.jsd, G.jed, nz), dzRegrid(G.isd, G.ied, G.jsd, G.jed, nz + 1);
  A3 h_old_u(G.isd - 1, G.ied, G.jsd, G.jed, nz), h_new_u(G.isd - 1, G.ied, G.jsd, G.jed, nz);
  A3 h_old_v(G.isd, G.ied, G.jsd - 1, G.jed, nz), h_new_v(G.isd, G.ied, G.jsd - 1, G.jed, nz);
  int rc = oracle_ale_regrid(d, Gp, GV, US, &CS->regridCS, a->h, h_new.p, dzRegrid.p);  // :1834
  if (rc) return rc;
  for (int m = 0; m < a->ntr; ++m)  // ALE_remap_tracers :1839
    if ((rc = oracle_ale_remap_scalar(d, Gp, &CS->remapCS, a->h, h_new.p, a->tr[m], a->conc_underflow ? a->conc_underflow[m] : 0.0, nthreads))) return rc;
  oracle_ale_remap_set_h_vel(d, Gp, a->h, h_old_u.p, h_old_v.p);    // :1842
  oracle_ale_remap_set_h_vel(d, Gp, h_new.p, h_new_u.p, h_new_v.p);  // :1846
  if ((rc = oracle_ale_remap_velocities(d, Gp, &CS->vel_remapCS, h_old_u.p, h_old_v.p, h_new_u.p, h_new_v.p, a->u, a->v, nthreads))) return rc;  // :1850
  if (CS->remap_aux_vars) {  // :1855-1871
    if ((rc = oracle_remap_dyn_split_rk2_aux_vars(d, Gp, &CS->vel_remapCS, dynCS, h_old_u.p, h_old_v.p, h_new_u.p, h_new_v.p, nthreads))) return rc;
    if (a->Kd_shear) oracle_ale_remap_interface_vals(d, Gp, a->h, h_new.p, a->Kd_shear);
    if (a->Kv_shear) oracle_ale_remap_interface_vals(d, Gp, a->h, h_new.p, a->Kv_shear);
    if (a->Kv_shear_Bu) oracle_ale_remap_vertex_vals(d, Gp, a->h, h_new.p, a->Kv_shear_Bu);
    if (a->Kv_shear) fillH(a->Kv_shear, nz + 1);
  }
  const V3 h = G.H3(a->h);
  for (int k = 1; k <= nz; ++k) for (int j = G.jsc - 1; j <= G.jec + 1; ++j) for (int i = G.isc - 1; i <= G.iec + 1; ++i)
    h(i, j, k) = h_new(i, j, k);  // :1875-1878
  return 0;
}
