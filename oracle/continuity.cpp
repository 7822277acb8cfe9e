// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of continuity_PPM, /root/reference/src/core/MOM_continuity_PPM.F90:
//   continuity_PPM :86-194, continuity_*_convergence :348-421, *_edge_thickness :425-515,
//   zonal_mass_flux :519-819 (meridional :1412-1709), zonal_flux_layer :896-971 (:1787-1869),
//   zonal_flux_thickness :975-1089 (:1873-1989), zonal_flux_adjust :1093-1242 (:1992-2139),
//   set_zonal_BT_cont :1246-1409 (:2143-2304), PPM_reconstruction_x/y :2307-2572,
//   PPM_limit_pos :2578, PPM_limit_CW84 :2620, ratio_max :2660, set_continuity_loop_bounds :2766.
// The meridional routines of the reference are the index-transposed mirror of the zonal ones
// (checked line by line); both are instantiated here from one template with the row structure
// (row arrays, row-level `domore`) of the Fortran kept intact.
// Frozen options: no OBCs (OBC pointer not associated).
#include "oracle.h"
#include "farray.hpp"
#include <algorithm>
#include <vector>
#include <cmath>
#include <omp.h>

using namespace orc;

namespace {

struct LB { int ish, ieh, jsh, jeh; };

struct Env {
  const mom6cu_domain* d;
  const mom6cu_continuity_cs* CS;
  double Angstrom_H, H_subroundoff;
  int nz;
  // grid (G-sized)
  V2 mask2dT, mask2dCu, mask2dCv, dxT, dyT, IdxT, IdyT, areaT, IareaT, dy_Cu, dx_Cv, dxCu, dyCv;
};

// Z = true: zonal (face (I=n, j=o), cells (n,o),(n+1,o)); Z = false: meridional (face (i=n, J=o),
// cells (n,o),(n,o+1)).  c(A,n,o,m) is the cell / face value displaced by m along the flow direction.
template <bool Z> inline double& c2(const V2& A, int n, int o, int m) { return Z ? A(n + m, o) : A(n, o + m); }
template <bool Z> inline double& c3(const V3& A, int n, int o, int m, int k) { return Z ? A(n + m, o, k) : A(n, o + m, k); }

// PPM_reconstruction_x :2307-2439 / _y :2442-2572 for one layer; hL = h_W|h_S, hR = h_E|h_N
template <bool Z>
void PPM_reconstruction(const Env& E, const V3& h_in, const V3& hL, const V3& hR, int k, const LB& lb, double h_min) {
  const bool monotonic = E.CS->monotonic, simple_2nd = E.CS->simple_2nd;
  int isl, iel, jsl, jel;
  if (Z) { isl = lb.ish - 1; iel = lb.ieh + 1; jsl = lb.jsh; jel = lb.jeh; }
  else { isl = lb.ish; iel = lb.ieh; jsl = lb.jsh - 1; jel = lb.jeh + 1; }
  const V2& mask = E.mask2dT;
  const double oneSixth = 1. / 6.;
  A2 slp(h_in.ilo, h_in.ilo + h_in.ni - 1, h_in.jlo, h_in.jlo + h_in.nj - 1);
  if (simple_2nd) {
    for (int j = jsl; j <= jel; ++j) for (int i = isl; i <= iel; ++i) {
      const double h_im1 = c2<Z>(mask, i, j, -1) * c3<Z>(h_in, i, j, -1, k) + (1.0 - c2<Z>(mask, i, j, -1)) * h_in(i, j, k);
      const double h_ip1 = c2<Z>(mask, i, j, +1) * c3<Z>(h_in, i, j, +1, k) + (1.0 - c2<Z>(mask, i, j, +1)) * h_in(i, j, k);
      hL(i, j, k) = 0.5 * (h_im1 + h_in(i, j, k));
      hR(i, j, k) = 0.5 * (h_ip1 + h_in(i, j, k));
    }
  } else {
    const int ja = Z ? jsl : jsl - 1, jb = Z ? jel : jel + 1, ia = Z ? isl - 1 : isl, ib = Z ? iel + 1 : iel;
    for (int j = ja; j <= jb; ++j) for (int i = ia; i <= ib; ++i) {
      if ((c2<Z>(mask, i, j, -1) * mask(i, j) * c2<Z>(mask, i, j, +1)) == 0.0) {
        slp(i, j) = 0.0;
      } else {
        const double hp = c3<Z>(h_in, i, j, +1, k), hm = c3<Z>(h_in, i, j, -1, k), hc = h_in(i, j, k);
        // This uses a simple 2nd order slope.
        double s = 0.5 * (hp - hm);
        // Monotonic constraint, see Eq. B2 in Lin 1994, MWR (132)
        const double dMx = fmax2(fmax2(hp, hm), hc) - hc;
        const double dMn = hc - fmin2(fmin2(hp, hm), hc);
        s = fsign(1., s) * fmin2(std::fabs(s), 2. * fmin2(dMx, dMn));
        slp(i, j) = s;
      }
    }
    for (int j = jsl; j <= jel; ++j) for (int i = isl; i <= iel; ++i) {
      const double h_im1 = c2<Z>(mask, i, j, -1) * c3<Z>(h_in, i, j, -1, k) + (1.0 - c2<Z>(mask, i, j, -1)) * h_in(i, j, k);
      const double h_ip1 = c2<Z>(mask, i, j, +1) * c3<Z>(h_in, i, j, +1, k) + (1.0 - c2<Z>(mask, i, j, +1)) * h_in(i, j, k);
      // Left/right values following Eq. B2 in Lin 1994, MWR (132)
      hL(i, j, k) = 0.5 * (h_im1 + h_in(i, j, k)) + oneSixth * (c2<Z>(slp, i, j, -1) - slp(i, j));
      hR(i, j, k) = 0.5 * (h_ip1 + h_in(i, j, k)) + oneSixth * (slp(i, j) - c2<Z>(slp, i, j, +1));
    }
  }
  if (monotonic) {
    // PPM_limit_CW84 :2620-2657
    for (int j = jsl; j <= jel; ++j) for (int i = isl; i <= iel; ++i) {
      const double h_i = h_in(i, j, k);
      if ((hR(i, j, k) - h_i) * (h_i - hL(i, j, k)) <= 0.) {
        hL(i, j, k) = h_i; hR(i, j, k) = h_i;
      } else {
        const double RLdiff = hR(i, j, k) - hL(i, j, k);
        const double RLmean = 0.5 * (hR(i, j, k) + hL(i, j, k));
        const double FunFac = 6. * RLdiff * (h_i - RLmean);
        const double RLdiff2 = RLdiff * RLdiff;
        if (FunFac > RLdiff2) hL(i, j, k) = 3. * h_i - 2. * hR(i, j, k);
        if (FunFac < -RLdiff2) hR(i, j, k) = 3. * h_i - 2. * hL(i, j, k);
      }
    }
  } else {
    // PPM_limit_pos :2578-2616
    for (int j = jsl; j <= jel; ++j) for (int i = isl; i <= iel; ++i) {
      const double curv = 3.0 * ((hL(i, j, k) + hR(i, j, k)) - 2.0 * h_in(i, j, k));
      if (curv > 0.0) {  // Only minima are limited.
        const double dh = hR(i, j, k) - hL(i, j, k);
        if (std::fabs(dh) < curv) {  // The parabola's minimum is within the cell.
          if (h_in(i, j, k) <= h_min) {
            hL(i, j, k) = h_in(i, j, k); hR(i, j, k) = h_in(i, j, k);
          } else if (12.0 * curv * (h_in(i, j, k) - h_min) < (curv * curv + 3.0 * (dh * dh))) {
            const double scale = 12.0 * curv * (h_in(i, j, k) - h_min) / (curv * curv + 3.0 * (dh * dh));
            hL(i, j, k) = h_in(i, j, k) + scale * (hL(i, j, k) - h_in(i, j, k));
            hR(i, j, k) = h_in(i, j, k) + scale * (hR(i, j, k) - h_in(i, j, k));
          }
        }
      }
    }
  }
}

// zonal_edge_thickness :425-468 / meridional_edge_thickness :472-515
template <bool Z>
void edge_thickness(const Env& E, const V3& h_in, const V3& hL, const V3& hR, const LB& lb) {
  const int nz = E.nz;
  if (E.CS->upwind_1st) {
    const int ia = Z ? lb.ish - 1 : lb.ish, ib = Z ? lb.ieh + 1 : lb.ieh;
    const int ja = Z ? lb.jsh : lb.jsh - 1, jb = Z ? lb.jeh : lb.jeh + 1;
    for (int k = 1; k <= nz; ++k) for (int j = ja; j <= jb; ++j) for (int i = ia; i <= ib; ++i) {
      hL(i, j, k) = h_in(i, j, k); hR(i, j, k) = h_in(i, j, k);
    }
  } else {
    _Pragma("omp parallel for")
    for (int k = 1; k <= nz; ++k) PPM_reconstruction<Z>(E, h_in, hL, hR, k, lb, 2.0 * E.Angstrom_H);
  }
}

inline double ratio_max(double a, double b, double maxrat) {  // :2660-2671
  if (std::fabs(a) > std::fabs(maxrat * b)) return maxrat;
  return a / b;
}

// per-row helper holding what the Fortran passes around: metrics of face (n,o)
template <bool Z>
struct Met {
  const Env& E;
  explicit Met(const Env& e) : E(e) {}
  inline double dy(int n, int o) const { return Z ? E.dy_Cu(n, o) : E.dx_Cv(n, o); }           // G%dy_Cu(I,j) | G%dx_Cv(i,J)
  inline double IdxT(int n, int o, int m) const { return Z ? E.IdxT(n + m, o) : E.IdyT(n, o + m); }
  inline double dxT(int n, int o, int m) const { return Z ? E.dxT(n + m, o) : E.dyT(n, o + m); }
  inline double areaT(int n, int o, int m) const { return Z ? E.areaT(n + m, o) : E.areaT(n, o + m); }
  inline double IareaT(int n, int o, int m) const { return Z ? E.IareaT(n + m, o) : E.IareaT(n, o + m); }
  inline double dxC(int n, int o) const { return Z ? E.dxCu(n, o) : E.dyCv(n, o); }            // G%dxCu(I,j) | G%dyCv(i,J)
  inline double maskC(int n, int o) const { return Z ? E.mask2dCu(n, o) : E.mask2dCv(n, o); }
};

// zonal_flux_layer :896-971 / merid_flux_layer :1787-1869 for the row o, inner index n = nlo..nhi
template <bool Z>
void flux_layer(const Env& E, const double* u /*by n*/, const V3& h, const V3& hL, const V3& hR, int k, double* uh,
                double* duhdu, const double* visc_rem, double dt, int o, int nlo, int nhi, const char* do_I,
                const V3* por, int n0) {
  Met<Z> M(E);
  const bool vol_CFL = E.CS->vol_CFL;
  for (int n = nlo; n <= nhi; ++n) if (do_I[n - n0]) {
    const double un = u[n - n0];
    const double face = por ? (M.dy(n, o) * (*por)(n, o, k)) : M.dy(n, o);
    double CFL, curv_3, h_marg;
    if (un > 0.0) {
      if (vol_CFL) CFL = (un * dt) * (M.dy(n, o) * M.IareaT(n, o, 0));
      else CFL = un * dt * M.IdxT(n, o, 0);
      curv_3 = (hL(n, o, k) + hR(n, o, k)) - 2.0 * h(n, o, k);
      uh[n - n0] = face * un * (hR(n, o, k) + CFL * (0.5 * (hL(n, o, k) - hR(n, o, k)) + curv_3 * (CFL - 1.5)));
      h_marg = hR(n, o, k) + CFL * ((hL(n, o, k) - hR(n, o, k)) + 3.0 * curv_3 * (CFL - 1.0));
    } else if (un < 0.0) {
      if (vol_CFL) CFL = (-un * dt) * (M.dy(n, o) * M.IareaT(n, o, 1));
      else CFL = -un * dt * M.IdxT(n, o, 1);
      const double hL1 = c3<Z>(hL, n, o, 1, k), hR1 = c3<Z>(hR, n, o, 1, k), h1 = c3<Z>(h, n, o, 1, k);
      curv_3 = (hL1 + hR1) - 2.0 * h1;
      uh[n - n0] = face * un * (hL1 + CFL * (0.5 * (hR1 - hL1) + curv_3 * (CFL - 1.5)));
      h_marg = hL1 + CFL * ((hR1 - hL1) + 3.0 * curv_3 * (CFL - 1.0));
    } else {
      uh[n - n0] = 0.0;
      h_marg = 0.5 * (c3<Z>(hL, n, o, 1, k) + hR(n, o, k));
    }
    duhdu[n - n0] = face * h_marg * visc_rem[n - n0];
  }
}

// zonal_flux_adjust :1093-1242 / meridional_flux_adjust :1992-2139
template <bool Z>
void flux_adjust(const Env& E, const V3& u, const V3& h_in, const V3& hL, const V3& hR, const double* uhbt,
                 const double* uh_tot_0, const double* duhdu_tot_0, double* du, const double* du_max_CFL,
                 const double* du_min_CFL, double dt, const std::vector<std::vector<double>>& visc_rem /*[k][n]*/,
                 int o, int nlo, int nhi, const char* do_I_in, const V3* por, const V3* uh_3d, int n0, int nw) {
  Met<Z> M(E);
  const int nz = E.nz, max_itts = 20;
  std::vector<std::vector<double>> uh_aux(nz + 1, std::vector<double>(nw, 0.0)), duhdu(nz + 1, std::vector<double>(nw, 0.0));
  std::vector<double> uh_err(nw), uh_err_best(nw), u_new(nw), duhdu_tot(nw), du_min(nw), du_max(nw);
  std::vector<char> do_I(nw);
  const mom6cu_continuity_cs* CS = E.CS;
  if (uh_3d) for (int k = 1; k <= nz; ++k) for (int n = nlo; n <= nhi; ++n) uh_aux[k][n - n0] = (*uh_3d)(n, o, k);
  for (int n = nlo; n <= nhi; ++n) {
    const int x = n - n0;
    du[x] = 0.0; do_I[x] = do_I_in[x];
    du_max[x] = du_max_CFL[x]; du_min[x] = du_min_CFL[x];
    uh_err[x] = uh_tot_0[x] - uhbt[x]; duhdu_tot[x] = duhdu_tot_0[x];
    uh_err_best[x] = std::fabs(uh_err[x]);
  }
  for (int itt = 1; itt <= max_itts; ++itt) {
    double tol_eta;
    if (itt <= 1) tol_eta = 1e-6 * CS->tol_eta;
    else if (itt == 2) tol_eta = 1e-4 * CS->tol_eta;
    else if (itt == 3) tol_eta = 1e-2 * CS->tol_eta;
    else tol_eta = CS->tol_eta;
    const double tol_vel = CS->tol_vel;
    for (int n = nlo; n <= nhi; ++n) {
      const int x = n - n0;
      if (uh_err[x] > 0.0) du_max[x] = du[x];
      else if (uh_err[x] < 0.0) du_min[x] = du[x];
      else do_I[x] = 0;
    }
    bool domore = false;
    for (int n = nlo; n <= nhi; ++n) {
      const int x = n - n0;
      if (!do_I[x]) continue;
      if ((dt * fmin2(M.IareaT(n, o, 0), M.IareaT(n, o, 1)) * std::fabs(uh_err[x]) > tol_eta) ||
          (CS->better_iter && ((std::fabs(uh_err[x]) > tol_vel * duhdu_tot[x]) || (std::fabs(uh_err[x]) > uh_err_best[x])))) {
        // Use Newton's method, provided it stays bounded.  Otherwise bisect the value with the appropriate bound.
        const double ddu = -uh_err[x] / duhdu_tot[x];
        const double du_prev = du[x];
        du[x] = du[x] + ddu;
        if (std::fabs(ddu) < 1.0e-15 * std::fabs(du[x])) {
          do_I[x] = 0;  // ddu is small enough to quit.
        } else if (ddu > 0.0) {
          if (du[x] >= du_max[x]) {
            du[x] = 0.5 * (du_prev + du_max[x]);
            if (du_max[x] - du_prev < 1.0e-15 * std::fabs(du[x])) do_I[x] = 0;
          }
        } else {  // ddu < 0.0
          if (du[x] <= du_min[x]) {
            du[x] = 0.5 * (du_prev + du_min[x]);
            if (du_prev - du_min[x] < 1.0e-15 * std::fabs(du[x])) do_I[x] = 0;
          }
        }
        if (do_I[x]) domore = true;
      } else {
        do_I[x] = 0;
      }
    }
    if (!domore) break;
    if ((itt < max_itts) || uh_3d) {
      for (int k = 1; k <= nz; ++k) {
        for (int n = nlo; n <= nhi; ++n) u_new[n - n0] = u(n, o, k) + du[n - n0] * visc_rem[k][n - n0];
        flux_layer<Z>(E, u_new.data(), h_in, hL, hR, k, uh_aux[k].data(), duhdu[k].data(), visc_rem[k].data(), dt, o,
                      nlo, nhi, do_I.data(), por, n0);
      }
    }
    if (itt < max_itts) {
      for (int n = nlo; n <= nhi; ++n) { uh_err[n - n0] = -uhbt[n - n0]; duhdu_tot[n - n0] = 0.0; }
      for (int k = 1; k <= nz; ++k) for (int n = nlo; n <= nhi; ++n) {
        uh_err[n - n0] = uh_err[n - n0] + uh_aux[k][n - n0];
        duhdu_tot[n - n0] = duhdu_tot[n - n0] + duhdu[k][n - n0];
      }
      for (int n = nlo; n <= nhi; ++n) uh_err_best[n - n0] = fmin2(uh_err_best[n - n0], std::fabs(uh_err[n - n0]));
    }
  }
  if (uh_3d) for (int k = 1; k <= nz; ++k) for (int n = nlo; n <= nhi; ++n) (*uh_3d)(n, o, k) = uh_aux[k][n - n0];
}

// set_zonal_BT_cont :1246-1409 / set_merid_BT_cont :2143-2304; the six outputs are
// (W0|S0, WW|SS, E0|N0, EE|NN, uBT_WW|vBT_SS, uBT_EE|vBT_NN)
template <bool Z>
void set_BT_cont(const Env& E, const V3& u, const V3& h_in, const V3& hL, const V3& hR, const V2& FA_W0, const V2& FA_WW,
                 const V2& FA_E0, const V2& FA_EE, const V2& uBT_WW, const V2& uBT_EE, const double* uh_tot_0,
                 const double* duhdu_tot_0, const double* du_max_CFL, const double* du_min_CFL, double dt,
                 const std::vector<std::vector<double>>& visc_rem, const double* visc_rem_max, int o, int nlo, int nhi,
                 const char* do_I, const V3* por, int n0, int nw) {
  Met<Z> M(E);
  const int nz = E.nz;
  const double Idt = 1.0 / dt;
  const double min_visc_rem = 0.1, CFL_min = 1e-6;
  std::vector<double> du0(nw), duL(nw), duR(nw), zeros(nw, 0.0), du_CFL(nw), u_L(nw), u_R(nw), u_0(nw), duhdu_L(nw),
      duhdu_R(nw), duhdu_0(nw), uh_L(nw), uh_R(nw), uh_0(nw), FAmt_L(nw), FAmt_R(nw), FAmt_0(nw), uhtot_L(nw), uhtot_R(nw);
  // Diagnose the zero-transport correction, du0.
  flux_adjust<Z>(E, u, h_in, hL, hR, zeros.data(), uh_tot_0, duhdu_tot_0, du0.data(), du_max_CFL, du_min_CFL, dt, visc_rem,
                 o, nlo, nhi, do_I, por, nullptr, n0, nw);
  bool domore = false;
  for (int n = nlo; n <= nhi; ++n) {
    const int x = n - n0;
    if (do_I[x]) domore = true;
    du_CFL[x] = (CFL_min * Idt) * M.dxC(n, o);
    duR[x] = fmin2(0.0, du0[x] - du_CFL[x]);
    duL[x] = fmax2(0.0, du0[x] + du_CFL[x]);
    FAmt_L[x] = 0.0; FAmt_R[x] = 0.0; FAmt_0[x] = 0.0;
    uhtot_L[x] = 0.0; uhtot_R[x] = 0.0;
  }
  if (!domore) {
    for (int n = nlo; n <= nhi; ++n) {
      FA_W0(n, o) = 0.0; FA_WW(n, o) = 0.0; FA_E0(n, o) = 0.0; FA_EE(n, o) = 0.0; uBT_WW(n, o) = 0.0; uBT_EE(n, o) = 0.0;
    }
    return;
  }
  for (int k = 1; k <= nz; ++k) for (int n = nlo; n <= nhi; ++n) {
    const int x = n - n0;
    if (!do_I[x]) continue;
    const double visc_rem_lim = fmax2(visc_rem[k][x], min_visc_rem * visc_rem_max[x]);
    if (visc_rem_lim > 0.0) {  // This is almost always true for ocean points.
      if (u(n, o, k) + duR[x] * visc_rem_lim > -du_CFL[x] * visc_rem[k][x])
        duR[x] = -(u(n, o, k) + du_CFL[x] * visc_rem[k][x]) / visc_rem_lim;
      if (u(n, o, k) + duL[x] * visc_rem_lim < du_CFL[x] * visc_rem[k][x])
        duL[x] = -(u(n, o, k) - du_CFL[x] * visc_rem[k][x]) / visc_rem_lim;
    }
  }
  for (int k = 1; k <= nz; ++k) {
    for (int n = nlo; n <= nhi; ++n) {
      const int x = n - n0;
      if (!do_I[x]) continue;
      u_L[x] = u(n, o, k) + duL[x] * visc_rem[k][x];
      u_R[x] = u(n, o, k) + duR[x] * visc_rem[k][x];
      u_0[x] = u(n, o, k) + du0[x] * visc_rem[k][x];
    }
    flux_layer<Z>(E, u_0.data(), h_in, hL, hR, k, uh_0.data(), duhdu_0.data(), visc_rem[k].data(), dt, o, nlo, nhi, do_I, por, n0);
    flux_layer<Z>(E, u_L.data(), h_in, hL, hR, k, uh_L.data(), duhdu_L.data(), visc_rem[k].data(), dt, o, nlo, nhi, do_I, por, n0);
    flux_layer<Z>(E, u_R.data(), h_in, hL, hR, k, uh_R.data(), duhdu_R.data(), visc_rem[k].data(), dt, o, nlo, nhi, do_I, por, n0);
    for (int n = nlo; n <= nhi; ++n) {
      const int x = n - n0;
      if (!do_I[x]) continue;
      FAmt_0[x] = FAmt_0[x] + duhdu_0[x];
      FAmt_L[x] = FAmt_L[x] + duhdu_L[x];
      FAmt_R[x] = FAmt_R[x] + duhdu_R[x];
      uhtot_L[x] = uhtot_L[x] + uh_L[x];
      uhtot_R[x] = uhtot_R[x] + uh_R[x];
    }
  }
  for (int n = nlo; n <= nhi; ++n) {
    const int x = n - n0;
    if (do_I[x]) {
      double FA_0 = FAmt_0[x], FA_avg = FAmt_0[x];
      if ((duL[x] - du0[x]) != 0.0) FA_avg = uhtot_L[x] / (duL[x] - du0[x]);
      if (FA_avg > fmax2(FA_0, FAmt_L[x])) FA_avg = fmax2(FA_0, FAmt_L[x]);
      else if (FA_avg < fmin2(FA_0, FAmt_L[x])) FA_0 = FA_avg;
      FA_W0(n, o) = FA_0; FA_WW(n, o) = FAmt_L[x];
      if (std::fabs(FA_0 - FAmt_L[x]) <= 1e-12 * FA_0) uBT_WW(n, o) = 0.0;
      else uBT_WW(n, o) = (1.5 * (duL[x] - du0[x])) * ((FAmt_L[x] - FA_avg) / (FAmt_L[x] - FA_0));

      FA_0 = FAmt_0[x]; FA_avg = FAmt_0[x];
      if ((duR[x] - du0[x]) != 0.0) FA_avg = uhtot_R[x] / (duR[x] - du0[x]);
      if (FA_avg > fmax2(FA_0, FAmt_R[x])) FA_avg = fmax2(FA_0, FAmt_R[x]);
      else if (FA_avg < fmin2(FA_0, FAmt_R[x])) FA_0 = FA_avg;
      FA_E0(n, o) = FA_0; FA_EE(n, o) = FAmt_R[x];
      if (std::fabs(FAmt_R[x] - FA_0) <= 1e-12 * FA_0) uBT_EE(n, o) = 0.0;
      else uBT_EE(n, o) = (1.5 * (duR[x] - du0[x])) * ((FAmt_R[x] - FA_avg) / (FAmt_R[x] - FA_0));
    } else {
      FA_W0(n, o) = 0.0; FA_WW(n, o) = 0.0; FA_E0(n, o) = 0.0; FA_EE(n, o) = 0.0; uBT_WW(n, o) = 0.0; uBT_EE(n, o) = 0.0;
    }
  }
}

// zonal_flux_thickness :975-1089 / meridional_flux_thickness :1873-1989
template <bool Z>
void flux_thickness(const Env& E, const V3& u, const V3& h, const V3& hL, const V3& hR, const V3& h_u, double dt,
                    const LB& lb, const V3* por, const V3* visc_rem_u) {
  Met<Z> M(E);
  const int nz = E.nz;
  const bool vol_CFL = E.CS->vol_CFL, marginal = E.CS->marginal_faces;
  const int olo = Z ? lb.jsh : lb.jsh - 1, ohi = lb.jeh, nlo = Z ? lb.ish - 1 : lb.ish, nhi = lb.ieh;
  _Pragma("omp parallel for")
  for (int k = 1; k <= nz; ++k) for (int o = olo; o <= ohi; ++o) for (int n = nlo; n <= nhi; ++n) {
    double CFL, curv_3, h_avg, h_marg;
    const double un = u(n, o, k);
    if (un > 0.0) {
      if (vol_CFL) CFL = (un * dt) * (M.dy(n, o) * M.IareaT(n, o, 0));
      else CFL = un * dt * M.IdxT(n, o, 0);
      curv_3 = (hL(n, o, k) + hR(n, o, k)) - 2.0 * h(n, o, k);
      h_avg = hR(n, o, k) + CFL * (0.5 * (hL(n, o, k) - hR(n, o, k)) + curv_3 * (CFL - 1.5));
      h_marg = hR(n, o, k) + CFL * ((hL(n, o, k) - hR(n, o, k)) + 3.0 * curv_3 * (CFL - 1.0));
    } else if (un < 0.0) {
      if (vol_CFL) CFL = (-un * dt) * (M.dy(n, o) * M.IareaT(n, o, 1));
      else CFL = -un * dt * M.IdxT(n, o, 1);
      const double hL1 = c3<Z>(hL, n, o, 1, k), hR1 = c3<Z>(hR, n, o, 1, k), h1 = c3<Z>(h, n, o, 1, k);
      curv_3 = (hL1 + hR1) - 2.0 * h1;
      h_avg = hL1 + CFL * (0.5 * (hR1 - hL1) + curv_3 * (CFL - 1.5));
      h_marg = hL1 + CFL * ((hR1 - hL1) + 3.0 * curv_3 * (CFL - 1.0));
    } else {
      h_avg = 0.5 * (c3<Z>(hL, n, o, 1, k) + hR(n, o, k));
      h_marg = 0.5 * (c3<Z>(hL, n, o, 1, k) + hR(n, o, k));
    }
    if (marginal) h_u(n, o, k) = h_marg;
    else h_u(n, o, k) = h_avg;
  }
  if (visc_rem_u) {
    for (int k = 1; k <= nz; ++k) for (int o = olo; o <= ohi; ++o) for (int n = nlo; n <= nhi; ++n)
      h_u(n, o, k) = h_u(n, o, k) * (por ? ((*visc_rem_u)(n, o, k) * (*por)(n, o, k)) : (*visc_rem_u)(n, o, k));
  } else if (por) {
    for (int k = 1; k <= nz; ++k) for (int o = olo; o <= ohi; ++o) for (int n = nlo; n <= nhi; ++n)
      h_u(n, o, k) = h_u(n, o, k) * (*por)(n, o, k);
  }
}

// zonal_mass_flux :519-819 / meridional_mass_flux :1412-1709
template <bool Z>
void mass_flux_impl(const Env& E, const V3& u, const V3& h_in, const V3& hL, const V3& hR, const V3& uh, double dt,
                    const V3* por, const LB& lb, const V2* uhbt, const V3* visc_rem_u, const V3* u_cor,
                    const mom6cu_bt_cont* BT_cont, const V2* du_cor, const V3* h_u_out) {
  Met<Z> M(E);
  const mom6cu_continuity_cs* CS = E.CS;
  const mom6cu_domain* d = E.d;
  const int nz = E.nz;
  const bool use_visc_rem = visc_rem_u != nullptr;
  const bool set_BT = BT_cont != nullptr;
  const int ish = lb.ish, ieh = lb.ieh, jsh = lb.jsh, jeh = lb.jeh;
  const int olo = Z ? jsh : jsh - 1, ohi = jeh, nlo = Z ? ish - 1 : ish, nhi = ieh;
  const int n0 = nlo, nw = nhi - nlo + 1;
  if (du_cor) du_cor->fill(0.0);
  double CFL_dt = CS->CFL_limit_adjust / dt;
  const double I_dt = 1.0 / dt;
  if (CS->aggress_adjust) CFL_dt = I_dt;
  // BT_cont views
  V2 FA_W0, FA_WW, FA_E0, FA_EE, uBT_WW, uBT_EE;
  if (set_BT) {
    const int il = Z ? d->isd - 1 : d->isd, jl = Z ? d->jsd : d->jsd - 1;
    if (Z) {
      FA_W0 = V2(BT_cont->FA_u_W0, il, d->ied, jl, d->jed); FA_WW = V2(BT_cont->FA_u_WW, il, d->ied, jl, d->jed);
      FA_E0 = V2(BT_cont->FA_u_E0, il, d->ied, jl, d->jed); FA_EE = V2(BT_cont->FA_u_EE, il, d->ied, jl, d->jed);
      uBT_WW = V2(BT_cont->uBT_WW, il, d->ied, jl, d->jed); uBT_EE = V2(BT_cont->uBT_EE, il, d->ied, jl, d->jed);
    } else {
      FA_W0 = V2(BT_cont->FA_v_S0, il, d->ied, jl, d->jed); FA_WW = V2(BT_cont->FA_v_SS, il, d->ied, jl, d->jed);
      FA_E0 = V2(BT_cont->FA_v_N0, il, d->ied, jl, d->jed); FA_EE = V2(BT_cont->FA_v_NN, il, d->ied, jl, d->jed);
      uBT_WW = V2(BT_cont->vBT_SS, il, d->ied, jl, d->jed); uBT_EE = V2(BT_cont->vBT_NN, il, d->ied, jl, d->jed);
    }
  }
  _Pragma("omp parallel for")
  for (int o = olo; o <= ohi; ++o) {
    std::vector<std::vector<double>> duhdu(nz + 1, std::vector<double>(nw)), visc_rem(nz + 1, std::vector<double>(nw, 1.0));
    std::vector<double> du(nw), du_min_CFL(nw), du_max_CFL(nw), duhdu_tot_0(nw), uh_tot_0(nw), visc_rem_max(nw), urow(nw),
        uhrow(nw);
    std::vector<char> do_I(nw, 1);
    // Set uh and duhdu.
    for (int k = 1; k <= nz; ++k) {
      if (use_visc_rem) for (int n = nlo; n <= nhi; ++n) visc_rem[k][n - n0] = (*visc_rem_u)(n, o, k);
      for (int n = nlo; n <= nhi; ++n) urow[n - n0] = u(n, o, k);
      flux_layer<Z>(E, urow.data(), h_in, hL, hR, k, uhrow.data(), duhdu[k].data(), visc_rem[k].data(), dt, o, nlo, nhi,
                    do_I.data(), por, n0);
      for (int n = nlo; n <= nhi; ++n) uh(n, o, k) = uhrow[n - n0];
    }
    if (uhbt || set_BT) {
      if (use_visc_rem && CS->use_visc_rem_max) {
        for (int x = 0; x < nw; ++x) visc_rem_max[x] = 0.0;
        for (int k = 1; k <= nz; ++k) for (int x = 0; x < nw; ++x) visc_rem_max[x] = fmax2(visc_rem_max[x], visc_rem[k][x]);
      } else {
        for (int x = 0; x < nw; ++x) visc_rem_max[x] = 1.0;
      }
      // Set limits on du that will keep the CFL number between -1 and 1.
      for (int n = nlo; n <= nhi; ++n) {
        const int x = n - n0;
        double I_vrm = 0.0;
        if (visc_rem_max[x] > 0.0) I_vrm = 1.0 / visc_rem_max[x];
        double dx_W, dx_E;
        if (CS->vol_CFL) {
          dx_W = ratio_max(M.areaT(n, o, 0), M.dy(n, o), 1000.0 * M.dxT(n, o, 0));
          dx_E = ratio_max(M.areaT(n, o, 1), M.dy(n, o), 1000.0 * M.dxT(n, o, 1));
        } else { dx_W = M.dxT(n, o, 0); dx_E = M.dxT(n, o, 1); }
        du_max_CFL[x] = 2.0 * (CFL_dt * dx_W) * I_vrm;
        du_min_CFL[x] = -2.0 * (CFL_dt * dx_E) * I_vrm;
        uh_tot_0[x] = 0.0; duhdu_tot_0[x] = 0.0;
      }
      for (int k = 1; k <= nz; ++k) for (int n = nlo; n <= nhi; ++n) {
        duhdu_tot_0[n - n0] = duhdu_tot_0[n - n0] + duhdu[k][n - n0];
        uh_tot_0[n - n0] = uh_tot_0[n - n0] + uh(n, o, k);
      }
      for (int k = 1; k <= nz; ++k) for (int n = nlo; n <= nhi; ++n) {
        const int x = n - n0;
        double dx_W, dx_E;
        if (CS->vol_CFL) {
          dx_W = ratio_max(M.areaT(n, o, 0), M.dy(n, o), 1000.0 * M.dxT(n, o, 0));
          dx_E = ratio_max(M.areaT(n, o, 1), M.dy(n, o), 1000.0 * M.dxT(n, o, 1));
        } else { dx_W = M.dxT(n, o, 0); dx_E = M.dxT(n, o, 1); }
        const double uk = u(n, o, k);
        if (use_visc_rem) {
          if (CS->aggress_adjust) {
            double du_lim = 0.499 * ((dx_W * I_dt - uk) + fmin2(0.0, c3<Z>(u, n, o, -1, k)));
            if (du_max_CFL[x] * visc_rem[k][x] > du_lim) du_max_CFL[x] = du_lim / visc_rem[k][x];
            du_lim = 0.499 * ((-dx_E * I_dt - uk) + fmax2(0.0, c3<Z>(u, n, o, +1, k)));
            if (du_min_CFL[x] * visc_rem[k][x] < du_lim) du_min_CFL[x] = du_lim / visc_rem[k][x];
          } else {
            if (du_max_CFL[x] * visc_rem[k][x] > dx_W * CFL_dt - uk * M.maskC(n, o))
              du_max_CFL[x] = (dx_W * CFL_dt - uk) / visc_rem[k][x];
            if (du_min_CFL[x] * visc_rem[k][x] < -dx_E * CFL_dt - uk * M.maskC(n, o))
              du_min_CFL[x] = -(dx_E * CFL_dt + uk) / visc_rem[k][x];
          }
        } else {
          if (CS->aggress_adjust) {
            du_max_CFL[x] = fmin2(du_max_CFL[x], 0.499 * ((dx_W * I_dt - uk) + fmin2(0.0, c3<Z>(u, n, o, -1, k))));
            du_min_CFL[x] = fmax2(du_min_CFL[x], 0.499 * ((-dx_E * I_dt - uk) + fmax2(0.0, c3<Z>(u, n, o, +1, k))));
          } else {
            du_max_CFL[x] = fmin2(du_max_CFL[x], dx_W * CFL_dt - uk);
            du_min_CFL[x] = fmax2(du_min_CFL[x], -(dx_E * CFL_dt + uk));
          }
        }
      }
      for (int x = 0; x < nw; ++x) {
        du_max_CFL[x] = fmax2(du_max_CFL[x], 0.0);
        du_min_CFL[x] = fmin2(du_min_CFL[x], 0.0);
      }
      for (int x = 0; x < nw; ++x) do_I[x] = 1;
      if (uhbt) {
        std::vector<double> uhbt_row(nw);
        for (int n = nlo; n <= nhi; ++n) uhbt_row[n - n0] = (*uhbt)(n, o);
        flux_adjust<Z>(E, u, h_in, hL, hR, uhbt_row.data(), uh_tot_0.data(), duhdu_tot_0.data(), du.data(), du_max_CFL.data(),
                       du_min_CFL.data(), dt, visc_rem, o, nlo, nhi, do_I.data(), por, &uh, n0, nw);
        if (u_cor) for (int k = 1; k <= nz; ++k) for (int n = nlo; n <= nhi; ++n)
          (*u_cor)(n, o, k) = u(n, o, k) + du[n - n0] * visc_rem[k][n - n0];
        if (du_cor) for (int n = nlo; n <= nhi; ++n) (*du_cor)(n, o) = du[n - n0];
      }
      if (set_BT)
        set_BT_cont<Z>(E, u, h_in, hL, hR, FA_W0, FA_WW, FA_E0, FA_EE, uBT_WW, uBT_EE, uh_tot_0.data(), duhdu_tot_0.data(),
                       du_max_CFL.data(), du_min_CFL.data(), dt, visc_rem, visc_rem_max.data(), o, nlo, nhi, do_I.data(), por,
                       n0, nw);
    }
  }
  if (set_BT && h_u_out) {
    if (u_cor) flux_thickness<Z>(E, *u_cor, h_in, hL, hR, *h_u_out, dt, lb, por, visc_rem_u);
    else flux_thickness<Z>(E, u, h_in, hL, hR, *h_u_out, dt, lb, por, visc_rem_u);
  }
}

// continuity_zonal_convergence :348-383 / continuity_merdional_convergence :386-421
template <bool Z>
void convergence(const Env& E, const V3& h, const V3& uh, double dt, const LB& lb, const V3* hin, double h_min) {
  const int nz = E.nz;
  _Pragma("omp parallel for")
  for (int k = 1; k <= nz; ++k) for (int j = lb.jsh; j <= lb.jeh; ++j) for (int i = lb.ish; i <= lb.ieh; ++i) {
    const double h0 = hin ? (*hin)(i, j, k) : h(i, j, k);
    h(i, j, k) = fmax2(h0 - dt * E.IareaT(i, j) * (uh(i, j, k) - c3<Z>(uh, i, j, -1, k)), h_min);
  }
}

LB set_continuity_loop_bounds(const Env& E, bool i_stencil, bool j_stencil) {  // :2766-2796
  int stencil = 3; if (E.CS->simple_2nd) stencil = 2; if (E.CS->upwind_1st) stencil = 1;
  LB lb;
  if (i_stencil) { lb.ish = E.d->isc - stencil; lb.ieh = E.d->iec + stencil; } else { lb.ish = E.d->isc; lb.ieh = E.d->iec; }
  if (j_stencil) { lb.jsh = E.d->jsc - stencil; lb.jeh = E.d->jec + stencil; } else { lb.jsh = E.d->jsc; lb.jeh = E.d->jec; }
  return lb;
}

}  // namespace

extern "C" int oracle_continuity(const mom6cu_domain* d, const mom6cu_grid* G, const mom6cu_vgrid* GV,
                                 const mom6cu_continuity_cs* CS, const mom6cu_continuity_args* a, int nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  const int isd = d->isd, ied = d->ied, jsd = d->jsd, jed = d->jed, nz = d->nk;
  Env E;
  E.d = d; E.CS = CS; E.Angstrom_H = GV->Angstrom_H; E.H_subroundoff = GV->H_subroundoff; E.nz = nz;
  auto H2 = [&](const double* p) { return V2((double*)p, isd, ied, jsd, jed); };
  auto U2 = [&](const double* p) { return V2((double*)p, isd - 1, ied, jsd, jed); };
  auto Vv2 = [&](const double* p) { return V2((double*)p, isd, ied, jsd - 1, jed); };
  auto H3 = [&](const double* p) { return V3((double*)p, isd, ied, jsd, jed, nz); };
  auto U3 = [&](const double* p) { return V3((double*)p, isd - 1, ied, jsd, jed, nz); };
  auto Vv3 = [&](const double* p) { return V3((double*)p, isd, ied, jsd - 1, jed, nz); };
  E.mask2dT = H2(G->mask2dT); E.mask2dCu = U2(G->mask2dCu); E.mask2dCv = Vv2(G->mask2dCv);
  E.dxT = H2(G->dxT); E.dyT = H2(G->dyT); E.IdxT = H2(G->IdxT); E.IdyT = H2(G->IdyT); E.areaT = H2(G->areaT);
  E.IareaT = H2(G->IareaT); E.dy_Cu = U2(G->dy_Cu); E.dx_Cv = Vv2(G->dx_Cv); E.dxCu = U2(G->dxCu); E.dyCv = Vv2(G->dyCv);

  V3 u = U3(a->u), v = Vv3(a->v), hin = H3(a->hin), h = H3(a->h), uh = U3(a->uh), vh = Vv3(a->vh);
  V3 porU, porV, vru, vrv, ucor, vcor, hu, hv;
  V2 uhbt, vhbt, ducor, dvcor;
  const V3 *pporU = nullptr, *pporV = nullptr, *pvru = nullptr, *pvrv = nullptr, *pucor = nullptr, *pvcor = nullptr,
           *phu = nullptr, *phv = nullptr;
  const V2 *puhbt = nullptr, *pvhbt = nullptr, *pducor = nullptr, *pdvcor = nullptr;
  if (a->por_face_areaU) { porU = U3(a->por_face_areaU); pporU = &porU; }
  if (a->por_face_areaV) { porV = Vv3(a->por_face_areaV); pporV = &porV; }
  if ((a->visc_rem_u != nullptr) != (a->visc_rem_v != nullptr)) return 2;  // :159-161 FATAL
  if (a->visc_rem_u) { vru = U3(a->visc_rem_u); pvru = &vru; vrv = Vv3(a->visc_rem_v); pvrv = &vrv; }
  if (a->u_cor) { ucor = U3(a->u_cor); pucor = &ucor; }
  if (a->v_cor) { vcor = Vv3(a->v_cor); pvcor = &vcor; }
  if (a->uhbt) { uhbt = U2(a->uhbt); puhbt = &uhbt; }
  if (a->vhbt) { vhbt = Vv2(a->vhbt); pvhbt = &vhbt; }
  if (a->du_cor) { ducor = U2(a->du_cor); pducor = &ducor; }
  if (a->dv_cor) { dvcor = Vv2(a->dv_cor); pdvcor = &dvcor; }
  if (a->BT_cont && a->BT_cont->h_u) { hu = U3(a->BT_cont->h_u); phu = &hu; }
  if (a->BT_cont && a->BT_cont->h_v) { hv = Vv3(a->BT_cont->h_v); phv = &hv; }

  // Local variables :144-147
  A3 h_W(isd, ied, jsd, jed, nz), h_E(isd, ied, jsd, jed, nz), h_S(isd, ied, jsd, jed, nz), h_N(isd, ied, jsd, jed, nz);
  const double h_min = GV->Angstrom_H;  // :152
  const bool x_first = ((d->first_direction % 2) == 0);  // :157
  const double dt = a->dt;
  if (x_first) {
    // First advect zonally, with loop bounds that accomodate the subsequent meridional advection.
    LB lb = set_continuity_loop_bounds(E, false, true);
    edge_thickness<true>(E, hin, h_W, h_E, lb);
    mass_flux_impl<true>(E, u, hin, h_W, h_E, uh, dt, pporU, lb, puhbt, pvru, pucor, a->BT_cont, pducor, phu);
    convergence<true>(E, h, uh, dt, lb, &hin, 0.0);
    // Now advect meridionally, using the updated thicknesses to determine the fluxes.
    lb = set_continuity_loop_bounds(E, false, false);
    edge_thickness<false>(E, h, h_S, h_N, lb);
    mass_flux_impl<false>(E, v, h, h_S, h_N, vh, dt, pporV, lb, pvhbt, pvrv, pvcor, a->BT_cont, pdvcor, phv);
    convergence<false>(E, h, vh, dt, lb, nullptr, h_min);
  } else {
    LB lb = set_continuity_loop_bounds(E, true, false);
    edge_thickness<false>(E, hin, h_S, h_N, lb);
    mass_flux_impl<false>(E, v, hin, h_S, h_N, vh, dt, pporV, lb, pvhbt, pvrv, pvcor, a->BT_cont, pdvcor, phv);
    convergence<false>(E, h, vh, dt, lb, &hin, 0.0);
    lb = set_continuity_loop_bounds(E, false, false);
    edge_thickness<true>(E, h, h_W, h_E, lb);
    mass_flux_impl<true>(E, u, h, h_W, h_E, uh, dt, pporU, lb, puhbt, pvru, pucor, a->BT_cont, pducor, phu);
    convergence<true>(E, h, uh, dt, lb, nullptr, h_min);
  }
  return 0;
}
