// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of advect_tracer, /root/reference/src/tracer/MOM_tracer_advect.F90:53-350, with advect_x :355-744 and
// advect_y :748-1152 (PLM, PPM:H3 and PPM reconstructions).  No OBCs; the flux diagnostics (ad_x, ad_y, ad2d_x, ad2d_y,
// advection_xy) are not produced.  Single tile: do_group_pass is the periodic wrap of oracle_fill_halo_2d and
// sum_across_PEs is the identity.  The loop structure (valid ranges that march inward between halo updates, the
// domore_u / domore_v / domore_k flags, max_iter) is kept exactly, because it decides how many passes are made.
// PARITY: PINNED BY A REFERENCE RUN -- the reference's own MOM_tracer_advect.F90, executed by oracle/f90run, agrees bit for bit
// (tests/test_reference_f90.py, advect_tracer/*).
#include "oracle.h"
#include "ogrid.hpp"
#include <cfloat>
#include <cmath>
#include <vector>
#include <algorithm>

using namespace orc;

namespace {

inline double min3(double a, double b, double c) { return fmin2(fmin2(a, b), c); }
inline double max3(double a, double b, double c) { return fmax2(fmax2(a, b), c); }

struct Adv {
  const OGrid& G;
  int ntr;
  std::vector<V3> Tr;
  std::vector<int> scheme;
  std::vector<double> underflow;
  double Angstrom_H, H_subroundoff;
  std::vector<char> domore_u, domore_v;  // (j - jsd, k-1), (J - JsdB, k-1)
  char& du(int j, int k) { return domore_u[(size_t)(k - 1) * (G.jed - G.jsd + 1) + (j - G.jsd)]; }
  char& dv(int J, int k) { return domore_v[(size_t)(k - 1) * (G.JedB - G.JsdB + 1) + (J - G.JsdB)]; }
};

// the monotonic slope of :455-459 (x) and :814-818 (y)
inline double plm_slope(double Tp, double Tc, double Tm, double maskprod) {
  const double dMx = max3(Tp, Tc, Tm) - Tc;
  const double dMn = Tc - min3(Tp, Tc, Tm);
  return maskprod * fsign(min3(0.5 * std::fabs(Tp - Tm), 2.0 * dMx, 2.0 * dMn), Tp - Tm);
}

// the PPM / PPM:H3 face flux of :546-590 given the upstream triplet
inline double ppm_flux(int scheme, double Tp, double Tc, double Tm, double sl_m, double sl_c, double sl_p, double maskprod, double uhh,
                       double CFL) {
  double aL, aR;
  if (scheme == MOM6CU_ADVECT_PPMH3) {
    aL = (5. * Tc + (2. * Tm - Tp)) / 6.;
    aL = fmax2(fmin2(Tc, Tm), aL); aL = fmin2(fmax2(Tc, Tm), aL);
    aR = (5. * Tc + (2. * Tp - Tm)) / 6.;
    aR = fmax2(fmin2(Tc, Tp), aR); aR = fmin2(fmax2(Tc, Tp), aR);
  } else {
    aL = 0.5 * ((Tm + Tc) + (sl_m - sl_c) / 3.);
    aR = 0.5 * ((Tc + Tp) + (sl_c - sl_p) / 3.);
  }
  const double dA = aR - aL, mA = 0.5 * (aR + aL);
  if (maskprod * (Tp - Tc) * (Tc - Tm) <= 0.) { aL = Tc; aR = Tc; }
  else if (dA * (Tc - mA) > (dA * dA) / 6.) aL = (3. * Tc) - 2. * aR;
  else if (dA * (Tc - mA) < -(dA * dA) / 6.) aR = (3. * Tc) - 2. * aL;
  const double a6 = 6. * Tc - 3. * (aR + aL);
  if (uhh >= 0.0) return uhh * (aR - 0.5 * CFL * ((aR - aL) - a6 * (1. - 2. / 3. * CFL)));
  return uhh * (aL + 0.5 * CFL * ((aR - aL) + a6 * (1. - 2. / 3. * CFL)));
}

// advect_x :355-744
void advect_x(Adv& A, const V3& hprev, const V3& uhr, const V2& uh_neglect, int is, int ie, int js, int je, int k) {
  const OGrid& G = A.G;
  const int ntr = A.ntr;
  bool usePLMslope = false;
  int stencil = 1;
  for (int m = 0; m < ntr; ++m) {
    if (A.scheme[m] == MOM6CU_ADVECT_PLM || A.scheme[m] == MOM6CU_ADVECT_PPM) usePLMslope = true;
    if (A.scheme[m] == MOM6CU_ADVECT_PPM) stencil = 2;
  }
  const double min_h = 0.1 * A.Angstrom_H, tiny_h = DBL_MIN, h_neglect = A.H_subroundoff;
  const int n = G.ied - G.isd + 3;
  auto X = [&](int i) { return i - (G.isd - 1); };  // index into the row arrays (covers IsdB..ied)
  std::vector<double> slope_x((size_t)n * ntr, 0.), T_tmp((size_t)n * ntr, 0.), flux_x((size_t)n * ntr, 0.), uhh(n, 0.), CFL(n, 0.),
      hlst(n, 0.), Ihnew(n, 0.);
  std::vector<char> do_i(n, 0);
  for (int j = js; j <= je; ++j) if (A.du(j, k)) {
    A.du(j, k) = 0;
    if (usePLMslope)
      for (int m = 0; m < ntr; ++m) for (int i = is - stencil; i <= ie + stencil; ++i)
        slope_x[(size_t)m * n + X(i)] = plm_slope(A.Tr[m](i + 1, j, k), A.Tr[m](i, j, k), A.Tr[m](i - 1, j, k), G.mask2dCu(i, j) * G.mask2dCu(i - 1, j));
    for (int m = 0; m < ntr; ++m) for (int i = G.isd; i <= G.ied; ++i) T_tmp[(size_t)m * n + X(i)] = A.Tr[m](i, j, k);
    for (int I = is - 1; I <= ie; ++I) {
      const double u = uhr(I, j, k);
      if ((u == 0.0) || ((u < 0.0) && (hprev(I + 1, j, k) <= tiny_h)) || ((u > 0.0) && (hprev(I, j, k) <= tiny_h))) {
        uhh[X(I)] = 0.0; CFL[X(I)] = 0.0;
      } else if (u < 0.0) {
        const double hup = hprev(I + 1, j, k) - G.areaT(I + 1, j) * min_h;
        const double hlos = fmax2(0.0, uhr(I + 1, j, k));
        if ((((hup - hlos) + u) < 0.0) && ((0.5 * hup + u) < 0.0)) { uhh[X(I)] = min3(-0.5 * hup, -hup + hlos, 0.0); A.du(j, k) = 1; }
        else uhh[X(I)] = u;
        CFL[X(I)] = -uhh[X(I)] / (hprev(I + 1, j, k));
      } else {
        const double hup = hprev(I, j, k) - G.areaT(I, j) * min_h;
        const double hlos = fmax2(0.0, -uhr(I - 1, j, k));
        if ((((hup - hlos) - u) < 0.0) && ((0.5 * hup - u) < 0.0)) { uhh[X(I)] = max3(0.5 * hup, hup - hlos, 0.0); A.du(j, k) = 1; }
        else uhh[X(I)] = u;
        CFL[X(I)] = uhh[X(I)] / (hprev(I, j, k));
      }
    }
    for (int m = 0; m < ntr; ++m) {
      const double* T = &T_tmp[(size_t)m * n];
      const double* sl = &slope_x[(size_t)m * n];
      double* fl = &flux_x[(size_t)m * n];
      if (A.scheme[m] == MOM6CU_ADVECT_PPM || A.scheme[m] == MOM6CU_ADVECT_PPMH3) {
        for (int I = is - 1; I <= ie; ++I) {
          const int i_up = (uhh[X(I)] >= 0.0) ? I : I + 1;
          fl[X(I)] = ppm_flux(A.scheme[m], T[X(i_up + 1)], T[X(i_up)], T[X(i_up - 1)], sl[X(i_up - 1)], sl[X(i_up)], sl[X(i_up + 1)],
                              G.mask2dCu(i_up, j) * G.mask2dCu(i_up - 1, j), uhh[X(I)], CFL[X(I)]);
        }
      } else {
        for (int I = is - 1; I <= ie; ++I) {
          if (uhh[X(I)] >= 0.0) fl[X(I)] = uhh[X(I)] * (T[X(I)] + 0.5 * sl[X(I)] * (1. - CFL[X(I)]));
          else fl[X(I)] = uhh[X(I)] * (T[X(I + 1)] - 0.5 * sl[X(I + 1)] * (1. - CFL[X(I)]));
        }
      }
    }
    for (int I = is - 1; I <= ie; ++I) {
      uhr(I, j, k) = uhr(I, j, k) - uhh[X(I)];
      if (std::fabs(uhr(I, j, k)) < uh_neglect(I, j)) uhr(I, j, k) = 0.0;
    }
    for (int i = is; i <= ie; ++i) {
      if ((uhh[X(i)] != 0.0) || (uhh[X(i - 1)] != 0.0)) {
        do_i[X(i)] = 1;
        hlst[X(i)] = hprev(i, j, k);
        hprev(i, j, k) = hprev(i, j, k) - (uhh[X(i)] - uhh[X(i - 1)]);
        if (hprev(i, j, k) <= 0.0) do_i[X(i)] = 0;
        else if (hprev(i, j, k) < h_neglect * G.areaT(i, j)) {
          hlst[X(i)] = hlst[X(i)] + (h_neglect * G.areaT(i, j) - hprev(i, j, k));
          Ihnew[X(i)] = 1.0 / (h_neglect * G.areaT(i, j));
        } else Ihnew[X(i)] = 1.0 / hprev(i, j, k);
      } else do_i[X(i)] = 0;
    }
    for (int m = 0; m < ntr; ++m) {
      const double* fl = &flux_x[(size_t)m * n];
      for (int i = is; i <= ie; ++i) if (do_i[X(i)]) {
        if (Ihnew[X(i)] > 0.0) A.Tr[m](i, j, k) = (A.Tr[m](i, j, k) * hlst[X(i)] - (fl[X(i)] - fl[X(i - 1)])) * Ihnew[X(i)];
      }
    }
  }
  for (int m = 0; m < ntr; ++m) if (A.underflow[m] > 0.0)
    for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i)
      if (std::fabs(A.Tr[m](i, j, k)) < A.underflow[m]) A.Tr[m](i, j, k) = 0.0;
}

// advect_y :748-1152
void advect_y(Adv& A, const V3& hprev, const V3& vhr, const V2& vh_neglect, int is, int ie, int js, int je, int k) {
  const OGrid& G = A.G;
  const int ntr = A.ntr;
  bool usePLMslope = false;
  int stencil = 1;
  for (int m = 0; m < ntr; ++m) {
    if (A.scheme[m] == MOM6CU_ADVECT_PLM || A.scheme[m] == MOM6CU_ADVECT_PPM) usePLMslope = true;
    if (A.scheme[m] == MOM6CU_ADVECT_PPM) stencil = 2;
  }
  const double min_h = 0.1 * A.Angstrom_H, tiny_h = DBL_MIN, h_neglect = A.H_subroundoff;
  const int ni = G.ied - G.isd + 1, nj = G.jed - G.jsd + 3;  // rows JsdB-1 .. jed+... (guard row each side)
  auto XI = [&](int i) { return i - G.isd; };
  auto YJ = [&](int j) { return j - (G.jsd - 1); };
  auto at = [&](int i, int j) { return (size_t)YJ(j) * ni + XI(i); };
  std::vector<std::vector<double>> slope_y(ntr, std::vector<double>((size_t)ni * nj, 0.)), flux_y(ntr, std::vector<double>((size_t)ni * nj, 0.)),
      T_tmp(ntr, std::vector<double>((size_t)ni * nj, 0.));
  std::vector<double> vhh((size_t)ni * nj, 0.), CFL(ni, 0.), hlst(ni, 0.), Ihnew(ni, 0.);
  std::vector<char> do_j_tr(nj + 4, 0), do_i(ni, 0);
  for (int J = js - 1; J <= je; ++J) if (A.dv(J, k)) for (int j2 = 1 - stencil; j2 <= stencil; ++j2) do_j_tr[YJ(J + j2) + 2] = 1;
  if (usePLMslope)
    for (int j = js - stencil; j <= je + stencil; ++j) if (do_j_tr[YJ(j) + 2]) for (int m = 0; m < ntr; ++m) for (int i = is; i <= ie; ++i)
      slope_y[m][at(i, j)] = plm_slope(A.Tr[m](i, j + 1, k), A.Tr[m](i, j, k), A.Tr[m](i, j - 1, k), G.mask2dCv(i, j) * G.mask2dCv(i, j - 1));
  for (int j = G.jsd; j <= G.jed; ++j) for (int m = 0; m < ntr; ++m) for (int i = G.isd; i <= G.ied; ++i) T_tmp[m][at(i, j)] = A.Tr[m](i, j, k);
  for (int J = js - 1; J <= je; ++J) {
    if (A.dv(J, k)) {
      A.dv(J, k) = 0;
      for (int i = is; i <= ie; ++i) {
        const double v = vhr(i, J, k);
        double& vh_ = vhh[at(i, J)];
        if ((v == 0.0) || ((v < 0.0) && (hprev(i, J + 1, k) <= tiny_h)) || ((v > 0.0) && (hprev(i, J, k) <= tiny_h))) {
          vh_ = 0.0; CFL[XI(i)] = 0.0;
        } else if (v < 0.0) {
          const double hup = hprev(i, J + 1, k) - G.areaT(i, J + 1) * min_h;
          const double hlos = fmax2(0.0, vhr(i, J + 1, k));
          if ((((hup - hlos) + v) < 0.0) && ((0.5 * hup + v) < 0.0)) { vh_ = min3(-0.5 * hup, -hup + hlos, 0.0); A.dv(J, k) = 1; }
          else vh_ = v;
          CFL[XI(i)] = -vh_ / hprev(i, J + 1, k);
        } else {
          const double hup = hprev(i, J, k) - G.areaT(i, J) * min_h;
          const double hlos = fmax2(0.0, -vhr(i, J - 1, k));
          if ((((hup - hlos) - v) < 0.0) && ((0.5 * hup - v) < 0.0)) { vh_ = max3(0.5 * hup, hup - hlos, 0.0); A.dv(J, k) = 1; }
          else vh_ = v;
          CFL[XI(i)] = vh_ / hprev(i, J, k);
        }
      }
      for (int m = 0; m < ntr; ++m) {
        if (A.scheme[m] == MOM6CU_ADVECT_PPM || A.scheme[m] == MOM6CU_ADVECT_PPMH3) {
          for (int i = is; i <= ie; ++i) {
            const double vh_ = vhh[at(i, J)];
            const int j_up = (vh_ >= 0.0) ? J : J + 1;
            flux_y[m][at(i, J)] = ppm_flux(A.scheme[m], T_tmp[m][at(i, j_up + 1)], T_tmp[m][at(i, j_up)], T_tmp[m][at(i, j_up - 1)],
                                           slope_y[m][at(i, j_up - 1)], slope_y[m][at(i, j_up)], slope_y[m][at(i, j_up + 1)],
                                           G.mask2dCv(i, j_up) * G.mask2dCv(i, j_up - 1), vh_, CFL[XI(i)]);
          }
        } else {
          for (int i = is; i <= ie; ++i) {
            const double vh_ = vhh[at(i, J)];
            if (vh_ >= 0.0) flux_y[m][at(i, J)] = vh_ * (T_tmp[m][at(i, J)] + 0.5 * slope_y[m][at(i, J)] * (1. - CFL[XI(i)]));
            else flux_y[m][at(i, J)] = vh_ * (T_tmp[m][at(i, J + 1)] - 0.5 * slope_y[m][at(i, J + 1)] * (1. - CFL[XI(i)]));
          }
        }
      }
    } else {
      for (int i = is; i <= ie; ++i) vhh[at(i, J)] = 0.0;
      for (int m = 0; m < ntr; ++m) for (int i = is; i <= ie; ++i) flux_y[m][at(i, J)] = 0.0;
    }
  }
  for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
    vhr(i, J, k) = vhr(i, J, k) - vhh[at(i, J)];
    if (std::fabs(vhr(i, J, k)) < vh_neglect(i, J)) vhr(i, J, k) = 0.0;
  }
  for (int j = js; j <= je; ++j) if (do_j_tr[YJ(j) + 2]) {
    for (int i = is; i <= ie; ++i) {
      if ((vhh[at(i, j)] != 0.0) || (vhh[at(i, j - 1)] != 0.0)) {
        do_i[XI(i)] = 1;
        hlst[XI(i)] = hprev(i, j, k);
        hprev(i, j, k) = fmax2(hprev(i, j, k) - (vhh[at(i, j)] - vhh[at(i, j - 1)]), 0.0);
        if (hprev(i, j, k) <= 0.0) do_i[XI(i)] = 0;
        else if (hprev(i, j, k) < h_neglect * G.areaT(i, j)) {
          hlst[XI(i)] = hlst[XI(i)] + (h_neglect * G.areaT(i, j) - hprev(i, j, k));
          Ihnew[XI(i)] = 1.0 / (h_neglect * G.areaT(i, j));
        } else Ihnew[XI(i)] = 1.0 / hprev(i, j, k);
      } else do_i[XI(i)] = 0;
    }
    for (int m = 0; m < ntr; ++m)
      for (int i = is; i <= ie; ++i) if (do_i[XI(i)])
        A.Tr[m](i, j, k) = (A.Tr[m](i, j, k) * hlst[XI(i)] - (flux_y[m][at(i, j)] - flux_y[m][at(i, j - 1)])) * Ihnew[XI(i)];
  }
  for (int m = 0; m < ntr; ++m) if (A.underflow[m] > 0.0)
    for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i)
      if (std::fabs(A.Tr[m](i, j, k)) < A.underflow[m]) A.Tr[m](i, j, k) = 0.0;
}

void fill3(const mom6cu_domain* d, const V3& f, int st) {
  for (int k = 1; k <= f.nk; ++k) oracle_fill_halo_2d(d, &f(f.ilo, f.jlo, k), st, 0);
}

}  // namespace

extern "C" int oracle_advect_tracer(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV, const mom6cu_tracer_advect_cs* CS,
                                    const mom6cu_advect_tracer_args* a, int* iterations) {
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, nz = G.ke;
  const int isd = G.isd, ied = G.ied, jsd = G.jsd, jed = G.jed;
  const int ntr = a->ntr;
  if (iterations) *iterations = 0;
  if (ntr == 0) return 0;
  Adv A{G};
  A.ntr = ntr; A.Angstrom_H = GV->Angstrom_H; A.H_subroundoff = GV->H_subroundoff;
  A.domore_u.assign((size_t)(jed - jsd + 1) * nz, 0); A.domore_v.assign((size_t)(G.JedB - G.JsdB + 1) * nz, 0);
  int stencil = 2;
  for (int m = 0; m < ntr; ++m) {
    A.Tr.push_back(G.H3(a->tr[m]));
    int s = a->advect_scheme ? a->advect_scheme[m] : -1;
    if (s < 0) s = CS->default_advect_scheme;
    A.scheme.push_back(s);
    A.underflow.push_back(a->conc_underflow ? a->conc_underflow[m] : 0.0);
    int stencil_local = 2;
    if (s == MOM6CU_ADVECT_PPM) stencil_local = 3;
    else if (s == MOM6CU_ADVECT_PPMH3) stencil_local = CS->useHuynhStencilBug ? 2 : 3;
    else if (s != MOM6CU_ADVECT_PLM) return 3;
    stencil = std::max(stencil, stencil_local);
  }
  if (std::min(std::min(is - isd, ied - ie), std::min(js - jsd, jed - je)) < stencil) return 2;  // FATAL :172
  const double dt = a->dt;
  int max_iter = 2 * (int)std::ceil(dt / CS->dt) + 1;
  if (a->max_iter_in >= 0) max_iter = a->max_iter_in;
  bool x_first = (d->first_direction % 2 == 0);
  if (a->x_first_in >= 0) x_first = a->x_first_in != 0;
  const V3 h_end = G.H3(a->h_end), uhtr = G.U3(a->uhtr), vhtr = G.V3_(a->vhtr);
  A3 hprev(isd, ied, jsd, jed, nz), uhr(isd - 1, ied, jsd, jed, nz), vhr(isd, ied, jsd - 1, jed, nz);
  A2 uh_neglect = G.aU(), vh_neglect = G.aV();
  std::vector<int> domore_k(nz + 1, 1);
  for (int k = 1; k <= nz; ++k) {
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) uhr(I, j, k) = uhtr(I, j, k);
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) vhr(i, J, k) = vhtr(i, J, k);
    if (!a->vol_prev) {
      for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
        hprev(i, j, k) = fmax2(0.0, G.areaT(i, j) * h_end(i, j, k) + ((uhr(i, j, k) - uhr(i - 1, j, k)) + (vhr(i, j, k) - vhr(i, j - 1, k))));
        hprev(i, j, k) = hprev(i, j, k) + fmax2(0.0, 1.0e-13 * hprev(i, j, k) - G.areaT(i, j) * h_end(i, j, k));
      }
    } else {
      const V3 vp = G.H3(a->vol_prev);
      for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) hprev(i, j, k) = vp(i, j, k);
    }
  }
  for (int j = jsd; j <= jed; ++j) for (int I = isd; I <= ied - 1; ++I) uh_neglect(I, j) = GV->H_subroundoff * fmin2(G.areaT(I, j), G.areaT(I + 1, j));
  for (int J = jsd; J <= jed - 1; ++J) for (int i = isd; i <= ied; ++i) vh_neglect(i, J) = GV->H_subroundoff * fmin2(G.areaT(i, J), G.areaT(i, J + 1));

  int isv = is, iev = ie, jsv = js, jev = je;
  int itt;
  for (itt = 1; itt <= max_iter; ++itt) {
    if (isv > is - stencil) {
      fill3(d, uhr, 1); fill3(d, vhr, 2); fill3(d, hprev, 0);
      for (int m = 0; m < ntr; ++m) fill3(d, A.Tr[m], 0);
      const int nsten_halo = std::min(std::min(is - isd, ied - ie), std::min(js - jsd, jed - je)) / stencil;
      isv = is - nsten_halo * stencil; jsv = js - nsten_halo * stencil;
      iev = ie + nsten_halo * stencil; jev = je + nsten_halo * stencil;
      if ((nsten_halo > 1) || (itt == 1)) {
        for (int k = 1; k <= nz; ++k) if (domore_k[k] > 0) {
          for (int j = jsv; j <= jev; ++j) if (!A.du(j, k))
            for (int I = isv + stencil - 1; I <= iev - stencil; ++I) if (uhr(I, j, k) != 0.0) { A.du(j, k) = 1; break; }
          for (int J = jsv + stencil - 1; J <= jev - stencil; ++J) if (!A.dv(J, k))
            for (int i = isv + stencil; i <= iev - stencil; ++i) if (vhr(i, J, k) != 0.0) { A.dv(J, k) = 1; break; }
          domore_k[k] = 0;
          for (int j = jsv; j <= jev; ++j) if (A.du(j, k)) domore_k[k] = 1;
          for (int J = jsv + stencil - 1; J <= jev - stencil; ++J) if (A.dv(J, k)) domore_k[k] = 1;
        }
      }
    }
    isv = isv + stencil; iev = iev - stencil; jsv = jsv + stencil; jev = jev - stencil;
    if (x_first) {
      for (int k = 1; k <= nz; ++k) if (domore_k[k] > 0) advect_x(A, hprev, uhr, uh_neglect, isv, iev, jsv - stencil, jev + stencil, k);
      for (int k = 1; k <= nz; ++k) if (domore_k[k] > 0) {
        advect_y(A, hprev, vhr, vh_neglect, isv, iev, jsv, jev, k);
        domore_k[k] = 0;
        for (int j = jsv - stencil; j <= jev + stencil; ++j) if (A.du(j, k)) domore_k[k] = 1;
        for (int J = jsv - 1; J <= jev; ++J) if (A.dv(J, k)) domore_k[k] = 1;
      }
    } else {
      for (int k = 1; k <= nz; ++k) if (domore_k[k] > 0) advect_y(A, hprev, vhr, vh_neglect, isv - stencil, iev + stencil, jsv, jev, k);
      for (int k = 1; k <= nz; ++k) if (domore_k[k] > 0) {
        advect_x(A, hprev, uhr, uh_neglect, isv, iev, jsv, jev, k);
        domore_k[k] = 0;
        for (int j = jsv; j <= jev; ++j) if (A.du(j, k)) domore_k[k] = 1;
        for (int J = jsv - 1; J <= jev; ++J) if (A.dv(J, k)) domore_k[k] = 1;
      }
    }
    if (itt >= max_iter) break;
    if (isv > is - stencil) {
      int do_any = 0;
      for (int k = 1; k <= nz; ++k) do_any += domore_k[k];
      if (do_any == 0) break;
    }
  }
  if (iterations) *iterations = std::min(itt, max_iter);
  if (a->uhr_out) { const V3 o = G.U3(a->uhr_out); for (size_t n = 0; n < o.size(); ++n) o.p[n] = uhr.p[n]; }
  if (a->vhr_out) { const V3 o = G.V3_(a->vhr_out); for (size_t n = 0; n < o.size(); ++n) o.p[n] = vhr.p[n]; }
  if (a->vol_prev && a->update_vol_prev) { const V3 o = G.H3(a->vol_prev); for (size_t n = 0; n < o.size(); ++n) o.p[n] = hprev.p[n]; }
  return 0;
}
