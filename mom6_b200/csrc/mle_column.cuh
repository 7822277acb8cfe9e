// The per-column and per-face work of mixedlayer_restrat_OM4 (src/parameterizations/lateral/MOM_mixed_layer_restrat.F90:189-714)
// as host/device code on the unified plane layout (common.cuh): the kernels of mle.cu call it with one thread per column / face,
// tests/harness/mle_host.cpp compiles the same functions with g++ and loops over the tile, so the code the GPU threads run is
// checked bit for bit against the oracle without a GPU (tests/test_mle.py).
//
// What differs from the reference's loop structure, and why every bit is the same:
//  * h_avail(i,j,k) = max(I4dt*areaT*(h - Angstrom_H), 0) (:382) is evaluated where a face needs it instead of being stored.
//  * a(k), b(k) = mu(top) - mu(bottom) (:528-531, :541-545) are recomputed in each of the three passes instead of stored; the value of
//    mu at a layer's bottom interface is kept for the next layer's top (the reference evaluates mu twice at the same argument).
//  * With MLE_TAIL_DH = 0, mu(sigma) = +0 for sigma <= -1, so once the top of a layer is at or below the base of the mixed layer
//    a(k) = (+0) - (+0): the limiter tests `a*uDml > 0`, `a*uDml < 0` are both false and the transport is a(k)*uDml + b(k)*uDml_slow =
//    (+0)*uDml + (+0)*uDml_slow, a signed zero.  The passes stop there; the remaining layers get that zero, and uhtr + zero*dt is
//    uhtr itself except that (-0) + (+0) = +0, which is applied where it occurs.
#pragma once
#include "mle_mu.cuh"

namespace m6mle {

struct Par {
  int nk;
  double dt, Z_to_H, Angstrom_H, h_neglect, g_Rho0, I4dt, h_min, vonKar_x_pi2, ustar_min, coef, coef2, front_length, stretch, tail_dh;
  double aFac1, bFac1, aFac2, bFac2;
  int filt1, filt2, res_upscale;
  Eos eos;
  int detect;           // MLE_DENSITY_DIFF > 0: detect_mld (:1503-1569) instead of MLE_MLD_STRETCH * h_MLD
  double density_diff;  // CS%MLE_density_diff
};

// :302-346 (the MLD filters) and :375-412 (mixed-layer thickness and thickness-weighted density of the column at plane offset g)
M6M_HD void column(const Par& P, const long long g, const long long pl, const double* h, const double* T, const double* S,
                   const double* h_MLD, double* MLD_filtered, double* MLD_filtered_slow, double* htot_fast, double* htot_slow,
                   double* Rml_av_fast, double* Rml_av_slow) {
  double MLD_fast;
  if (P.detect) {  // detect_mld :1503-1569: the depth at which sigma-0 exceeds its surface value by MLE_DENSITY_DIFF
    double hk = h[g];
    double dK = 0.5 * hk, dKm1;
    const double rhoSurf = density(P.eos, T[g], S[g], 0.);
    double dRk = 0., dRkm1;
    double mld = 0.;
    for (int k = 1; k < P.nk; ++k) {
      const long long gk = g + (long long)k * pl;
      const double hk1 = h[gk];
      dKm1 = dK;
      dK = dK + 0.5 * (hk1 + hk);
      hk = hk1;
      dRkm1 = dRk;
      dRk = density(P.eos, T[gk], S[gk], 0.);
      dRk = dRk - rhoSurf;
      const double ddRho = dRk - dRkm1;
      if ((mld == 0.) && (ddRho > 0.) && (dRkm1 < P.density_diff) && (dRk >= P.density_diff)) {
        const double aFac = (P.density_diff - dRkm1) / ddRho;
        mld = dK * aFac + dKm1 * (1. - aFac);
      }
    }
    mld = P.stretch * mld;
    if ((mld == 0.) && (dRk < P.density_diff)) mld = dK;  // assume mixing to the bottom
    MLD_fast = mld;
  } else {
    MLD_fast = P.stretch * h_MLD[g];
  }
  if (P.filt1) {
    const double f = fmx(MLD_fast, P.bFac1 * MLD_fast + P.aFac1 * MLD_filtered[g]);
    MLD_filtered[g] = f;
    MLD_fast = f;
  }
  double MLD_slow = MLD_fast;
  if (P.filt2) {
    const double f = fmx(MLD_fast, P.bFac2 * MLD_fast + P.aFac2 * MLD_filtered_slow[g]);
    MLD_filtered_slow[g] = f;
    MLD_slow = f;
  }
  double hf = 0.0, hs = 0.0, rf = 0.0, rs = 0.0;
  for (int k = 0; k < P.nk; ++k) {
    const bool nf = hf < MLD_fast, ns = hs < MLD_slow;
    if (!(nf || ns)) break;  // htot only grows while the test holds: nothing below can contribute (the row's keep_going, per column)
    const long long gk = g + (long long)k * pl;
    const double hk = h[gk];
    const double rho = density(P.eos, T[gk], S[gk], 0.0);
    if (nf) { const double dh = fmn(hk, MLD_fast - hf); rf = rf + dh * rho; hf = hf + dh; }
    if (ns) { const double dh = fmn(hk, MLD_slow - hs); rs = rs + dh * rho; hs = hs + dh; }
  }
  htot_fast[g] = hf; htot_slow[g] = hs;
  Rml_av_fast[g] = -(P.g_Rho0 * rf) / (hf + P.h_neglect);
  Rml_av_slow[g] = -(P.g_Rho0 * rs) / (hs + P.h_neglect);
}

// One velocity face (:464-540 for u, :543-621 for v).  g: plane offset of the face and of its western / southern cell; sd: offset
// to the eastern / northern cell (1 | pitch); qd: offset from the face's upper q point to its other one (pitch for u: (I,J-1);
// 1 for v: (I-1,J)).  len = G%dyCu | G%dxCv, Idn = G%IdxCu | G%IdyCv, dxC / dyC the face's own metrics.
M6M_HD void face(const Par& P, const long long g, const long long sd, const long long qd, const long long pl, const double* h,
                 const double* areaT, const double* ustar, const double* Rd_dx_h, const double* htot_fast, const double* htot_slow,
                 const double* Rml_av_fast, const double* Rml_av_slow, const double* CoriolisBu, const double* maskC, const double* dxC,
                 const double* dyC, const double* len, const double* Idn, double* hml, double* htr) {
  const double u_star = fmx(P.ustar_min, 0.5 * (P.Z_to_H * ustar[g] + P.Z_to_H * ustar[g + sd]));
  const double absf = 0.5 * (fabs(CoriolisBu[g - qd]) + fabs(CoriolisBu[g]));
  double res_scaling_fac = 0.0;
  if (P.res_upscale) {
    const double lfront = 0.5 * (P.front_length + P.front_length);
    double I_LFront = 0.0;
    if (lfront != 0.0) I_LFront = 1.0 / lfront;
    const double dx = dxC[g], dy = dyC[g];
    res_scaling_fac = (sqrt(0.5 * ((dx * dx) + (dy * dy))) * I_LFront) * fmn(1., 0.5 * (Rd_dx_h[g] + Rd_dx_h[g + sd]));
  }
  const double hfs = htot_fast[g] + htot_fast[g + sd], hss = htot_slow[g] + htot_slow[g + sd];
  double h_vel = 0.5 * (hfs + P.h_neglect);
  double mom_mixrate = P.vonKar_x_pi2 * (u_star * u_star) / (absf * (h_vel * h_vel) + 4.0 * (h_vel + P.h_neglect) * u_star);
  double timescale = 0.0625 * (absf + 2.0 * mom_mixrate) / ((absf * absf) + (mom_mixrate * mom_mixrate));
  timescale = timescale * P.coef;
  if (P.res_upscale) timescale = timescale * res_scaling_fac;
  double Dml = timescale * maskC[g] * len[g] * Idn[g] * (Rml_av_fast[g + sd] - Rml_av_fast[g]) * (h_vel * h_vel);
  h_vel = 0.5 * (hss + P.h_neglect);
  mom_mixrate = P.vonKar_x_pi2 * (u_star * u_star) / (absf * (h_vel * h_vel) + 4.0 * (h_vel + P.h_neglect) * u_star);
  timescale = 0.0625 * (absf + 2.0 * mom_mixrate) / ((absf * absf) + (mom_mixrate * mom_mixrate));
  timescale = timescale * P.coef2;
  if (P.res_upscale) timescale = timescale * res_scaling_fac;
  double Dml_slow = timescale * maskC[g] * len[g] * Idn[g] * (Rml_av_slow[g + sd] - Rml_av_slow[g]) * (h_vel * h_vel);
  if (Dml + Dml_slow == 0.) {
    for (int k = 0; k < P.nk; ++k) hml[g + (long long)k * pl] = 0.0;
    return;
  }
  const double IhTot = 2.0 / (hfs + P.h_neglect), IhTot_slow = 2.0 / (hss + P.h_neglect);
  const double tail = P.tail_dh;
  const bool cut = (tail == 0.0);  // mu(sigma <= -1, 0) = +0: the passes may stop at the base of the mixed layer
  const double cA = P.I4dt * areaT[g], cB = P.I4dt * areaT[g + sd];  // h_avail = max(I4dt*areaT*(h - Angstrom_H), 0)  :382
  // pass 1: limit Dml by the volume available on the upwind side of each layer (:527-537)
  double zpa = 0.0, mua = mu(zpa, tail);
  for (int k = 0; k < P.nk; ++k) {
    if (cut && zpa <= -1.0) break;
    const long long gk = g + (long long)k * pl;
    const double h0 = h[gk], h1 = h[gk + sd];
    const double hAtVel = 0.5 * (h0 + h1);
    double a = mua;
    zpa = zpa - (hAtVel * IhTot);
    mua = mu(zpa, tail);
    a = a - mua;
    if (a * Dml > 0.0) { const double av = fmx(cA * (h0 - P.Angstrom_H), 0.0); if (a * Dml > av) Dml = av / a; }
    else if (a * Dml < 0.0) { const double av = fmx(cB * (h1 - P.Angstrom_H), 0.0); if (-a * Dml > av) Dml = -av / a; }
  }
  // pass 2: limit Dml_slow by what Dml leaves (:538-553)
  zpa = 0.0; mua = mu(zpa, tail);
  double zpb = 0.0, mub = mua;
  for (int k = 0; k < P.nk; ++k) {
    if (cut && zpb <= -1.0) break;  // b(k) = +0 from here on: neither test can hold
    const long long gk = g + (long long)k * pl;
    const double h0 = h[gk], h1 = h[gk + sd];
    const double hAtVel = 0.5 * (h0 + h1);
    double a = mua;
    zpa = zpa - (hAtVel * IhTot);
    mua = mu(zpa, tail);
    a = a - mua;
    double b = mub;
    zpb = zpb - (hAtVel * IhTot_slow);
    mub = mu(zpb, tail);
    b = b - mub;
    if (b * Dml_slow > 0.0) {
      const double room = fmx(cA * (h0 - P.Angstrom_H), 0.0) - a * Dml;
      if (b * Dml_slow > room) Dml_slow = fmx(0., room) / b;
    } else if (b * Dml_slow < 0.0) {
      const double room = fmx(cB * (h1 - P.Angstrom_H), 0.0) + a * Dml;
      if (-b * Dml_slow > room) Dml_slow = -fmx(0., room) / b;
    }
  }
  // pass 3: the transports (:554-557)
  zpa = 0.0; zpb = 0.0; mua = mu(zpa, tail); mub = mua;
  int k = 0;
  for (; k < P.nk; ++k) {
    if (cut && zpa <= -1.0 && zpb <= -1.0) break;
    const long long gk = g + (long long)k * pl;
    const double hAtVel = 0.5 * (h[gk] + h[gk + sd]);
    double a = mua;
    zpa = zpa - (hAtVel * IhTot);
    mua = mu(zpa, tail);
    a = a - mua;
    double b = mub;
    zpb = zpb - (hAtVel * IhTot_slow);
    mub = mu(zpb, tail);
    b = b - mub;
    const double t = a * Dml + b * Dml_slow;
    hml[gk] = t;
    htr[gk] = htr[gk] + t * P.dt;
  }
  if (k < P.nk) {
    const double t0 = 0.0 * Dml + 0.0 * Dml_slow;  // a(k) = b(k) = +0 below the mixed layer
    const double t0dt = t0 * P.dt;
    const bool zero = (t0dt == 0.0);  // false only for a non-finite uDml (0*inf): then every layer takes the general update
    const bool fix = !signbit(t0dt);  // x + (+0) = x except for x = -0
    for (; k < P.nk; ++k) {
      const long long gk = g + (long long)k * pl;
      hml[gk] = t0;
      if (!zero) htr[gk] = htr[gk] + t0dt;
      else if (fix) { const double x = htr[gk]; if (x == 0.0 && signbit(x)) htr[gk] = x + t0dt; }
    }
  }
}

// :623-627
M6M_HD void update(const Par& P, const long long g, const long long gk, const long long pitch, const double* uhml, const double* vhml,
                   const double* IareaT, double* h) {
  double hn = h[gk] - P.dt * IareaT[g] * ((uhml[gk] - uhml[gk - 1]) + (vhml[gk] - vhml[gk - pitch]));
  if (hn < P.h_min) hn = P.h_min;
  h[gk] = hn;
}

}  // namespace m6mle
