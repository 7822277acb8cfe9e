// The per-column and per-face work of thickness_diffuse / thickness_diffuse_full (src/parameterizations/lateral/MOM_thickness_diffuse.F90
// :134-1670, the density-gradient path) as host/device code on the unified plane layout (common.cuh): the kernels of thickdiff.cu call it
// with one thread per column / face, tests/harness/thickdiff_host.cpp compiles the same functions with g++ and loops over the tile, so
// the code the GPU threads run is checked bit for bit against the oracle without a GPU (tests/test_thickness_diffuse.py).
// What is arranged differently from the reference, with the same operations on the same operands (hence the same bits):
//  * h_avail(i,j,k) (:869, :876) is evaluated where it is used instead of stored; KH_u / KH_v (:226-443) are evaluated per face;
//  * the two bottom-up loops of a face (the unlimited streamfunction :913-1100, then the limited transports :1124-1176) are one loop:
//    the first never reads what the second writes;
//  * dz, N2_unlim and dzN2_u only feed the FGNV streamfunction and the GM work diagnostics, which are outside the frozen option set.
#pragma once
#include <math.h>
#if defined(__CUDACC__)
#define M6T_HD __host__ __device__ __forceinline__
#else
#define M6T_HD inline
#endif

namespace m6td {

M6T_HD double fmx(double a, double b) { return (a > b) ? a : b; }
M6T_HD double fmn(double a, double b) { return (a < b) ? a : b; }

struct Par {
  int nk, eos_form, Resoln_scaled, have_p_surf;
  double dt, I4dt, Angstrom_H, h_neglect, h_neglect2, dz_neglect, H_to_Z, Z_to_H, g_H_to_RZ /* GV%g_Earth*GV%H_to_RZ */, Z_to_L;
  double Khth, Khth_Min, Khth_Max, max_Khth_CFL, I_slope_max2, kap_dt_x2, h0;
  double dRho_dT, dRho_dS;
  // the OM4-style selection (face_ext): stored slopes, the FGNV elliptic streamfunction, the MEKE diffusivity
  int stored_slopes, FGNV, use_MEKE_Kh;
  double G_rho0, dz_neglect2, N2_floor, FGNV_scale, KhTh_fac;
};
constexpr int EOS_LINEAR = 1;

// calculate_density_derivs: EOS_LINEAR, or the "buggy" Wright fit (MOM_EOS_Wright.F90:178-206)
M6T_HD void density_derivs(const Par& P, double T, double S, double pressure, double& drho_dT, double& drho_dS) {
  if (P.eos_form == EOS_LINEAR) { drho_dT = P.dRho_dT; drho_dS = P.dRho_dS; return; }
  const double a0 = 7.057924e-4, a1 = 3.480336e-7, a2 = -1.112733e-7;
  const double b0 = 5.790749e8, b1 = 3.516535e6, b2 = -4.002714e4, b3 = 2.084372e2, b4 = 5.944068e5, b5 = -9.643486e3;
  const double c0 = 1.704853e5, c1 = 7.904722e2, c2 = -7.984422, c3 = 5.140652e-2, c4 = -2.302158e2, c5 = -3.079464;
  const double al0 = (a0 + a1 * T) + a2 * S;
  const double p0 = (b0 + b4 * S) + T * (b1 + T * ((b2 + b3 * T)) + b5 * S);
  const double lambda = (c0 + c4 * S) + T * (c1 + T * ((c2 + c3 * T)) + c5 * S);
  double I_denom2 = 1.0 / (lambda + al0 * (pressure + p0));
  I_denom2 = I_denom2 * I_denom2;
  drho_dT = I_denom2 * (lambda * (b1 + T * (2.0 * b2 + 3.0 * b3 * T) + b5 * S) -
                        (pressure + p0) * ((pressure + p0) * a1 + (c1 + T * (c2 * 2.0 + c3 * 3.0 * T) + c5 * S)));
  drho_dS = I_denom2 * (lambda * (b4 + b5 * T) - (pressure + p0) * ((pressure + p0) * a2 + (c4 + c5 * T)));
}

// One column (at plane offset g): find_eta (MOM_interface_heights.F90:91-97), the available-volume sums and pressures (:864-883) and
// vert_fill_TS (MOM_isopycnal_slopes.F90:661-697).  e, pres, rsum: nk+1 planes; h_frac, Tf, Sf, c1 (scratch): nk planes.
M6T_HD void column(const Par& P, const long long g, const long long pl, const double* h, const double* T_in, const double* S_in,
                   const double* p_surf, const double* areaT, const double* bathyT, double* e, double* pres, double* rsum, double* h_frac,
                   double* Tf, double* Sf, double* c1) {
  const int nz = P.nk;
  // find_eta
  double ek = -(bathyT[g] + 0.0);
  e[g + (long long)nz * pl] = ek;
  for (int k = nz - 1; k >= 0; --k) { ek = ek + h[g + (long long)k * pl] * P.H_to_Z; e[g + (long long)k * pl] = ek; }
  // h_avail_rsum, h_frac, pres
  const double cA = P.I4dt * areaT[g];
  double rs = 0.0, pr = 0.0;
  if (P.have_p_surf) pr = p_surf[g];
  rsum[g] = rs; pres[g] = pr;
  for (int k = 0; k < nz; ++k) {
    const long long gk = g + (long long)k * pl;
    const double hk = h[gk];
    const double hav = fmx(cA * (hk - P.Angstrom_H), 0.0);
    if (k == 0) { rs = hav; h_frac[gk] = 1.0; }
    else {
      rs = rs + hav;
      double f = 0.0;
      if (hav > 0.0) f = hav / rs;
      h_frac[gk] = f;
    }
    rsum[gk + pl] = rs;
    pr = pr + P.g_H_to_RZ * hk;
    pres[gk + pl] = pr;
  }
  // vert_fill_TS
  if (P.kap_dt_x2 <= 0.0) {
    for (int k = 0; k < nz; ++k) { const long long gk = g + (long long)k * pl; Tf[gk] = T_in[gk]; Sf[gk] = S_in[gk]; }
    return;
  }
  double hk = h[g], hk1 = h[g + pl];
  double ent = P.kap_dt_x2 / ((hk + hk1) + P.h0);  // ent(K=2)
  double h_tr = hk + P.h_neglect;
  double b1 = 1.0 / (h_tr + ent);
  double d1 = b1 * h_tr;
  double Tp = (b1 * h_tr) * T_in[g], Sp = (b1 * h_tr) * S_in[g];
  Tf[g] = Tp; Sf[g] = Sp;
  for (int k = 1; k < nz - 1; ++k) {  // Fortran k = 2 .. nz-1
    const long long gk = g + (long long)k * pl;
    hk = hk1; hk1 = h[gk + pl];
    const double ent1 = P.kap_dt_x2 / ((hk + hk1) + P.h0);  // ent(K+1)
    h_tr = hk + P.h_neglect;
    c1[gk] = ent * b1;
    b1 = 1.0 / ((h_tr + d1 * ent) + ent1);
    d1 = b1 * (h_tr + d1 * ent);
    Tp = b1 * (h_tr * T_in[gk] + ent * Tp);
    Sp = b1 * (h_tr * S_in[gk] + ent * Sp);
    Tf[gk] = Tp; Sf[gk] = Sp;
    ent = ent1;
  }
  {
    const long long gk = g + (long long)(nz - 1) * pl;
    c1[gk] = ent * b1;
    h_tr = h[gk] + P.h_neglect;
    b1 = 1.0 / (h_tr + d1 * ent);
    Tp = b1 * (h_tr * T_in[gk] + ent * Tp);
    Sp = b1 * (h_tr * S_in[gk] + ent * Sp);
    Tf[gk] = Tp; Sf[gk] = Sp;
  }
  for (int k = nz - 2; k >= 0; --k) {
    const long long gk = g + (long long)k * pl;
    const double c = c1[gk + pl];
    Tp = Tf[gk] + c * Tp;
    Sp = Sf[gk] + c * Sp;
    Tf[gk] = Tp; Sf[gk] = Sp;
  }
}

// One velocity face (:913-1229 for u, :1236-1529 for v, and the layer-1 condition :1532-1536).  g: plane offset of the face and of its
// western / southern cell (L); sd: offset to the eastern / northern cell (R).  IdC = G%IdxCu | G%IdyCv (along the face normal),
// lenC = G%dy_Cu | G%dx_Cv, IdxC / IdyC the face's own inverse spacings (for KH_[uv]_CFL), Res_fn = VarMix%Res_fn_u | Res_fn_v.
// Writes hD (the diffusive transport), adds hD*dt to htr, and copies hD to hGM if it is not null.
M6T_HD void face(const Par& P, const long long g, const long long sd, const long long pl, const double* h, const double* e, const double* pres,
                 const double* rsum, const double* h_frac, const double* T, const double* S, const double* areaT, const double* IdC,
                 const double* lenC, const double* IdxC, const double* IdyC, const double* Res_fn, double* hD, double* htr, double* hGM) {
  const int nz = P.nk;
  // the diffusivity of the face :226-305 (the same at every interface)
  const double KH_CFL = (0.25 * P.max_Khth_CFL) / (P.dt * ((IdxC[g] * IdxC[g]) + (IdyC[g] * IdyC[g])));
  double Khth_loc = P.Khth;
  if (P.Resoln_scaled) Khth_loc = Khth_loc * Res_fn[g];
  if (P.Khth_Max > 0) Khth_loc = fmx(P.Khth_Min, fmn(Khth_loc, P.Khth_Max));
  else Khth_loc = fmx(P.Khth_Min, Khth_loc);
  const double KH = fmn(KH_CFL, Khth_loc);
  const double KHlen = KH * lenC[g];
  const double Id = IdC[g];
  const double cL = P.I4dt * areaT[g], cR = P.I4dt * areaT[g + sd];
  const double ebotL = e[g + (long long)nz * pl], ebotR = e[g + sd + (long long)nz * pl];
  double tot = 0.0;  // uhtot | vhtot
  for (int k = nz - 1; k >= 1; --k) {  // Fortran K = k+1 = nz .. 2: the interface between layers k-1 and k (0-based)
    const long long gk = g + (long long)k * pl, gm = gk - pl;
    const double hLk = h[gk], hRk = h[gk + sd], hLm = h[gm], hRm = h[gm + sd];
    const double TLk = T[gk], TRk = T[gk + sd], TLm = T[gm], TRm = T[gm + sd];
    const double SLk = S[gk], SRk = S[gk + sd], SLm = S[gm], SRm = S[gm + sd];
    const double eL = e[gk], eR = e[gk + sd];
    const double pres_f = 0.5 * (pres[gk] + pres[gk + sd]);
    const double T_f = 0.25 * ((TLk + TRk) + (TLm + TRm));
    const double S_f = 0.25 * ((SLk + SRk) + (SLm + SRm));
    double dT, dS;
    density_derivs(P, T_f, S_f, pres_f, dT, dS);
    const double drdiA = dT * (TRm - TLm) + dS * (SRm - SLm);
    const double drdiB = dT * (TRk - TLk) + dS * (SRk - SLk);
    const double drdkL = (dT * (TLk - TLm) + dS * (SLk - SLm));
    const double drdkR = (dT * (TRk - TRm) + dS * (SRk - SRm));
    const double hg2L = hLm * hLk + P.h_neglect2, hg2R = hRm * hRk + P.h_neglect2;
    const double haL = 0.5 * (hLm + hLk) + P.h_neglect, haR = 0.5 * (hRm + hRk) + P.h_neglect;
    const double dzaL = haL * P.H_to_Z, dzaR = haR * P.H_to_Z;
    const double wtL = hg2L * (haR * dzaR), wtR = hg2R * (haL * dzaL);
    const double drdz = ((wtL * drdkL) + (wtR * drdkR)) / ((dzaL * wtL) + (dzaR * wtR));
    const double hg2A = hLm * hRm + P.h_neglect2, hg2B = hLk * hRk + P.h_neglect2;
    const double haA = 0.5 * (hLm + hRm) + P.h_neglect, haB = 0.5 * (hLk + hRk) + P.h_neglect;
    const double wtA = hg2A * haB, wtB = hg2B * haA;
    const double drdx = ((wtA * drdiA + wtB * drdiB) / (wtA + wtB) - drdz * (eL - eR)) * Id;
    const double mag_grad2 = (P.Z_to_L * drdx) * (P.Z_to_L * drdx) + drdz * drdz;
    double Slope, ratio;
    if (mag_grad2 > 0.0) { Slope = drdx / sqrt(mag_grad2); ratio = (Slope * Slope) * P.I_slope_max2; }
    else { Slope = 0.0; ratio = 1.0e20; }
    // int_slope = 0 (:470-473): Slope = (1 - 0)*Slope + 0*(...), ratio = (1 - 0)*ratio
    Slope = (1.0 - 0.0) * Slope + 0.0 * ((eR - eL) * Id);
    ratio = (1.0 - 0.0) * ratio;
    double Sfn = -(KHlen)*Slope;
    if (Sfn > 0.0) {
      if (eL < ebotR) Sfn = 0.0;
      else { const double eLb = e[gk + pl]; if (ebotR > eLb) Sfn = Sfn * ((eL - ebotR) / ((eL - eLb) + P.dz_neglect)); }
    } else {
      if (eR < ebotL) Sfn = 0.0;
      else { const double eRb = e[gk + sd + pl]; if (ebotL > eRb) Sfn = Sfn * ((eR - ebotL) / ((eR - eRb) + P.dz_neglect)); }
    }
    // the limited transport of layer k :1138-1160
    double Sfn_safe;
    if (tot <= 0.0) Sfn_safe = tot * (1.0 - h_frac[gk]);
    else Sfn_safe = tot * (1.0 - h_frac[gk + sd]);
    const double Sfn_est = (P.Z_to_H * Sfn + ratio * Sfn_safe) / (1.0 + ratio);
    const double Sfn_in_H = fmn(fmx(Sfn_est, -rsum[gk]), rsum[gk + sd]);
    const double havL = fmx(cL * (hLk - P.Angstrom_H), 0.0), havR = fmx(cR * (hRk - P.Angstrom_H), 0.0);
    const double t = fmx(fmn((Sfn_in_H - tot), havL), -havR);
    tot = tot + t;
    hD[gk] = t;
    htr[gk] = htr[gk] + t * P.dt;
    if (hGM) hGM[gk] = t;
  }
  const double t1 = -tot;  // :1533-1534
  hD[g] = t1;
  htr[g] = htr[g] + t1 * P.dt;
  if (hGM) hGM[g] = t1;
}

// The same face with USE_STORED_SLOPES (:1024-1026), KHTH_USE_FGNV_STREAMFUNCTION (:980-1022 dzN2, :1103-1122, streamfn_solver :1674-1707) and
// / or the MEKE diffusivity (:281-284).  The elliptic solve needs every interface of the face before the transports can be formed, so
// this is the reference's three-loop structure with the per-interface values in scratch fields (nk+1 planes each: sfn, ratio, hN2, c2,
// c1; shared by the u and v kernels, which run one after the other).  slope: VarMix%slope_x | slope_y (nk+1 planes), maskC = G%OBCmaskCu|Cv.
M6T_HD void face_ext(const Par& P, const long long g, const long long sd, const long long pl, const double* h, const double* e, const double* pres,
                     const double* rsum, const double* h_frac, const double* T, const double* S, const double* areaT, const double* IdC,
                     const double* lenC, const double* IdxC, const double* IdyC, const double* Res_fn, const double* maskC, const double* slope,
                     const double* cg1, const double* MEKE_Kh, double* sfn_s, double* ratio_s, double* hN2_s, double* c2_s, double* c1_s,
                     double* hD, double* htr, double* hGM) {
  const int nz = P.nk;
  const double KH_CFL = (0.25 * P.max_Khth_CFL) / (P.dt * ((IdxC[g] * IdxC[g]) + (IdyC[g] * IdyC[g])));
  double Khth_loc = P.Khth;
  if (P.use_MEKE_Kh) Khth_loc = Khth_loc + P.KhTh_fac * sqrt(MEKE_Kh[g] * MEKE_Kh[g + sd]);
  if (P.Resoln_scaled) Khth_loc = Khth_loc * Res_fn[g];
  if (P.Khth_Max > 0) Khth_loc = fmx(P.Khth_Min, fmn(Khth_loc, P.Khth_Max));
  else Khth_loc = fmx(P.Khth_Min, Khth_loc);
  const double KH = fmn(KH_CFL, Khth_loc);
  const double KHlen = KH * lenC[g];
  const double Id = IdC[g];
  const double cL = P.I4dt * areaT[g], cR = P.I4dt * areaT[g + sd];
  const double ebotL = e[g + (long long)nz * pl], ebotR = e[g + sd + (long long)nz * pl];
  const bool derivs = (!P.stored_slopes) || P.FGNV;
  // ---- loop 1: the slope and the unlimited streamfunction at every interior interface
  for (int k = nz - 1; k >= 1; --k) {
    const long long gk = g + (long long)k * pl, gm = gk - pl;
    const double hLk = h[gk], hRk = h[gk + sd], hLm = h[gm], hRm = h[gm + sd];
    const double eL = e[gk], eR = e[gk + sd];
    double drdiA = 0., drdiB = 0., drdkL = 0., drdkR = 0.;
    if (derivs) {
      const double TLk = T[gk], TRk = T[gk + sd], TLm = T[gm], TRm = T[gm + sd];
      const double SLk = S[gk], SRk = S[gk + sd], SLm = S[gm], SRm = S[gm + sd];
      const double pres_f = 0.5 * (pres[gk] + pres[gk + sd]);
      const double T_f = 0.25 * ((TLk + TRk) + (TLm + TRm));
      const double S_f = 0.25 * ((SLk + SRk) + (SLm + SRm));
      double dT, dS;
      density_derivs(P, T_f, S_f, pres_f, dT, dS);
      drdiA = dT * (TRm - TLm) + dS * (SRm - SLm);
      drdiB = dT * (TRk - TLk) + dS * (SRk - SLk);
      drdkL = (dT * (TLk - TLm) + dS * (SLk - SLm));
      drdkR = (dT * (TRk - TRm) + dS * (SRk - SRm));
    }
    double drdz = 0., hg2A = 0., hg2B = 0., haA = 0., haB = 0.;
    if (derivs) {
      const double hg2L = hLm * hLk + P.h_neglect2, hg2R = hRm * hRk + P.h_neglect2;
      const double haL = 0.5 * (hLm + hLk) + P.h_neglect, haR = 0.5 * (hRm + hRk) + P.h_neglect;
      const double dzaL = haL * P.H_to_Z, dzaR = haR * P.H_to_Z;
      const double wtL = hg2L * (haR * dzaR), wtR = hg2R * (haL * dzaL);
      drdz = ((wtL * drdkL) + (wtR * drdkR)) / ((dzaL * wtL) + (dzaR * wtR));
      hg2A = hLm * hRm + P.h_neglect2; hg2B = hLk * hRk + P.h_neglect2;
      haA = 0.5 * (hLm + hRm) + P.h_neglect; haB = 0.5 * (hLk + hRk) + P.h_neglect;
      const double N2_unlim = drdz * P.G_rho0;
      const double dzL1 = P.H_to_Z * hLm, dzR1 = P.H_to_Z * hRm, dzL0 = P.H_to_Z * hLk, dzR0 = P.H_to_Z * hRk;  // thickness_to_dz, Boussinesq
      const double dzg2A = dzL1 * dzR1 + P.dz_neglect2, dzg2B = dzL0 * dzR0 + P.dz_neglect2;
      const double dzaA = 0.5 * (dzL1 + dzR1) + P.dz_neglect, dzaB = 0.5 * (dzL0 + dzR0) + P.dz_neglect;
      if (P.FGNV) hN2_s[gk] = (0.5 * (dzg2A / dzaA + dzg2B / dzaB)) * fmx(N2_unlim, P.N2_floor);
    }
    double Slope, ratio;
    if (P.stored_slopes) {
      Slope = slope[gk];
      ratio = (Slope * Slope) * P.I_slope_max2;
    } else {
      const double wtA = hg2A * haB, wtB = hg2B * haA;
      const double drdx = ((wtA * drdiA + wtB * drdiB) / (wtA + wtB) - drdz * (eL - eR)) * Id;
      const double mag_grad2 = (P.Z_to_L * drdx) * (P.Z_to_L * drdx) + drdz * drdz;
      if (mag_grad2 > 0.0) { Slope = drdx / sqrt(mag_grad2); ratio = (Slope * Slope) * P.I_slope_max2; }
      else { Slope = 0.0; ratio = 1.0e20; }
    }
    Slope = (1.0 - 0.0) * Slope + 0.0 * ((eR - eL) * Id);  // int_slope = 0 (:470-473)
    ratio = (1.0 - 0.0) * ratio;
    double Sfn = -(KHlen)*Slope;
    if (Sfn > 0.0) {
      if (eL < ebotR) Sfn = 0.0;
      else { const double eLb = e[gk + pl]; if (ebotR > eLb) Sfn = Sfn * ((eL - ebotR) / ((eL - eLb) + P.dz_neglect)); }
    } else {
      if (eR < ebotL) Sfn = 0.0;
      else { const double eRb = e[gk + sd + pl]; if (ebotL > eRb) Sfn = Sfn * ((eR - ebotL) / ((eR - eRb) + P.dz_neglect)); }
    }
    sfn_s[gk] = Sfn; ratio_s[gk] = ratio;
  }
  // ---- loop 2: the elliptic equation of Ferrari et al. (2010) for the streamfunction (plane p holds interface K = p+1 / layer k = p+1)
  if (P.FGNV) {
    if (maskC[g] > 0.) {
      const double cg = 0.5 * (cg1[g] + cg1[g + sd]);
      for (int k = 0; k < nz; ++k) {
        const long long gk = g + (long long)k * pl;
        const double dzL = P.H_to_Z * h[gk], dzR = P.H_to_Z * h[gk + sd];
        const double dz_harm = fmx(P.dz_neglect, 2. * dzL * dzR / ((dzL + dzR) + P.dz_neglect));
        c2_s[gk] = P.FGNV_scale * (cg * cg) / dz_harm;
      }
      for (int k = 1; k < nz; ++k) { const long long gk = g + (long long)k * pl; sfn_s[gk] = (1. + P.FGNV_scale) * sfn_s[gk]; }
      // streamfn_solver(nz, c2_dz, dzN2, Sfn_unlim)
      sfn_s[g] = 0.;
      double b_denom = hN2_s[g + pl] + c2_s[g];
      double beta = 1.0 / (b_denom + c2_s[g + pl]);
      double d1 = beta * b_denom;
      double sprev = (beta * hN2_s[g + pl]) * sfn_s[g + pl];
      sfn_s[g + pl] = sprev;
      for (int p = 2; p < nz; ++p) {  // Fortran K = p+1 = 3 .. nz
        const long long gp = g + (long long)p * pl;
        const double c2m = c2_s[gp - pl];  // c2_h(k-1)
        c1_s[gp - pl] = beta * c2m;        // c1(k-1)
        b_denom = hN2_s[gp] + d1 * c2m;
        beta = 1.0 / (b_denom + c2_s[gp]);
        d1 = beta * b_denom;
        sprev = beta * (hN2_s[gp] * sfn_s[gp] + c2m * sprev);
        sfn_s[gp] = sprev;
      }
      c1_s[g + (long long)(nz - 1) * pl] = beta * c2_s[g + (long long)(nz - 1) * pl];  // c1(nk)
      double snext = 0.;  // sfn(nk+1)
      for (int p = nz - 1; p >= 1; --p) {  // Fortran K = p+1 = nk .. 2
        const long long gp = g + (long long)p * pl;
        snext = sfn_s[gp] + c1_s[gp] * snext;
        sfn_s[gp] = snext;
      }
    } else {
      for (int k = 1; k < nz; ++k) sfn_s[g + (long long)k * pl] = 0.;
    }
  }
  // ---- loop 3: the limited transports :1124-1176 and the layer-1 condition :1533
  double tot = 0.0;
  for (int k = nz - 1; k >= 1; --k) {
    const long long gk = g + (long long)k * pl;
    const double hLk = h[gk], hRk = h[gk + sd];
    const double ratio = ratio_s[gk];
    double Sfn_safe;
    if (tot <= 0.0) Sfn_safe = tot * (1.0 - h_frac[gk]);
    else Sfn_safe = tot * (1.0 - h_frac[gk + sd]);
    const double Sfn_est = (P.Z_to_H * sfn_s[gk] + ratio * Sfn_safe) / (1.0 + ratio);
    const double Sfn_in_H = fmn(fmx(Sfn_est, -rsum[gk]), rsum[gk + sd]);
    const double havL = fmx(cL * (hLk - P.Angstrom_H), 0.0), havR = fmx(cR * (hRk - P.Angstrom_H), 0.0);
    const double t = fmx(fmn((Sfn_in_H - tot), havL), -havR);
    tot = tot + t;
    hD[gk] = t;
    htr[gk] = htr[gk] + t * P.dt;
    if (hGM) hGM[gk] = t;
  }
  const double t1 = -tot;
  hD[g] = t1;
  htr[g] = htr[g] + t1 * P.dt;
  if (hGM) hGM[g] = t1;
}

// :611-615
M6T_HD void update(const Par& P, const long long g, const long long gk, const long long pitch, const double* uhD, const double* vhD,
                   const double* IareaT, double* h) {
  double hn = h[gk] - P.dt * IareaT[g] * ((uhD[gk] - uhD[gk - 1]) + (vhD[gk] - vhD[gk - pitch]));
  if (hn < P.Angstrom_H) hn = P.Angstrom_H;
  h[gk] = hn;
}

}  // namespace m6td
