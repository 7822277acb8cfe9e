// PressureForce_FV_Bouss for sm_100a: two column kernels for the whole routine.
//
// Replaces src/core/MOM_PressureForce_FV.F90:947-2017 with int_density_dz_linear / int_density_dz_wright
// (src/equation_of_state/MOM_EOS_linear.F90:275, MOM_EOS_Wright.F90:389) and Set_pbce_Bouss
// (src/core/MOM_PressureForce_Montgomery.F90:649-748).  The reference materialises 10 3-D temporaries
// (e, pa, dpa, intz_dpa, intx_pa, intx_dpa, inty_pa, inty_dpa, al0/p0/lambda per layer ...) and sweeps them ~15 times.
//
// Design (DESIGN.md "K11"):
//  * pgf_e_kernel: interface heights e(:,:,K), the bottom-up k recursion (:1150-1202), one thread per column.  e is the
//    only 3-D temporary kept (nk+1 planes): the top-down sweep needs e(K), e(K+1) of three neighbouring columns.
//  * pgf_main_kernel: one thread per (i,j) owns the tracer column and the u face to its east and the v face to its north.
//    It marches k top-down carrying pa of the three columns (own, east, north), intx_pa and inty_pa in registers; per
//    layer it evaluates the analytic layer integrals (dpa, intz_dpa) of the three columns, the 3-point Boole quadrature
//    of intx_dpa / inty_dpa with the mass-weighted interpolation, PFu/PFv (:1795-1813), and Set_pbce_Bouss's own-column
//    recursion.  Recomputing the two neighbour columns' dpa (~25 flops + 2 divides each) replaces 6 3-D arrays of
//    traffic and all inter-thread synchronisation; every sum over k is sequential (bitwise parity).
#include "ctx.h"
#include "stage.h"
#include <cmath>
#include <cstdlib>

using m6::Geom;
using m6::fmax2;
using m6::fmin2;

namespace {

constexpr int PGF_MINB_DEFAULT = 4;  // measured at 1440x1080x75: 10.0 ms (2 CTAs/SM), 8.6 (3), 8.45 (4)

// Wright (1997) fit used by EOS_WRIGHT, MOM_EOS_Wright.F90:23-37
#define W_a0 7.057924e-4
#define W_a1 3.480336e-7
#define W_a2 -1.112733e-7
#define W_b0 5.790749e8
#define W_b1 3.516535e6
#define W_b2 -4.002714e4
#define W_b3 2.084372e2
#define W_b4 5.944068e5
#define W_b5 -9.643486e3
#define W_c0 1.704853e5
#define W_c1 7.904722e2
#define W_c2 -7.984422
#define W_c3 5.140652e-2
#define W_c4 -2.302158e2
#define W_c5 -3.079464

struct PgfK {
  int EOS_form, MassWghtInterp, use_SSH_in_Z0p, nk, have_p_atm;
  double rho_ref, rho0_int, rho0_pbce, GxRho_ref, I_g_rho, GFS_scale, Z_ref, dz_neglect, h_neglect;
  double g_Earth, Rho0, H_to_Z, Z_to_H;
  double Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp;
  const double *h, *T, *S, *p_atm, *bathyT, *IdxCu, *IdyCv, *Rlay, *g_prime;
  double *e, *PFu, *PFv, *pbce, *eta;
};

__device__ __forceinline__ double max3(double a, double b, double c) { return fmax2(fmax2(a, b), c); }

__global__ void __launch_bounds__(128) pgf_e_kernel(const Geom G, const PgfK K) {  // :1150-1152, :1200-1202
  const int i = G.isc - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsc - 1 + blockIdx.y;
  if (i > G.iec + 1 || j > G.jec + 1) return;
  const long long g = G.idx(i, j);
  double e = -__ldg(K.bathyT + g);
  K.e[g + (long long)K.nk * G.plane] = e;
  for (int k = K.nk - 1; k >= 0; --k) {
    e = e + __ldg(K.h + g + (long long)k * G.plane) * K.H_to_Z;
    K.e[g + (long long)k * G.plane] = e;
  }
  if (K.eta) K.eta[g] = e * K.Z_to_H;  // :1885-1887
}

// Wright coefficients of one cell
struct WCo { double al0, p0, lambda; };
__device__ __forceinline__ WCo wright_co(double T, double S) {
  WCo c;
  c.al0 = (W_a0 + W_a1 * T) + W_a2 * S;
  c.p0 = (W_b0 + W_b4 * S) + T * (W_b1 + T * ((W_b2 + W_b3 * T)) + W_b5 * S);
  c.lambda = (W_c0 + W_c4 * S) + T * (W_c1 + T * ((W_c2 + W_c3 * T)) + W_c5 * S);
  return c;
}

// one column's layer: inputs and the analytic integrals dpa, intz_dpa (MOM_EOS_Wright.F90:524-548, MOM_EOS_linear.F90:337-345)
struct Lay { double T, S, zt, zb, z0, dpa, intz; WCo w; };

__device__ __forceinline__ void layer_integrals(const PgfK& K, Lay& L, double GxRho, double hk, int k) {
  const double C1_3 = 1.0 / 3.0, C1_7 = 1.0 / 7.0, C1_9 = 1.0 / 9.0, C1_6 = 1.0 / 6.0;
  if (K.EOS_form == MOM6CU_EOS_WRIGHT) {
    L.w = wright_co(L.T, L.S);
    const double dz = L.zt - L.zb;
    const double p_ave = -GxRho * (0.5 * (L.zt + L.zb) - L.z0);
    const double I_al0 = 1.0 / L.w.al0;
    const double I_Lzz = 1.0 / (L.w.p0 + (L.w.lambda * I_al0) + p_ave);
    const double eps = 0.5 * GxRho * dz * I_Lzz, eps2 = eps * eps;
    const double rho_anom = (L.w.p0 + p_ave) * (I_Lzz * I_al0) - K.rho_ref;
    const double rem = (1.0 / K.rho0_int) * (L.w.lambda * (I_al0 * I_al0)) * eps2 * (C1_3 + eps2 * (0.2 + eps2 * (C1_7 + C1_9 * eps2)));
    L.dpa = 1.0 * (K.g_Earth * rho_anom * dz - 2.0 * eps * rem);
    L.intz = 1.0 * (0.5 * K.g_Earth * rho_anom * (dz * dz) - dz * (1.0 + eps) * rem);
  } else if (K.EOS_form == MOM6CU_EOS_LINEAR) {
    const double dz = L.zt - L.zb;
    const double p_ave = -GxRho * (0.5 * (L.zt + L.zb) - L.z0);
    const double rho_anom = (K.Rho_T0_S0 - K.rho_ref) + K.dRho_dT * L.T + K.dRho_dS * L.S + K.dRho_dp * p_ave;
    L.dpa = K.g_Earth * rho_anom * dz;
    L.intz = 0.5 * K.g_Earth * (rho_anom - C1_6 * K.dRho_dp * (GxRho * dz)) * (dz * dz);
  } else {  // no EOS: :1318-1324 (L.T carries dz_geo)
    const double Rlay = K.Rlay[k];
    const double dz_geo = K.g_Earth * K.H_to_Z * hk;
    L.T = dz_geo;
    L.dpa = (Rlay - K.rho_ref) * dz_geo;
    L.intz = 0.5 * (Rlay - K.rho_ref) * dz_geo * hk;
  }
  if (K.EOS_form != MOM6CU_EOS_NONE && K.Z_to_H != 1.0) L.intz = L.intz * K.Z_to_H;  // :1306-1311
}

// intx_dpa / inty_dpa between the columns L (left) and R (right): MOM_EOS_Wright.F90:550-597, MOM_EOS_linear.F90:347-394
__device__ __forceinline__ double face_integral(const PgfK& K, const Lay& L, const Lay& R, double GxRho, double bathyL, double bathyR,
                                                double sshL, double sshR, int k) {
  const double C1_3 = 1.0 / 3.0, C1_7 = 1.0 / 7.0, C1_9 = 1.0 / 9.0, C1_6 = 1.0 / 6.0, C1_90 = 1.0 / 90.0;
  if (K.EOS_form == MOM6CU_EOS_NONE) return 0.5 * (K.Rlay[k] - K.rho_ref) * (L.T + R.T);  // :1326-1333
  double hWght = 0.0;
  if (K.MassWghtInterp & 1) hWght = max3(0., -bathyL - R.zt, -bathyR - L.zt);
  if (K.MassWghtInterp & 2) hWght = max3(hWght, R.zb - sshL, L.zb - sshR);
  double hWt_LL = 1.0, hWt_LR = 0.0, hWt_RR = 1.0, hWt_RL = 0.0;
  if (hWght > 0.) {
    const double hL = (L.zt - L.zb) + K.dz_neglect, hR = (R.zt - R.zb) + K.dz_neglect;
    const double r = (hL - hR) / (hL + hR);
    hWght = hWght * (r * r);
    const double iDenom = 1.0 / (hWght * (hR + hL) + hL * hR);
    hWt_LL = (hWght * hL + hR * hL) * iDenom; hWt_LR = (hWght * hR) * iDenom;
    hWt_RR = (hWght * hR + hR * hL) * iDenom; hWt_RL = (hWght * hL) * iDenom;
  } else if (K.EOS_form == MOM6CU_EOS_LINEAR) {
    const double dzL = L.zt - L.zb, dzR = R.zt - R.zb;
    double p_ave = -GxRho * (0.5 * (L.zt + L.zb) - L.z0);
    const double raL = (K.Rho_T0_S0 - K.rho_ref) + ((K.dRho_dT * L.T + K.dRho_dS * L.S) + K.dRho_dp * p_ave);
    p_ave = -GxRho * (0.5 * (R.zt + R.zb) - R.z0);
    const double raR = (K.Rho_T0_S0 - K.rho_ref) + ((K.dRho_dT * R.T + K.dRho_dS * R.S) + K.dRho_dp * p_ave);
    return K.g_Earth * C1_6 * ((dzL * (2.0 * raL + raR)) + (dzR * (2.0 * raR + raL)));
  }
  double intz[6];
  intz[1] = L.dpa; intz[5] = R.dpa;
#pragma unroll
  for (int m = 2; m <= 4; ++m) {
    const double wt_L = 0.25 * (double)(5 - m), wt_R = 1.0 - wt_L;
    const double wtT_L = (wt_L * hWt_LL) + (wt_R * hWt_RL), wtT_R = (wt_L * hWt_LR) + (wt_R * hWt_RR);
    const double dz = (wt_L * (L.zt - L.zb)) + (wt_R * (R.zt - R.zb));
    const double p_ave = -GxRho * ((wt_L * (0.5 * (L.zt + L.zb) - L.z0)) + (wt_R * (0.5 * (R.zt + R.zb) - R.z0)));
    if (K.EOS_form == MOM6CU_EOS_WRIGHT) {
      const double al0 = (wtT_L * L.w.al0) + (wtT_R * R.w.al0);
      const double p0 = (wtT_L * L.w.p0) + (wtT_R * R.w.p0);
      const double lambda = (wtT_L * L.w.lambda) + (wtT_R * R.w.lambda);
      const double I_al0 = 1.0 / al0;
      const double I_Lzz = 1.0 / (p0 + (lambda * I_al0) + p_ave);
      const double eps = 0.5 * GxRho * dz * I_Lzz, eps2 = eps * eps;
      intz[m] = 1.0 * (K.g_Earth * dz * ((p0 + p_ave) * (I_Lzz * I_al0) - K.rho_ref) -
                       2.0 * eps * (1.0 / K.rho0_int) * (lambda * (I_al0 * I_al0)) * eps2 * (C1_3 + eps2 * (0.2 + eps2 * (C1_7 + C1_9 * eps2))));
    } else {
      const double rho_anom = (K.Rho_T0_S0 - K.rho_ref) + ((K.dRho_dT * ((wtT_L * L.T) + (wtT_R * R.T)) +
                                                            K.dRho_dS * ((wtT_L * L.S) + (wtT_R * R.S))) + K.dRho_dp * p_ave);
      intz[m] = K.g_Earth * rho_anom * dz;
    }
  }
  return C1_90 * (7.0 * (intz[1] + intz[5]) + 32.0 * (intz[2] + intz[4]) + 12.0 * intz[3]);
}

__device__ __forceinline__ double eos_density(const PgfK& K, double T, double S, double p) {
  if (K.EOS_form == MOM6CU_EOS_LINEAR) return K.Rho_T0_S0 + K.dRho_dT * T + K.dRho_dS * S + K.dRho_dp * p;
  const double al0 = (W_a0 + W_a1 * T) + W_a2 * S;
  const double p0 = (W_b0 + W_b4 * S) + T * (W_b1 + T * (W_b2 + W_b3 * T) + W_b5 * S);
  const double lambda = (W_c0 + W_c4 * S) + T * (W_c1 + T * (W_c2 + W_c3 * T) + W_c5 * S);
  return (p + p0) / (lambda + al0 * (p + p0));
}
__device__ __forceinline__ void eos_derivs(const PgfK& K, double T, double S, double p, double& drho_dT, double& drho_dS) {
  if (K.EOS_form == MOM6CU_EOS_LINEAR) { drho_dT = K.dRho_dT; drho_dS = K.dRho_dS; return; }
  const double al0 = (W_a0 + W_a1 * T) + W_a2 * S;
  const double p0 = (W_b0 + W_b4 * S) + T * (W_b1 + T * ((W_b2 + W_b3 * T)) + W_b5 * S);
  const double lambda = (W_c0 + W_c4 * S) + T * (W_c1 + T * ((W_c2 + W_c3 * T)) + W_c5 * S);
  double I_denom2 = 1.0 / (lambda + al0 * (p + p0));
  I_denom2 = I_denom2 * I_denom2;
  drho_dT = I_denom2 * (lambda * (W_b1 + T * (2.0 * W_b2 + 3.0 * W_b3 * T) + W_b5 * S) -
                        (p + p0) * ((p + p0) * W_a1 + (W_c1 + T * (W_c2 * 2.0 + W_c3 * 3.0 * T) + W_c5 * S)));
  drho_dS = I_denom2 * (lambda * (W_b4 + W_b5 * T) - (p + p0) * ((p + p0) * W_a2 + (W_c4 + W_c5 * T)));
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) pgf_main_kernel(const Geom G, const PgfK K) {
  const int i = G.isc - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsc - 1 + blockIdx.y;
  if (i > G.iec + 1 || j > G.jec + 1) return;
  const long long g = G.idx(i, j), P = G.pitch;
  const int nz = K.nk;
  const bool use_EOS = K.EOS_form != MOM6CU_EOS_NONE;
  const bool do_u = (i <= G.iec) && (j >= G.jsc) && (j <= G.jec);  // face I=i, j=js..je
  const bool do_v = (j <= G.jec) && (i >= G.isc) && (i <= G.iec);  // face J=j, i=is..ie
  const bool east_ok = i <= G.iec, north_ok = j <= G.jec;          // neighbour columns inside Isq:Ieq+1 / Jsq:Jeq+1
  const long long gE = east_ok ? g + 1 : g, gN = north_ok ? g + P : g;
  const double GxRho = K.g_Earth * K.rho0_int;
  const double I_Rho0 = 1.0 / K.Rho0, G_Rho0 = K.g_Earth / K.Rho0;
  // surface values of the three columns (:1252-1276)
  const double e1_0 = K.e[g], e1_E = K.e[gE], e1_N = K.e[gN];
  const double pat0 = K.have_p_atm ? __ldg(K.p_atm + g) : 0.0, patE = K.have_p_atm ? __ldg(K.p_atm + gE) : 0.0,
               patN = K.have_p_atm ? __ldg(K.p_atm + gN) : 0.0;
  double pa0, paE, paN;
  if (K.have_p_atm) { pa0 = K.GxRho_ref * (e1_0 - K.Z_ref) + pat0; paE = K.GxRho_ref * (e1_E - K.Z_ref) + patE; paN = K.GxRho_ref * (e1_N - K.Z_ref) + patN; }
  else { pa0 = K.GxRho_ref * (e1_0 - K.Z_ref); paE = K.GxRho_ref * (e1_E - K.Z_ref); paN = K.GxRho_ref * (e1_N - K.Z_ref); }
  double z00, z0E, z0N;
  if (K.use_SSH_in_Z0p && K.have_p_atm) { z00 = e1_0 + pat0 * K.I_g_rho; z0E = e1_E + patE * K.I_g_rho; z0N = e1_N + patN * K.I_g_rho; }
  else if (K.use_SSH_in_Z0p) { z00 = e1_0; z0E = e1_E; z0N = e1_N; }
  else { z00 = K.Z_ref; z0E = K.Z_ref; z0N = K.Z_ref; }
  const double bathy0 = __ldg(K.bathyT + g), bathyE = __ldg(K.bathyT + gE), bathyN = __ldg(K.bathyT + gN);
  const double IdxCu = __ldg(K.IdxCu + g), IdyCv = __ldg(K.IdyCv + g);
  double intx_pa = 0.5 * (pa0 + paE), inty_pa = 0.5 * (pa0 + paN);  // :1538-1544
  // GFS_scale < 1 (:1843-1875)
  double dMx = 0.0, dMy = 0.0;
  if (K.GFS_scale < 1.0) {
    double r0, rE, rN;
    if (use_EOS) {
      r0 = eos_density(K, __ldg(K.T + g), __ldg(K.S + g), pat0); rE = eos_density(K, __ldg(K.T + gE), __ldg(K.S + gE), patE);
      rN = eos_density(K, __ldg(K.T + gN), __ldg(K.S + gN), patN);
    } else { r0 = rE = rN = K.Rlay[0]; }
    const double dM0 = (K.GFS_scale - 1.0) * (G_Rho0 * r0) * (e1_0 - K.Z_ref);
    const double dME = (K.GFS_scale - 1.0) * (G_Rho0 * rE) * (e1_E - K.Z_ref);
    const double dMN = (K.GFS_scale - 1.0) * (G_Rho0 * rN) * (e1_N - K.Z_ref);
    dMx = (dME - dM0) * IdxCu; dMy = (dMN - dM0) * IdyCv;
  }
  // Set_pbce_Bouss own-column state (MOM_PressureForce_Montgomery.F90:685-745)
  const double e_bot = K.e[g + (long long)nz * G.plane];
  const double Rho0xG = K.rho0_pbce * K.g_Earth;
  double Ihtot = 0.0, pbce = 0.0, T_prev = 0.0, S_prev = 0.0;
  if (K.pbce) {
    if (use_EOS) Ihtot = K.H_to_Z / ((e1_0 - e_bot) + K.dz_neglect);
    else Ihtot = 1.0 / ((e1_0 - e_bot) + K.dz_neglect);
  }
  double zt0 = e1_0, ztE = e1_E, ztN = e1_N;
  for (int k = 0; k < nz; ++k) {
    const long long ko = (long long)k * G.plane;
    Lay c, E, N;
    c.zt = zt0; E.zt = ztE; N.zt = ztN;
    c.zb = K.e[g + ko + G.plane]; E.zb = K.e[gE + ko + G.plane]; N.zb = K.e[gN + ko + G.plane];
    c.z0 = z00; E.z0 = z0E; N.z0 = z0N;
    const double h0 = __ldg(K.h + g + ko), hE = __ldg(K.h + gE + ko), hN = __ldg(K.h + gN + ko);
    if (use_EOS) {
      c.T = __ldg(K.T + g + ko); c.S = __ldg(K.S + g + ko);
      E.T = __ldg(K.T + gE + ko); E.S = __ldg(K.S + gE + ko);
      N.T = __ldg(K.T + gN + ko); N.S = __ldg(K.S + gN + ko);
    }
    layer_integrals(K, c, GxRho, h0, k);
    layer_integrals(K, E, GxRho, hE, k);
    layer_integrals(K, N, GxRho, hN, k);
    if (do_u) {
      const double intx_dpa = face_integral(K, c, E, GxRho, bathy0, bathyE, e1_0, e1_E, k);
      double PF = (((pa0 * h0 + c.intz) - (paE * hE + E.intz)) + ((hE - h0) * intx_pa - (E.zb - c.zb) * intx_dpa * K.Z_to_H)) *
                  ((2.0 * I_Rho0 * IdxCu) / ((h0 + hE) + K.h_neglect));
      if (K.GFS_scale < 1.0) PF = PF - dMx;
      K.PFu[g + ko] = PF;
      intx_pa = intx_pa + intx_dpa;
    }
    if (do_v) {
      const double inty_dpa = face_integral(K, c, N, GxRho, bathy0, bathyN, e1_0, e1_N, k);
      double PF = (((pa0 * h0 + c.intz) - (paN * hN + N.intz)) + ((hN - h0) * inty_pa - (N.zb - c.zb) * inty_dpa * K.Z_to_H)) *
                  ((2.0 * I_Rho0 * IdyCv) / ((h0 + hN) + K.h_neglect));
      if (K.GFS_scale < 1.0) PF = PF - dMy;
      K.PFv[g + ko] = PF;
      inty_pa = inty_pa + inty_dpa;
    }
    if (K.pbce) {
      if (use_EOS) {
        const double press = -Rho0xG * (c.zt - K.Z_ref);
        if (k == 0) {
          pbce = G_Rho0 * (K.GFS_scale * eos_density(K, c.T, c.S, press)) * K.H_to_Z;
        } else {
          const double T_int = 0.5 * (T_prev + c.T), S_int = 0.5 * (S_prev + c.S);
          double dR_dT, dR_dS;
          eos_derivs(K, T_int, S_int, press, dR_dT, dR_dS);
          pbce = pbce + G_Rho0 * ((c.zt - e_bot) * Ihtot) * (dR_dT * (c.T - T_prev) + dR_dS * (c.S - S_prev));
        }
        T_prev = c.T; S_prev = c.S;
      } else {
        if (k == 0) pbce = K.g_prime[0] * K.H_to_Z;
        else pbce = pbce + (K.g_prime[k] * K.H_to_Z) * ((c.zt - e_bot) * Ihtot);
      }
      K.pbce[g + ko] = pbce;
    }
    pa0 = pa0 + c.dpa; paE = paE + E.dpa; paN = paN + N.dpa;  // :1340-1345
    zt0 = c.zb; ztE = E.zb; ztN = N.zb;
  }
}

}  // namespace

#include "remap_column.cuh"
#include "pgf_recon.cuh"

int m6_pressure_force_run(mom6cu_ctx* c, const PgfDev& D) {
  if (!c->have_pgf_cs) return c->fail(MOM6CU_ERR_BAD_ARG, "MOM_PressureForce_FV_Bouss: Module must be initialized before it is used.");
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "PressureForce: grid / vertical grid not set");
  const mom6cu_pressureforce_cs& S = c->pgf_cs;
  if (!c->vgrid.Boussinesq) return c->fail(MOM6CU_ERR_UNSUPPORTED, "PressureForce: the non-Boussinesq PGF is outside the frozen option set");
  const bool use_EOS = S.EOS_form != MOM6CU_EOS_NONE;
  if (use_EOS && (!D.T || !D.S)) return c->fail(MOM6CU_ERR_BAD_ARG, "PressureForce: tv%%T and tv%%S are required with an equation of state");
  const mom6cu_domain& d = c->dom;
  if ((d.isc - d.isd) < 1 || (d.jsc - d.jsd) < 1) return c->fail(MOM6CU_ERR_BAD_ARG, "PressureForce needs a halo of at least 1");
  const Geom& G = c->g;
  PgfK K = {};
  K.EOS_form = S.EOS_form; K.MassWghtInterp = S.MassWghtInterp; K.use_SSH_in_Z0p = S.use_SSH_in_Z0p; K.nk = G.nk;
  K.have_p_atm = D.p_atm ? 1 : 0;
  const mom6cu_vgrid& GV = c->vgrid;
  K.rho_ref = S.rho_ref; K.GFS_scale = S.GFS_scale; K.Z_ref = S.Z_ref; K.dz_neglect = S.dZ_subroundoff; K.h_neglect = GV.H_subroundoff;
  K.g_Earth = GV.g_Earth; K.Rho0 = GV.Rho0; K.H_to_Z = GV.H_to_Z; K.Z_to_H = GV.Z_to_H;
  if (S.rho_ref_bug) { K.rho0_int = S.rho_ref; K.rho0_pbce = S.rho_ref; K.GxRho_ref = GV.g_Earth * GV.Rho0; K.I_g_rho = 1.0 / (S.rho_ref * GV.g_Earth); }
  else { K.rho0_int = GV.Rho0; K.rho0_pbce = GV.Rho0; K.GxRho_ref = GV.g_Earth * S.rho_ref; K.I_g_rho = 1.0 / (GV.Rho0 * GV.g_Earth); }
  K.Rho_T0_S0 = S.Rho_T0_S0; K.dRho_dT = S.dRho_dT; K.dRho_dS = S.dRho_dS; K.dRho_dp = S.dRho_dp;
  K.h = D.h; K.T = D.T; K.S = D.S; K.p_atm = D.p_atm; K.bathyT = c->grid.bathyT; K.IdxCu = c->grid.IdxCu; K.IdyCv = c->grid.IdyCv;
  K.Rlay = c->pgf_Rlay; K.g_prime = c->pgf_gprime;
  K.e = c->plane3k("pgf.e", G.nk + 1);
  if (!K.e) return MOM6CU_ERR_CUDA;
  K.PFu = D.PFu; K.PFv = D.PFv; K.pbce = D.pbce; K.eta = D.eta;
  dim3 grid((d.iec - d.isc + 3 + 127) / 128, d.jec - d.jsc + 3);
  M6_LAUNCH(c, pgf_e_kernel, grid, 128, 0, G, K);
  if (S.reconstruct && use_EOS && S.Recon_Scheme > 0) {  // use_ALE (:1120-1122) with PRESSURE_RECONSTRUCTION_SCHEME 1 (PLM) or 2 (PPM)
    PgfRecon R = {};
    R.scheme = S.Recon_Scheme; R.boundary_extrap = S.boundary_extrap; R.inaccurate = S.use_inaccurate_pgf_rho_anom;
    R.van_only = S.MassWghtInterpVanOnly; R.h_nv = GV.H_to_Z * S.h_nonvanished; R.H_subroundoff = GV.H_subroundoff;
    R.T_t = c->plane3k("pgf.T_t", G.nk); R.T_b = c->plane3k("pgf.T_b", G.nk);
    R.S_t = c->plane3k("pgf.S_t", G.nk); R.S_b = c->plane3k("pgf.S_b", G.nk);
    if (!R.T_t || !R.T_b || !R.S_t || !R.S_b) return MOM6CU_ERR_CUDA;
    const dim3 ge(grid.x, grid.y, 2);
    static int edges_array = -1;   // MOM6CU_PGF_EDGES_ARRAY=1: the array form of the PLM edge values (testing; PPM always uses it)
    if (edges_array < 0) { const char* e = getenv("MOM6CU_PGF_EDGES_ARRAY"); edges_array = (e && atoi(e)) ? 1 : 0; }
    if (S.Recon_Scheme == 1 && !edges_array) M6_LAUNCH(c, pgf_ts_edges_plm_kernel, grid, 128, 0, G, K, R);
    else if (G.nk <= 40) M6_LAUNCH(c, pgf_ts_edges_kernel<40>, ge, 128, 0, G, K, R);
    else if (G.nk <= 80) M6_LAUNCH(c, pgf_ts_edges_kernel<80>, ge, 128, 0, G, K, R);
    else M6_LAUNCH(c, pgf_ts_edges_kernel<128>, ge, 128, 0, G, K, R);
    constexpr int TX = 32, TY = 8;
    const dim3 gr((d.iec - d.isc + 2 + TX - 2) / (TX - 1), (d.jec - d.jsc + 2 + TY - 2) / (TY - 1));
    static int var = -1;
    if (var < 0) { const char* e = getenv("MOM6CU_PGF_VAR"); var = e ? atoi(e) : 0; }
    if (S.Recon_Scheme == 2) M6_LAUNCH(c, (pgf_recon_kernel<true, TX, TY>), gr, TX * TY, 0, G, K, R);
    else if (var == 1) M6_LAUNCH(c, (pgf_recon_kernel<false, TX, TY, 1>), gr, TX * TY, 0, G, K, R);
    else if (var == 2) M6_LAUNCH(c, (pgf_recon_kernel<false, TX, TY, 2>), gr, TX * TY, 0, G, K, R);
    else if (var == 3) M6_LAUNCH(c, (pgf_recon_kernel<false, TX, TY, 3>), gr, TX * TY, 0, G, K, R);
    else M6_LAUNCH(c, (pgf_recon_kernel<false, TX, TY>), gr, TX * TY, 0, G, K, R);
    M6_CUDA(c, cudaGetLastError());
    return 0;
  }
  {  // resident CTAs per SM: 2 (194 registers, no spills), 3 (168) or 4 (128, ~30 doubles spilled); MOM6CU_PGF_MINB overrides
    static int minb = -1;
    if (minb < 0) { const char* e = getenv("MOM6CU_PGF_MINB"); minb = e ? atoi(e) : PGF_MINB_DEFAULT; }
    if (minb >= 4) M6_LAUNCH(c, pgf_main_kernel<4>, grid, 128, 0, G, K);
    else if (minb == 3) M6_LAUNCH(c, pgf_main_kernel<3>, grid, 128, 0, G, K);
    else M6_LAUNCH(c, pgf_main_kernel<1>, grid, 128, 0, G, K);
  }
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

extern "C" int mom6cu_set_cs_pressureforce(mom6cu_ctx* c, const mom6cu_pressureforce_cs* CS) {
  if (!c || !CS) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (CS->unsupported)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "PressureForce_init: tides/SAL, Stanley SGS, intxpa resets/corrections, bulk mixed layers and "
                                           "non-Boussinesq dynamics are outside the frozen option set of this build");
  if (CS->EOS_form != MOM6CU_EOS_NONE && CS->EOS_form != MOM6CU_EOS_LINEAR && CS->EOS_form != MOM6CU_EOS_WRIGHT)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "PressureForce: No analytic integration option is available with this EOS!");
  if (CS->reconstruct && CS->EOS_form != MOM6CU_EOS_NONE && CS->Recon_Scheme > 0) {
    if (CS->Recon_Scheme > 2) return c->fail(MOM6CU_ERR_BAD_ARG, "PressureForce: PRESSURE_RECONSTRUCTION_SCHEME must be 1 (PLM) or 2 (PPM)");
    if (CS->ALE_answer_date < 20190101)
      return c->fail(MOM6CU_ERR_UNSUPPORTED, "PressureForce: ALE answer_date %d < 20190101 is not implemented", CS->ALE_answer_date);
    if (c->g.nk > 128) return c->fail(MOM6CU_ERR_UNSUPPORTED, "PressureForce: %d levels exceed the 128-level column capacity of the T,S reconstruction", c->g.nk);
    if (c->g.nk < (CS->Recon_Scheme == 2 ? 4 : 2)) return c->fail(MOM6CU_ERR_BAD_ARG, "PressureForce: too few layers for the T,S reconstruction");
  }
  {  // EOS_type unit conversion factors: the device path is the unscaled one (0 is read as 1)
    const double sc[4] = {CS->kg_m3_to_R, CS->RL2_T2_to_Pa, CS->C_to_degC, CS->S_to_ppt};
    for (double v : sc)
      if (v != 0.0 && v != 1.0) return c->fail(MOM6CU_ERR_UNSUPPORTED, "PressureForce: rescaled EOS units (EOS%%kg_m3_to_R etc. /= 1) are not implemented on the device");
  }
  c->pgf_cs = *CS;
  c->pgf_Rlay = nullptr; c->pgf_gprime = nullptr;
  if (CS->EOS_form == MOM6CU_EOS_NONE) {
    if (!CS->Rlay || !CS->g_prime) return c->fail(MOM6CU_ERR_BAD_ARG, "PressureForce: GV%%Rlay and GV%%g_prime are required without an EOS");
    double* p = c->buf("PGF.Rlay", (size_t)c->g.nk);
    double* q = c->buf("PGF.g_prime", (size_t)c->g.nk + 1);
    if (!p || !q) return MOM6CU_ERR_CUDA;
    M6_CUDA(c, cudaMemcpyAsync(p, CS->Rlay, sizeof(double) * c->g.nk, cudaMemcpyDefault, c->stream));
    M6_CUDA(c, cudaMemcpyAsync(q, CS->g_prime, sizeof(double) * (c->g.nk + 1), cudaMemcpyDefault, c->stream));
    M6_CUDA(c, cudaStreamSynchronize(c->stream));
    c->pgf_Rlay = p; c->pgf_gprime = q;
  }
  c->have_pgf_cs = true;
  return 0;
}

extern "C" int mom6cu_pressure_force(mom6cu_ctx* c, const mom6cu_pressureforce_args* a) {
  if (!c || !a) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!a->h || !a->PFu || !a->PFv) return c->fail(MOM6CU_ERR_BAD_ARG, "PressureForce: null required argument");
  Stager S(c, "pgf.");
  PgfDev D = {};
  int rc;
  if ((rc = S.in3(a->h, ST_H, "h", &D.h)) || (rc = S.in3(a->T, ST_H, "T", &D.T)) || (rc = S.in3(a->S, ST_H, "S", &D.S)) ||
      (rc = S.in2(a->p_atm, ST_H, "p_atm", &D.p_atm)) || (rc = S.io3(a->PFu, ST_U, "PFu", &D.PFu)) ||
      (rc = S.io3(a->PFv, ST_V, "PFv", &D.PFv)) || (rc = S.io3(a->pbce, ST_H, "pbce", &D.pbce)) || (rc = S.io2(a->eta, ST_H, "eta", &D.eta)))
    return rc;
  if ((rc = S.begin())) return rc;
  if ((rc = m6_pressure_force_run(c, D))) return rc;
  return S.finish();
}
