// RECONSTRUCT_FOR_PRESSURE for PressureForce_FV_Bouss on sm_100a (included by pgf.cu).
//
// Replaces, for the default ALE configuration of the reference (src/core/MOM_PressureForce_FV.F90:2172-2190):
//   TS_PLM_edge_values / TS_PPM_edge_values      src/ALE/MOM_ALE.F90:1495-1660
//   int_density_dz_generic_plm / _generic_ppm    src/core/MOM_density_integrals.F90:418-868 / :874-1310
//   and the rest of the routine (:1150-1276 surface values, :1340-1345 pa, :1538-1558 intx_pa / inty_pa, :1793-1812 PFu / PFv,
//   :1843-1875 GFS_scale, Set_pbce_Bouss MOM_PressureForce_Montgomery.F90:685-745) as in pgf.cu.
//
// Design:
//  * pgf_ts_edges_kernel: one thread per (column, field): the PLM / PPM reconstruction of the column with the ALE column code the
//    remapping kernels use (remap_column.cuh, bit-exact with the reference's PLM_functions / PPM_functions / regrid_edge_values),
//    writing the top / bottom edge values T_t, T_b, S_t, S_b (4 scratch 3-D fields).
//  * pgf_recon_kernel: a CTA owns a TX x TY patch of tracer columns and marches the layers top-down.  Per layer every thread evaluates
//    its own column's five-point vertical quadrature (dpa, intz_dpa: 5 equation-of-state evaluations) and publishes the column's layer
//    data (interface heights, edge values, dpa, intz_dpa, pa, h) in shared memory; after ONE barrier (the planes are double-buffered)
//    it evaluates the 3 x 5 point quadratures of the face to its east and of the face to its north from its neighbours' shared data
//    (30 equation-of-state evaluations) and the accelerations PFu, PFv.  The patches overlap by one column / row (the last thread column
//    and row only publish), so no column integral is exchanged between CTAs and none is recomputed per face.
//  * Everything is evaluated with the reference's own expression order (bitwise parity, -fmad=false); the k-recursions for pa, intx_pa,
//    inty_pa and pbce stay sequential in registers.
//  * Roofline: fp64 pipe.  ~35 density evaluations (one IEEE division each with the Wright form) + the interpolation arithmetic per cell,
//    ~2.9 k fp64 instructions per cell (ncu), against 3 + 4 + 3 doubles of compulsory traffic: ~36 instructions per byte, far to the right of the B200 ridge
//    for fp64 (~5 flop/byte), so the kernel is measured against the fp64 issue rate, not HBM.
#pragma once

namespace {

struct PgfRecon {
  int scheme, boundary_extrap, inaccurate, van_only;
  double h_nv;  // GV%H_to_Z * CS%h_nonvanished
  double H_subroundoff;
  double *T_t, *T_b, *S_t, *S_b;
};

template <int KCAP>
__global__ void __launch_bounds__(128) pgf_ts_edges_kernel(const Geom G, const PgfK K, const PgfRecon R) {
  const int i = G.isc - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsc - 1 + blockIdx.y, f = blockIdx.z;
  if (i > G.iec + 1 || j > G.jec + 1) return;
  const long long g = G.idx(i, j), pl = G.plane;
  const int nk = K.nk;
  const double* Q = (f == 0) ? K.S : K.T;
  double* Qt = (f == 0) ? R.S_t : R.T_t;
  double* Qb = (f == 0) ? R.S_b : R.T_b;
  double h[KCAP + 1];
  m6remap::Recon<KCAP> C;
  for (int k = 1; k <= nk; ++k) { h[k] = __ldg(K.h + g + (long long)(k - 1) * pl); C.u[k] = __ldg(Q + g + (long long)(k - 1) * pl); }
  if (R.scheme == 1) {
    m6remap::PLM_reconstruction<KCAP>(nk, h, C, R.H_subroundoff, R.boundary_extrap != 0);
  } else {
    m6remap::edge_values_implicit_h4<KCAP>(nk, h, C, R.H_subroundoff);
    m6remap::PPM_reconstruction<KCAP>(nk, h, C, R.H_subroundoff, R.boundary_extrap != 0);
  }
  for (int k = 1; k <= nk; ++k) { Qt[g + (long long)(k - 1) * pl] = C.E1[k]; Qb[g + (long long)(k - 1) * pl] = C.E2[k]; }
}

// TS_PLM_edge_values (ALE_PLM_edge_values, MOM_ALE.F90:1518-1576) as a streaming sweep: one thread per column marches the layers once with the
// three-layer windows of h, T, S and the slopes slp(k-1), slp(k), slp(k+1) in registers -- no thread-local column arrays (the array form above
// took 4.6 ms at 1440x1080x75 in local memory; this one reads h, T, S once and writes the four edge fields once: 56 B/cell).  The arithmetic
// is PLM_slope_wa / PLM_monotonized_slope / PLM_extrapolate_slope of remap_column.cuh, the same functions in the same order.
__global__ void __launch_bounds__(128) pgf_ts_edges_plm_kernel(const Geom G, const PgfK K, const PgfRecon R) {
  const int i = G.isc - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsc - 1 + blockIdx.y;
  if (i > G.iec + 1 || j > G.jec + 1) return;
  const long long g = G.idx(i, j), pl = G.plane;
  const int nk = K.nk;
  const double hneg = R.H_subroundoff;
  const bool ext = R.boundary_extrap != 0;
  auto H = [&](int k) { return __ldg(K.h + g + (long long)(k - 1) * pl); };
  auto Tq = [&](int k) { return __ldg(K.T + g + (long long)(k - 1) * pl); };
  auto Sq = [&](int k) { return __ldg(K.S + g + (long long)(k - 1) * pl); };
  auto put = [&](int k, double Tc, double mT, double Sc, double mS) {
    const long long o = g + (long long)(k - 1) * pl;
    R.T_t[o] = Tc - 0.5 * mT; R.T_b[o] = Tc + 0.5 * mT;
    R.S_t[o] = Sc - 0.5 * mS; R.S_b[o] = Sc + 0.5 * mS;
  };
  double hm = H(1), hc = H(2), Tm = Tq(1), Tc = Tq(2), Sm = Sq(1), Sc = Sq(2);
  if (ext) put(1, Tm, -m6remap::PLM_extrapolate_slope(hc, hm, hneg, Tc, Tm), Sm, -m6remap::PLM_extrapolate_slope(hc, hm, hneg, Sc, Sm));
  else { const long long o = g; R.T_t[o] = Tm; R.T_b[o] = Tm; R.S_t[o] = Sm; R.S_b[o] = Sm; }
  double hp = 0., Tp = 0., Sp = 0.;
  double sT_m = 0., sS_m = 0., sT_c = 0., sS_c = 0.;   // slp(k-1), slp(k)
  if (nk >= 3) {
    hp = H(3); Tp = Tq(3); Sp = Sq(3);
    sT_c = m6remap::PLM_slope_wa(hm, hc, hp, hneg, Tm, Tc, Tp);
    sS_c = m6remap::PLM_slope_wa(hm, hc, hp, hneg, Sm, Sc, Sp);
  }
  for (int k = 2; k <= nk - 1; ++k) {
    double hpp = 0., Tpp = 0., Spp = 0., sT_p = 0., sS_p = 0.;
    if (k + 1 <= nk - 1) {
      hpp = H(k + 2); Tpp = Tq(k + 2); Spp = Sq(k + 2);
      sT_p = m6remap::PLM_slope_wa(hc, hp, hpp, hneg, Tc, Tp, Tpp);
      sS_p = m6remap::PLM_slope_wa(hc, hp, hpp, hneg, Sc, Sp, Spp);
    }
    put(k, Tc, m6remap::PLM_monotonized_slope(Tm, Tc, Tp, sT_m, sT_c, sT_p), Sc, m6remap::PLM_monotonized_slope(Sm, Sc, Sp, sS_m, sS_c, sS_p));
    hm = hc; hc = hp; hp = hpp; Tm = Tc; Tc = Tp; Tp = Tpp; Sm = Sc; Sc = Sp; Sp = Spp;
    sT_m = sT_c; sT_c = sT_p; sS_m = sS_c; sS_c = sS_p;
  }
  // layer nk: (hm, hc) = (h(nk-1), h(nk))
  if (ext) put(nk, Tc, m6remap::PLM_extrapolate_slope(hm, hc, hneg, Tm, Tc), Sc, m6remap::PLM_extrapolate_slope(hm, hc, hneg, Sm, Sc));
  else { const long long o = g + (long long)(nk - 1) * pl; R.T_t[o] = Tc; R.T_b[o] = Tc; R.S_t[o] = Sc; R.S_b[o] = Sc; }
}

// calculate_density with / without rho_ref for an unscaled EOS (MOM_EOS.F90:332-334): density_anomaly_elem / density_elem of
// EOS_WRIGHT (MOM_EOS_Wright.F90:102-130, :80-97) and EOS_LINEAR (MOM_EOS_linear.F90:74-84, :60-68).
struct RhoFn {
  int form, use_ref;
  double rho_ref, pa_000, Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp;
  __device__ __forceinline__ double operator()(double T, double S, double pressure) const {
    if (form == MOM6CU_EOS_LINEAR) {
      if (use_ref) return (Rho_T0_S0 - rho_ref) + ((dRho_dT * T + dRho_dS * S) + dRho_dp * pressure);
      return Rho_T0_S0 + dRho_dT * T + dRho_dS * S + dRho_dp * pressure;
    }
    if (use_ref) {
      const double al_TS = W_a1 * T + W_a2 * S;
      const double al0 = W_a0 + al_TS;
      const double p_TSp = pressure + (W_b4 * S + T * (W_b1 + (T * (W_b2 + W_b3 * T) + W_b5 * S)));
      const double lam_TS = W_c4 * S + T * (W_c1 + (T * (W_c2 + W_c3 * T) + W_c5 * S));
      return (pa_000 + (p_TSp - rho_ref * (p_TSp * al0 + (W_b0 * al_TS + lam_TS)))) / ((W_c0 + lam_TS) + al0 * (W_b0 + p_TSp));
    }
    const double al0 = (W_a0 + W_a1 * T) + W_a2 * S;
    const double p0 = (W_b0 + W_b4 * S) + T * (W_b1 + T * (W_b2 + W_b3 * T) + W_b5 * S);
    const double lambda = (W_c0 + W_c4 * S) + T * (W_c1 + T * (W_c2 + W_c3 * T) + W_c5 * S);
    return (pressure + p0) / (lambda + al0 * (pressure + p0));
  }
};

// what a column publishes per layer
enum { Q_ET = 0, Q_EB, Q_TT, Q_TB, Q_ST, Q_SB, Q_TM, Q_SM, Q_DPA, Q_INTZ, Q_PA, Q_H, Q_N };

// Inlined at both call sites with the sub-column loop unrolled: the out-of-line variant (one copy of the 15 evaluations instead of 30, to
// relieve the instruction cache) measured 34.5 ms against 30.6 ms at 1440x1080x75 -- the call overhead and the lost overlap between
// sub-columns cost more than the instruction-cache misses it removed.
template <bool PPM, int MUNR>
__device__ __forceinline__ double recon_face_integral(const PgfK& K, const PgfRecon& R, const RhoFn& rho, double GxRho, const double* L,
                                                      const double* Rt, int sq /* stride between quantities */, double bathyL, double bathyR,
                                                      double e1L, double e1R, double z0L, double z0R) {
  const double C1_90 = 1.0 / 90.0;
  const double eLt = L[Q_ET * sq], eLb = L[Q_EB * sq], eRt = Rt[Q_ET * sq], eRb = Rt[Q_EB * sq];
  const double massWeightToggle = (K.MassWghtInterp & 1) ? 1. : 0., TopWeightToggle = (K.MassWghtInterp & 2) ? 1. : 0.;
  const double massWeightNVonlyToggle = R.van_only ? 0. : 1.;
  double hWght = massWeightToggle * max3(0., -bathyL - eRt, -bathyR - eLt);
  const double hWghtTop = TopWeightToggle * max3(0., eRb - e1L, eLb - e1R);
  hWght = fmax2(hWght, hWghtTop);
  if (((eLt - eLb) > R.h_nv) && ((eRt - eRb) > R.h_nv)) hWght = massWeightNVonlyToggle * hWght;
  const double TtL = L[Q_TT * sq], TbL = L[Q_TB * sq], StL = L[Q_ST * sq], SbL = L[Q_SB * sq];
  const double TtR = Rt[Q_TT * sq], TbR = Rt[Q_TB * sq], StR = Rt[Q_ST * sq], SbR = Rt[Q_SB * sq];
  double Ttl, Tbl, Tml = 0., Ttr, Tbr, Tmr = 0., Stl, Sbl, Sml = 0., Str, Sbr, Smr = 0.;
  if (hWght > 0.) {
    const double hL = (eLt - eLb) + K.dz_neglect;
    const double hR = (eRt - eRb) + K.dz_neglect;
    const double r = (hL - hR) / (hL + hR);
    hWght = hWght * (r * r);
    const double iDenom = 1. / (hWght * (hR + hL) + hL * hR);
    Ttl = ((hWght * hR) * TtR + (hWght * hL + hR * hL) * TtL) * iDenom;
    Ttr = ((hWght * hL) * TtL + (hWght * hR + hR * hL) * TtR) * iDenom;
    Tbl = ((hWght * hR) * TbR + (hWght * hL + hR * hL) * TbL) * iDenom;
    Tbr = ((hWght * hL) * TbL + (hWght * hR + hR * hL) * TbR) * iDenom;
    Stl = ((hWght * hR) * StR + (hWght * hL + hR * hL) * StL) * iDenom;
    Str = ((hWght * hL) * StL + (hWght * hR + hR * hL) * StR) * iDenom;
    Sbl = ((hWght * hR) * SbR + (hWght * hL + hR * hL) * SbL) * iDenom;
    Sbr = ((hWght * hL) * SbL + (hWght * hR + hR * hL) * SbR) * iDenom;
    if (PPM) {
      const double TmL = L[Q_TM * sq], SmL = L[Q_SM * sq], TmR = Rt[Q_TM * sq], SmR = Rt[Q_SM * sq];
      Tml = ((hWght * hR) * TmR + (hWght * hL + hR * hL) * TmL) * iDenom;
      Tmr = ((hWght * hL) * TmL + (hWght * hR + hR * hL) * TmR) * iDenom;
      Sml = ((hWght * hR) * SmR + (hWght * hL + hR * hL) * SmL) * iDenom;
      Smr = ((hWght * hL) * SmL + (hWght * hR + hR * hL) * SmR) * iDenom;
    }
  } else {
    Ttl = TtL; Tbl = TbL; Ttr = TtR; Tbr = TbR;
    Stl = StL; Sbl = SbL; Str = StR; Sbr = SbR;
    if (PPM) { Tml = L[Q_TM * sq]; Tmr = Rt[Q_TM * sq]; Sml = L[Q_SM * sq]; Smr = Rt[Q_SM * sq]; }
  }
  double intz[6];
  intz[1] = L[Q_DPA * sq]; intz[5] = Rt[Q_DPA * sq];
  double i2 = 0., i3 = 0., i4 = 0.;
#pragma unroll MUNR
  for (int m = 2; m <= 4; ++m) {
    const double w_left = 0.25 * (double)(5 - m), w_right = 1.0 - w_left;
    const double dz_x = (w_left * (eLt - eLb)) + (w_right * (eRt - eRb));
    double p15 = -GxRho * ((w_left * (eLt - z0L)) + (w_right * (eRt - z0R)));
    double T_top, T_bot, S_top, S_bot, s6 = 0., t6 = 0.;
    T_top = (w_left * Ttl) + (w_right * Ttr); T_bot = (w_left * Tbl) + (w_right * Tbr);
    S_top = (w_left * Stl) + (w_right * Str); S_bot = (w_left * Sbl) + (w_right * Sbr);
    if (PPM) {
      const double T_mn = (w_left * Tml) + (w_right * Tmr), S_mn = (w_left * Sml) + (w_right * Smr);
      s6 = 3.0 * (2.0 * S_mn - (S_top + S_bot));
      t6 = 3.0 * (2.0 * T_mn - (T_top + T_bot));
    }
    double r15[6];
#pragma unroll
    for (int n = 1; n <= 5; ++n) {
      const double wt_t = 0.25 * (double)(5 - n), wt_b = 1.0 - wt_t;
      double Sn, Tn;
      if (PPM) {
        Sn = wt_t * S_top + wt_b * (S_bot + s6 * wt_t);
        Tn = wt_t * T_top + wt_b * (T_bot + t6 * wt_t);
      } else if (n == 1) { Sn = S_top; Tn = T_top; }
      else if (n == 5) { Sn = S_bot; Tn = T_bot; }
      else {
        Sn = wt_t * S_top + wt_b * S_bot;
        Tn = wt_t * T_top + wt_b * T_bot;
      }
      if (n > 1) p15 = p15 + GxRho * 0.25 * dz_x;
      r15[n] = rho(Tn, Sn, p15);
    }
    double iz;
    if (rho.use_ref) iz = (K.g_Earth * dz_x * (C1_90 * (7.0 * (r15[1] + r15[5]) + 32.0 * (r15[2] + r15[4]) + 12.0 * r15[3])));
    else iz = (K.g_Earth * dz_x * (C1_90 * (7.0 * (r15[1] + r15[5]) + 32.0 * (r15[2] + r15[4]) + 12.0 * r15[3]) - rho.rho_ref));
    if (m == 2) i2 = iz; else if (m == 3) i3 = iz; else i4 = iz;
  }
  return C1_90 * (7.0 * (intz[1] + intz[5]) + 32.0 * (i2 + i4) + 12.0 * i3);
}

// VAR (A/B measurements, MOM6CU_PGF_VAR): bit 0 = the sub-column loop of the face integrals not unrolled (3 x less code), bit 1 = compiled for
// 3 CTAs/SM (<= 85 registers).
template <bool PPM, int TX, int TY, int VAR = 0>
__global__ void __launch_bounds__(TX* TY, (VAR & 2) ? 3 : ((TX * TY <= 256) ? 2 : 1)) pgf_recon_kernel(const Geom G, const PgfK K, const PgfRecon R) {
  constexpr int NT = TX * TY;
  __shared__ double sm[2][Q_N][NT];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX, t = threadIdx.x;
  // patches overlap by one column / row: thread (tx, ty) owns the column (i, j); its east / north faces exist when tx < TX-1 / ty < TY-1
  const int i = G.isc - 1 + blockIdx.x * (TX - 1) + tx, j = G.jsc - 1 + blockIdx.y * (TY - 1) + ty;
  const bool in = (i <= G.iec + 1) && (j <= G.jec + 1);
  const int ic = in ? i : G.iec + 1, jc = (j <= G.jec + 1) ? j : G.jec + 1;
  const long long g = G.idx(i <= G.iec + 1 ? i : G.iec + 1, jc), pl = G.plane;
  (void)ic;
  const int nz = K.nk;
  const bool do_u = in && (tx < TX - 1) && (i <= G.iec) && (j >= G.jsc) && (j <= G.jec);
  const bool do_v = in && (ty < TY - 1) && (j <= G.jec) && (i >= G.isc) && (i <= G.iec);
  const int tE = t + 1, tN = t + TX;
  const double GxRho = K.g_Earth * K.rho0_int;
  const double I_Rho0 = 1.0 / K.Rho0, G_Rho0 = K.g_Earth / K.Rho0;
  const double C1_90 = 1.0 / 90.0;
  RhoFn rho;
  rho.form = K.EOS_form; rho.use_ref = (PPM || !R.inaccurate) ? 1 : 0; rho.rho_ref = K.rho_ref;
  rho.pa_000 = (W_b0 * (1.0 - W_a0 * K.rho_ref) - K.rho_ref * W_c0);
  rho.Rho_T0_S0 = K.Rho_T0_S0; rho.dRho_dT = K.dRho_dT; rho.dRho_dS = K.dRho_dS; rho.dRho_dp = K.dRho_dp;
  // surface values (:1252-1276); the neighbours' through shared memory
  const double e1 = K.e[g];
  const double pat = K.have_p_atm ? __ldg(K.p_atm + g) : 0.0;
  double pa = K.have_p_atm ? (K.GxRho_ref * (e1 - K.Z_ref) + pat) : (K.GxRho_ref * (e1 - K.Z_ref));
  double z0;
  if (K.use_SSH_in_Z0p && K.have_p_atm) z0 = e1 + pat * K.I_g_rho;
  else if (K.use_SSH_in_Z0p) z0 = e1;
  else z0 = K.Z_ref;
  const double bathy = __ldg(K.bathyT + g);
  const double IdxCu = __ldg(K.IdxCu + g), IdyCv = __ldg(K.IdyCv + g);
  // per-column constants of the neighbours: (e1, z0, bathy, dM) staged once in buffer 1, pa of the surface in buffer 0
  double dM = 0.0;
  if (K.GFS_scale < 1.0) {  // :1843-1875
    const double r0 = eos_density(K, __ldg(K.T + g), __ldg(K.S + g), pat);
    dM = (K.GFS_scale - 1.0) * (G_Rho0 * r0) * (e1 - K.Z_ref);
  }
  sm[1][0][t] = e1; sm[1][1][t] = z0; sm[1][2][t] = bathy; sm[1][3][t] = dM; sm[1][4][t] = pa;
  __syncthreads();
  const double e1E = sm[1][0][do_u ? tE : t], z0E = sm[1][1][do_u ? tE : t], bathyE = sm[1][2][do_u ? tE : t];
  const double e1N = sm[1][0][do_v ? tN : t], z0N = sm[1][1][do_v ? tN : t], bathyN = sm[1][2][do_v ? tN : t];
  const double dMx = do_u ? (sm[1][3][tE] - dM) * IdxCu : 0.0, dMy = do_v ? (sm[1][3][tN] - dM) * IdyCv : 0.0;
  double intx_pa = 0.5 * (pa + sm[1][4][do_u ? tE : t]), inty_pa = 0.5 * (pa + sm[1][4][do_v ? tN : t]);  // :1538-1544
  __syncthreads();
  // Set_pbce_Bouss own-column state (MOM_PressureForce_Montgomery.F90:685-745)
  const double e_bot = K.e[g + (long long)nz * pl];
  const double Rho0xG = K.rho0_pbce * K.g_Earth;
  double Ihtot = 0.0, pbce = 0.0, T_prev = 0.0, S_prev = 0.0;
  if (K.pbce) Ihtot = K.H_to_Z / ((e1 - e_bot) + K.dz_neglect);
  double zt = e1;
  for (int k = 0; k < nz; ++k) {
    const long long ko = (long long)k * pl;
    double* my = &sm[k & 1][0][t];
    const double zb = K.e[g + ko + pl];
    const double hk = __ldg(K.h + g + ko);
    const double Tt = R.T_t[g + ko], Tb = R.T_b[g + ko], St = R.S_t[g + ko], Sb = R.S_b[g + ko];
    const double Tm = __ldg(K.T + g + ko), Sm = __ldg(K.S + g + ko);
    // 1. the vertical integrals of the own column (:563-614 / :1030-1077)
    double dpa, intz_dpa;
    {
      double s6 = 0., t6 = 0.;
      if (PPM) {
        s6 = 3.0 * (2.0 * Sm - (St + Sb));
        t6 = 3.0 * (2.0 * Tm - (Tt + Tb));
      }
      const double dz = zt - zb;
      double r5[6];
#pragma unroll
      for (int n = 1; n <= 5; ++n) {
        const double wt_t = 0.25 * (double)(5 - n), wt_b = 1.0 - wt_t;
        const double p5 = -GxRho * ((zt - z0) - 0.25 * (double)(n - 1) * dz);
        double S5, T5;
        if (PPM) {
          S5 = wt_t * St + wt_b * (Sb + s6 * wt_t);
          T5 = wt_t * Tt + wt_b * (Tb + t6 * wt_t);
        } else {
          S5 = wt_t * St + wt_b * Sb;
          T5 = wt_t * Tt + wt_b * Tb;
        }
        r5[n] = rho(T5, S5, p5);
      }
      if (rho.use_ref) {
        const double rho_anom = C1_90 * (7.0 * (r5[1] + r5[5]) + 32.0 * (r5[2] + r5[4]) + 12.0 * r5[3]);
        dpa = K.g_Earth * dz * rho_anom;
        intz_dpa = 0.5 * K.g_Earth * (dz * dz) * (rho_anom - C1_90 * (16.0 * (r5[4] - r5[2]) + 7.0 * (r5[5] - r5[1])));
      } else {
        const double rho_anom = C1_90 * (7.0 * (r5[1] + r5[5]) + 32.0 * (r5[2] + r5[4]) + 12.0 * r5[3]) - K.rho_ref;
        dpa = K.g_Earth * dz * rho_anom;
        intz_dpa = 0.5 * K.g_Earth * (dz * dz) *
                   (rho_anom - C1_90 * (16.0 * ((r5[4] - K.rho_ref) - (r5[2] - K.rho_ref)) + 7.0 * ((r5[5] - K.rho_ref) - (r5[1] - K.rho_ref))));
      }
      if (K.Z_to_H != 1.0) intz_dpa = intz_dpa * K.Z_to_H;  // :1306-1311
    }
    my[Q_ET * NT] = zt; my[Q_EB * NT] = zb; my[Q_TT * NT] = Tt; my[Q_TB * NT] = Tb; my[Q_ST * NT] = St; my[Q_SB * NT] = Sb;
    if (PPM) { my[Q_TM * NT] = Tm; my[Q_SM * NT] = Sm; }
    my[Q_DPA * NT] = dpa; my[Q_INTZ * NT] = intz_dpa; my[Q_PA * NT] = pa; my[Q_H * NT] = hk;
    __syncthreads();  // the only barrier of the layer: the next layer writes the other buffer
    if (do_u) {
      const double* Rt = &sm[k & 1][0][tE];
      const double intx_dpa = recon_face_integral<PPM, (VAR & 1) ? 1 : 3>(K, R, rho, GxRho, my, Rt, NT, bathy, bathyE, e1, e1E, z0, z0E);
      const double hE = Rt[Q_H * NT], paE = Rt[Q_PA * NT], intzE = Rt[Q_INTZ * NT], zbE = Rt[Q_EB * NT];
      double PF = (((pa * hk + intz_dpa) - (paE * hE + intzE)) + ((hE - hk) * intx_pa - (zbE - zb) * intx_dpa * K.Z_to_H)) *
                  ((2.0 * I_Rho0 * IdxCu) / ((hk + hE) + K.h_neglect));
      if (K.GFS_scale < 1.0) PF = PF - dMx;
      K.PFu[g + ko] = PF;
      intx_pa = intx_pa + intx_dpa;
    }
    if (do_v) {
      const double* Rt = &sm[k & 1][0][tN];
      const double inty_dpa = recon_face_integral<PPM, (VAR & 1) ? 1 : 3>(K, R, rho, GxRho, my, Rt, NT, bathy, bathyN, e1, e1N, z0, z0N);
      const double hN = Rt[Q_H * NT], paN = Rt[Q_PA * NT], intzN = Rt[Q_INTZ * NT], zbN = Rt[Q_EB * NT];
      double PF = (((pa * hk + intz_dpa) - (paN * hN + intzN)) + ((hN - hk) * inty_pa - (zbN - zb) * inty_dpa * K.Z_to_H)) *
                  ((2.0 * I_Rho0 * IdyCv) / ((hk + hN) + K.h_neglect));
      if (K.GFS_scale < 1.0) PF = PF - dMy;
      K.PFv[g + ko] = PF;
      inty_pa = inty_pa + inty_dpa;
    }
    // Set_pbce_Bouss: only by the thread that owns the column in the non-overlapping partition
    if (K.pbce && in && (tx < TX - 1 || i == G.iec + 1) && (ty < TY - 1 || j == G.jec + 1)) {
      const double press = -Rho0xG * (zt - K.Z_ref);
      if (k == 0) {
        pbce = G_Rho0 * (K.GFS_scale * eos_density(K, Tm, Sm, press)) * K.H_to_Z;
      } else {
        const double T_int = 0.5 * (T_prev + Tm), S_int = 0.5 * (S_prev + Sm);
        double dR_dT, dR_dS;
        eos_derivs(K, T_int, S_int, press, dR_dT, dR_dS);
        pbce = pbce + G_Rho0 * ((zt - e_bot) * Ihtot) * (dR_dT * (Tm - T_prev) + dR_dS * (Sm - S_prev));
      }
      T_prev = Tm; S_prev = Sm;
      K.pbce[g + ko] = pbce;
    }
    pa = pa + dpa;  // :1340-1345
    zt = zb;
  }
}

}  // namespace
