// btstep for sm_100a: everything around the substep loop, device resident.
//
// Replaces src/core/MOM_barotropic.F90: btstep :455-2172 (setup :868-1795, post :1814-1913),
// btstep_find_Cor :2836, btstep_ubt_from_layer :3388, btstep_layer_accel :3432, set_local_BT_cont_types :4876,
// and btcalc :4360, bt_mass_source :5243.  The substep loop is bt_timeloop.cu (m6_bt_run).
//
// Design (DESIGN.md "K1/K2/K4/K5"):
//  * The reference makes ~12 passes over wt_u/wt_v and the 3-D inputs to build its 2-D forcing, weights and
//    transports (:1011-1089, :1159-1178, :1324-1330, :1479-1485, :3411-3413).  Here ONE column kernel per direction
//    owns a velocity-point column: it forms the normalisation of the weights in a first sweep over k and every
//    k-weighted sum (ubt_Cor, gtot_E/W, uhbt0, ubt, BT_force, av_rem) in a second, each sum sequential in k
//    (bitwise parity); the normalised weights wt_u/wt_v are never stored.
//  * Everything 2-D (BTCL fits, f_4 Coriolis weights, Cor_ref, eta_src and its bounds, e_anom, ...) is written
//    straight into the resident planes the substep kernel reads; no host round trip between setup, loop and post.
//  * Halo updates of the setup (pass_gtot, pass_ubt_Cor, pass_eta_bt_rem, pass_force_hbt0_Cor_ref, BT_cont passes,
//    :811-859) are m6_halo_update calls on the same planes (cyclic wrap on one tile, NCCL between tiles).
//  * av_rem**Instep (:1502) is the only non-IEEE-exact operation on the whole path.  With BT_STRONG_DRAG=False it is
//    evaluated by the host's libm (what the Fortran runtime calls) on the 2-D av_rem planes so that answers match
//    the reference bit for bit; with BT_STRONG_DRAG=True everything stays on the device.
#include "ctx.h"
#include "pow_glibc.cuh"
#include "bt_planes.h"
#include "stage.h"
#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

using m6::Geom;
using m6::fmax2;
using m6::fmin2;


namespace {

enum { FA_EE = 0, FA_E0, FA_W0, FA_WW, UBT_WW, UBT_EE, CRV_W, CRV_E, UH_WW, UH_EE };

struct Pl10 { double* p[10]; };
struct Pl4 { double* p[4]; };

__device__ __forceinline__ double find_hbt10(double u, const Pl10& b, long long g) {  // find_uhbt :4610-4631
  if (u == 0.0) return 0.0;
  const double uEE = b.p[UBT_EE][g];
  if (u < uEE) return (u - uEE) * b.p[FA_EE][g] + b.p[UH_EE][g];
  if (u < 0.0) return u * (b.p[FA_E0][g] + b.p[CRV_E][g] * (u * u));
  const double uWW = b.p[UBT_WW][g];
  if (u <= uWW) return u * (b.p[FA_W0][g] + b.p[CRV_W][g] * (u * u));
  return (u - uWW) * b.p[FA_WW][g] + b.p[UH_WW][g];
}

// ---- set_local_BT_cont_types :4897-4925 (copy into wide arrays) and :4950-5002 (derived fields)
struct BtclCopy { const double* src[6]; double* dst[6]; };
__global__ void btcl_copy_kernel(const Geom G, const BtclCopy U, const BtclCopy V) {
  const int i = G.isc - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsc - 1 + blockIdx.y;
  if (i > G.iec || j > G.jec) return;
  const long long g = G.idx(i, j);
  if (j >= G.jsc) for (int m = 0; m < 6; ++m) U.dst[m][g] = U.src[m][g];
  if (i >= G.isc) for (int m = 0; m < 6; ++m) V.dst[m][g] = V.src[m][g];
}
__global__ void btcl_derive_kernel(const Geom G, Pl10 bu, Pl10 bv, int hs, double dt) {
  const int i = G.isc - hs - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsc - hs - 1 + blockIdx.y;
  if (i > G.iec + hs || j > G.jec + hs) return;
  const long long g = G.idx(i, j);
  const double C1_3 = 1.0 / 3.0;
  for (int dir = 0; dir < 2; ++dir) {
    if (dir == 0 && j < G.jsc - hs) continue;
    if (dir == 1 && i < G.isc - hs) continue;
    Pl10& b = dir == 0 ? bu : bv;
    const double uEE = dt * b.p[UBT_EE][g], uWW = dt * b.p[UBT_WW][g];
    const double fEE = b.p[FA_EE][g], fE0 = b.p[FA_E0][g], fW0 = b.p[FA_W0][g], fWW = b.p[FA_WW][g];
    b.p[UBT_EE][g] = uEE; b.p[UBT_WW][g] = uWW;
    b.p[UH_EE][g] = uEE * (C1_3 * (2.0 * fE0 + fEE));
    b.p[UH_WW][g] = uWW * (C1_3 * (2.0 * fW0 + fWW));
    double cW = 0.0, cE = 0.0;
    if (fabs(uWW) > 0.0) cW = (C1_3 * (fWW - fW0)) / (uWW * uWW);
    if (fabs(uEE) > 0.0) cE = (C1_3 * (fEE - fE0)) / (uEE * uEE);
    b.p[CRV_W][g] = cW; b.p[CRV_E][g] = cE;
  }
}

// ---- the column kernel: weights and every k-weighted sum of the setup
struct ColArgs {
  // 3-D inputs of this direction
  const double *frhat, *visc_rem, *vel_Cor, *vel_in, *bc_accel, *pbce, *uh0, *u_uh0;
  // 2-D inputs
  const double *mask, *tau, *tau_bot, *IDat;
  Pl10 btcl;
  // 2-D outputs
  double *vbt_Cor, *gtot_A, *gtot_B, *hbt0, *bt, *BT_force, *bt_rem, *av_rem;
  int nlo, nhi, olo, ohi, nk;
  int wt_uv_bug, visc_rem_uh0, strong_drag, nstep;
  double Instep, RZ_to_H, vel_underflow;
};

template <bool U>
__global__ void __launch_bounds__(128) bt_col_kernel(const Geom G, const ColArgs A) {
  const int n = A.nlo + blockIdx.x * blockDim.x + threadIdx.x, o = A.olo + blockIdx.y;
  if (n > A.nhi || o > A.ohi) return;
  const long long g = G.idx(n, o), sd = U ? 1 : G.pitch;
  const int nz = A.nk;
  const double subroundoff = 1e-30;  // MOM_barotropic.F90:445
  auto wraw = [&](long long gk) {    // :1011-1034
    double visc_rem = fmin2(__ldg(A.visc_rem + gk), 1.);
    visc_rem = fmax2(visc_rem, 1. - 0.5 * A.Instep / (visc_rem + subroundoff));
    visc_rem = fmax2(visc_rem, 0.);
    return __ldg(A.frhat + gk) * visc_rem;
  };
  const double mask = __ldg(A.mask + g);
  double Iwt = 1.0;
  if (!A.wt_uv_bug) {  // :1036-1058
    Iwt = wraw(g);
    for (int k = 1; k < nz; ++k) Iwt = Iwt + wraw(g + (long long)k * G.plane);
    if (fabs(Iwt) > 0.0) Iwt = mask / Iwt;
  }
  double vCor = 0.0, gA = 0.0, gB = 0.0, hbt = 0.0, bt0 = 0.0, bt = 0.0, av_rem = 0.0, force = 0.0;
  if (mask > 0.0) force = __ldg(A.tau + g) * A.RZ_to_H * __ldg(A.IDat + g) * __ldg(A.visc_rem + g);  // :1280
  if (A.tau_bot && mask > 0.0) force = force - __ldg(A.tau_bot + g) * A.RZ_to_H * __ldg(A.IDat + g);  // :1312
  for (int k = 0; k < nz; ++k) {   // (#pragma unroll 4 measured slower: 124 registers, 2.76 vs 1.96 ms for the v columns)
    const long long gk = g + (long long)k * G.plane;
    double wt = wraw(gk);
    if (!A.wt_uv_bug) wt = wt * Iwt;
    const double fr = __ldg(A.frhat + gk);
    vCor = vCor + wt * __ldg(A.vel_Cor + gk);                    // :1066-1073
    gA = gA + __ldg(A.pbce + gk) * wt;                            // :1079-1080 / :1086-1087
    gB = gB + __ldg(A.pbce + gk + sd) * wt;
    if (A.uh0) {                                                  // :1159-1178
      hbt = hbt + __ldg(A.uh0 + gk);
      bt0 = bt0 + (A.visc_rem_uh0 ? wt : fr) * __ldg(A.u_uh0 + gk);
    }
    bt = bt + wt * __ldg(A.vel_in + gk);                          // :3411-3417
    force = force + wt * __ldg(A.bc_accel + gk);                  // :1324-1330
    av_rem = av_rem + fr * __ldg(A.visc_rem + gk);                // :1479-1485
  }
  A.vbt_Cor[g] = vCor;
  A.gtot_A[g] = gA;
  A.gtot_B[g + sd] = gB;
  if (A.uh0) A.hbt0[g] = hbt - find_hbt10(bt0, A.btcl, g);       // :1212-1216
  if (fabs(bt) < A.vel_underflow) bt = 0.0;                        // :3420-3425
  A.bt[g] = bt;
  A.BT_force[g] = force;
  if (A.strong_drag) A.bt_rem[g] = mask * ((A.nstep * av_rem) / (1.0 + (A.nstep - 1) * av_rem));  // :1489-1491
  else A.av_rem[g] = av_rem;
}

// ---- 2-D setup on the wide planes
struct Setup2D {
  const double *q_D, *D_u_Cor, *D_v_Cor, *OBCmask_u, *OBCmask_v;  // wide CS arrays
  Pl4 f4u, f4v;
  int Sadourny, isvf, ievf, jsvf, jevf;
};
__global__ void bt_find_Cor_kernel(const Geom G, const Setup2D S) {  // btstep_find_Cor :2866-2895 with q = q_D etc. (:868-881)
  const int i = S.isvf - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = S.jsvf - 1 + blockIdx.y;
  if (i > S.ievf + 1 || j > S.jevf + 1) return;
  const long long g = G.idx(i, j), P = G.pitch;
  const double* q = S.q_D; const double* Du = S.D_u_Cor; const double* Dv = S.D_v_Cor;
  if (j <= S.jevf) {  // f_4_v(:,i,J), J=jsvf-1..jevf, i=isvf-1..ievf+1
    const double m = S.OBCmask_v[g];
    if (S.Sadourny) {
      S.f4v.p[0][g] = m * Du[g - 1] * q[g - 1];
      S.f4v.p[1][g] = m * Du[g] * q[g];
      S.f4v.p[3][g] = m * Du[g + P] * q[g];
      S.f4v.p[2][g] = m * Du[g + P - 1] * q[g - 1];
    } else {
      S.f4v.p[0][g] = m * Du[g - 1] * ((q[g] + q[g - P - 1]) + q[g - 1]) / 3.0;
      S.f4v.p[1][g] = m * Du[g] * (q[g] + (q[g - 1] + q[g - P])) / 3.0;
      S.f4v.p[3][g] = m * Du[g + P] * (q[g] + (q[g - 1] + q[g + P])) / 3.0;
      S.f4v.p[2][g] = m * Du[g + P - 1] * ((q[g] + q[g + P - 1]) + q[g - 1]) / 3.0;
    }
  }
  if (i <= S.ievf) {  // f_4_u(:,I,j), j=jsvf-1..jevf+1, I=isvf-1..ievf
    const double m = S.OBCmask_u[g];
    if (S.Sadourny) {
      S.f4u.p[3][g] = m * Dv[g + 1] * q[g];
      S.f4u.p[2][g] = m * Dv[g] * q[g];
      S.f4u.p[0][g] = m * Dv[g - P] * q[g - P];
      S.f4u.p[1][g] = m * Dv[g - P + 1] * q[g - P];
    } else {
      S.f4u.p[3][g] = m * Dv[g + 1] * (q[g] + (q[g + 1] + q[g - P])) / 3.0;
      S.f4u.p[2][g] = m * Dv[g] * (q[g] + (q[g - 1] + q[g - P])) / 3.0;
      S.f4u.p[0][g] = m * Dv[g - P] * ((q[g] + q[g - P - 1]) + q[g - P]) / 3.0;
      S.f4u.p[1][g] = m * Dv[g - P + 1] * ((q[g] + q[g - P + 1]) + q[g - P]) / 3.0;
    }
  }
}

// bt_rem = mask * av_rem**Instep (:1497-1509) over the faces (is-1:ie, js:je) and (is:ie, js-1:je), on the device: the real power is the
// reference platform's own libm routine restated operation for operation (pow_glibc.cuh, bit-identical to glibc's pow on 6e8 tested
// arguments), so no value leaves the GPU.  An argument outside that routine's validated domain (impossible for 0 < av_rem, Instep <= 1
// unless av_rem < 1e-222) yields NaN, which the solver propagates into every output.
__global__ void bt_rem_pow_kernel(m6::Geom G, const double* __restrict__ mu, const double* __restrict__ mv,
                                  const double* __restrict__ au, const double* __restrict__ av, int is, int js, int nx, int ny,
                                  double Instep, double* __restrict__ ru, double* __restrict__ rv) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nx * ny) return;
  const int i = is - 1 + n % nx, j = js - 1 + n / nx;
  const size_t g = (size_t)G.idx(i, j);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  if (j >= js) {
    double r = 0.0;
    const double m = mu[g], a = au[g];
    if (m * a > 0.0) { bool ok = true; const double pw = m6pow::pow_glibc<true>(a, Instep, &ok); r = ok ? m * pw : qnan; }
    ru[g] = r;
  }
  if (i >= is) {
    double r = 0.0;
    const double m = mv[g], a = av[g];
    if (m * a > 0.0) { bool ok = true; const double pw = m6pow::pow_glibc<true>(a, Instep, &ok); r = ok ? m * pw : qnan; }
    rv[g] = r;
  }
}

__global__ void bt_Cor_ref_kernel(const Geom G, Pl4 f4u, Pl4 f4v, const double* ubt_Cor, const double* vbt_Cor,
                                  double* Cor_ref_u, double* Cor_ref_v) {  // :1452-1461
  const int i = G.isc - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsc - 1 + blockIdx.y;
  if (i > G.iec || j > G.jec) return;
  const long long g = G.idx(i, j), P = G.pitch;
  if (j >= G.jsc)
    Cor_ref_u[g] = (((f4u.p[3][g] * vbt_Cor[g + 1]) + (f4u.p[0][g] * vbt_Cor[g - P])) +
                    ((f4u.p[2][g] * vbt_Cor[g]) + (f4u.p[1][g] * vbt_Cor[g - P + 1])));
  if (i >= G.isc)
    Cor_ref_v[g] = -1.0 * (((f4v.p[0][g] * ubt_Cor[g - 1]) + (f4v.p[3][g] * ubt_Cor[g + P])) +
                           ((f4v.p[1][g] * ubt_Cor[g]) + (f4v.p[2][g] * ubt_Cor[g + P - 1])));
}

// copy eta_in / eta_PF_in over G's data domain into the zeroed wide planes (:997-1003)
__global__ void bt_copy_G_kernel(const Geom G, const double* a, double* wa, const double* b, double* wb) {
  const int i = G.isd + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsd + blockIdx.y;
  if (i > G.ied || j > G.jed) return;
  const long long g = G.idx(i, j);
  wa[g] = a[g]; wb[g] = b[g];
}

struct SrcArgs {
  int bound_BT_corr, BT_cont_bounds, Boussinesq;
  double dt, Idt, Instep, maxCFL_BT_cont, Z_to_H;
  const double *maskT, *dxT, *dyT, *IareaT, *bathyT, *eta, *eta_cor_bound, *uhbt0, *vhbt0;
  Pl10 bu, bv;
  double *eta_cor, *eta_src;
};
__global__ void bt_eta_src_kernel(const Geom G, const SrcArgs S) {  // :1549-1587
  const int i = G.isc + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsc + blockIdx.y;
  if (i > G.iec || j > G.jec) return;
  const long long g = G.idx(i, j), P = G.pitch;
  double ec = S.eta_cor[g];
  if (S.bound_BT_corr) {
    if (S.BT_cont_bounds) {
      if (S.maskT[g] > 0.0) {
        if (ec > 0.0) {
          const double u_max_cor = S.dxT[g] * (S.maxCFL_BT_cont * S.Idt);
          const double v_max_cor = S.dyT[g] * (S.maxCFL_BT_cont * S.Idt);
          const double eta_cor_max = S.dt * (S.IareaT[g] *
              (((find_hbt10(u_max_cor, S.bu, g) + S.uhbt0[g]) - (find_hbt10(-u_max_cor, S.bu, g - 1) + S.uhbt0[g - 1])) +
               ((find_hbt10(v_max_cor, S.bv, g) + S.vhbt0[g]) - (find_hbt10(-v_max_cor, S.bv, g - P) + S.vhbt0[g - P]))));
          ec = fmin2(ec, fmax2(0.0, eta_cor_max));
        } else {
          double Htot = S.eta[g];
          if (S.Boussinesq) Htot = S.bathyT[g] * S.Z_to_H + S.eta[g];
          ec = fmax2(ec, -fmax2(0.0, Htot));
        }
      }
    } else {
      const double b = S.dt * S.eta_cor_bound[g];
      if (fabs(ec) > b) ec = copysign(b, ec);
    }
    S.eta_cor[g] = ec;
  }
  S.eta_src[g] = S.maskT[g] * (S.Instep * ec);
}

// e_anom, etaav, eta_out (:1814-1847)
__global__ void bt_post2d_kernel(const Geom G, double dgeo_de, const double* eta, const double* eta_in, const double* eta_PF,
                                 const double* eta_sum, const double* eta_wtd, double* e_anom, double* etaav, double* eta_out) {
  const int i = G.isc + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsc + blockIdx.y;
  if (i > G.iec || j > G.jec) return;
  const long long g = G.idx(i, j);
  if (etaav) etaav[g] = eta_sum[g] * 1.0;
  e_anom[g] = dgeo_de * (0.5 * (eta[g] + eta_in[g]) - eta_PF[g]);
  eta_out[g] = eta_wtd[g] * 1.0;
}

// btstep_layer_accel :3480-3497
__global__ void bt_layer_accel_kernel(const Geom G, const double* u_accel_bt, const double* v_accel_bt, const double* pbce,
                                      const double* gtot_E, const double* gtot_W, const double* gtot_N, const double* gtot_S,
                                      const double* e_anom, const double* IdxCu, const double* IdyCv, double accel_underflow,
                                      double* accel_layer_u, double* accel_layer_v, int nk) {
  const int i = G.isc - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsc - 1 + blockIdx.y;
  if (i > G.iec || j > G.jec) return;
  const long long g = G.idx(i, j), P = G.pitch;
  const bool do_u = j >= G.jsc, do_v = i >= G.isc;
  const double ea0 = e_anom[g], eaE = e_anom[g + 1], eaN = e_anom[g + P];
  const double gE = gtot_E[g], gWe = gtot_W[g + 1], gN = gtot_N[g], gSn = gtot_S[g + P];
  const double ua = u_accel_bt[g], va = v_accel_bt[g], Idx = IdxCu[g], Idy = IdyCv[g];
  for (int k = blockIdx.z; k < nk; k += gridDim.z) {
    const long long gk = g + (long long)k * G.plane;
    const double p0 = __ldg(pbce + gk);
    if (do_u) {
      double a = (ua - (((__ldg(pbce + gk + 1) - gWe) * eaE) - ((p0 - gE) * ea0)) * Idx);
      if (fabs(a) < accel_underflow) a = 0.0;
      accel_layer_u[gk] = a;
    }
    if (do_v) {
      double a = (va - (((__ldg(pbce + gk + P) - gSn) * eaN) - ((p0 - gN) * ea0)) * Idy);
      if (fabs(a) < accel_underflow) a = 0.0;
      accel_layer_v[gk] = a;
    }
  }
}

// ---- btcalc :4360-4605: one thread per velocity-point column
struct BtcalcArgs {
  const double *h, *h_vel, *bathyT, *mask;
  double* frhat;
  int scheme, nlo, nhi, olo, ohi, nk;  // scheme: 0 = from h_u/h_v, 1 HARMONIC, 2 ARITHMETIC, 3 HYBRID
  double h_neglect, Z_to_H;
};
template <bool U>
__global__ void __launch_bounds__(128) btcalc_kernel(const Geom G, const BtcalcArgs A) {
  const int n = A.nlo + blockIdx.x * blockDim.x + threadIdx.x, o = A.olo + blockIdx.y;
  if (n > A.nhi || o > A.ohi) return;
  const long long g = G.idx(n, o), sd = U ? 1 : G.pitch;
  const int nz = A.nk;
  const double hn = A.h_neglect;
  double D_shallow = 0.0, e_bot = 0.0;
  if (A.scheme == 3) {
    e_bot = -0.5 * A.Z_to_H * (__ldg(A.bathyT + g + sd) + __ldg(A.bathyT + g));
    D_shallow = -A.Z_to_H * fmin2(__ldg(A.bathyT + g + sd), __ldg(A.bathyT + g));
  }
  // hat(k) for one layer; the HYBRID branch needs e(K+1), passed in and updated bottom-up
  auto hat = [&](int k, double& e_below) -> double {
    const long long gk = g + (long long)k * G.plane;
    if (A.scheme == 0) return __ldg(A.h_vel + gk);
    const double hp = __ldg(A.h + gk + sd), h0 = __ldg(A.h + gk);
    if (A.scheme == 2) return 0.5 * (hp + h0);
    if (A.scheme == 1) return 2.0 * (hp * h0) / ((hp + h0) + hn);
    const double e_above = e_below + 0.5 * (hp + h0);
    const double h_arith = 0.5 * (hp + h0);
    double r;
    if (e_below >= D_shallow) r = h_arith;
    else {
      const double h_harm = (hp * h0) / (h_arith + hn);
      if (e_above <= D_shallow) r = h_harm;
      else {
        const double wt_arith = (e_above - D_shallow) / (h_arith + hn);
        r = wt_arith * h_arith + (1.0 - wt_arith) * h_harm;
      }
    }
    e_below = e_above;
    return r;
  };
  double tot = 0.0;
  if (A.scheme == 3) {  // bottom-up recursion, sum in the same (bottom-up) order as the reference (:4458-4473)
    double e = e_bot;
    for (int k = nz - 1; k >= 0; --k) { const double v = hat(k, e); A.frhat[g + (long long)k * G.plane] = v; tot = tot + v; }
  } else {
    double e = 0.0;
    for (int k = 0; k < nz; ++k) { const double v = hat(k, e); A.frhat[g + (long long)k * G.plane] = v; tot = tot + v; }
  }
  const double Ihat = __ldg(A.mask + g) / (tot + hn);
  for (int k = 0; k < nz; ++k) { const long long gk = g + (long long)k * G.plane; A.frhat[gk] = A.frhat[gk] * Ihat; }
}

__global__ void bt_mass_source_kernel(const Geom G, const double* h, const double* eta, const double* bathyT, int Boussinesq,
                                      double Z_to_H, int set_cor, double* eta_cor, int nk) {  // :5268-5292
  const int i = G.isc + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsc + blockIdx.y;
  if (i > G.iec || j > G.jec) return;
  const long long g = G.idx(i, j);
  double eta_h = Boussinesq ? __ldg(h + g) - bathyT[g] * Z_to_H : __ldg(h + g);
  for (int k = 1; k < nk; ++k) eta_h = eta_h + __ldg(h + g + (long long)k * G.plane);
  const double d_eta = eta_h - eta[g];
  if (set_cor) eta_cor[g] = d_eta; else eta_cor[g] = eta_cor[g] + d_eta;
}

inline dim3 grid2(int ni, int nj, int bx) { return dim3((ni + bx - 1) / bx, nj); }

}  // namespace

// Device-resident btstep.  All pointers of D are unified planes.
int m6_btstep_run(mom6cu_ctx* c, const mom6cu_barotropic_cs& CS, const BtstepDev& D) {
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: grid / vertical grid not set");
  if (CS.unsupported) return c->fail(MOM6CU_ERR_UNSUPPORTED, "btstep: an option outside the frozen option set is enabled");
  if (CS.adjust_BT_cont)  // adjust_local_BT_cont_types, MOM_barotropic.F90:1180-1198, 5010-5103
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "btstep: ADJUST_BT_CONT=True is outside the frozen option set");
  if (!D.have_BT_cont) return c->fail(MOM6CU_ERR_UNSUPPORTED, "btstep: USE_BT_CONT_TYPE=False is outside the frozen option set");
  const bool add_uh0 = D.uh0 != nullptr;
  if (add_uh0 && !(D.vh0 && D.u_uh0 && D.v_vh0))
    return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: vh0, u_uh0, and v_vh0 must be associated if uh0 is used.");
  const Geom& G = c->g;
  const mom6cu_domain& d = c->dom;
  const GridDev& M = c->grid;
  const int is = d.isc, ie = d.iec, js = d.jsc, je = d.jec, nz = G.nk;
  const double dt = D.dt, Idt = 1.0 / dt;
  // :765-802
  const int stencil = std::max(1, CS.min_stencil);
  int num_cycles = 1;
  if (CS.use_wide_halos) num_cycles = std::min((is - d.isdw) / stencil, (js - d.jsdw) / stencil);
  if (num_cycles < 1) return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: wide halo narrower than the stencil");
  const int isvf = is - (num_cycles - 1) * stencil, ievf = ie + (num_cycles - 1) * stencil;
  const int jsvf = js - (num_cycles - 1) * stencil, jevf = je + (num_cycles - 1) * stencil;
  if (isvf - 2 < d.isdw - 1 || jsvf - 2 < d.jsdw - 1) return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: wide halo too narrow for the Coriolis stencil");
  const int nstep = (int)std::ceil(dt / CS.dtbt - 0.0001);
  if (nstep < 1) return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: nstep < 1");
  const double Instep = 1.0 / (double)nstep;
  const double dtbt = dt * Instep;
  const double dgeo_de = 1.0 + CS.G_extra;

  BtDevice B;
  int rc;
  if ((rc = m6_bt_alloc(c, B))) return rc;
  BtPlanes& P = B.P;
  const size_t pb = (size_t)G.plane * sizeof(double);
  auto pl = [&](const char* n) { return c->plane2(std::string("bts.") + n); };
  double *ubt_Cor = pl("ubt_Cor"), *vbt_Cor = pl("vbt_Cor"), *av_rem_u = pl("av_rem_u"), *av_rem_v = pl("av_rem_v"), *e_anom = pl("e_anom");
  if (!e_anom) return MOM6CU_ERR_CUDA;
  // zero the wide planes the reference zeroes (:938-965) and the accumulators (:1248-1249)
  double* zero_list[] = {(double*)P.gtot_E, (double*)P.gtot_W, (double*)P.gtot_N, (double*)P.gtot_S, B.eta[0], (double*)P.eta_PF,
                         (double*)P.Cor_ref_u, (double*)P.BT_force_u, B.ubt[0], (double*)P.bt_rem_u, (double*)P.uhbt0,
                         (double*)P.Cor_ref_v, (double*)P.BT_force_v, B.vbt[0], (double*)P.bt_rem_v, (double*)P.vhbt0,
                         (double*)P.eta_src, P.u_accel_bt, P.v_accel_bt, ubt_Cor, vbt_Cor, e_anom, P.eta_sum, P.eta_wtd};
  for (double* z : zero_list) M6_CUDA(c, cudaMemsetAsync(z, 0, pb, c->stream));
  Pl10 bu, bv; Pl4 f4u, f4v;
  for (int m = 0; m < 10; ++m) { bu.p[m] = (double*)P.bu[m]; bv.p[m] = (double*)P.bv[m]; M6_CUDA(c, cudaMemsetAsync(bu.p[m], 0, pb, c->stream)); M6_CUDA(c, cudaMemsetAsync(bv.p[m], 0, pb, c->stream)); }
  for (int m = 0; m < 4; ++m) { f4u.p[m] = (double*)P.f4u[m]; f4v.p[m] = (double*)P.f4v[m]; M6_CUDA(c, cudaMemsetAsync(f4u.p[m], 0, pb, c->stream)); M6_CUDA(c, cudaMemsetAsync(f4v.p[m], 0, pb, c->stream)); }
  P.IareaT = CS.IareaT_OBCmask; P.IdxCu = CS.IdxCu; P.IdyCv = CS.IdyCv;  // resident wide CS planes
  P.ubtav = CS.ubtav; P.vbtav = CS.vbtav; P.uhbtav = D.uhbtav; P.vhbtav = D.vhbtav;

  // ---- set_local_BT_cont_types (:1132) ----
  {
    const int hs = std::max(1 + ievf - ie, 0);
    BtclCopy U = {{D.FA_u_EE, D.FA_u_E0, D.FA_u_W0, D.FA_u_WW, D.uBT_WW, D.uBT_EE}, {bu.p[FA_EE], bu.p[FA_E0], bu.p[FA_W0], bu.p[FA_WW], bu.p[UBT_WW], bu.p[UBT_EE]}};
    BtclCopy V = {{D.FA_v_NN, D.FA_v_N0, D.FA_v_S0, D.FA_v_SS, D.vBT_SS, D.vBT_NN}, {bv.p[FA_EE], bv.p[FA_E0], bv.p[FA_W0], bv.p[FA_WW], bv.p[UBT_WW], bv.p[UBT_EE]}};
    M6_LAUNCH(c, btcl_copy_kernel, grid2(ie - is + 2, je - js + 2, 128), 128, 0, G, U, V);
    double* hf[12]; int hst[12];
    for (int m = 0; m < 6; ++m) { hf[m] = U.dst[m]; hst[m] = ST_U; hf[6 + m] = V.dst[m]; hst[6 + m] = ST_V; }
    if ((rc = m6_halo_update(c, hf, hst, 12, 1, 1))) return rc;
    M6_LAUNCH(c, btcl_derive_kernel, grid2(ie - is + 2 * hs + 2, je - js + 2 * hs + 2, 128), 128, 0, G, bu, bv, hs, 1.0);
  }
  // ---- column sums ----
  ColArgs A = {};
  A.nk = nz; A.wt_uv_bug = CS.wt_uv_bug; A.visc_rem_uh0 = CS.visc_rem_u_uh0; A.strong_drag = CS.strong_drag; A.nstep = nstep;
  A.Instep = Instep; A.RZ_to_H = c->vgrid.RZ_to_H; A.vel_underflow = CS.vel_underflow;
  A.frhat = CS.frhatu; A.visc_rem = D.visc_rem_u; A.vel_Cor = D.U_Cor; A.vel_in = D.U_in; A.bc_accel = D.bc_accel_u; A.pbce = D.pbce;
  A.uh0 = D.uh0; A.u_uh0 = D.u_uh0; A.mask = M.mask2dCu; A.tau = D.taux; A.tau_bot = (D.taux_bot && D.tauy_bot) ? D.taux_bot : nullptr;
  A.IDat = CS.IDatu; A.btcl = bu;
  A.vbt_Cor = ubt_Cor; A.gtot_A = (double*)P.gtot_E; A.gtot_B = (double*)P.gtot_W; A.hbt0 = (double*)P.uhbt0; A.bt = B.ubt[0];
  A.BT_force = (double*)P.BT_force_u; A.bt_rem = (double*)P.bt_rem_u; A.av_rem = av_rem_u;
  A.nlo = is - 1; A.nhi = ie; A.olo = js; A.ohi = je;
  M6_LAUNCH(c, bt_col_kernel<true>, grid2(ie - is + 2, je - js + 1, 128), 128, 0, G, A);
  const int pnx = ie - is + 2, pny = je - js + 2;  // covers (is-1:ie, js-1:je)
  const size_t npk = (size_t)pnx * pny;
  A.frhat = CS.frhatv; A.visc_rem = D.visc_rem_v; A.vel_Cor = D.V_Cor; A.vel_in = D.V_in; A.bc_accel = D.bc_accel_v;
  A.uh0 = D.vh0; A.u_uh0 = D.v_vh0; A.mask = M.mask2dCv; A.tau = D.tauy; A.tau_bot = (D.taux_bot && D.tauy_bot) ? D.tauy_bot : nullptr;
  A.IDat = CS.IDatv; A.btcl = bv;
  A.vbt_Cor = vbt_Cor; A.gtot_A = (double*)P.gtot_N; A.gtot_B = (double*)P.gtot_S; A.hbt0 = (double*)P.vhbt0; A.bt = B.vbt[0];
  A.BT_force = (double*)P.BT_force_v; A.bt_rem = (double*)P.bt_rem_v; A.av_rem = av_rem_v;
  A.nlo = is; A.nhi = ie; A.olo = js - 1; A.ohi = je;
  M6_LAUNCH(c, bt_col_kernel<false>, grid2(ie - is + 1, je - js + 2, 128), 128, 0, G, A);

  // ---- Coriolis weights (:1419), halo updates (:1436-1441), Cor_ref (:1452-1461) ----
  Setup2D S = {CS.q_D, CS.D_u_Cor, CS.D_v_Cor, CS.OBCmask_u, CS.OBCmask_v, f4u, f4v, CS.Sadourny, isvf, ievf, jsvf, jevf};
  M6_LAUNCH(c, bt_find_Cor_kernel, grid2(ievf - isvf + 3, jevf - jsvf + 3, 128), 128, 0, G, S);
  {
    double* hf[4] = {(double*)P.gtot_E, (double*)P.gtot_N, (double*)P.gtot_W, (double*)P.gtot_S};
    const int hst[4] = {ST_H, ST_H, ST_H, ST_H};
    if ((rc = m6_halo_update(c, hf, hst, 4, 1, 1))) return rc;
    double* hg[2] = {ubt_Cor, vbt_Cor};
    const int gst[2] = {ST_U, ST_V};
    if ((rc = m6_halo_update(c, hg, gst, 2, 0, 1))) return rc;
  }
  M6_LAUNCH(c, bt_Cor_ref_kernel, grid2(ie - is + 2, je - js + 2, 128), 128, 0, G, f4u, f4v, ubt_Cor, vbt_Cor,
            (double*)P.Cor_ref_u, (double*)P.Cor_ref_v);
  if (!CS.strong_drag) {  // the viscous remnant with BT_STRONG_DRAG=False (:1497-1509)
    // rows/columns outside the computational faces stay zero, as in the reference's zero-initialised wide arrays
    M6_CUDA(c, cudaMemsetAsync((double*)P.bt_rem_u, 0, pb, c->stream));
    M6_CUDA(c, cudaMemsetAsync((double*)P.bt_rem_v, 0, pb, c->stream));
    M6_LAUNCH(c, bt_rem_pow_kernel, dim3((unsigned)((npk + 127) / 128)), 128, 0, G, M.mask2dCu, M.mask2dCv, av_rem_u, av_rem_v, is, js, pnx, pny,
              Instep, (double*)P.bt_rem_u, (double*)P.bt_rem_v);
  }
  // ---- eta, eta_PF (:997-1003) and the mass source (:1549-1587) ----
  M6_LAUNCH(c, bt_copy_G_kernel, grid2(d.ied - d.isd + 1, d.jed - d.jsd + 1, 128), 128, 0, G, D.eta_in, B.eta[0], D.eta_PF_in, (double*)P.eta_PF);
  SrcArgs Sr = {CS.bound_BT_corr, CS.BT_cont_bounds, c->vgrid.Boussinesq, dt, Idt, Instep, CS.maxCFL_BT_cont, c->vgrid.Z_to_H,
                M.mask2dT, M.dxT, M.dyT, CS.IareaT, CS.bathyT, B.eta[0], CS.eta_cor_bound, P.uhbt0, P.vhbt0, bu, bv, CS.eta_cor, (double*)P.eta_src};
  if (CS.bound_BT_corr && !CS.BT_cont_bounds && !CS.eta_cor_bound) return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: CS%%eta_cor_bound is required");
  M6_LAUNCH(c, bt_eta_src_kernel, grid2(ie - is + 1, je - js + 1, 128), 128, 0, G, Sr);
  {  // pass_eta_bt_rem, pass_force_hbt0_Cor_ref (:1627-1643)
    double* hf[10] = {(double*)P.eta_PF, (double*)P.eta_src, (double*)P.bt_rem_u, (double*)P.bt_rem_v, (double*)P.BT_force_u,
                      (double*)P.BT_force_v, (double*)P.Cor_ref_u, (double*)P.Cor_ref_v, (double*)P.uhbt0, (double*)P.vhbt0};
    const int hst[10] = {ST_H, ST_H, ST_U, ST_V, ST_U, ST_V, ST_U, ST_V, ST_U, ST_V};
    if ((rc = m6_halo_update(c, hf, hst, add_uh0 ? 10 : 8, 1, 1))) return rc;
  }
  // ---- filter weights (:1727-1795) ----
  double dt_filt;
  if (CS.dt_bt_filter >= 0.0) dt_filt = 0.5 * std::max(0.0, std::min(CS.dt_bt_filter, 2.0 * dt));
  else dt_filt = 0.5 * std::max(0.0, dt * std::min(-CS.dt_bt_filter, 2.0));
  const int nfilter = (int)std::ceil(dt_filt / dtbt);
  if (nstep + nfilter == 0) return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: number of barotropic step (nstep+nfilter) is 0");
  const int nt = nstep + nfilter;
  std::vector<double> wt_vel(nt), wt_eta(nt), wt_trans(nt + 1), wt_accel(nt + 1), wt_accel2(nt + 1);
  double sum_wt_vel = 0.0, sum_wt_eta = 0.0, sum_wt_accel = 0.0, sum_wt_trans = 0.0;
  for (int n = 1; n <= nt; ++n) {
    if ((n == nstep) || (dt_filt - std::abs(n - nstep) * dtbt >= 0.0)) { wt_vel[n - 1] = 1.0; wt_eta[n - 1] = 1.0; }
    else if (dtbt + dt_filt - std::abs(n - nstep) * dtbt > 0.0) { wt_vel[n - 1] = 1.0 + (dt_filt / dtbt) - std::abs(n - nstep); wt_eta[n - 1] = wt_vel[n - 1]; }
    else { wt_vel[n - 1] = 0.0; wt_eta[n - 1] = 0.0; }
    sum_wt_vel = sum_wt_vel + wt_vel[n - 1]; sum_wt_eta = sum_wt_eta + wt_eta[n - 1];
  }
  wt_trans[nt] = 0.0; wt_accel[nt] = 0.0;
  for (int n = nt; n >= 1; --n) {
    wt_trans[n - 1] = wt_trans[n] + wt_eta[n - 1];
    wt_accel[n - 1] = wt_accel[n] + wt_vel[n - 1];
    sum_wt_accel = sum_wt_accel + wt_accel[n - 1]; sum_wt_trans = sum_wt_trans + wt_trans[n - 1];
  }
  const double I_sum_wt_vel = 1.0 / sum_wt_vel, I_sum_wt_accel = 1.0 / sum_wt_accel;
  const double I_sum_wt_eta = 1.0 / sum_wt_eta, I_sum_wt_trans = 1.0 / sum_wt_trans;
  for (int n = 1; n <= nt; ++n) {
    wt_vel[n - 1] = wt_vel[n - 1] * I_sum_wt_vel;
    wt_accel2[n - 1] = wt_accel[n - 1] * I_sum_wt_accel;
    wt_trans[n - 1] = wt_trans[n - 1] * I_sum_wt_trans;
    wt_accel[n - 1] = wt_accel[n - 1] * I_sum_wt_accel;
    wt_eta[n - 1] = wt_eta[n - 1] * I_sum_wt_eta;
  }
  // ---- the substep loop (:1803) ----
  mom6cu_bt_timeloop_args T = {};
  T.wt_vel = wt_vel.data(); T.wt_eta = wt_eta.data(); T.wt_accel = wt_accel.data(); T.wt_trans = wt_trans.data(); T.wt_accel2 = wt_accel2.data();
  T.dtbt = dtbt; T.dgeo_de = dgeo_de; T.bebt = CS.bebt; T.vel_underflow = CS.vel_underflow;
  T.nstep = nstep; T.nfilter = nfilter; T.use_BT_cont = 1; T.find_etaav = D.etaav ? 1 : 0;
  T.BT_project_velocity = CS.BT_project_velocity; T.use_old_coriolis_bracket_bug = CS.use_old_coriolis_bracket_bug;
  T.use_wide_halos = CS.use_wide_halos; T.min_stencil = CS.min_stencil;
  int slot = 0;
  if ((rc = m6_bt_run(c, B, &T, &slot))) return rc;
  // ---- post (:1814-1913) ----
  M6_LAUNCH(c, bt_post2d_kernel, grid2(ie - is + 1, je - js + 1, 128), 128, 0, G, dgeo_de, B.eta[slot], D.eta_in, P.eta_PF,
            P.eta_sum, P.eta_wtd, e_anom, D.etaav, D.eta_out);
  {
    double* hg[6] = {e_anom, CS.ubtav, CS.vbtav, D.uhbtav, D.vhbtav, D.etaav};
    const int gst[6] = {ST_H, ST_U, ST_V, ST_U, ST_V, ST_H};
    if ((rc = m6_halo_update(c, hg, gst, D.etaav ? 6 : 5, 0, 1))) return rc;
  }
  {
    dim3 grid((ie - is + 2 + 127) / 128, je - js + 2, nz < 32 ? nz : 32);
    M6_LAUNCH(c, bt_layer_accel_kernel, grid, 128, 0, G, P.u_accel_bt, P.v_accel_bt, D.pbce, P.gtot_E, P.gtot_W, P.gtot_N, P.gtot_S,
              e_anom, CS.IdxCu, CS.IdyCv, CS.vel_underflow * Idt, D.accel_layer_u, D.accel_layer_v, nz);
  }
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

int m6_btcalc_run(mom6cu_ctx* c, const double* h, const double* h_u, const double* h_v, const double* bathyT, int hvel_scheme,
                  int may_use_default, double* frhatu, double* frhatv) {
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "btcalc: grid / vertical grid not set");
  const bool have = h_u && h_v;
  int scheme = have ? 0 : hvel_scheme;
  if (!have && !(hvel_scheme == 1 || hvel_scheme == 2 || hvel_scheme == 3)) {
    if (may_use_default) scheme = 3;
    else return c->fail(MOM6CU_ERR_BAD_ARG, "btcalc: Inconsistent settings of optional arguments and hvel_scheme.");
  }
  const mom6cu_domain& d = c->dom;
  BtcalcArgs A = {h, h_u, bathyT, c->grid.mask2dCu, frhatu, scheme, d.isc - 1, d.iec, d.jsc, d.jec, c->g.nk,
                  c->vgrid.H_subroundoff, c->vgrid.Z_to_H};
  M6_LAUNCH(c, btcalc_kernel<true>, grid2(d.iec - d.isc + 2, d.jec - d.jsc + 1, 128), 128, 0, c->g, A);
  A.h_vel = h_v; A.mask = c->grid.mask2dCv; A.frhat = frhatv; A.nlo = d.isc; A.nhi = d.iec; A.olo = d.jsc - 1; A.ohi = d.jec;
  M6_LAUNCH(c, btcalc_kernel<false>, grid2(d.iec - d.isc + 1, d.jec - d.jsc + 2, 128), 128, 0, c->g, A);
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

int m6_bt_mass_source_run(mom6cu_ctx* c, const double* h, const double* eta, int set_cor, double* eta_cor) {
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "bt_mass_source: grid / vertical grid not set");
  const mom6cu_domain& d = c->dom;
  M6_LAUNCH(c, bt_mass_source_kernel, grid2(d.iec - d.isc + 1, d.jec - d.jsc + 1, 128), 128, 0, c->g, h, eta, c->grid.bathyT,
            c->vgrid.Boussinesq, c->vgrid.Z_to_H, set_cor, eta_cor, c->g.nk);
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------- C ABI
int m6_stage_barotropic_cs(mom6cu_ctx* c, Stager& S, const mom6cu_barotropic_cs* CSh, mom6cu_barotropic_cs* CSp) {
  mom6cu_barotropic_cs& CS = *CSp;
  int rc;
  // control-structure arrays (wide / G-sized)
#define W2(f, st) if ((rc = S.in(CSh->f, st, 1, 1, "cs." #f, &CS.f))) return rc; if (!CS.f) return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: CS%%" #f " is null")
  W2(IareaT, ST_H); W2(IareaT_OBCmask, ST_H); W2(IdxCu, ST_U); W2(IdyCv, ST_V);
  W2(q_D, ST_Q); W2(D_u_Cor, ST_U); W2(D_v_Cor, ST_V); W2(OBCmask_u, ST_U); W2(OBCmask_v, ST_V);
#undef W2
  if ((rc = S.in(CSh->bathyT, ST_H, 1, 1, "cs.bathyT", &CS.bathyT))) return rc;
  if (CSh->ua_polarity || CSh->va_polarity) {
    // tripolar polarity reversal is outside the frozen option set: the arrays must be all +1; checked on the host copy
    int ilo, ihi, jlo, jhi; m6_extent(c, ST_H, 1, &ilo, &ihi, &jlo, &jhi);
    const size_t n = (size_t)(ihi - ilo + 1) * (jhi - jlo + 1);
    // (device-resident arrays are copied back and checked once per pointer; host arrays every call)
    const double* pol[2] = {CSh->ua_polarity, CSh->va_polarity};
    for (const double* p : pol) {
      if (!p) continue;
      const double* hp = p;
      if (m6_is_device_ptr(p)) {
        if (c->polarity_checked.count(p)) continue;
        double* tmp = c->host_scratch("bt.polarity", n);
        if (!tmp) return MOM6CU_ERR_CUDA;
        M6_CUDA(c, cudaMemcpyAsync(tmp, p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        M6_CUDA(c, cudaStreamSynchronize(c->stream));
        hp = tmp;
      }
      for (size_t q = 0; q < n; ++q)
        if (hp[q] < 0.0) return c->fail(MOM6CU_ERR_UNSUPPORTED, "btstep: reversed polarity (tripolar fold) is outside the frozen option set");
      if (hp != p) c->polarity_checked.insert(p);
    }
  }
  if ((rc = S.in3(CSh->frhatu, ST_U, "cs.frhatu", &CS.frhatu)) || (rc = S.in3(CSh->frhatv, ST_V, "cs.frhatv", &CS.frhatv)) ||
      (rc = S.in2(CSh->IDatu, ST_U, "cs.IDatu", &CS.IDatu)) || (rc = S.in2(CSh->IDatv, ST_V, "cs.IDatv", &CS.IDatv)) ||
      (rc = S.in2(CSh->eta_cor_bound, ST_H, "cs.eta_cor_bound", &CS.eta_cor_bound)) ||
      (rc = S.io2(CSh->eta_cor, ST_H, "cs.eta_cor", &CS.eta_cor)) || (rc = S.io2(CSh->ubtav, ST_U, "cs.ubtav", &CS.ubtav)) ||
      (rc = S.io2(CSh->vbtav, ST_V, "cs.vbtav", &CS.vbtav)))
    return rc;
  if (!CS.frhatu || !CS.frhatv || !CS.IDatu || !CS.IDatv || !CS.eta_cor || !CS.ubtav || !CS.vbtav)
    return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: a required control-structure array is null");
  return 0;
}

extern "C" int mom6cu_btstep(mom6cu_ctx* c, const mom6cu_barotropic_cs* CSh, const mom6cu_btstep_args* a) {
  if (!c || !CSh || !a) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!a->U_in || !a->V_in || !a->eta_in || !a->bc_accel_u || !a->bc_accel_v || !a->taux || !a->tauy || !a->pbce || !a->eta_PF_in ||
      !a->U_Cor || !a->V_Cor || !a->accel_layer_u || !a->accel_layer_v || !a->eta_out || !a->uhbtav || !a->vhbtav ||
      !a->visc_rem_u || !a->visc_rem_v)
    return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: null required argument");
  if (!a->BT_cont) return c->fail(MOM6CU_ERR_UNSUPPORTED, "btstep: USE_BT_CONT_TYPE=False is outside the frozen option set");
  Stager S(c, "btstep.");
  const int nk = c->g.nk;
  mom6cu_barotropic_cs CS = *CSh;
  BtstepDev D = {};
  D.dt = a->dt;
  int rc;
  if ((rc = m6_stage_barotropic_cs(c, S, CSh, &CS))) return rc;
  // arguments
  if ((rc = S.in3(a->U_in, ST_U, "U_in", &D.U_in)) || (rc = S.in3(a->V_in, ST_V, "V_in", &D.V_in)) ||
      (rc = S.in2(a->eta_in, ST_H, "eta_in", &D.eta_in)) || (rc = S.in3(a->bc_accel_u, ST_U, "bcu", &D.bc_accel_u)) ||
      (rc = S.in3(a->bc_accel_v, ST_V, "bcv", &D.bc_accel_v)) || (rc = S.in2(a->taux, ST_U, "taux", &D.taux)) ||
      (rc = S.in2(a->tauy, ST_V, "tauy", &D.tauy)) || (rc = S.in3(a->pbce, ST_H, "pbce", &D.pbce)) ||
      (rc = S.in2(a->eta_PF_in, ST_H, "eta_PF_in", &D.eta_PF_in)) || (rc = S.in3(a->U_Cor, ST_U, "U_Cor", &D.U_Cor)) ||
      (rc = S.in3(a->V_Cor, ST_V, "V_Cor", &D.V_Cor)) || (rc = S.in3(a->visc_rem_u, ST_U, "vru", &D.visc_rem_u)) ||
      (rc = S.in3(a->visc_rem_v, ST_V, "vrv", &D.visc_rem_v)) || (rc = S.in2(a->taux_bot, ST_U, "taux_bot", &D.taux_bot)) ||
      (rc = S.in2(a->tauy_bot, ST_V, "tauy_bot", &D.tauy_bot)) || (rc = S.in3(a->uh0, ST_U, "uh0", &D.uh0)) ||
      (rc = S.in3(a->vh0, ST_V, "vh0", &D.vh0)) || (rc = S.in3(a->u_uh0, ST_U, "u_uh0", &D.u_uh0)) ||
      (rc = S.in3(a->v_vh0, ST_V, "v_vh0", &D.v_vh0)) ||
      (rc = S.io3(a->accel_layer_u, ST_U, "accel_u", &D.accel_layer_u)) || (rc = S.io3(a->accel_layer_v, ST_V, "accel_v", &D.accel_layer_v)) ||
      (rc = S.io2(a->eta_out, ST_H, "eta_out", &D.eta_out)) || (rc = S.io2(a->uhbtav, ST_U, "uhbtav", &D.uhbtav)) ||
      (rc = S.io2(a->vhbtav, ST_V, "vhbtav", &D.vhbtav)) || (rc = S.io2(a->etaav, ST_H, "etaav", &D.etaav)))
    return rc;
  const mom6cu_bt_cont* B = a->BT_cont;
  D.have_BT_cont = 1;
  if ((rc = S.in2(B->FA_u_EE, ST_U, "FA_u_EE", &D.FA_u_EE)) || (rc = S.in2(B->FA_u_E0, ST_U, "FA_u_E0", &D.FA_u_E0)) ||
      (rc = S.in2(B->FA_u_W0, ST_U, "FA_u_W0", &D.FA_u_W0)) || (rc = S.in2(B->FA_u_WW, ST_U, "FA_u_WW", &D.FA_u_WW)) ||
      (rc = S.in2(B->uBT_WW, ST_U, "uBT_WW", &D.uBT_WW)) || (rc = S.in2(B->uBT_EE, ST_U, "uBT_EE", &D.uBT_EE)) ||
      (rc = S.in2(B->FA_v_NN, ST_V, "FA_v_NN", &D.FA_v_NN)) || (rc = S.in2(B->FA_v_N0, ST_V, "FA_v_N0", &D.FA_v_N0)) ||
      (rc = S.in2(B->FA_v_S0, ST_V, "FA_v_S0", &D.FA_v_S0)) || (rc = S.in2(B->FA_v_SS, ST_V, "FA_v_SS", &D.FA_v_SS)) ||
      (rc = S.in2(B->vBT_SS, ST_V, "vBT_SS", &D.vBT_SS)) || (rc = S.in2(B->vBT_NN, ST_V, "vBT_NN", &D.vBT_NN)))
    return rc;
  if (!D.FA_u_EE || !D.FA_u_E0 || !D.FA_u_W0 || !D.FA_u_WW || !D.uBT_WW || !D.uBT_EE || !D.FA_v_NN || !D.FA_v_N0 || !D.FA_v_S0 ||
      !D.FA_v_SS || !D.vBT_SS || !D.vBT_NN)
    return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: a BT_cont array is not allocated");
  if ((rc = S.begin())) return rc;
  if ((rc = m6_btstep_run(c, CS, D))) return rc;
  return S.finish();
}

extern "C" int mom6cu_btcalc(mom6cu_ctx* c, const mom6cu_btcalc_args* a) {
  if (!c || !a) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!a->h || !a->frhatu || !a->frhatv || !a->bathyT) return c->fail(MOM6CU_ERR_BAD_ARG, "btcalc: null required argument");
  Stager S(c, "btcalc.");
  const double *h, *hu, *hv, *bathyT; double *fu, *fv;
  int rc;
  if ((rc = S.in3(a->h, ST_H, "h", &h)) || (rc = S.in3(a->h_u, ST_U, "h_u", &hu)) || (rc = S.in3(a->h_v, ST_V, "h_v", &hv)) ||
      (rc = S.in2(a->bathyT, ST_H, "bathyT", &bathyT)) || (rc = S.io3(a->frhatu, ST_U, "frhatu", &fu)) ||
      (rc = S.io3(a->frhatv, ST_V, "frhatv", &fv)))
    return rc;
  if ((rc = S.begin())) return rc;
  if ((rc = m6_btcalc_run(c, h, hu, hv, bathyT, a->hvel_scheme, a->may_use_default, fu, fv))) return rc;
  return S.finish();
}

extern "C" int mom6cu_bt_mass_source(mom6cu_ctx* c, const double* h, const double* eta, int set_cor, double* eta_cor) {
  if (!c || !h || !eta || !eta_cor) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  Stager S(c, "btms.");
  const double *dh, *de; double* dc;
  int rc;
  if ((rc = S.in3(h, ST_H, "h", &dh)) || (rc = S.in2(eta, ST_H, "eta", &de)) || (rc = S.io2(eta_cor, ST_H, "eta_cor", &dc))) return rc;
  if ((rc = S.begin())) return rc;
  if ((rc = m6_bt_mass_source_run(c, dh, de, set_cor, dc))) return rc;
  return S.finish();
}

// ---------------------------------------------------------------------------------------------------------------
// set_dtbt, MOM_barotropic.F90:3509-3633.  One thread per h-point: the four k-ordered sums gtot_[EWNS] (:3593-3598), the
// face areas of its four faces (BT_cont_to_face_areas :5107 / find_face_areas :5146, halo 0) and Idt_max2 (:3611-3616).
// The reference then scans the points in memory order with `if (Idt_max2*min_max_dt2 > 1.) min_max_dt2 = 1./Idt_max2`,
// which is not exactly a minimum of 1/Idt_max2 in floating point, so that scan is done in the same order on the host
// from a pinned copy of the 2-D field (one 2-D download per call; set_dtbt runs once per coupling step at most).
namespace {
struct DtbtK {
  int is, ie, js, je, nz, mode;  // mode 0: BT_cont, 1: eta (Nonlinear_continuity), 2: bathymetry + add_max
  double gtot_est, Z_to_H, zadd, bebt, cor_scale2;
  const double *pbce, *frhatu, *frhatv, *bathyT, *eta;
  const double *EE, *E0, *W0, *WW, *NN, *N0, *S0, *SS;
  const double *dy_Cu, *dx_Cv, *IdxCu, *IdyCv, *IareaT, *Coriolis2Bu;
  double* out;  // packed (je-js+1) x (ie-is+1)
};
__device__ __forceinline__ double dtbt_face(const DtbtK& P, const Geom& G, long long g, bool uface) {
  const long long s = uface ? 1 : G.pitch;
  if (P.mode == 0) {
    const double* a = uface ? P.EE : P.NN; const double* b = uface ? P.E0 : P.N0; const double* c = uface ? P.W0 : P.S0;
    const double* d = uface ? P.WW : P.SS;
    return fmax2(fmax2(fmax2(a[g], b[g]), c[g]), d[g]);
  }
  const double len = uface ? P.dy_Cu[g] : P.dx_Cv[g];
  if (P.mode == 1) {
    const double H1 = P.bathyT[g] * P.Z_to_H + P.eta[g], H2 = P.bathyT[g + s] * P.Z_to_H + P.eta[g + s];
    return ((H1 > 0.0) && (H2 > 0.0)) ? len * (2.0 * H1 * H2) / (H1 + H2) : 0.0;
  }
  return len * P.Z_to_H * fmax2(fmax2(P.bathyT[g + s], P.bathyT[g]) + P.zadd, 0.0);
}
__global__ void set_dtbt_kernel(Geom G, DtbtK P) {
  const int i = P.is + blockIdx.x * blockDim.x + threadIdx.x, j = P.js + blockIdx.y;
  if (i > P.ie) return;
  const long long g = G.idx(i, j), pl = G.plane, pt = G.pitch;
  double gE = 0.0, gW = 0.0, gN = 0.0, gS = 0.0;
  if (P.pbce) {
    for (int k = 0; k < P.nz; ++k) {
      const long long o = (long long)k * pl + g;
      const double pb = P.pbce[o];
      gE = gE + pb * P.frhatu[o]; gW = gW + pb * P.frhatu[o - 1];
      gN = gN + pb * P.frhatv[o]; gS = gS + pb * P.frhatv[o - pt];
    }
  } else { gE = gW = gN = gS = P.gtot_est; }
  const double DuE = dtbt_face(P, G, g, true), DuW = dtbt_face(P, G, g - 1, true);
  const double DvN = dtbt_face(P, G, g, false), DvS = dtbt_face(P, G, g - pt, false);
  const double* C2 = P.Coriolis2Bu;
  const double Idt_max2 = 0.5 * (1.0 + 2.0 * P.bebt) * (P.IareaT[g] *
      (((gE * DuE * P.IdxCu[g]) + (gW * DuW * P.IdxCu[g - 1])) + ((gN * DvN * P.IdyCv[g]) + (gS * DvS * P.IdyCv[g - pt]))) +
      ((C2[g] + C2[g - 1 - pt]) + (C2[g - 1] + C2[g - pt])) * P.cor_scale2);
  P.out[(size_t)(j - P.js) * (P.ie - P.is + 1) + (i - P.is)] = Idt_max2;
}
}  // namespace

// device-level: every array pointer is a resident plane
int m6_set_dtbt_run(mom6cu_ctx* c, const mom6cu_set_dtbt_args& a, double* dtbt, double* dtbt_max) {
  const Geom& G = c->g;
  const mom6cu_domain& d = c->dom;
  const int ni = d.iec - d.isc + 1, nj = d.jec - d.jsc + 1;
  const size_t n = (size_t)ni * nj;
  double* dev = c->buf("dtbt.Idt", n);
  double* host = c->host_scratch("dtbt.Idt", n);
  if (!dev || !host) return MOM6CU_ERR_CUDA;
  DtbtK P = {};
  P.is = d.isc; P.ie = d.iec; P.js = d.jsc; P.je = d.jec; P.nz = G.nk;
  P.gtot_est = a.gtot_est; P.Z_to_H = c->vgrid.Z_to_H; P.zadd = a.Z_ref + a.SSH_add; P.bebt = a.bebt;
  P.cor_scale2 = a.BT_Coriolis_scale * a.BT_Coriolis_scale;
  P.pbce = a.pbce; P.frhatu = a.frhatu; P.frhatv = a.frhatv; P.bathyT = a.bathyT; P.eta = a.eta;
  if (a.BT_cont) {
    const mom6cu_bt_cont* B = a.BT_cont;
    P.mode = 0; P.EE = B->FA_u_EE; P.E0 = B->FA_u_E0; P.W0 = B->FA_u_W0; P.WW = B->FA_u_WW; P.NN = B->FA_v_NN; P.N0 = B->FA_v_N0;
    P.S0 = B->FA_v_S0; P.SS = B->FA_v_SS;
  } else if (a.Nonlinear_continuity && a.eta) P.mode = 1;
  else P.mode = 2;
  P.dy_Cu = c->grid.dy_Cu; P.dx_Cv = c->grid.dx_Cv; P.IdxCu = c->grid.IdxCu; P.IdyCv = c->grid.IdyCv; P.IareaT = c->grid.IareaT;
  P.Coriolis2Bu = c->grid.Coriolis2Bu; P.out = dev;
  M6_LAUNCH(c, set_dtbt_kernel, dim3((ni + 127) / 128, nj), 128, 0, G, P);
  M6_CUDA(c, cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  double min_max_dt2 = 1.0e38 * (c->US.s_to_T * c->US.s_to_T);
  for (size_t q = 0; q < n; ++q) if (host[q] * min_max_dt2 > 1.0) min_max_dt2 = 1.0 / host[q];
  const double dgeo_de = 1.0 + std::max(0.0, a.G_extra - 0.0);
  double mx = std::sqrt(min_max_dt2 / dgeo_de);
  int rc;
  if (c->nranks > 1 && (rc = m6_allreduce_min_double(c, &mx))) return rc;  // min_across_PEs :3621
  *dtbt = a.dtbt_fraction * mx;
  *dtbt_max = mx;
  return 0;
}

extern "C" int mom6cu_set_dtbt(mom6cu_ctx* c, const mom6cu_set_dtbt_args* a, double* dtbt, double* dtbt_max) {
  if (!c || !a || !dtbt || !dtbt_max) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "set_dtbt: mom6cu_set_grid / mom6cu_set_vgrid have not been called");
  if (!(a->pbce || a->have_gtot_est)) return c->fail(MOM6CU_ERR_BAD_ARG, "set_dtbt: Either pbce or gtot_est must be present.");
  if (!c->vgrid.Boussinesq) return c->fail(MOM6CU_ERR_UNSUPPORTED, "set_dtbt: non-Boussinesq face areas are not implemented");
  if (a->pbce && (!a->frhatu || !a->frhatv)) return c->fail(MOM6CU_ERR_BAD_ARG, "set_dtbt: CS%%frhatu / frhatv are null");
  if (!a->bathyT) return c->fail(MOM6CU_ERR_BAD_ARG, "set_dtbt: CS%%bathyT is null");
  Stager S(c, "dtbt.");
  mom6cu_set_dtbt_args D = *a;
  mom6cu_bt_cont B = {};
  int rc;
  if ((rc = S.in3(a->pbce, ST_H, "pbce", &D.pbce)) || (rc = S.in3(a->frhatu, ST_U, "frhatu", &D.frhatu)) ||
      (rc = S.in3(a->frhatv, ST_V, "frhatv", &D.frhatv)) || (rc = S.in2(a->bathyT, ST_H, "bathyT", &D.bathyT)) ||
      (rc = S.in2(a->eta, ST_H, "eta", &D.eta))) return rc;
  if (a->BT_cont) {
    const mom6cu_bt_cont* H = a->BT_cont;
    const double* p;
#define FA(f, st) if (!H->f) return c->fail(MOM6CU_ERR_BAD_ARG, "set_dtbt: BT_cont%%" #f " is not allocated"); if ((rc = S.in2(H->f, st, #f, &p))) return rc; B.f = (double*)p
    FA(FA_u_EE, ST_U); FA(FA_u_E0, ST_U); FA(FA_u_W0, ST_U); FA(FA_u_WW, ST_U); FA(FA_v_NN, ST_V); FA(FA_v_N0, ST_V); FA(FA_v_S0, ST_V); FA(FA_v_SS, ST_V);
#undef FA
    D.BT_cont = &B;
  }
  if ((rc = S.begin())) return rc;
  if ((rc = m6_set_dtbt_run(c, D, dtbt, dtbt_max))) return rc;
  return S.finish();
}
