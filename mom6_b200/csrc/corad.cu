// CorAdCalc for sm_100a: ONE fused kernel for the whole routine.
//
// Replaces src/core/MOM_CoriolisAdv.F90: CorAdCalc :125-965 and gradKE :969-1051, which are ~25 separate
// 2-D sweeps per layer in the reference (dvdx, dudy, hArea_u/v, rel_vort, abs_vort, q, a/b/c/d, KE, KEx,
// KEy, CAu, CAv, ...), each a round trip through memory.
//
// Design (DESIGN.md "K10"):
//  * One CTA owns a TX x TY tile of one layer.  Phase 1 builds the potential vorticity q (and Ih_q /
//    abs_vort where the selected scheme needs them) on the (TX+2)x(TY+2) q-points around the tile and
//    the kinetic energy KE on (TX+1)x(TY+1) h-points in shared memory; phase 2 turns them into CAu and
//    CAv.  None of the reference's 2-D temporaries ever exists in HBM: 5 arrays in, 2 out.
//  * Layers are independent (the reference's OpenMP loop, :281).  blockIdx.x runs over k so that the
//    CTAs resident at one time share the same few (i,j) tiles: the ~15 metric planes they read are then
//    served by L2 and cost HBM traffic once per tile, not once per layer.
//  * Every expression keeps the reference's parenthesisation (no FMA contraction), so halo q/KE values
//    recomputed by neighbouring CTAs are bit-identical.
#include "ctx.h"
#include <cstdlib>
#include "stage.h"
#include <cmath>

using m6::Geom;
using m6::fmax2;
using m6::fmin2;

namespace {

struct CorAdK {
  mom6cu_coriolisadv_cs CS;
  double vol_neglect, eps_vel, h_tiny;
  const double *u, *v, *h, *uh, *vh, *porU, *porV;
  double *CAu, *CAv, *RV, *PV, *gKEu, *gKEv;
  GridDev M;
};

__device__ __forceinline__ double max4(double a, double b, double c, double d) { return fmax2(fmax2(fmax2(a, b), c), d); }
__device__ __forceinline__ double min4(double a, double b, double c, double d) { return fmin2(fmin2(fmin2(a, b), c), d); }

// uh_min/uh_max of CORIOLIS_EN_DIS (:594-614) at the u-point offset g; DIR 0: zonal, 1: meridional (:615-635)
template <int DIR>
__device__ __forceinline__ void en_dis_minmax(const CorAdK& K, const Geom& G, long long g, long long gk, double& fmin_,
                                              double& fmax_) {
  const double c1 = 1.0 - 1.5 * 0.5, c2 = 1.0 - 0.5, c3 = 2.0, slope = 0.5;
  const long long sd = DIR == 0 ? 1 : G.pitch;
  const double* vel = DIR == 0 ? K.u : K.v;
  const double* por = DIR == 0 ? K.porU : K.porV;
  const double dy = __ldg((DIR == 0 ? K.M.dy_Cu : K.M.dx_Cv) + g);
  const double p = por ? __ldg(por + gk) : 1.0;
  double uhc = 0.5 * ((dy * p) * __ldg(vel + gk)) * (__ldg(K.h + gk) + __ldg(K.h + gk + sd));
  double uhm = __ldg((DIR == 0 ? K.uh : K.vh) + gk);
  if (dy == 0.0) uhc = uhm;
  if (fabs(uhc) < 0.1 * fabs(uhm)) uhm = 10.0 * uhc;
  else if (fabs(uhc) > c1 * fabs(uhm)) {
    if (fabs(uhc) < c2 * fabs(uhm)) uhc = (3.0 * uhc + (1.0 - c2 * 3.0) * uhm);
    else if (fabs(uhc) <= c3 * fabs(uhm)) uhc = uhm;
    else uhc = slope * uhc + (1.0 - c3 * slope) * uhm;
  }
  if (uhc > uhm) { fmin_ = uhm; fmax_ = uhc; }
  else { fmax_ = uhm; fmin_ = uhc; }
}

template <int TX, int TY, int MINB = 3>
__global__ void __launch_bounds__(TX* TY, MINB) corad_kernel(const Geom G, const CorAdK K) {
  constexpr int NT = TX * TY, QW = TX + 2, QH = TY + 2, QN = QW * QH, KW = TX + 1, KH = TY + 1, KN = KW * KH;
  __shared__ double sq[QN];    // q(I,J),   I = ti0-1 .. ti0+TX, J = tj0-1 .. tj0+TY
  __shared__ double saux[QN];  // Ih_q (AL_BLEND) or abs_vort (ROBUST_ENSTRO / bound_Coriolis)
  __shared__ double sKE[KN];   // KE(i,j),  i = ti0 .. ti0+TX,  j = tj0 .. tj0+TY
  const int k = blockIdx.x;
  const int Isq = G.isc - 1, Ieq = G.iec, Jsq = G.jsc - 1, Jeq = G.jec;
  const int ti0 = Isq + blockIdx.y * TX, tj0 = Jsq + blockIdx.z * TY;
  const long long koff = (long long)k * G.plane;
  const double* __restrict__ u = K.u + koff;
  const double* __restrict__ v = K.v + koff;
  const double* __restrict__ h = K.h + koff;
  const int scheme = K.CS.Coriolis_Scheme;
  const bool want_absv = (scheme == MOM6CU_ROBUST_ENSTRO) || K.CS.bound_Coriolis;
  const long long P = G.pitch;

  // ---- phase 1a: q on the extended tile (:314-324, :459-491)
  for (int p = threadIdx.x; p < QN; p += NT) {
    const int qy = p / QW, qx = p - qy * QW;
    const int I = ti0 - 1 + qx, J = tj0 - 1 + qy;
    double qv = 0.0, aux = 0.0;
    if (I >= Isq - 1 && I <= Ieq + 1 && J >= Jsq - 1 && J <= Jeq + 1) {
      const long long g = G.idx(I, J);
      const double dvdx = (__ldg(v + g + 1) * __ldg(K.M.dyCv + g + 1)) - (__ldg(v + g) * __ldg(K.M.dyCv + g));
      const double dudy = (__ldg(u + g + P) * __ldg(K.M.dxCu + g + P)) - (__ldg(u + g) * __ldg(K.M.dxCu + g));
      const double A00 = __ldg(K.M.mask2dT + g) * __ldg(K.M.areaT + g);
      const double A10 = __ldg(K.M.mask2dT + g + 1) * __ldg(K.M.areaT + g + 1);
      const double A01 = __ldg(K.M.mask2dT + g + P) * __ldg(K.M.areaT + g + P);
      const double A11 = __ldg(K.M.mask2dT + g + P + 1) * __ldg(K.M.areaT + g + P + 1);
      const double h00 = __ldg(h + g), h10 = __ldg(h + g + 1), h01 = __ldg(h + g + P), h11 = __ldg(h + g + P + 1);
      const double hArea_v0 = 0.5 * ((A00 * h00) + (A01 * h01));  // hArea_v(i,J)
      const double hArea_v1 = 0.5 * ((A10 * h10) + (A11 * h11));  // hArea_v(i+1,J)
      const double hArea_u0 = 0.5 * ((A00 * h00) + (A10 * h10));  // hArea_u(I,j)
      const double hArea_u1 = 0.5 * ((A01 * h01) + (A11 * h11));  // hArea_u(I,j+1)
      const double Area_q = (A00 + A11) + (A10 + A01);
      const double mB = __ldg(K.M.mask2dBu + g);
      double rel_vort;
      if (K.CS.no_slip) rel_vort = (2.0 - mB) * (dvdx - dudy) * __ldg(K.M.IareaBu + g);
      else rel_vort = mB * (dvdx - dudy) * __ldg(K.M.IareaBu + g);
      const double abs_vort = __ldg(K.M.CoriolisBu + g) + rel_vort;
      const double hArea_q = (hArea_u0 + hArea_u1) + (hArea_v0 + hArea_v1);
      const double Ih_q = Area_q / (hArea_q + K.vol_neglect);
      qv = abs_vort * Ih_q;
      aux = want_absv ? abs_vort : Ih_q;
      // diagnostics are owned by the CTA whose tile holds the point; the first / last tiles also own
      // the extra row and column of the reference's Isq-1:Ieq+1 range
      const bool own_x = (qx >= 1 && qx <= TX) || (qx == 0 && blockIdx.y == 0) || (qx == TX + 1 && blockIdx.y == gridDim.y - 1);
      const bool own_y = (qy >= 1 && qy <= TY) || (qy == 0 && blockIdx.z == 0) || (qy == TY + 1 && blockIdx.z == gridDim.z - 1);
      if ((K.RV || K.PV) && own_x && own_y) {
        if (K.RV) K.RV[g + koff] = rel_vort;
        if (K.PV) K.PV[g + koff] = qv;
      }
    }
    sq[p] = qv;
    saux[p] = aux;
  }
  // ---- phase 1b: KE (gradKE :995-1025)
  for (int p = threadIdx.x; p < KN; p += NT) {
    const int ky = p / KW, kx = p - ky * KW;
    const int i = ti0 + kx, j = tj0 + ky;
    double KE = 0.0;
    if (i >= Isq && i <= Ieq + 1 && j >= Jsq && j <= Jeq + 1) {
      const long long g = G.idx(i, j);
      const double u0 = __ldg(u + g), um = __ldg(u + g - 1), v0 = __ldg(v + g), vm = __ldg(v + g - P);
      if (K.CS.KE_Scheme == MOM6CU_KE_ARAKAWA) {
        KE = (((__ldg(K.M.areaCu + g) * (u0 * u0)) + (__ldg(K.M.areaCu + g - 1) * (um * um))) +
              ((__ldg(K.M.areaCv + g) * (v0 * v0)) + (__ldg(K.M.areaCv + g - P) * (vm * vm)))) *
             0.25 * __ldg(K.M.IareaT + g);
      } else if (K.CS.KE_Scheme == MOM6CU_KE_SIMPLE_GUDONOV) {
        const double up = 0.5 * (um + fabs(um)), up2 = up * up;
        const double umm = 0.5 * (u0 - fabs(u0)), um2 = umm * umm;
        const double vp = 0.5 * (vm + fabs(vm)), vp2 = vp * vp;
        const double vmm = 0.5 * (v0 - fabs(v0)), vm2 = vmm * vmm;
        KE = (fmax2(up2, um2) + fmax2(vp2, vm2)) * 0.5;
      } else {
        const double up = 0.5 * (um + fabs(um)), up2a = up * up * __ldg(K.M.areaCu + g - 1);
        const double umm = 0.5 * (u0 - fabs(u0)), um2a = umm * umm * __ldg(K.M.areaCu + g);
        const double vp = 0.5 * (vm + fabs(vm)), vp2a = vp * vp * __ldg(K.M.areaCv + g - P);
        const double vmm = 0.5 * (v0 - fabs(v0)), vm2a = vmm * vmm * __ldg(K.M.areaCv + g);
        KE = (fmax2(um2a, up2a) + fmax2(vm2a, vp2a)) * 0.5 * __ldg(K.M.IareaT + g);
      }
    }
    sKE[p] = KE;
  }
  __syncthreads();

  // ---- phase 2: one (I=i, j) u-point and one (i, J=j) v-point per thread
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int i = ti0 + tx, j = tj0 + ty;
  if (i > Ieq || j > Jeq) return;
  const long long g = G.idx(i, j), gk = g + koff;
  const double* __restrict__ uh = K.uh + koff;
  const double* __restrict__ vh = K.vh + koff;
  const int qc = (ty + 1) * QW + (tx + 1);  // q(I=i, J=j)
#define Q(di, dj) sq[qc + (dj)*QW + (di)]
#define AUX(di, dj) saux[qc + (dj)*QW + (di)]
  const double C1_12 = 1.0 / 12.0, C1_24 = 1.0 / 24.0;
  // a, b, c, d, ep_u, ep_v of :523-588 at arbitrary offsets, rebuilt from q in shared memory
  double Fe_m2 = 0.0, rat_lin = 0.0;
  if (scheme == MOM6CU_AL_BLEND) {
    Fe_m2 = K.CS.F_eff_max_blend - 2.0;
    rat_lin = 1.5 * Fe_m2 / fmax2(K.CS.wt_lin_blend, 1.0e-16);
    if (K.CS.F_eff_max_blend <= 2.0) { Fe_m2 = -1.; rat_lin = -1.0; }
  }
  // blend weights of the h-cell whose NE corner is q(I+di, J+dj)  (:551-571)
  auto blend = [&](int di, int dj, double& AL_wt, double& Sad_wt) {
    const double i00 = AUX(di - 1, dj - 1), i10 = AUX(di, dj - 1), i01 = AUX(di - 1, dj), i11 = AUX(di, dj);
    const double min_Ihq = min4(i00, i10, i01, i11), max_Ihq = max4(i00, i10, i01, i11);
    double rat_m1 = 1.0e15;
    if (max_Ihq < 1.0e15 * min_Ihq) rat_m1 = max_Ihq / min_Ihq - 1.0;
    if (rat_m1 <= Fe_m2) AL_wt = 1.0;
    else if (rat_m1 < 1.5 * Fe_m2) AL_wt = 3.0 * Fe_m2 / rat_m1 - 2.0;
    else AL_wt = 0.0;
    if (rat_m1 <= 1.5 * Fe_m2) Sad_wt = 0.0;
    else if (rat_m1 <= rat_lin) Sad_wt = 1.0 - (1.5 * Fe_m2) / rat_m1;
    else if (rat_m1 < 2.0 * rat_lin) Sad_wt = 1.0 - (K.CS.wt_lin_blend / rat_lin) * (rat_m1 - 2.0 * rat_lin);
    else Sad_wt = 1.0;
  };
  // The coefficient arrays use the h-cell (ic,jc) = cell whose NE corner is q(I+di,J+dj):
  //   a(Ic-1,jc), d(Ic-1,jc), b(Ic,jc), c(Ic,jc), ep_u(ic,jc), ep_v(ic,jc)
  enum { CA, CB, CC, CD, CEU, CEV };
  auto coef = [&](int which, int di, int dj) -> double {
    const double q11 = Q(di, dj), q00 = Q(di - 1, dj - 1), q01 = Q(di - 1, dj), q10 = Q(di, dj - 1);
    if (scheme == MOM6CU_ARAKAWA_LAMB81) {
      switch (which) {
        case CA: return (2.0 * (q11 + q00) + (q01 + q10)) * C1_24;
        case CD: return ((q11 + q00) + 2.0 * (q01 + q10)) * C1_24;
        case CB: return ((q11 + q00) + 2.0 * (q01 + q10)) * C1_24;
        case CC: return (2.0 * (q11 + q00) + (q01 + q10)) * C1_24;
        case CEU: return ((q11 - q00) + (q01 - q10)) * C1_24;
        default: return (-(q11 - q00) + (q01 - q10)) * C1_24;
      }
    } else {  // AL_BLEND
      double AL_wt, Sad_wt;
      blend(di, dj, AL_wt, Sad_wt);
      switch (which) {
        case CA: return Sad_wt * 0.25 * q01 + (1.0 - Sad_wt) * (((2.0 - AL_wt) * q01 + AL_wt * q10) + 2.0 * (q11 + q00)) * C1_24;
        case CD: return Sad_wt * 0.25 * q00 + (1.0 - Sad_wt) * (((2.0 - AL_wt) * q00 + AL_wt * q11) + 2.0 * (q01 + q10)) * C1_24;
        case CB: return Sad_wt * 0.25 * q11 + (1.0 - Sad_wt) * (((2.0 - AL_wt) * q11 + AL_wt * q00) + 2.0 * (q01 + q10)) * C1_24;
        case CC: return Sad_wt * 0.25 * q10 + (1.0 - Sad_wt) * (((2.0 - AL_wt) * q10 + AL_wt * q01) + 2.0 * (q11 + q00)) * C1_24;
        case CEU: return AL_wt * ((q11 - q00) + (q01 - q10)) * C1_24;
        default: return AL_wt * (-(q11 - q00) + (q01 - q10)) * C1_24;
      }
    }
  };
  const bool AL = (scheme == MOM6CU_ARAKAWA_LAMB81) || (scheme == MOM6CU_AL_BLEND);

  // ---- CAu(I,j), j >= js  (:644-758)
  if (j >= G.jsc) {
    const double IdxCu = __ldg(K.M.IdxCu + g);
    const double vh00 = __ldg(vh + g), vh10 = __ldg(vh + g + 1), vh0m = __ldg(vh + g - P), vh1m = __ldg(vh + g - P + 1);
    double CAu = 0.0;
    if (scheme == MOM6CU_SADOURNY75_ENERGY) {
      if (K.CS.Coriolis_En_Dis) {
        const double uk = __ldg(u + g);
        double vmin00, vmax00, vmin10, vmax10, vmin0m, vmax0m, vmin1m, vmax1m;
        en_dis_minmax<1>(K, G, g, gk, vmin00, vmax00);
        en_dis_minmax<1>(K, G, g + 1, gk + 1, vmin10, vmax10);
        en_dis_minmax<1>(K, G, g - P, gk - P, vmin0m, vmax0m);
        en_dis_minmax<1>(K, G, g - P + 1, gk - P + 1, vmin1m, vmax1m);
        double temp1, temp2;
        const double qJ = Q(0, 0), qJm = Q(0, -1);
        if (qJ * uk == 0.0) temp1 = qJ * ((vmax00 + vmax10) + (vmin00 + vmin10)) * 0.5;
        else if (qJ * uk < 0.0) temp1 = qJ * (vmax00 + vmax10);
        else temp1 = qJ * (vmin00 + vmin10);
        if (qJm * uk == 0.0) temp2 = qJm * ((vmax0m + vmax1m) + (vmin0m + vmin1m)) * 0.5;
        else if (qJm * uk < 0.0) temp2 = qJm * (vmax0m + vmax1m);
        else temp2 = qJm * (vmin0m + vmin1m);
        CAu = 0.25 * IdxCu * (temp1 + temp2);
      } else {
        CAu = 0.25 * ((Q(0, 0) * (vh10 + vh00)) + (Q(0, -1) * (vh0m + vh1m))) * IdxCu;
      }
    } else if (scheme == MOM6CU_SADOURNY75_ENSTRO) {
      CAu = 0.125 * (IdxCu * (Q(0, 0) + Q(0, -1))) * ((vh10 + vh00) + (vh0m + vh1m));
    } else if (scheme == MOM6CU_ARAKAWA_HSU90) {
      const double a = (Q(0, 0) + (Q(1, 0) + Q(0, -1))) * C1_12;
      const double d = ((Q(0, 0) + Q(1, -1)) + Q(0, -1)) * C1_12;
      const double b = (Q(0, 0) + (Q(-1, 0) + Q(0, -1))) * C1_12;
      const double c = ((Q(0, 0) + Q(-1, -1)) + Q(0, -1)) * C1_12;
      CAu = (((a * vh10) + (c * vh0m)) + ((b * vh00) + (d * vh1m))) * IdxCu;
    } else if (AL) {
      // a(I,j), d(I,j) belong to the cell (i+1,j) [NE corner q(I+1,J)]; b(I,j), c(I,j) to the cell (i,j)
      const double a = coef(CA, 1, 0), d = coef(CD, 1, 0), b = coef(CB, 0, 0), c = coef(CC, 0, 0);
      CAu = (((a * vh10) + (c * vh0m)) + ((b * vh00) + (d * vh1m))) * IdxCu;
    } else if (scheme == MOM6CU_ROBUST_ENSTRO) {
      const double h00 = __ldg(h + g), h01 = __ldg(h + g + P), h0m = __ldg(h + g - P);
      const double h10 = __ldg(h + g + 1), h11 = __ldg(h + g + P + 1), h1m = __ldg(h + g - P + 1);
      double Heff1 = fabs(vh00 * __ldg(K.M.IdxCv + g)) / (K.eps_vel + fabs(__ldg(v + g)));
      Heff1 = fmax2(Heff1, fmin2(h00, h01)); Heff1 = fmin2(Heff1, fmax2(h00, h01));
      double Heff2 = fabs(vh0m * __ldg(K.M.IdxCv + g - P)) / (K.eps_vel + fabs(__ldg(v + g - P)));
      Heff2 = fmax2(Heff2, fmin2(h0m, h00)); Heff2 = fmin2(Heff2, fmax2(h0m, h00));
      double Heff3 = fabs(vh10 * __ldg(K.M.IdxCv + g + 1)) / (K.eps_vel + fabs(__ldg(v + g + 1)));
      Heff3 = fmax2(Heff3, fmin2(h10, h11)); Heff3 = fmin2(Heff3, fmax2(h10, h11));
      double Heff4 = fabs(vh1m * __ldg(K.M.IdxCv + g - P + 1)) / (K.eps_vel + fabs(__ldg(v + g - P + 1)));
      Heff4 = fmax2(Heff4, fmin2(h1m, h10)); Heff4 = fmin2(Heff4, fmax2(h1m, h10));
      const double av0 = AUX(0, 0), avm = AUX(0, -1);
      if (K.CS.PV_Adv_Scheme == MOM6CU_PV_ADV_CENTERED) {
        CAu = 0.5 * (av0 + avm) * ((vh00 + vh1m) + (vh0m + vh10)) / (K.h_tiny + ((Heff1 + Heff4) + (Heff2 + Heff3))) * IdxCu;
      } else {
        const double VHeff = ((vh00 + vh1m) + (vh0m + vh10));
        const double QVHeff = 0.5 * (((av0 + avm) * VHeff) - ((av0 - avm) * fabs(VHeff)));
        CAu = (QVHeff / (K.h_tiny + ((Heff1 + Heff4) + (Heff2 + Heff3)))) * IdxCu;
      }
    }
    if (AL) CAu = CAu + ((coef(CEU, 0, 0) * __ldg(uh + g - 1)) - (coef(CEU, 1, 0) * __ldg(uh + g + 1))) * IdxCu;
    if (K.CS.bound_Coriolis) {
      const double av0 = AUX(0, 0), avm = AUX(0, -1);
      const double fv1 = av0 * __ldg(v + g + 1), fv2 = av0 * __ldg(v + g);
      const double fv3 = avm * __ldg(v + g - P + 1), fv4 = avm * __ldg(v + g - P);
      CAu = fmin2(CAu, max4(fv1, fv2, fv3, fv4));
      CAu = fmax2(CAu, min4(fv1, fv2, fv3, fv4));
    }
    const double KEx = (sKE[ty * KW + tx + 1] - sKE[ty * KW + tx]) * IdxCu;
    K.CAu[gk] = CAu - KEx;
    if (K.gKEu) K.gKEu[gk] = -KEx;
  }

  // ---- CAv(i,J), i >= is  (:763-881)
  if (i >= G.isc) {
    const double IdyCv = __ldg(K.M.IdyCv + g);
    const double uh00 = __ldg(uh + g), uhm0 = __ldg(uh + g - 1), uh01 = __ldg(uh + g + P), uhm1 = __ldg(uh + g + P - 1);
    double CAv = 0.0;
    if (scheme == MOM6CU_SADOURNY75_ENERGY) {
      if (K.CS.Coriolis_En_Dis) {
        const double vk = __ldg(v + g);
        double umin00, umax00, uminm0, umaxm0, umin01, umax01, uminm1, umaxm1;
        en_dis_minmax<0>(K, G, g, gk, umin00, umax00);
        en_dis_minmax<0>(K, G, g - 1, gk - 1, uminm0, umaxm0);
        en_dis_minmax<0>(K, G, g + P, gk + P, umin01, umax01);
        en_dis_minmax<0>(K, G, g + P - 1, gk + P - 1, uminm1, umaxm1);
        double temp1, temp2;
        const double qm = Q(-1, 0), q0 = Q(0, 0);
        if (qm * vk == 0.0) temp1 = qm * ((umaxm0 + umaxm1) + (uminm0 + uminm1)) * 0.5;
        else if (qm * vk > 0.0) temp1 = qm * (umaxm0 + umaxm1);
        else temp1 = qm * (uminm0 + uminm1);
        if (q0 * vk == 0.0) temp2 = q0 * ((umax00 + umax01) + (umin00 + umin01)) * 0.5;
        else if (q0 * vk > 0.0) temp2 = q0 * (umax00 + umax01);
        else temp2 = q0 * (umin00 + umin01);
        CAv = -0.25 * IdyCv * (temp1 + temp2);
      } else {
        CAv = -0.25 * ((Q(-1, 0) * (uhm0 + uhm1)) + (Q(0, 0) * (uh00 + uh01))) * IdyCv;
      }
    } else if (scheme == MOM6CU_SADOURNY75_ENSTRO) {
      CAv = -0.125 * (IdyCv * (Q(-1, 0) + Q(0, 0))) * ((uhm0 + uhm1) + (uh00 + uh01));
    } else if (scheme == MOM6CU_ARAKAWA_HSU90) {
      // a(I-1,j), c(I,j+1), b(I,j), d(I-1,j+1) with the :524-532 definitions
      const double a = (Q(-1, 0) + (Q(0, 0) + Q(-1, -1))) * C1_12;
      const double c = ((Q(0, 1) + Q(-1, 0)) + Q(0, 0)) * C1_12;
      const double b = (Q(0, 0) + (Q(-1, 0) + Q(0, -1))) * C1_12;
      const double d = ((Q(-1, 1) + Q(0, 0)) + Q(-1, 0)) * C1_12;
      CAv = -(((a * uhm0) + (c * uh01)) + ((b * uh00) + (d * uhm1))) * IdyCv;
    } else if (AL) {
      // a(I-1,j): cell (i,j); c(I,j+1): cell (i,j+1); b(I,j): cell (i,j); d(I-1,j+1): cell (i,j+1)
      const double a = coef(CA, 0, 0), c = coef(CC, 0, 1), b = coef(CB, 0, 0), d = coef(CD, 0, 1);
      CAv = -(((a * uhm0) + (c * uh01)) + ((b * uh00) + (d * uhm1))) * IdyCv;
    } else if (scheme == MOM6CU_ROBUST_ENSTRO) {
      const double h00 = __ldg(h + g), h10 = __ldg(h + g + 1), hm0 = __ldg(h + g - 1);
      const double h01 = __ldg(h + g + P), h11 = __ldg(h + g + P + 1), hm1 = __ldg(h + g + P - 1);
      double Heff1 = fabs(uh00 * __ldg(K.M.IdyCu + g)) / (K.eps_vel + fabs(__ldg(u + g)));
      Heff1 = fmax2(Heff1, fmin2(h00, h10)); Heff1 = fmin2(Heff1, fmax2(h00, h10));
      double Heff2 = fabs(uhm0 * __ldg(K.M.IdyCu + g - 1)) / (K.eps_vel + fabs(__ldg(u + g - 1)));
      Heff2 = fmax2(Heff2, fmin2(hm0, h00)); Heff2 = fmin2(Heff2, fmax2(hm0, h00));
      double Heff3 = fabs(uh01 * __ldg(K.M.IdyCu + g + P)) / (K.eps_vel + fabs(__ldg(u + g + P)));
      Heff3 = fmax2(Heff3, fmin2(h01, h11)); Heff3 = fmin2(Heff3, fmax2(h01, h11));
      double Heff4 = fabs(uhm1 * __ldg(K.M.IdyCu + g + P - 1)) / (K.eps_vel + fabs(__ldg(u + g + P - 1)));
      Heff4 = fmax2(Heff4, fmin2(hm1, h01)); Heff4 = fmin2(Heff4, fmax2(hm1, h01));
      const double av0 = AUX(0, 0), avm = AUX(-1, 0);
      if (K.CS.PV_Adv_Scheme == MOM6CU_PV_ADV_CENTERED) {
        CAv = -0.5 * (av0 + avm) * ((uh00 + uhm1) + (uhm0 + uh01)) / (K.h_tiny + ((Heff1 + Heff4) + (Heff2 + Heff3))) * IdyCv;
      } else {
        const double UHeff = ((uh00 + uhm1) + (uhm0 + uh01));
        const double QUHeff = 0.5 * (((av0 + avm) * UHeff) - ((av0 - avm) * fabs(UHeff)));
        CAv = -QUHeff / (K.h_tiny + ((Heff1 + Heff4) + (Heff2 + Heff3))) * IdyCv;
      }
    }
    if (AL) CAv = CAv + ((coef(CEV, 0, 0) * __ldg(vh + g - P)) - (coef(CEV, 0, 1) * __ldg(vh + g + P))) * IdyCv;
    if (K.CS.bound_Coriolis) {
      const double av0 = AUX(0, 0), avm = AUX(-1, 0);
      const double fu1 = -av0 * __ldg(u + g + P), fu2 = -av0 * __ldg(u + g);
      const double fu3 = -avm * __ldg(u + g + P - 1), fu4 = -avm * __ldg(u + g - 1);
      CAv = fmin2(CAv, max4(fu1, fu2, fu3, fu4));
      CAv = fmax2(CAv, min4(fu1, fu2, fu3, fu4));
    }
    const double KEy = (sKE[(ty + 1) * KW + tx] - sKE[ty * KW + tx]) * IdyCv;
    K.CAv[gk] = CAv - KEy;
    if (K.gKEv) K.gKEv[gk] = -KEy;
  }
#undef Q
#undef AUX
}

constexpr int CA_TX = 32, CA_TY = 8;

}  // namespace

int m6_coradcalc_run(mom6cu_ctx* c, const CorAdDev& D) {
  if (!c->have_corad_cs) return c->fail(MOM6CU_ERR_BAD_ARG, "MOM_CoriolisAdv: Module must be initialized before it is used.");
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "CorAdCalc: grid / vertical grid not set");
  const mom6cu_domain& d = c->dom;
  if ((d.isc - d.isd) < 2 || (d.jsc - d.jsd) < 2)
    return c->fail(MOM6CU_ERR_BAD_ARG, "CorAdCalc needs h(is-1:ie+2,js-1:je+2): halo must be at least 2 wide");
  const mom6cu_coriolisadv_cs& S = c->corad_cs;
  if (S.Coriolis_Scheme < MOM6CU_SADOURNY75_ENERGY || S.Coriolis_Scheme > MOM6CU_AL_BLEND)
    return c->fail(MOM6CU_ERR_BAD_ARG, "CoriolisAdv_init: Unrecognized setting of CORIOLIS_SCHEME");
  if (S.KE_Scheme < MOM6CU_KE_ARAKAWA || S.KE_Scheme > MOM6CU_KE_GUDONOV)
    return c->fail(MOM6CU_ERR_BAD_ARG, "CoriolisAdv_init: #define KE_SCHEME in input file is invalid.");
  CorAdK K;
  K.CS = S;
  // CoriolisAdv_init :1111, :1158-1160: these combinations are forced off
  if (S.Coriolis_Scheme == MOM6CU_ROBUST_ENSTRO) K.CS.Coriolis_En_Dis = 0;
  if ((K.CS.Coriolis_En_Dis && S.Coriolis_Scheme == MOM6CU_SADOURNY75_ENERGY) || S.Coriolis_Scheme == MOM6CU_ROBUST_ENSTRO)
    K.CS.bound_Coriolis = 0;
  const double m_to_L = c->US.m_to_L;
  K.vol_neglect = c->vgrid.H_subroundoff * ((1e-4 * m_to_L) * (1e-4 * m_to_L));
  K.eps_vel = 1.0e-10 * c->US.m_s_to_L_T;
  K.h_tiny = c->vgrid.Angstrom_H;
  K.u = D.u; K.v = D.v; K.h = D.h; K.uh = D.uh; K.vh = D.vh; K.porU = D.por_face_areaU; K.porV = D.por_face_areaV;
  K.CAu = D.CAu; K.CAv = D.CAv; K.RV = D.RV; K.PV = D.PV; K.gKEu = D.gradKEu; K.gKEv = D.gradKEv;
  K.M = c->grid;
  const int nI = d.iec - (d.isc - 1) + 1, nJ = d.jec - (d.jsc - 1) + 1;
  dim3 grid(c->g.nk, (nI + CA_TX - 1) / CA_TX, (nJ + CA_TY - 1) / CA_TY);
  // 4 CTAs/SM (64 registers, a few spills to L1) measured 4.37 ms against 5.24 ms for 3 CTAs/SM (80 registers) at 1440x1080x75
  // (profiles/r02_corad_minb.log); MOM6CU_CORAD_MINB=3 selects the latter for A/B measurements
  static int minb = -1;
  if (minb < 0) { const char* e = getenv("MOM6CU_CORAD_MINB"); minb = e ? atoi(e) : 4; }
  if (minb == 4) M6_LAUNCH(c, (corad_kernel<CA_TX, CA_TY, 4>), grid, CA_TX * CA_TY, 0, c->g, K);
  else M6_LAUNCH(c, (corad_kernel<CA_TX, CA_TY>), grid, CA_TX * CA_TY, 0, c->g, K);
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

extern "C" int mom6cu_set_cs_coriolisadv(mom6cu_ctx* c, const mom6cu_coriolisadv_cs* CS) {
  if (!c || !CS) return MOM6CU_ERR_BAD_ARG;
  c->corad_cs = *CS;
  c->have_corad_cs = true;
  return 0;
}

extern "C" int mom6cu_coradcalc(mom6cu_ctx* c, const mom6cu_coradcalc_args* a) {
  if (!c || !a) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!a->u || !a->v || !a->h || !a->uh || !a->vh || !a->CAu || !a->CAv)
    return c->fail(MOM6CU_ERR_BAD_ARG, "CorAdCalc: null required argument");
  Stager S(c, "corad.");
  CorAdDev D = {};
  int rc;
  if ((rc = S.in3(a->u, ST_U, "u", &D.u)) || (rc = S.in3(a->v, ST_V, "v", &D.v)) || (rc = S.in3(a->h, ST_H, "h", &D.h)) ||
      (rc = S.in3(a->uh, ST_U, "uh", &D.uh)) || (rc = S.in3(a->vh, ST_V, "vh", &D.vh)) ||
      (rc = S.in3(a->por_face_areaU, ST_U, "porU", &D.por_face_areaU)) ||
      (rc = S.in3(a->por_face_areaV, ST_V, "porV", &D.por_face_areaV)) ||
      (rc = S.io3(a->CAu, ST_U, "CAu", &D.CAu)) || (rc = S.io3(a->CAv, ST_V, "CAv", &D.CAv)) ||
      (rc = S.io3(a->RV, ST_Q, "RV", &D.RV)) || (rc = S.io3(a->PV, ST_Q, "PV", &D.PV)) ||
      (rc = S.io3(a->gradKEu, ST_U, "gKEu", &D.gradKEu)) || (rc = S.io3(a->gradKEv, ST_V, "gKEv", &D.gradKEv)))
    return rc;
  if ((rc = S.begin())) return rc;
  if ((rc = m6_coradcalc_run(c, D))) return rc;
  return S.finish();
}
