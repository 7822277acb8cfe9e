// horizontal_viscosity for sm_100a: ONE fused kernel for the whole routine.
//
// Replaces src/parameterizations/lateral/MOM_hor_visc.F90:266-2317 (the `do k` body :707-2218), which is ~30
// separate 2-D sweeps per layer in the reference (dudx, dvdy, sh_xx, dvdx, dudy, sh_xy, h_u, h_v, Del2u, Del2v,
// Shear_mag, hrat_min, Kh, Ah, str_xx, dDel2vdx, dDel2udy, hq, str_xy, diffu, diffv ...).
//
// Design (DESIGN.md "K12"):
//  * One CTA owns a TX x TY tile of one layer and walks the 4-stage stencil chain
//        (u,v,h) -> (sh_xx, sh_xy, h_u, h_v) -> (Del2u, Del2v) -> (str_xx, str_xy) -> (diffu, diffv)
//    entirely in shared memory on a (TX+5) x (TY+5) extended tile (the biharmonic operator reaches 2 points
//    either way).  3 arrays are read from HBM and 2 written; none of the reference's temporaries exists in HBM.
//  * Layers are independent (the reference's OpenMP loop, :669).  blockIdx.x runs over k so the CTAs resident at
//    one time share the same few (i,j) tiles and the ~45 metric / coefficient planes are served by L2.
//  * Every phase is restricted to the index range of the corresponding reference loop and keeps its
//    parenthesisation (no FMA contraction), so halo values recomputed by neighbouring CTAs are bit-identical.
#include "ctx.h"
#include <cstdlib>
#include "stage.h"
#include <cmath>

using m6::Geom;
using m6::fmax2;
using m6::fmin2;

namespace {

struct HorViscK {
  mom6cu_hor_visc_cs CS;  // flags + device pointers of the control-structure arrays
  double h_neglect, h_neglect3;
  int use_cont_huv;
  const double *u, *v, *h, *hu_cont, *hv_cont;
  double *diffu, *diffv;
  GridDev M;
};

__device__ __forceinline__ double min4(double a, double b, double c, double d) { return fmin2(fmin2(fmin2(a, b), c), d); }
__device__ __forceinline__ bool rng(int v, int lo, int hi) { return v >= lo && v <= hi; }

template <int TX, int TY>
__global__ void __launch_bounds__(TX* TY) hor_visc_kernel(const Geom G, const HorViscK K) {
  constexpr int NT = TX * TY, EW = TX + 5, EH = TY + 5, EN = EW * EH, NP = (EN + NT - 1) / NT;
  extern __shared__ double sm[];
  double* su = sm;             // u(I,j)
  double* sv = su + EN;        // v(i,J)
  double* sh = sv + EN;        // h(i,j)
  double* shu = sh + EN;       // h_u(I,j)
  double* shv = shu + EN;      // h_v(i,J)
  double* sxx = shv + EN;      // sh_xx(i,j)
  double* sxy = sxx + EN;      // sh_xy(I,J)
  double* sd2u = sxy + EN;     // Del2u(I,j)
  double* sd2v = sd2u + EN;    // Del2v(i,J)
  double* stxx = sd2v + EN;    // str_xx(i,j)
  double* stxy = stxx + EN;    // str_xy(I,J)
  const mom6cu_hor_visc_cs& CS = K.CS;
  const int k = blockIdx.x;
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec;
  const int Isq = is - 1, Ieq = ie, Jsq = js - 1, Jeq = je;
  const int ti0 = Isq + blockIdx.y * TX, tj0 = Jsq + blockIdx.z * TY;
  const long long koff = (long long)k * G.plane;
  const long long P = G.pitch;
  const bool smag = CS.Smagorinsky_Kh || CS.Smagorinsky_Ah;
  const bool bbound = CS.better_bound_Ah || CS.better_bound_Kh;
  const bool legacy_bound = CS.Smagorinsky_Kh && (CS.bound_Kh && !CS.better_bound_Kh);

  // the extended-tile points this thread owns
  int pe[NP], pi[NP], pj[NP];
  long long pg[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const int p = threadIdx.x + q * NT;
    pe[q] = (p < EN) ? p : -1;
    const int ey = p / EW, ex = p - ey * EW;
    pi[q] = ti0 - 2 + ex; pj[q] = tj0 - 2 + ey;
    pg[q] = G.idx(pi[q], pj[q]);
  }

  // ---- phase A: stage u, v, h (zero outside the memory domain)
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    if (pe[q] < 0) continue;
    const int i = pi[q], j = pj[q];
    const long long g = pg[q] + koff;
    const bool ini = rng(i, G.isd, G.ied), inj = rng(j, G.jsd, G.jed);
    su[pe[q]] = (rng(i, G.isd - 1, G.ied) && inj) ? __ldg(K.u + g) : 0.0;
    sv[pe[q]] = (ini && rng(j, G.jsd - 1, G.jed)) ? __ldg(K.v + g) : 0.0;
    sh[pe[q]] = (ini && inj) ? __ldg(K.h + g) : 0.0;
  }
  __syncthreads();

  // ---- phase B: tension, shearing strain, thicknesses at velocity points (:720-785, :909-919)
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const int e = pe[q];
    if (e < 0) continue;
    const int i = pi[q], j = pj[q];
    const long long g = pg[q];
    const int ex = e % EW, ey = e / EW;
    double v_xx = 0.0, v_xy = 0.0, v_hu = 0.0, v_hv = 0.0;
    if (ex >= 1 && ey >= 1 && rng(i, Isq - 1, Ieq + 2) && rng(j, Jsq - 1, Jeq + 2)) {
      const double dudx = __ldg(CS.DY_dxT + g) * ((__ldg(K.M.IdyCu + g) * su[e]) - (__ldg(K.M.IdyCu + g - 1) * su[e - 1]));
      const double dvdy = __ldg(CS.DX_dyT + g) * ((__ldg(K.M.IdxCv + g) * sv[e]) - (__ldg(K.M.IdxCv + g - P) * sv[e - EW]));
      v_xx = dudx - dvdy;
    }
    if (ex <= EW - 2 && ey <= EH - 2 && rng(i, is - 2, Ieq + 1) && rng(j, js - 2, Jeq + 1)) {
      const double dvdx = __ldg(CS.DY_dxBu + g) * ((sv[e + 1] * __ldg(K.M.IdyCv + g + 1)) - (sv[e] * __ldg(K.M.IdyCv + g)));
      const double dudy = __ldg(CS.DX_dyBu + g) * ((su[e + EW] * __ldg(K.M.IdxCu + g + P)) - (su[e] * __ldg(K.M.IdxCu + g)));
      const double mB = __ldg(K.M.mask2dBu + g);
      v_xy = CS.no_slip ? (2.0 - mB) * (dvdx + dudy) : mB * (dvdx + dudy);
    }
    if (K.use_cont_huv) {
      if (rng(i, Isq - 1, Ieq + 1) && rng(j, js - 2, je + 2)) v_hu = __ldg(K.hu_cont + g + koff);
      if (rng(i, is - 2, ie + 2) && rng(j, Jsq - 1, Jeq + 1)) v_hv = __ldg(K.hv_cont + g + koff);
    } else {
      if (ex <= EW - 2 && rng(i, is - 2, Ieq + 1) && rng(j, js - 2, je + 2)) {
        if (CS.use_land_mask) v_hu = 0.5 * (__ldg(K.M.mask2dT + g) * sh[e] + __ldg(K.M.mask2dT + g + 1) * sh[e + 1]);
        else v_hu = 0.5 * (sh[e] + sh[e + 1]);
      }
      if (ey <= EH - 2 && rng(i, is - 2, ie + 2) && rng(j, js - 2, Jeq + 1)) {
        if (CS.use_land_mask) v_hv = 0.5 * (__ldg(K.M.mask2dT + g) * sh[e] + __ldg(K.M.mask2dT + g + P) * sh[e + EW]);
        else v_hv = 0.5 * (sh[e] + sh[e + EW]);
      }
    }
    sxx[e] = v_xx; sxy[e] = v_xy; shu[e] = v_hu; shv[e] = v_hv;
  }
  __syncthreads();

  // ---- phase C: Del2u, Del2v (:936-944)
  if (CS.biharmonic) {
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const int e = pe[q];
      if (e < 0) continue;
      const int i = pi[q], j = pj[q];
      const long long g = pg[q];
      const int ex = e % EW, ey = e / EW;
      double d2u = 0.0, d2v = 0.0;
      if (ex >= 1 && ex <= EW - 2 && ey >= 1 && ey <= EH - 2) {
        if (rng(i, Isq - 1, Ieq + 1) && rng(j, js - 1, Jeq + 1))
          d2u = __ldg(CS.Idx2dyCu + g) * ((__ldg(CS.dx2q + g) * sxy[e]) - (__ldg(CS.dx2q + g - P) * sxy[e - EW])) +
                __ldg(CS.Idxdy2u + g) * ((__ldg(CS.dy2h + g + 1) * sxx[e + 1]) - (__ldg(CS.dy2h + g) * sxx[e]));
        if (rng(i, is - 1, Ieq + 1) && rng(j, Jsq - 1, Jeq + 1))
          d2v = __ldg(CS.Idxdy2v + g) * ((__ldg(CS.dy2q + g) * sxy[e]) - (__ldg(CS.dy2q + g - 1) * sxy[e - 1])) -
                __ldg(CS.Idx2dyCv + g) * ((__ldg(CS.dx2h + g + P) * sxx[e + EW]) - (__ldg(CS.dx2h + g) * sxx[e]));
      }
      sd2u[e] = d2u; sd2v[e] = d2v;
    }
    __syncthreads();
  }

  // ---- phase D: viscosities and stresses at h points (:1114-1458) and at q points (:1490-1924)
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const int e = pe[q];
    if (e < 0) continue;
    const int i = pi[q], j = pj[q];
    const long long g = pg[q];
    const int ex = e % EW, ey = e / EW;
    double s_xx = 0.0, s_xy = 0.0;
    // h point (i,j): str_xx on Isq:Ieq+1, Jsq:Jeq+1 (is_Kh = Isq, ie_Kh = ie+1 without Leith)
    if (ex >= 2 && ex <= EW - 2 && ey >= 2 && ey <= EH - 2 && rng(i, Isq, Ieq + 1) && rng(j, Jsq, Jeq + 1)) {
      double Shear_mag = 0.0, hrat_min = 0.0, visc_bound_rem = 1.0;
      if (smag) {
        const double sh_xx_sq = sxx[e] * sxx[e];
        const double sh_xy_sq = 0.25 * (((sxy[e - EW - 1] * sxy[e - EW - 1]) + (sxy[e] * sxy[e])) +
                                        ((sxy[e - 1] * sxy[e - 1]) + (sxy[e - EW] * sxy[e - EW])));
        Shear_mag = sqrt(sh_xx_sq + sh_xy_sq);
      }
      if (bbound) {
        const double h_min = min4(shu[e], shu[e - 1], shv[e], shv[e - EW]);
        hrat_min = fmin2(1.0, h_min / (sh[e] + K.h_neglect));
      }
      if (CS.Laplacian) {
        double Kh = __ldg(CS.Kh_bg_xx + g);
        if (CS.Smagorinsky_Kh) {
          if (CS.add_LES_viscosity) Kh = Kh + __ldg(CS.Laplac2_const_xx + g) * Shear_mag;
          else Kh = fmax2(Kh, __ldg(CS.Laplac2_const_xx + g) * Shear_mag);
        }
        if (legacy_bound) Kh = fmin2(Kh, __ldg(CS.Kh_Max_xx + g));
        Kh = fmax2(Kh, CS.Kh_bg_min);
        if (CS.better_bound_Kh && CS.better_bound_Ah) {
          const double Kh_max_here = hrat_min * __ldg(CS.Kh_Max_xx + g);
          if (Kh >= Kh_max_here) { visc_bound_rem = 0.0; Kh = Kh_max_here; }
          else if ((Kh > 0.0) || (CS.backscatter_underbound && (Kh_max_here > 0.0))) visc_bound_rem = 1.0 - Kh / Kh_max_here;
        } else if (CS.better_bound_Kh) {
          Kh = fmin2(Kh, hrat_min * __ldg(CS.Kh_Max_xx + g));
        }
        s_xx = -Kh * sxx[e];
      }
      if (CS.biharmonic) {
        double Ah = __ldg(CS.Ah_bg_xx + g);
        if (CS.Smagorinsky_Ah) {
          double AhSm;
          if (CS.bound_Coriolis) AhSm = Shear_mag * (__ldg(CS.Biharm_const_xx + g) + __ldg(CS.Biharm_const2_xx + g) * Shear_mag);
          else AhSm = __ldg(CS.Biharm_const_xx + g) * Shear_mag;
          Ah = fmax2(Ah, AhSm);
          if (CS.bound_Ah && !CS.better_bound_Ah) Ah = fmin2(Ah, __ldg(CS.Ah_Max_xx + g));
        }
        if (CS.Re_Ah > 0.0) {
          const double s1 = su[e] + su[e - 1], s2 = sv[e] + sv[e - EW];
          const double KE = 0.125 * ((s1 * s1) + (s2 * s2));
          Ah = sqrt(KE) * __ldg(CS.Re_Ah_const_xx + g);
        }
        if (CS.better_bound_Ah) {
          if (CS.better_bound_Kh) Ah = fmin2(Ah, visc_bound_rem * hrat_min * __ldg(CS.Ah_Max_xx + g));
          else Ah = fmin2(Ah, hrat_min * __ldg(CS.Ah_Max_xx + g));
        }
        const double d_del2u = (__ldg(K.M.IdyCu + g) * sd2u[e]) - (__ldg(K.M.IdyCu + g - 1) * sd2u[e - 1]);
        const double d_del2v = (__ldg(K.M.IdxCv + g) * sd2v[e]) - (__ldg(K.M.IdxCv + g - P) * sd2v[e - EW]);
        const double d_str = Ah * ((__ldg(CS.DY_dxT + g) * d_del2u) - (__ldg(CS.DX_dyT + g) * d_del2v));
        s_xx = s_xx + d_str;
      }
      s_xx = s_xx * (sh[e] * __ldg(CS.reduction_xx + g));
    }
    // q point (I,J): str_xy on is-1:Ieq, js-1:Jeq
    if (ex >= 1 && ex <= EW - 3 && ey >= 1 && ey <= EH - 3 && rng(i, is - 1, Ieq) && rng(j, js - 1, Jeq)) {
      double Shear_mag = 0.0, hrat_min = 0.0, visc_bound_rem = 1.0;
      if (smag) {
        const double sh_xy_sq = sxy[e] * sxy[e];
        const double sh_xx_sq = 0.25 * (((sxx[e] * sxx[e]) + (sxx[e + EW + 1] * sxx[e + EW + 1])) +
                                        ((sxx[e + EW] * sxx[e + EW]) + (sxx[e + 1] * sxx[e + 1])));
        Shear_mag = sqrt(sh_xy_sq + sh_xx_sq);
      }
      const double hu0 = shu[e], hu1 = shu[e + EW], hv0 = shv[e], hv1 = shv[e + 1];
      const double h2uq = 4.0 * (hu0 * hu1);
      const double h2vq = 4.0 * (hv0 * hv1);
      double hq = (2.0 * (h2uq * h2vq)) / (K.h_neglect3 + (h2uq + h2vq) * ((hu0 + hu1) + (hv0 + hv1)));
      if (bbound) {
        const double h_min = min4(hu0, hu1, hv0, hv1);
        hrat_min = fmin2(1.0, h_min / (hq + K.h_neglect));
      }
      const double mB = __ldg(K.M.mask2dBu + g);
      if (CS.no_slip && (mB < 0.5)) {
        const double mu0 = __ldg(K.M.mask2dCu + g), mu1 = __ldg(K.M.mask2dCu + g + P);
        const double mv0 = __ldg(K.M.mask2dCv + g), mv1 = __ldg(K.M.mask2dCv + g + 1);
        if ((mu0 + mu1) + (mv0 + mv1) > 0.0) {
          const double hu = mu0 * hu0 + mu1 * hu1;
          const double hv = mv0 * hv0 + mv1 * hv1;
          if ((mu0 + mu1) * (mv0 + mv1) == 0.0) { hq = hu + hv; hrat_min = 1.0; }
          else { hq = 2.0 * (hu * hv) / ((hu + hv) + K.h_neglect); hrat_min = fmin2(1.0, fmin2(hu, hv) / (hq + K.h_neglect)); }
        }
      }
      if (CS.Laplacian) {
        double Kh = __ldg(CS.Kh_bg_xy + g);
        if (CS.Smagorinsky_Kh) {
          if (CS.add_LES_viscosity) Kh = Kh + __ldg(CS.Laplac2_const_xy + g) * Shear_mag;
          else Kh = fmax2(Kh, __ldg(CS.Laplac2_const_xy + g) * Shear_mag);
        }
        if (legacy_bound) Kh = fmin2(Kh, __ldg(CS.Kh_Max_xy + g));
        Kh = fmax2(Kh, CS.Kh_bg_min);
        if (CS.better_bound_Kh && CS.better_bound_Ah) {
          const double Kh_max_here = hrat_min * __ldg(CS.Kh_Max_xy + g);
          if (Kh >= Kh_max_here) { visc_bound_rem = 0.0; Kh = Kh_max_here; }
          else if ((Kh > 0.0) || (CS.backscatter_underbound && (Kh_max_here > 0.0))) visc_bound_rem = 1.0 - Kh / Kh_max_here;
        } else if (CS.better_bound_Kh) {
          Kh = fmin2(Kh, hrat_min * __ldg(CS.Kh_Max_xy + g));
        }
        s_xy = -Kh * sxy[e];
      }
      if (CS.biharmonic) {
        double Ah = __ldg(CS.Ah_bg_xy + g);
        if (CS.Smagorinsky_Ah) {
          double AhSm;
          if (CS.bound_Coriolis) AhSm = Shear_mag * (__ldg(CS.Biharm_const_xy + g) + __ldg(CS.Biharm_const2_xy + g) * Shear_mag);
          else AhSm = __ldg(CS.Biharm_const_xy + g) * Shear_mag;
          Ah = fmax2(Ah, AhSm);
          if (CS.bound_Ah && !CS.better_bound_Ah) Ah = fmin2(Ah, __ldg(CS.Ah_Max_xy + g));
        }
        if (CS.Re_Ah > 0.0) {
          const double s1 = su[e] + su[e + EW], s2 = sv[e] + sv[e + 1];
          const double KE = 0.125 * ((s1 * s1) + (s2 * s2));
          Ah = sqrt(KE) * __ldg(CS.Re_Ah_const_xy + g);
        }
        if (CS.better_bound_Ah) {
          if (CS.better_bound_Kh) Ah = fmin2(Ah, visc_bound_rem * hrat_min * __ldg(CS.Ah_Max_xy + g));
          else Ah = fmin2(Ah, hrat_min * __ldg(CS.Ah_Max_xy + g));
        }
        const double dDel2vdx = __ldg(CS.DY_dxBu + g) * ((sd2v[e + 1] * __ldg(K.M.IdyCv + g + 1)) - (sd2v[e] * __ldg(K.M.IdyCv + g)));
        const double dDel2udy = __ldg(CS.DX_dyBu + g) * ((sd2u[e + EW] * __ldg(K.M.IdxCu + g + P)) - (sd2u[e] * __ldg(K.M.IdxCu + g)));
        const double d_str = Ah * (dDel2vdx + dDel2udy);
        s_xy = s_xy + d_str;
      }
      if (CS.no_slip) s_xy = s_xy * (hq * __ldg(CS.reduction_xy + g));
      else s_xy = s_xy * (hq * mB * __ldg(CS.reduction_xy + g));
    }
    stxx[e] = s_xx; stxy[e] = s_xy;
  }
  __syncthreads();

  // ---- phase E: accelerations (:1929-1954)
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int i = ti0 + tx, j = tj0 + ty;
  if (i > Ieq || j > Jeq) return;
  const int e = (ty + 2) * EW + (tx + 2);
  const long long g = G.idx(i, j), gk = g + koff;
  if (j >= js) {
    K.diffu[gk] = ((__ldg(K.M.IdxCu + g) * ((__ldg(CS.dx2q + g - P) * stxy[e - EW]) - (__ldg(CS.dx2q + g) * stxy[e])) +
                    __ldg(K.M.IdyCu + g) * ((__ldg(CS.dy2h + g) * stxx[e]) - (__ldg(CS.dy2h + g + 1) * stxx[e + 1]))) *
                   __ldg(K.M.IareaCu + g)) / (shu[e] + K.h_neglect);
  }
  if (i >= is) {
    K.diffv[gk] = ((__ldg(K.M.IdyCv + g) * ((__ldg(CS.dy2q + g - 1) * stxy[e - 1]) - (__ldg(CS.dy2q + g) * stxy[e])) -
                    __ldg(K.M.IdxCv + g) * ((__ldg(CS.dx2h + g) * stxx[e]) - (__ldg(CS.dx2h + g + P) * stxx[e + EW]))) *
                   __ldg(K.M.IareaCv + g)) / (shv[e] + K.h_neglect);
  }
}

constexpr int HV_TX = 32, HV_TY = 16;
constexpr size_t HV_SMEM = (size_t)11 * (HV_TX + 5) * (HV_TY + 5) * sizeof(double);

}  // namespace

int m6_hor_visc_run(mom6cu_ctx* c, const HorViscDev& D) {
  if (!c->have_hv_cs) return c->fail(MOM6CU_ERR_BAD_ARG, "MOM_hor_visc: Module must be initialized before it is used.");
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "horizontal_viscosity: grid / vertical grid not set");
  const mom6cu_hor_visc_cs& S = c->hv_cs;
  if (!(S.Laplacian || S.biharmonic)) return 0;  // :507
  const mom6cu_domain& d = c->dom;
  if ((d.isc - d.isd) < 2 || (d.jsc - d.jsd) < 2)
    return c->fail(MOM6CU_ERR_BAD_ARG, "horizontal_viscosity needs u(is-2:ie+2,...): halo must be at least 2 wide");
  HorViscK K;
  K.CS = c->hv_cs_dev;
  K.h_neglect = c->vgrid.H_subroundoff;
  K.h_neglect3 = K.h_neglect * K.h_neglect * K.h_neglect;
  K.use_cont_huv = (S.use_cont_thick && D.hu_cont && D.hv_cont) ? 1 : 0;
  K.u = D.u; K.v = D.v; K.h = D.h; K.hu_cont = D.hu_cont; K.hv_cont = D.hv_cont; K.diffu = D.diffu; K.diffv = D.diffv;
  K.M = c->grid;
  const int nI = d.iec - (d.isc - 1) + 1, nJ = d.jec - (d.jsc - 1) + 1;
  static int ty = -1;  // tile height: 16 (512 threads, 68 KB smem) or 8 (256 threads, 42 KB); MOM6CU_HV_TY overrides
  if (ty < 0) { const char* e = getenv("MOM6CU_HV_TY"); ty = e ? atoi(e) : 8; }  // measured: 13.4 ms with 32x8 tiles, 14.0 with 32x16
  if (ty == 8) {
    constexpr size_t SM8 = (size_t)11 * (HV_TX + 5) * (8 + 5) * sizeof(double);
    auto kern = hor_visc_kernel<HV_TX, 8>;
    static bool attr8 = false;
    if (!attr8) { M6_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM8)); attr8 = true; }
    dim3 grid(c->g.nk, (nI + HV_TX - 1) / HV_TX, (nJ + 8 - 1) / 8);
    M6_LAUNCH(c, kern, grid, HV_TX * 8, SM8, c->g, K);
  } else {
    auto kern = hor_visc_kernel<HV_TX, HV_TY>;
    static bool attr_set = false;
    if (!attr_set) { M6_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HV_SMEM)); attr_set = true; }
    dim3 grid(c->g.nk, (nI + HV_TX - 1) / HV_TX, (nJ + HV_TY - 1) / HV_TY);
    M6_LAUNCH(c, kern, grid, HV_TX * HV_TY, HV_SMEM, c->g, K);
  }
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

// hor_visc_init's products, uploaded once and kept resident
extern "C" int mom6cu_set_cs_hor_visc(mom6cu_ctx* c, const mom6cu_hor_visc_cs* CS) {
  if (!c || !CS) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (CS->unsupported)
    return c->fail(MOM6CU_ERR_UNSUPPORTED, "hor_visc_init: Leith/GME/MEKE/anisotropic/ZB2020/resolution-scaled viscosities and "
                                           "OBCs are outside the frozen option set of this build");
  c->hv_cs = *CS;
  c->hv_cs_dev = *CS;
  struct F { const char* name; const double* src; int st; const double** dst; bool need; };
  mom6cu_hor_visc_cs& D = c->hv_cs_dev;
  const bool lap = CS->Laplacian, bih = CS->biharmonic;
  const bool kmax = lap && (CS->better_bound_Kh || (CS->Smagorinsky_Kh && CS->bound_Kh));
  const bool amax = bih && (CS->better_bound_Ah || (CS->Smagorinsky_Ah && CS->bound_Ah));
  const F f[MOM6CU_HOR_VISC_NARRAYS] = {
      {"dx2h", CS->dx2h, ST_H, &D.dx2h, true}, {"dy2h", CS->dy2h, ST_H, &D.dy2h, true},
      {"DX_dyT", CS->DX_dyT, ST_H, &D.DX_dyT, true}, {"DY_dxT", CS->DY_dxT, ST_H, &D.DY_dxT, true},
      {"reduction_xx", CS->reduction_xx, ST_H, &D.reduction_xx, true},
      {"Kh_bg_xx", CS->Kh_bg_xx, ST_H, &D.Kh_bg_xx, lap}, {"Ah_bg_xx", CS->Ah_bg_xx, ST_H, &D.Ah_bg_xx, bih},
      {"Kh_Max_xx", CS->Kh_Max_xx, ST_H, &D.Kh_Max_xx, kmax}, {"Ah_Max_xx", CS->Ah_Max_xx, ST_H, &D.Ah_Max_xx, amax},
      {"Laplac2_const_xx", CS->Laplac2_const_xx, ST_H, &D.Laplac2_const_xx, lap && CS->Smagorinsky_Kh},
      {"Biharm_const_xx", CS->Biharm_const_xx, ST_H, &D.Biharm_const_xx, bih && CS->Smagorinsky_Ah},
      {"Biharm_const2_xx", CS->Biharm_const2_xx, ST_H, &D.Biharm_const2_xx, bih && CS->Smagorinsky_Ah && CS->bound_Coriolis},
      {"Re_Ah_const_xx", CS->Re_Ah_const_xx, ST_H, &D.Re_Ah_const_xx, bih && CS->Re_Ah > 0.0},
      {"dx2q", CS->dx2q, ST_Q, &D.dx2q, true}, {"dy2q", CS->dy2q, ST_Q, &D.dy2q, true},
      {"DX_dyBu", CS->DX_dyBu, ST_Q, &D.DX_dyBu, true}, {"DY_dxBu", CS->DY_dxBu, ST_Q, &D.DY_dxBu, true},
      {"reduction_xy", CS->reduction_xy, ST_Q, &D.reduction_xy, true},
      {"Kh_bg_xy", CS->Kh_bg_xy, ST_Q, &D.Kh_bg_xy, lap}, {"Ah_bg_xy", CS->Ah_bg_xy, ST_Q, &D.Ah_bg_xy, bih},
      {"Kh_Max_xy", CS->Kh_Max_xy, ST_Q, &D.Kh_Max_xy, kmax}, {"Ah_Max_xy", CS->Ah_Max_xy, ST_Q, &D.Ah_Max_xy, amax},
      {"Laplac2_const_xy", CS->Laplac2_const_xy, ST_Q, &D.Laplac2_const_xy, lap && CS->Smagorinsky_Kh},
      {"Biharm_const_xy", CS->Biharm_const_xy, ST_Q, &D.Biharm_const_xy, bih && CS->Smagorinsky_Ah},
      {"Biharm_const2_xy", CS->Biharm_const2_xy, ST_Q, &D.Biharm_const2_xy, bih && CS->Smagorinsky_Ah && CS->bound_Coriolis},
      {"Re_Ah_const_xy", CS->Re_Ah_const_xy, ST_Q, &D.Re_Ah_const_xy, bih && CS->Re_Ah > 0.0},
      {"Idx2dyCu", CS->Idx2dyCu, ST_U, &D.Idx2dyCu, bih}, {"Idxdy2u", CS->Idxdy2u, ST_U, &D.Idxdy2u, bih},
      {"Idx2dyCv", CS->Idx2dyCv, ST_V, &D.Idx2dyCv, bih}, {"Idxdy2v", CS->Idxdy2v, ST_V, &D.Idxdy2v, bih}};
  for (int m = 0; m < MOM6CU_HOR_VISC_NARRAYS; ++m) {
    *f[m].dst = nullptr;
    if (!f[m].src) {
      if (f[m].need) return c->fail(MOM6CU_ERR_BAD_ARG, "mom6cu_set_cs_hor_visc: CS%%%s is required by the selected options", f[m].name);
      continue;
    }
    double* p = c->plane2(std::string("HV.") + f[m].name);
    if (!p) return MOM6CU_ERR_CUDA;
    int rc = m6_up(c, f[m].src, f[m].st, 0, 1, p);
    if (rc) return rc;
    *f[m].dst = p;
  }
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  c->have_hv_cs = true;
  return 0;
}

extern "C" int mom6cu_horizontal_viscosity(mom6cu_ctx* c, const mom6cu_hor_visc_args* a) {
  if (!c || !a) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!a->u || !a->v || !a->h || !a->diffu || !a->diffv)
    return c->fail(MOM6CU_ERR_BAD_ARG, "horizontal_viscosity: null required argument");
  Stager S(c, "hv.");
  HorViscDev D = {};
  int rc;
  if ((rc = S.in3(a->u, ST_U, "u", &D.u)) || (rc = S.in3(a->v, ST_V, "v", &D.v)) || (rc = S.in3(a->h, ST_H, "h", &D.h)) ||
      (rc = S.in3(a->hu_cont, ST_U, "hu_cont", &D.hu_cont)) || (rc = S.in3(a->hv_cont, ST_V, "hv_cont", &D.hv_cont)) ||
      (rc = S.io3(a->diffu, ST_U, "diffu", &D.diffu)) || (rc = S.io3(a->diffv, ST_V, "diffv", &D.diffv)))
    return rc;
  if ((rc = S.begin())) return rc;
  if ((rc = m6_hor_visc_run(c, D))) return rc;
  return S.finish();
}
