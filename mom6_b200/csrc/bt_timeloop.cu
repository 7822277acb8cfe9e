// Barotropic substep loop for sm_100a: ONE fused kernel per forward-backward substep.
//
// Replaces btstep_timeloop and its helpers, src/core/MOM_barotropic.F90:2175-2832
// (btloop_eta_predictor :2956, btloop_find_PF :3063, btloop_update_v :3209,
// btloop_update_u :3306, find_uhbt/find_vhbt :4610/:4744), which are ~12 separate 2-D
// sweeps per substep in the reference.
//
// Design (DESIGN.md "K3"):
//  * One CTA owns a TX x TY tile of (i,j) points.  The stencil chain of a substep
//    (transport -> eta_pred -> pressure force -> v -> u -> transport -> eta) reaches
//    2 points to the south/west and 1 to the north/east, so the CTA recomputes the
//    chain on a (TX+3) x (TY+3) extended tile held in shared memory; nothing but the
//    final state and the accumulators goes back to HBM.  Neighbouring CTAs recompute
//    the same halo points with the same instruction sequence (no FMA contraction),
//    so the redundant values are bit-identical.
//  * State (eta, ubt, vbt) is ping-ponged between two sets of planes so that a CTA
//    never reads a neighbour's half-updated halo; every point of the plane is written
//    each substep (untouched points are copied through), so the planes stay complete.
//  * Every phase is restricted to exactly the index range of the corresponding
//    reference loop (the shrinking "valid" region isv:iev, jsv:jev of the wide-halo
//    scheme, :2509-2518), so all arrays are bitwise identical to the reference's, halo
//    points included.
//  * All coefficient planes are read through the read-only path, structure-of-arrays
//    (the reference's BTCL_u/v array-of-structs is split into 10 planes at upload) so
//    each warp load is a contiguous 256-byte request.
#include "ctx.h"
#include "bt_planes.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace {

using m6::Geom;

// local_BT_cont_u_type field order, MOM_barotropic.F90:367-390
enum { FA_EE = 0, FA_E0, FA_W0, FA_WW, UBT_WW, UBT_EE, CRV_W, CRV_E, UH_WW, UH_EE };

struct BtStep {
  int isv, iev, jsv, jev;  // valid range of this substep
  int v_first, bracket_bug, add_eta_sum;
  double dtbt, dgeo_de, vel_underflow;
  double wt_vel, wt_eta, wt_accel, wt_trans, wt_accel2, trans_wt1, trans_wt2;
};

// find_uhbt / find_vhbt, MOM_barotropic.F90:4610-4631 / :4744-4765
__device__ __forceinline__ double find_hbt(double u, const double* const* __restrict__ b, long long g) {
  if (u == 0.0) return 0.0;
  const double uEE = __ldg(b[UBT_EE] + g);
  if (u < uEE) return (u - uEE) * __ldg(b[FA_EE] + g) + __ldg(b[UH_EE] + g);
  if (u < 0.0) return u * (__ldg(b[FA_E0] + g) + __ldg(b[CRV_E] + g) * (u * u));
  const double uWW = __ldg(b[UBT_WW] + g);
  if (u <= uWW) return u * (__ldg(b[FA_W0] + g) + __ldg(b[CRV_W] + g) * (u * u));
  return (u - uWW) * __ldg(b[FA_WW] + g) + __ldg(b[UH_WW] + g);
}

__device__ __forceinline__ bool in_rng(int v, int lo, int hi) { return (v >= lo) && (v <= hi); }

template <int TX, int TY, int NT, bool BT_CONT, bool PROJECT>
__global__ void __launch_bounds__(NT)
bt_substep_kernel(const Geom G, const BtPlanes P, const BtStep S) {
  constexpr int EW = TX + 3, EH = TY + 3, EN = EW * EH;
  extern __shared__ double smem[];
  double* ub = smem;            // ubt at the start of the substep
  double* vb = ub + EN;
  double* ubn = vb + EN;        // updated ubt
  double* vbn = ubn + EN;
  double* uh = vbn + EN;        // uhbt
  double* vh = uh + EN;         // vhbt
  double* ep = vh + EN;         // eta_pred (or eta when projecting)

  const int ti0 = G.i0 + blockIdx.x * TX, tj0 = G.j0 + blockIdx.y * TY;
  const int isv = S.isv, iev = S.iev, jsv = S.jsv, jev = S.jev;
  const double dtbt = S.dtbt;

  // ---- phase A: stage ubt/vbt, predictor transports (btloop_eta_predictor :3036-3043)
  for (int p = threadIdx.x; p < EN; p += NT) {
    const int ey = p / EW, ex = p - ey * EW;
    const int i = ti0 - 2 + ex, j = tj0 - 2 + ey;
    double u = 0.0, v = 0.0, e = 0.0, tu = 0.0, tv = 0.0;
    if (G.inside(i, j)) {
      const long long g = G.idx(i, j);
      u = __ldg(P.ubt_in + g);
      v = __ldg(P.vbt_in + g);
      if (PROJECT) {
        e = __ldg(P.eta_in + g);
      } else {
        if (in_rng(j, jsv - 1, jev + 1) && in_rng(i, isv - 2, iev + 1)) {
          if (BT_CONT) tu = find_hbt(u, P.bu, g) + __ldg(P.uhbt0 + g);
          else tu = __ldg(P.Datu + g) * u + __ldg(P.uhbt0 + g);
        }
        if (in_rng(j, jsv - 2, jev + 1) && in_rng(i, isv - 1, iev + 1)) {
          if (BT_CONT) tv = find_hbt(v, P.bv, g) + __ldg(P.vhbt0 + g);
          else tv = __ldg(P.Datv + g) * v + __ldg(P.vhbt0 + g);
        }
      }
    }
    ub[p] = u; vb[p] = v; ubn[p] = u; vbn[p] = v; uh[p] = tu; vh[p] = tv; ep[p] = e;
  }
  __syncthreads();

  // ---- phase B: eta_pred (:3045-3048 / linear form :3051-3057)
  if (!PROJECT) {
    for (int p = threadIdx.x; p < EN; p += NT) {
      const int ey = p / EW, ex = p - ey * EW;
      const int i = ti0 - 2 + ex, j = tj0 - 2 + ey;
      if (ex >= 1 && ey >= 1 && in_rng(i, isv - 1, iev + 1) && in_rng(j, jsv - 1, jev + 1)) {
        const long long g = G.idx(i, j);
        ep[p] = (__ldg(P.eta_in + g) + __ldg(P.eta_src + g)) + (dtbt * __ldg(P.IareaT + g)) *
                ((uh[p - 1] - uh[p]) + (vh[p - EW] - vh[p]));
      }
    }
    __syncthreads();
  }

  // ---- phases C/D: velocity updates in the order given by v_first (:2580-2600)
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const bool do_v = (S.v_first != 0) == (pass == 0);
    if (do_v) {
      // btloop_update_v :3264-3299 with PFv from btloop_find_PF :3134-3138
      const double* usrc = S.v_first ? ub : ubn;  // v second sees the updated u
      const int ilo = S.v_first ? isv - 1 : isv, ihi = S.v_first ? iev + 1 : iev;
      const int exlo = S.v_first ? 1 : 2, exhi = S.v_first ? TX + 2 : TX + 1;
      const bool bug = (!S.v_first) && S.bracket_bug;
      for (int p = threadIdx.x; p < EN; p += NT) {
        const int ey = p / EW, ex = p - ey * EW;
        const int i = ti0 - 2 + ex, J = tj0 - 2 + ey;
        if (ex >= exlo && ex <= exhi && ey >= 1 && ey <= TY + 1 && in_rng(i, ilo, ihi) && in_rng(J, jsv - 1, jev)) {
          const long long g = G.idx(i, J);
          const long long gn = g + G.pitch;
          const double PFv = (((ep[p] - __ldg(P.eta_PF + g)) * __ldg(P.gtot_N + g)) -
                              ((ep[p + EW] - __ldg(P.eta_PF + gn)) * __ldg(P.gtot_S + gn))) *
                             S.dgeo_de * __ldg(P.IdyCv + g);
          const double f1 = __ldg(P.f4v[0] + g), f2 = __ldg(P.f4v[1] + g), f3 = __ldg(P.f4v[2] + g),
                       f4 = __ldg(P.f4v[3] + g);
          double Cor_v;
          if (bug)
            Cor_v = -1.0 * (((f1 * usrc[p - 1]) + (f2 * usrc[p])) + ((f4 * usrc[p + EW]) + (f3 * usrc[p + EW - 1]))) -
                    __ldg(P.Cor_ref_v + g);
          else
            Cor_v = -1.0 * (((f1 * usrc[p - 1]) + (f4 * usrc[p + EW])) + ((f2 * usrc[p]) + (f3 * usrc[p + EW - 1]))) -
                    __ldg(P.Cor_ref_v + g);
          double vn = __ldg(P.bt_rem_v + g) * (vb[p] + dtbt * ((__ldg(P.BT_force_v + g) + Cor_v) + PFv));
          if (fabs(vn) < S.vel_underflow) vn = 0.0;
          vbn[p] = vn;
          if (ex >= 2 && ex <= TX + 1 && ey >= 2)  // owner (ey <= TY+1 already holds)
            P.v_accel_bt[g] = P.v_accel_bt[g] + S.wt_accel * (Cor_v + PFv);
        }
      }
    } else {
      // btloop_update_u :3358-3379 with PFu from btloop_find_PF :3126-3130
      const double* vsrc = S.v_first ? vbn : vb;  // u second sees the updated v
      const int jlo = S.v_first ? jsv : jsv - 1, jhi = S.v_first ? jev : jev + 1;
      const int eylo = S.v_first ? 2 : 1, eyhi = S.v_first ? TY + 1 : TY + 2;
      for (int p = threadIdx.x; p < EN; p += NT) {
        const int ey = p / EW, ex = p - ey * EW;
        const int I = ti0 - 2 + ex, j = tj0 - 2 + ey;
        if (ex >= 1 && ex <= TX + 1 && ey >= eylo && ey <= eyhi && in_rng(I, isv - 1, iev) && in_rng(j, jlo, jhi)) {
          const long long g = G.idx(I, j);
          const long long ge = g + 1;
          const double PFu = (((ep[p] - __ldg(P.eta_PF + g)) * __ldg(P.gtot_E + g)) -
                              ((ep[p + 1] - __ldg(P.eta_PF + ge)) * __ldg(P.gtot_W + ge))) *
                             S.dgeo_de * __ldg(P.IdxCu + g);
          const double f1 = __ldg(P.f4u[0] + g), f2 = __ldg(P.f4u[1] + g), f3 = __ldg(P.f4u[2] + g),
                       f4 = __ldg(P.f4u[3] + g);
          const double Cor_u = (((f4 * vsrc[p + 1]) + (f1 * vsrc[p - EW])) + ((f3 * vsrc[p]) + (f2 * vsrc[p - EW + 1]))) -
                               __ldg(P.Cor_ref_u + g);
          double un = __ldg(P.bt_rem_u + g) * (ub[p] + dtbt * ((__ldg(P.BT_force_u + g) + Cor_u) + PFu));
          if (fabs(un) < S.vel_underflow) un = 0.0;
          ubn[p] = un;
          if (ex >= 2 && ey >= 2 && ey <= TY + 1)  // owner
            P.u_accel_bt[g] = P.u_accel_bt[g] + S.wt_accel * (Cor_u + PFu);
        }
      }
    }
    __syncthreads();
  }

  // ---- phase E: transports from the time-weighted velocities (:2623-2646) and the
  //      running sums on the computational domain (:2691-2702)
  for (int p = threadIdx.x; p < EN; p += NT) {
    const int ey = p / EW, ex = p - ey * EW;
    const int i = ti0 - 2 + ex, j = tj0 - 2 + ey;
    if (ex >= 1 && ex <= TX + 1 && ey >= 2 && ey <= TY + 1 && in_rng(i, isv - 1, iev) && in_rng(j, jsv, jev)) {
      const long long g = G.idx(i, j);
      const double ut = S.trans_wt1 * ubn[p] + S.trans_wt2 * ub[p];
      double t;
      if (BT_CONT) t = find_hbt(ut, P.bu, g) + __ldg(P.uhbt0 + g);
      else t = __ldg(P.Datu + g) * ut + __ldg(P.uhbt0 + g);
      uh[p] = t;
      if (ex >= 2 && in_rng(i, G.isc - 1, G.iec) && in_rng(j, G.jsc, G.jec)) {
        P.ubtav[g] = P.ubtav[g] + S.wt_trans * ut;
        P.uhbtav[g] = P.uhbtav[g] + S.wt_trans * t;
        P.ubt_wtd[g] = P.ubt_wtd[g] + S.wt_vel * ubn[p];
      }
    }
    if (ex >= 2 && ex <= TX + 1 && ey >= 1 && ey <= TY + 1 && in_rng(i, isv, iev) && in_rng(j, jsv - 1, jev)) {
      const long long g = G.idx(i, j);
      const double vt = S.trans_wt1 * vbn[p] + S.trans_wt2 * vb[p];
      double t;
      if (BT_CONT) t = find_hbt(vt, P.bv, g) + __ldg(P.vhbt0 + g);
      else t = __ldg(P.Datv + g) * vt + __ldg(P.vhbt0 + g);
      vh[p] = t;
      if (ey >= 2 && in_rng(i, G.isc, G.iec) && in_rng(j, G.jsc - 1, G.jec)) {
        P.vbtav[g] = P.vbtav[g] + S.wt_trans * vt;
        P.vhbtav[g] = P.vhbtav[g] + S.wt_trans * t;
        P.vbt_wtd[g] = P.vbt_wtd[g] + S.wt_vel * vbn[p];
      }
    }
  }
  __syncthreads();

  // ---- phase F: corrector eta (:2722-2727), eta_sum (:3141-3146), write the new state
  for (int p = threadIdx.x; p < EN; p += NT) {
    const int ey = p / EW, ex = p - ey * EW;
    if (ex < 2 || ex > TX + 1 || ey < 2 || ey > TY + 1) continue;
    const int i = ti0 - 2 + ex, j = tj0 - 2 + ey;
    if (!G.inside(i, j)) continue;
    const long long g = G.idx(i, j);
    double e = __ldg(P.eta_in + g);
    if (S.add_eta_sum && in_rng(i, G.isc, G.iec) && in_rng(j, G.jsc, G.jec))
      P.eta_sum[g] = P.eta_sum[g] + S.wt_accel2 * (PROJECT ? e : ep[p]);
    if (in_rng(i, isv, iev) && in_rng(j, jsv, jev)) {
      e = (e + __ldg(P.eta_src + g)) + (dtbt * __ldg(P.IareaT + g)) * ((uh[p - 1] - uh[p]) + (vh[p - EW] - vh[p]));
      P.eta_wtd[g] = P.eta_wtd[g] + e * S.wt_eta;
    }
    P.eta_out[g] = e;
    P.ubt_out[g] = ubn[p];
    P.vbt_out[g] = vbn[p];
  }
}

__global__ void zero_rect_kernel(const Geom G, double* a, int ilo, int ihi, int jlo, int jhi) {
  const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = jlo + blockIdx.y;
  if (i <= ihi && j <= jhi) a[G.idx(i, j)] = 0.0;
}

// Tile of the substep kernel: 64 x 16 points, 512 threads, 71 KB of shared memory (3 CTAs/SM).  MOM6CU_BT_TILE selects alternatives for A/B
// measurements: 1 = 32 x 16 / 256 threads (37 KB: 6 CTAs/SM in different phases), 2 = 32 x 32 / 512 threads.
// Tried and measured slower at 4320x3240 (172 ms per 68 substeps for the plain kernel): fully unrolled phase loops (195 ms), L2 prefetch of
// the later phases' coefficient planes in phase A (186 ms), both (206 ms).
template <int TX, int TY, int NT, bool BT_CONT, bool PROJECT>
int launch_substep_tile(mom6cu_ctx* c, const BtPlanes& P, const BtStep& S) {
  constexpr size_t SMEM = (size_t)7 * (TX + 3) * (TY + 3) * sizeof(double);
  auto kern = bt_substep_kernel<TX, TY, NT, BT_CONT, PROJECT>;
  static bool attr_set = false;
  if (!attr_set) {
    M6_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    attr_set = true;
  }
  dim3 grid((c->g.nx + TX - 1) / TX, (c->g.ny + TY - 1) / TY);
  M6_LAUNCH(c, kern, grid, NT, SMEM, c->g, P, S);
  return 0;
}

template <bool BT_CONT, bool PROJECT>
int launch_substep(mom6cu_ctx* c, const BtPlanes& P, const BtStep& S) {
  static int tile = -1;
  if (tile < 0) { const char* e = getenv("MOM6CU_BT_TILE"); tile = e ? atoi(e) : 0; }
  if (tile == 1) return launch_substep_tile<32, 16, 256, BT_CONT, PROJECT>(c, P, S);
  if (tile == 2) return launch_substep_tile<32, 32, 512, BT_CONT, PROJECT>(c, P, S);
  return launch_substep_tile<64, 16, 512, BT_CONT, PROJECT>(c, P, S);
}

}  // namespace
int m6_bt_alloc(mom6cu_ctx* c, BtDevice& D) {
  auto pl = [&](const char* n) { return c->plane2(std::string("bt.") + n); };
  for (int s = 0; s < 2; ++s) {
    D.eta[s] = pl(s ? "eta1" : "eta0"); D.ubt[s] = pl(s ? "ubt1" : "ubt0"); D.vbt[s] = pl(s ? "vbt1" : "vbt0");
  }
  BtPlanes& P = D.P;
  P.uhbt0 = pl("uhbt0"); P.vhbt0 = pl("vhbt0"); P.Datu = pl("Datu"); P.Datv = pl("Datv");
  static const char* bn[10] = {"b0", "b1", "b2", "b3", "b4", "b5", "b6", "b7", "b8", "b9"};
  for (int m = 0; m < 10; ++m) {
    P.bu[m] = pl((std::string("btclu.") + bn[m]).c_str());
    P.bv[m] = pl((std::string("btclv.") + bn[m]).c_str());
  }
  P.eta_src = pl("eta_src"); P.eta_PF = pl("eta_PF");
  P.gtot_E = pl("gtot_E"); P.gtot_W = pl("gtot_W"); P.gtot_N = pl("gtot_N"); P.gtot_S = pl("gtot_S");
  for (int m = 0; m < 4; ++m) {
    P.f4u[m] = pl((std::string("f4u.") + bn[m]).c_str());
    P.f4v[m] = pl((std::string("f4v.") + bn[m]).c_str());
  }
  P.bt_rem_u = pl("bt_rem_u"); P.bt_rem_v = pl("bt_rem_v");
  P.BT_force_u = pl("BT_force_u"); P.BT_force_v = pl("BT_force_v");
  P.Cor_ref_u = pl("Cor_ref_u"); P.Cor_ref_v = pl("Cor_ref_v");
  P.IareaT = pl("IareaT_OBCmask"); P.IdxCu = pl("IdxCu"); P.IdyCv = pl("IdyCv");
  P.u_accel_bt = pl("u_accel_bt"); P.v_accel_bt = pl("v_accel_bt");
  P.eta_sum = pl("eta_sum"); P.eta_wtd = pl("eta_wtd");
  P.ubtav = pl("ubtav"); P.vbtav = pl("vbtav"); P.uhbtav = pl("uhbtav"); P.vhbtav = pl("vhbtav");
  P.ubt_wtd = pl("ubt_wtd"); P.vbt_wtd = pl("vbt_wtd");
  if (!P.vbt_wtd || !D.vbt[1]) return MOM6CU_ERR_CUDA;
  return 0;
}

namespace {
int bt_check(mom6cu_ctx* c, const mom6cu_bt_timeloop_args* a) {
  if (!a) return c->fail(MOM6CU_ERR_BAD_ARG, "btstep_timeloop: null args");
  if (a->nstep + a->nfilter <= 0)
    return c->fail(MOM6CU_ERR_BAD_ARG, "btstep: number of barotropic step (nstep+nfilter) is 0");
  if (a->use_BT_cont && (!a->BTCL_u || !a->BTCL_v))
    return c->fail(MOM6CU_ERR_BAD_ARG, "btstep_timeloop: use_BT_cont without BTCL_u/BTCL_v");
  if (!a->use_BT_cont && (!a->Datu || !a->Datv))
    return c->fail(MOM6CU_ERR_BAD_ARG, "btstep_timeloop: Datu/Datv required when use_BT_cont is false");
  const mom6cu_domain& d = c->dom;
  if (d.npi * d.npj > 1 && !c->comm)
    return c->fail(MOM6CU_ERR_BAD_ARG, "btstep_timeloop: multi-rank layout without a communicator");
  const int stencil = std::max(1, a->min_stencil);
  if ((d.isc - d.isdw) < stencil || (d.jsc - d.jsdw) < stencil)
    return c->fail(MOM6CU_ERR_BAD_ARG, "btstep_timeloop: wide halo narrower than the stencil");
  return 0;
}

int bt_upload(mom6cu_ctx* c, BtDevice& D, const mom6cu_bt_timeloop_args* a) {
  int rc;
  BtPlanes& P = D.P;
#define UP(src, st, dst) if ((rc = m6_up(c, (src), (st), 1, 1, (double*)(dst)))) return rc
  UP(a->eta, ST_H, D.eta[0]); UP(a->ubt, ST_U, D.ubt[0]); UP(a->vbt, ST_V, D.vbt[0]);
  UP(a->uhbt0, ST_U, P.uhbt0); UP(a->vhbt0, ST_V, P.vhbt0);
  if (a->use_BT_cont) {
    if ((rc = m6_up_aos(c, a->BTCL_u, 10, ST_U, 1, (double* const*)P.bu))) return rc;
    if ((rc = m6_up_aos(c, a->BTCL_v, 10, ST_V, 1, (double* const*)P.bv))) return rc;
  } else {
    UP(a->Datu, ST_U, P.Datu); UP(a->Datv, ST_V, P.Datv);
  }
  UP(a->eta_src, ST_H, P.eta_src); UP(a->eta_PF, ST_H, P.eta_PF);
  UP(a->gtot_E, ST_H, P.gtot_E); UP(a->gtot_W, ST_H, P.gtot_W);
  UP(a->gtot_N, ST_H, P.gtot_N); UP(a->gtot_S, ST_H, P.gtot_S);
  if ((rc = m6_up_aos(c, a->f_4_u, 4, ST_U, 1, (double* const*)P.f4u))) return rc;
  if ((rc = m6_up_aos(c, a->f_4_v, 4, ST_V, 1, (double* const*)P.f4v))) return rc;
  UP(a->bt_rem_u, ST_U, P.bt_rem_u); UP(a->bt_rem_v, ST_V, P.bt_rem_v);
  UP(a->BT_force_u, ST_U, P.BT_force_u); UP(a->BT_force_v, ST_V, P.BT_force_v);
  UP(a->Cor_ref_u, ST_U, P.Cor_ref_u); UP(a->Cor_ref_v, ST_V, P.Cor_ref_v);
  UP(a->IareaT_OBCmask, ST_H, P.IareaT); UP(a->IdxCu, ST_U, P.IdxCu); UP(a->IdyCv, ST_V, P.IdyCv);
  UP(a->u_accel_bt, ST_U, P.u_accel_bt); UP(a->v_accel_bt, ST_V, P.v_accel_bt);
  // intent(out) accumulators keep the caller's values outside the ranges the
  // reference zeroes, so start from the caller's arrays.
  UP(a->eta_wtd, ST_H, P.eta_wtd);
  if (a->find_etaav) UP(a->eta_sum, ST_H, P.eta_sum);
#undef UP
#define UPG(src, st, dst) if ((rc = m6_up(c, (src), (st), 0, 1, (double*)(dst)))) return rc
  UPG(a->ubtav, ST_U, P.ubtav); UPG(a->vbtav, ST_V, P.vbtav);
  UPG(a->uhbtav, ST_U, P.uhbtav); UPG(a->vhbtav, ST_V, P.vhbtav);
  UPG(a->ubt_wtd, ST_U, P.ubt_wtd); UPG(a->vbt_wtd, ST_V, P.vbt_wtd);
#undef UPG
  return 0;
}

}  // namespace

int m6_bt_halo_exchange(mom6cu_ctx* c, double* eta, double* ubt, double* vbt);  // halo.cu

namespace {

// The substep loop proper: everything inside is device work on c->stream.
}  // namespace
int m6_bt_run(mom6cu_ctx* c, BtDevice& D, const mom6cu_bt_timeloop_args* a, int* final_slot) {
  const mom6cu_domain& d = c->dom;
  const Geom& G = c->g;
  const int is = d.isc, ie = d.iec, js = d.jsc, je = d.jec;
  const int ntot = a->nstep + a->nfilter;
  // :2413-2422
  const int stencil = std::max(1, a->min_stencil);
  int num_cycles = 1;
  if (a->use_wide_halos) num_cycles = std::min((is - d.isdw) / stencil, (js - d.jsdw) / stencil);
  const int isvf = is - (num_cycles - 1) * stencil, ievf = ie + (num_cycles - 1) * stencil;
  const int jsvf = js - (num_cycles - 1) * stencil, jevf = je + (num_cycles - 1) * stencil;

  BtStep S;
  S.dtbt = a->dtbt; S.dgeo_de = a->dgeo_de; S.vel_underflow = a->vel_underflow;
  S.bracket_bug = a->use_old_coriolis_bracket_bug;
  if (a->BT_project_velocity) { S.trans_wt1 = (1.0 + a->bebt); S.trans_wt2 = -a->bebt; }
  else { S.trans_wt1 = a->bebt; S.trans_wt2 = (1.0 - a->bebt); }

  // :2455-2486 zero the time-averaged quantities on exactly the reference's ranges
  auto zero = [&](double* p, int ilo, int ihi, int jlo, int jhi) {
    dim3 grid((ihi - ilo + 128) / 128, jhi - jlo + 1);
    M6_LAUNCH(c, zero_rect_kernel, grid, 128, 0, G, p, ilo, ihi, jlo, jhi);
  };
  BtPlanes& P = D.P;
  zero(P.eta_wtd, isvf - 1, ievf + 1, jsvf - 1, jevf + 1);
  if (a->find_etaav) zero(P.eta_sum, isvf - 1, ievf + 1, jsvf - 1, jevf + 1);
  zero(P.ubtav, is - 1, ie, js, je); zero(P.uhbtav, is - 1, ie, js, je); zero(P.ubt_wtd, is - 1, ie, js, je);
  zero(P.vbtav, is, ie, js - 1, je); zero(P.vhbtav, is, ie, js - 1, je); zero(P.vbt_wtd, is, ie, js - 1, je);

  int cur = 0;
  int isv = is, iev = ie, jsv = js, jev = je;
  for (int n = 1; n <= ntot; ++n) {
    // :2509-2518
    if ((iev - stencil < ie) || (jev - stencil < je)) {
      int rc = m6_bt_halo_exchange(c, D.eta[cur], D.ubt[cur], D.vbt[cur]);
      if (rc) return rc;
      isv = isvf; iev = ievf; jsv = jsvf; jev = jevf;
    } else {
      isv += stencil; iev -= stencil; jsv += stencil; jev -= stencil;
    }
    S.isv = isv; S.iev = iev; S.jsv = jsv; S.jev = jev;
    S.v_first = (((n + d.first_direction) % 2) == 1) ? 1 : 0;  // :2561
    S.wt_vel = a->wt_vel[n - 1]; S.wt_eta = a->wt_eta[n - 1]; S.wt_accel = a->wt_accel[n - 1];
    S.wt_trans = a->wt_trans[n - 1]; S.wt_accel2 = a->wt_accel2[n - 1];
    S.add_eta_sum = (a->find_etaav && (fabs(S.wt_accel2) > 0.0)) ? 1 : 0;
    P.eta_in = D.eta[cur]; P.ubt_in = D.ubt[cur]; P.vbt_in = D.vbt[cur];
    P.eta_out = D.eta[cur ^ 1]; P.ubt_out = D.ubt[cur ^ 1]; P.vbt_out = D.vbt[cur ^ 1];
    int rc;
    if (a->use_BT_cont) rc = a->BT_project_velocity ? launch_substep<true, true>(c, P, S) : launch_substep<true, false>(c, P, S);
    else rc = a->BT_project_velocity ? launch_substep<false, true>(c, P, S) : launch_substep<false, false>(c, P, S);
    if (rc) return rc;
    cur ^= 1;
  }
  M6_CUDA(c, cudaGetLastError());
  *final_slot = cur;
  return 0;
}

namespace {
int bt_download(mom6cu_ctx* c, BtDevice& D, const mom6cu_bt_timeloop_args* a, int slot) {
  int rc;
  BtPlanes& P = D.P;
#define DN(src, st, w, dst) if ((rc = m6_down(c, (src), (st), (w), 1, (dst)))) return rc
  DN(D.eta[slot], ST_H, 1, a->eta); DN(D.ubt[slot], ST_U, 1, a->ubt); DN(D.vbt[slot], ST_V, 1, a->vbt);
  DN(P.u_accel_bt, ST_U, 1, a->u_accel_bt); DN(P.v_accel_bt, ST_V, 1, a->v_accel_bt);
  DN(P.eta_wtd, ST_H, 1, a->eta_wtd);
  if (a->find_etaav) DN(P.eta_sum, ST_H, 1, a->eta_sum);
  DN(P.ubtav, ST_U, 0, a->ubtav); DN(P.vbtav, ST_V, 0, a->vbtav);
  DN(P.uhbtav, ST_U, 0, a->uhbtav); DN(P.vhbtav, ST_V, 0, a->vhbtav);
  DN(P.ubt_wtd, ST_U, 0, a->ubt_wtd); DN(P.vbt_wtd, ST_V, 0, a->vbt_wtd);
#undef DN
  return 0;
}

}  // namespace

extern "C" int mom6cu_btstep_timeloop(mom6cu_ctx* c, const mom6cu_bt_timeloop_args* a) {
  return mom6cu_btstep_timeloop_resident(c, a, 1, 1);
}

extern "C" int mom6cu_btstep_timeloop_resident(mom6cu_ctx* c, const mom6cu_bt_timeloop_args* a, int reps,
                                               int download) {
  if (!c) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  int rc = bt_check(c, a);
  if (rc) return rc;
  BtDevice D;
  if ((rc = m6_bt_alloc(c, D))) return rc;
  if ((rc = bt_upload(c, D, a))) return rc;
  int slot = 0;
  // keep pristine copies of everything the loop modifies so every repetition starts from the
  // same state (device-to-device restores, outside the timed region)
  const size_t pb = (size_t)c->g.plane * sizeof(double);
  double* keep = nullptr;
  double* srcs[5] = {D.eta[0], D.ubt[0], D.vbt[0], D.P.u_accel_bt, D.P.v_accel_bt};
  if (reps > 1) {
    keep = c->buf("bt.__keep", (size_t)5 * c->g.plane);
    if (!keep) return MOM6CU_ERR_CUDA;
    for (int m = 0; m < 5; ++m)
      M6_CUDA(c, cudaMemcpyAsync(keep + (size_t)m * c->g.plane, srcs[m], pb, cudaMemcpyDeviceToDevice, c->stream));
  }
  c->total_ms = 0.0;
  for (int r = 0; r < (reps > 1 ? reps : 1); ++r) {
    if (reps > 1)
      for (int m = 0; m < 5; ++m)
        M6_CUDA(c, cudaMemcpyAsync(srcs[m], keep + (size_t)m * c->g.plane, pb, cudaMemcpyDeviceToDevice, c->stream));
    M6_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    if ((rc = m6_bt_run(c, D, a, &slot))) return rc;
    M6_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    M6_CUDA(c, cudaEventSynchronize(c->ev1));
    float ms1 = 0.f;
    M6_CUDA(c, cudaEventElapsedTime(&ms1, c->ev0, c->ev1));
    c->total_ms += ms1;
    c->last_ms = ms1;
  }
  if (download && (rc = bt_download(c, D, a, slot))) return rc;
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
