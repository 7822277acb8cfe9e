// advect_tracer (/root/reference/src/tracer/MOM_tracer_advect.F90:53-350) with advect_x :355-744 and advect_y :748-1152.
//
// Device formulation.  The reference sweeps rows under per-row `domore` flags and lets the valid range march inward
// between halo updates; both are CPU devices that do not change what a cell computes: a face with no remaining
// transport is a no-op (uhh = 0, no cell changes), a row is flagged exactly when a face of it was limited, and a value
// computed in the halo equals the one the halo update would deliver.  So each pass here is one kernel over every face
// of the pass's range, reading the state before the pass and writing the state after it into a second set of arrays
// (ping-pong: x pass A -> B, y pass B -> A), with a halo update per iteration; the number of iterations is the
// reference's: it stops at max_iter, or at an iteration the reference would test (every nsten_halo-th) in which no
// face was limited.  Per-cell arithmetic keeps the reference's expression order (bitwise with -fmad=false).
// Not produced: the flux diagnostics ad_x, ad_y, ad2d_x, ad2d_y, advection_xy.  OBCs are rejected by the caller.
#include "ctx.h"
#include "common.cuh"
#include "stage.h"
#include <cfloat>
#include <cmath>
#include <vector>
#include <algorithm>

using m6::Geom;
using m6::fmax2;
using m6::fmin2;
using m6::fsign;

namespace {

constexpr int NTB = 4;    // tracers per launch
constexpr int XB = 128;   // faces per block in the x pass (XB-1 cells)
constexpr int YSEG = 24;  // rows marched by one thread in the y pass

struct AdvPass {
  int is, ie, js, je;          // cells updated by the pass
  int nt;                      // tracers in this launch
  int first;                   // this launch also advances hprev / uhr (the first tracer group of a pass)
  double min_h, h_neglect, H_subroundoff;
  const double* areaT; const double* maskC;  // mask2dCu (x) / mask2dCv (y)
  const double* hp_old; double* hp_new;
  const double* tr_old; double* tr_new;      // remaining transport uhr / vhr
  const double* T_old[NTB]; double* T_new[NTB];
  int scheme[NTB]; double underflow[NTB];
  int* limited;                // set to 1 when a face of the pass was limited
  // x pass only: domore_u(j,k) of the reference (:433): rows whose flag is clear keep uhr untouched (so a -0.0 stays
  // -0.0); the flag of the next pass is raised where a face of the row was limited (:524,:535).  Indexed (k, j - jsd).
  const int* row_old; int* row_new; int nrow, jsd;
};

__device__ __forceinline__ double min3(double a, double b, double c) { return fmin2(fmin2(a, b), c); }
__device__ __forceinline__ double max3(double a, double b, double c) { return fmax2(fmax2(a, b), c); }

// :455-459 / :814-818
__device__ __forceinline__ double plm_slope(double Tp, double Tc, double Tm, double maskprod) {
  const double dMx = max3(Tp, Tc, Tm) - Tc;
  const double dMn = Tc - min3(Tp, Tc, Tm);
  return maskprod * fsign(min3(0.5 * fabs(Tp - Tm), 2.0 * dMx, 2.0 * dMn), Tp - Tm);
}

// the transport used this pass and its upwind CFL number, :513-542 / :872-901.  `o` is the offset of the face's western /
// southern cell, `s` the stride to the next cell along the pass direction.
__device__ __forceinline__ void face_transport(const AdvPass& P, long long o, long long o2, long long s, long long s2, double& uhh, double& CFL,
                                               bool& lim) {
  // o: 3-D offset (hp, tr);  o2: 2-D offset (areaT);  s, s2: strides along the direction
  const double tiny_h = DBL_MIN;
  const double u = __ldg(P.tr_old + o);
  lim = false;
  if ((u == 0.0) || ((u < 0.0) && (__ldg(P.hp_old + o + s) <= tiny_h)) || ((u > 0.0) && (__ldg(P.hp_old + o) <= tiny_h))) {
    uhh = 0.0; CFL = 0.0;
  } else if (u < 0.0) {
    const double hup = __ldg(P.hp_old + o + s) - __ldg(P.areaT + o2 + s2) * P.min_h;
    const double hlos = fmax2(0.0, __ldg(P.tr_old + o + s));
    if ((((hup - hlos) + u) < 0.0) && ((0.5 * hup + u) < 0.0)) { uhh = min3(-0.5 * hup, -hup + hlos, 0.0); lim = true; }
    else uhh = u;
    CFL = -uhh / (__ldg(P.hp_old + o + s));
  } else {
    const double hup = __ldg(P.hp_old + o) - __ldg(P.areaT + o2) * P.min_h;
    const double hlos = fmax2(0.0, -__ldg(P.tr_old + o - s));
    if ((((hup - hlos) - u) < 0.0) && ((0.5 * hup - u) < 0.0)) { uhh = max3(0.5 * hup, hup - hlos, 0.0); lim = true; }
    else uhh = u;
    CFL = uhh / (__ldg(P.hp_old + o));
  }
}

// the tracer flux through a face, :544-608 / :903-960
__device__ __forceinline__ double face_flux(int scheme, const double* __restrict__ T, const double* __restrict__ maskC, long long o, long long o2,
                                            long long s, long long s2, double uhh, double CFL) {
  if (scheme == MOM6CU_ADVECT_PLM) {
    if (uhh >= 0.0) {
      const double sl = plm_slope(__ldg(T + o + s), __ldg(T + o), __ldg(T + o - s), __ldg(maskC + o2) * __ldg(maskC + o2 - s2));
      return uhh * (__ldg(T + o) + 0.5 * sl * (1. - CFL));
    }
    const double sl = plm_slope(__ldg(T + o + 2 * s), __ldg(T + o + s), __ldg(T + o), __ldg(maskC + o2 + s2) * __ldg(maskC + o2));
    return uhh * (__ldg(T + o + s) - 0.5 * sl * (1. - CFL));
  }
  const long long u3 = (uhh >= 0.0) ? o : o + s, u2 = (uhh >= 0.0) ? o2 : o2 + s2;  // the upstream cell
  const double Tp = __ldg(T + u3 + s), Tc = __ldg(T + u3), Tm = __ldg(T + u3 - s);
  double aL, aR;
  if (scheme == MOM6CU_ADVECT_PPMH3) {
    aL = (5. * Tc + (2. * Tm - Tp)) / 6.;
    aL = fmax2(fmin2(Tc, Tm), aL); aL = fmin2(fmax2(Tc, Tm), aL);
    aR = (5. * Tc + (2. * Tp - Tm)) / 6.;
    aR = fmax2(fmin2(Tc, Tp), aR); aR = fmin2(fmax2(Tc, Tp), aR);
  } else {
    const double sl_m = plm_slope(Tc, Tm, __ldg(T + u3 - 2 * s), __ldg(maskC + u2 - s2) * __ldg(maskC + u2 - 2 * s2));
    const double sl_c = plm_slope(Tp, Tc, Tm, __ldg(maskC + u2) * __ldg(maskC + u2 - s2));
    const double sl_p = plm_slope(__ldg(T + u3 + 2 * s), Tp, Tc, __ldg(maskC + u2 + s2) * __ldg(maskC + u2));
    aL = 0.5 * ((Tm + Tc) + (sl_m - sl_c) / 3.);
    aR = 0.5 * ((Tc + Tp) + (sl_c - sl_p) / 3.);
  }
  const double dA = aR - aL, mA = 0.5 * (aR + aL);
  if (__ldg(maskC + u2) * __ldg(maskC + u2 - s2) * (Tp - Tc) * (Tc - Tm) <= 0.) { aL = Tc; aR = Tc; }
  else if (dA * (Tc - mA) > (dA * dA) / 6.) aL = (3. * Tc) - 2. * aR;
  else if (dA * (Tc - mA) < -(dA * dA) / 6.) aR = (3. * Tc) - 2. * aL;
  const double a6 = 6. * Tc - 3. * (aR + aL);
  if (uhh >= 0.0) return uhh * (aR - 0.5 * CFL * ((aR - aL) - a6 * (1. - 2. / 3. * CFL)));
  return uhh * (aL + 0.5 * CFL * ((aR - aL) + a6 * (1. - 2. / 3. * CFL)));
}

// the cell update of :612-703 / :1048-1090 given the transports and fluxes of its two faces; ymax: the y pass clips at 0
template <bool YPASS>
__device__ __forceinline__ void cell_update(const AdvPass& P, long long o, long long o2, double uhh_lo, double uhh_hi, const double* fl_lo,
                                            const double* fl_hi) {
  bool do_i = false;
  double hlst = 0., Ihnew = 0.;
  double hnew = __ldg(P.hp_old + o);
  if ((uhh_hi != 0.0) || (uhh_lo != 0.0)) {
    do_i = true;
    hlst = hnew;
    hnew = hnew - (uhh_hi - uhh_lo);
    if (YPASS) hnew = fmax2(hnew, 0.0);
    const double hmin = P.h_neglect * __ldg(P.areaT + o2);
    if (hnew <= 0.0) do_i = false;
    else if (hnew < hmin) { hlst = hlst + (hmin - hnew); Ihnew = 1.0 / hmin; }
    else Ihnew = 1.0 / hnew;
  }
  if (P.first) P.hp_new[o] = hnew;
#pragma unroll
  for (int m = 0; m < NTB; ++m) if (m < P.nt) {
    double T = __ldg(P.T_old[m] + o);
    if (do_i && (!YPASS ? (Ihnew > 0.0) : true)) T = (T * hlst - (fl_hi[m] - fl_lo[m])) * Ihnew;
    if (P.underflow[m] > 0.0 && fabs(T) < P.underflow[m]) T = 0.0;
    P.T_new[m][o] = T;
  }
}

// x pass: thread t of a block holds face I = ib + t of row j, layer k; cells ib+1 .. ib+XB-1 are updated by threads 1..XB-1
__global__ void __launch_bounds__(XB) advect_x_kernel(Geom G, AdvPass P) {
  __shared__ double s_uhh[XB];
  __shared__ double s_fl[NTB][XB];
  const int t = threadIdx.x;
  const int I = P.is - 1 + blockIdx.x * (XB - 1) + t;
  const int j = P.js + blockIdx.y, k = blockIdx.z;
  const long long o2 = G.idx(I, j), o = (long long)k * G.plane + o2;
  const bool face = (I <= P.ie);
  const int row = k * P.nrow + (j - P.jsd);
  const bool domore = P.row_old[row] != 0;
  double uhh = 0., CFL = 0., fl[NTB];
  bool lim = false;
  if (face && domore) face_transport(P, o, o2, 1, 1, uhh, CFL, lim);
#pragma unroll
  for (int m = 0; m < NTB; ++m) {
    fl[m] = (face && m < P.nt) ? face_flux(P.scheme[m], P.T_old[m] + (long long)k * G.plane, P.maskC, o2, o2, 1, 1, uhh, CFL) : 0.;
    s_fl[m][t] = fl[m];
  }
  s_uhh[t] = uhh;
  if (lim) { *P.limited = 1; P.row_new[row] = 1; }
  __syncthreads();
  if (!face) return;
  if (P.first && (t > 0 || blockIdx.x == 0)) {  // :609-612: each face is written by one block
    double r = __ldg(P.tr_old + o);
    if (domore) {
      r = r - uhh;
      if (fabs(r) < P.H_subroundoff * fmin2(__ldg(P.areaT + o2), __ldg(P.areaT + o2 + 1))) r = 0.0;
    }
    P.tr_new[o] = r;
  }
  if (t == 0) return;
  double fl_lo[NTB];
#pragma unroll
  for (int m = 0; m < NTB; ++m) fl_lo[m] = s_fl[m][t - 1];
  cell_update<false>(P, o, o2, s_uhh[t - 1], uhh, fl_lo, fl);
}

// y pass: one thread marches faces J = jb-1 .. jb+YSEG-1 of column i, layer k, carrying the southern face in registers
__global__ void __launch_bounds__(128) advect_y_kernel(Geom G, AdvPass P) {
  const int i = P.is + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > P.ie) return;
  const int jb = P.js + blockIdx.y * YSEG, k = blockIdx.z;
  const int jt = min(jb + YSEG - 1, P.je);
  const long long s2 = G.pitch, s = G.pitch, kp = (long long)k * G.plane;
  double uhh_lo = 0., fl_lo[NTB];
  bool any_lim = false;
  for (int J = jb - 1; J <= jt; ++J) {
    const long long o2 = G.idx(i, J), o = kp + o2;
    double uhh, CFL, fl[NTB];
    bool lim;
    face_transport(P, o, o2, s, s2, uhh, CFL, lim);
#pragma unroll
    for (int m = 0; m < NTB; ++m) fl[m] = (m < P.nt) ? face_flux(P.scheme[m], P.T_old[m] + kp, P.maskC, o2, o2, s, s2, uhh, CFL) : 0.;
    const bool own = (J >= jb) || (blockIdx.y == 0);  // face jb-1 belongs to the segment below, except for the first
    if (own) {
      any_lim |= lim;
      if (P.first) {
        double r = __ldg(P.tr_old + o) - uhh;
        if (fabs(r) < P.H_subroundoff * fmin2(__ldg(P.areaT + o2), __ldg(P.areaT + o2 + s2))) r = 0.0;
        P.tr_new[o] = r;
      }
    }
    if (J >= jb) cell_update<true>(P, o, o2, uhh_lo, uhh, fl_lo, fl);
    uhh_lo = uhh;
#pragma unroll
    for (int m = 0; m < NTB; ++m) fl_lo[m] = fl[m];
  }
  if (any_lim) *P.limited = 1;
}

// domore_u(j,k) = any(uhr(I,j,k) /= 0) over the faces of the x pass (:228-233); one warp per row
__global__ void advect_rowflag_kernel(Geom G, const double* __restrict__ uhr, int is, int ie, int js, int nrow, int jsd, int* __restrict__ flag) {
  const int j = js + blockIdx.x, k = blockIdx.y;
  int any = 0;
  for (int I = is - 1 + threadIdx.x; I <= ie; I += blockDim.x) any |= (uhr[(long long)k * G.plane + G.idx(I, j)] != 0.0);
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flag[k * nrow + (j - jsd)] = any ? 1 : 0;
}

// :152-200: uhr, vhr, hprev
struct AdvInit {
  int is, ie, js, je;
  const double *h_end, *uhtr, *vhtr, *vol_prev, *areaT;
  double *uhr, *vhr, *hprev;
};
__global__ void advect_init_kernel(Geom G, AdvInit A) {
  const int i = G.isd - 1 + blockIdx.x * blockDim.x + threadIdx.x, j = G.jsd - 1 + blockIdx.y, k = blockIdx.z;
  if (i > G.ied) return;
  const long long o2 = G.idx(i, j), o = (long long)k * G.plane + o2;
  const bool ci = (i >= A.is && i <= A.ie), cj = (j >= A.js && j <= A.je);
  const double u = (cj && i >= A.is - 1 && i <= A.ie) ? A.uhtr[o] : 0.0;
  const double v = (ci && j >= A.js - 1 && j <= A.je) ? A.vhtr[o] : 0.0;
  A.uhr[o] = u; A.vhr[o] = v;
  double hp = 0.0;
  if (ci && cj) {
    if (A.vol_prev) hp = A.vol_prev[o];
    else {
      const double uw = A.uhtr[o - 1], vs = A.vhtr[o - G.pitch];
      hp = fmax2(0.0, A.areaT[o2] * A.h_end[o] + ((u - uw) + (v - vs)));
      hp = hp + fmax2(0.0, 1.0e-13 * hp - A.areaT[o2] * A.h_end[o]);
    }
  }
  A.hprev[o] = hp;
}

}  // namespace

extern "C" int mom6cu_advect_tracer(mom6cu_ctx* c, const mom6cu_tracer_advect_cs* CS, const mom6cu_advect_tracer_args* a) {
  if (!c || !CS || !a) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "advect_tracer: mom6cu_set_grid / mom6cu_set_vgrid have not been called");
  if (!a->h_end || !a->uhtr || !a->vhtr || a->ntr < 0 || (a->ntr > 0 && !a->tr)) return c->fail(MOM6CU_ERR_BAD_ARG, "advect_tracer: null required argument");
  const int ntr = a->ntr;
  if (ntr == 0) return 0;  // :139
  const mom6cu_domain& d = c->dom;
  const Geom& G = c->g;
  const int is = d.isc, ie = d.iec, js = d.jsc, je = d.jec, nz = G.nk;
  std::vector<int> scheme(ntr);
  int stencil = 2;
  for (int m = 0; m < ntr; ++m) {
    int s = a->advect_scheme ? a->advect_scheme[m] : -1;
    if (s < 0) s = CS->default_advect_scheme;
    int sl = 2;
    if (s == MOM6CU_ADVECT_PPM) sl = 3;
    else if (s == MOM6CU_ADVECT_PPMH3) sl = CS->useHuynhStencilBug ? 2 : 3;
    else if (s != MOM6CU_ADVECT_PLM) return c->fail(MOM6CU_ERR_UNSUPPORTED, "advect_tracer: unknown advection scheme %d for tracer %d", s, m);
    scheme[m] = s;
    stencil = std::max(stencil, sl);
    if (!a->tr[m]) return c->fail(MOM6CU_ERR_BAD_ARG, "advect_tracer: tracer %d is null", m);
  }
  const int hmin = std::min(std::min(is - d.isd, d.ied - ie), std::min(js - d.jsd, d.jed - je));
  if (hmin < stencil) return c->fail(MOM6CU_ERR_BAD_ARG, "MOM_tracer_advect: stencil is wider than the halo.");  // :172
  int max_iter = 2 * (int)std::ceil(a->dt / CS->dt) + 1;
  if (a->max_iter_in >= 0) max_iter = a->max_iter_in;
  bool x_first = (d.first_direction % 2 == 0);
  if (a->x_first_in >= 0) x_first = a->x_first_in != 0;
  const int nsten_halo = hmin / stencil;

  Stager S(c, "adv.");
  int rc;
  const double *d_hend, *d_uhtr, *d_vhtr;
  double *d_vol = nullptr, *d_uo = nullptr, *d_vo = nullptr;
  if ((rc = S.in3(a->h_end, ST_H, "h_end", &d_hend)) || (rc = S.in3(a->uhtr, ST_U, "uhtr", &d_uhtr)) || (rc = S.in3(a->vhtr, ST_V, "vhtr", &d_vhtr)))
    return rc;
  if (a->vol_prev) {
    if (a->update_vol_prev) { if ((rc = S.io3(a->vol_prev, ST_H, "vol_prev", &d_vol))) return rc; }
    else { const double* p; if ((rc = S.in3(a->vol_prev, ST_H, "vol_prev", &p))) return rc; d_vol = (double*)p; }
  }
  if (a->uhr_out && (rc = S.io3(a->uhr_out, ST_U, "uhr_out", &d_uo))) return rc;
  if (a->vhr_out && (rc = S.io3(a->vhr_out, ST_V, "vhr_out", &d_vo))) return rc;
  std::vector<double*> TA(ntr), TB(ntr);
  for (int m = 0; m < ntr; ++m) {
    char nm[32];
    snprintf(nm, sizeof nm, "tr%d", m);
    if ((rc = S.io3(a->tr[m], ST_H, nm, &TA[m]))) return rc;
    snprintf(nm, sizeof nm, "adv.trB%d", m);
    if (!(TB[m] = c->plane3(nm))) return MOM6CU_ERR_CUDA;
  }
  double *hp[2] = {c->plane3("adv.hprevA"), c->plane3("adv.hprevB")}, *uhr[2] = {c->plane3("adv.uhrA"), c->plane3("adv.uhrB")},
         *vhr[2] = {c->plane3("adv.vhrA"), c->plane3("adv.vhrB")};
  const int nrow = d.jed - d.jsd + 1;
  int* d_row[2] = {(int*)c->buf("adv.rowA", (size_t)(nrow * nz + 1) / 2 + 1), (int*)c->buf("adv.rowB", (size_t)(nrow * nz + 1) / 2 + 1)};
  if (!d_row[0] || !d_row[1]) return MOM6CU_ERR_CUDA;
  int irow = 0;
  int* d_flags = (int*)c->buf("adv.flags", 64);
  int* h_flags = (int*)c->host_scratch("adv.flags", 64);
  if (!hp[0] || !hp[1] || !uhr[0] || !uhr[1] || !vhr[0] || !vhr[1] || !d_flags || !h_flags) return MOM6CU_ERR_CUDA;
  if ((rc = S.begin())) return rc;

  {
    AdvInit I = {is, ie, js, je, d_hend, d_uhtr, d_vhtr, d_vol, c->grid.areaT, uhr[0], vhr[0], hp[0]};
    const dim3 grid((d.ied - d.isd + 2 + 127) / 128, d.jed - d.jsd + 2, nz);
    M6_LAUNCH(c, advect_init_kernel, grid, 128, 0, G, I);
  }
  // the second copies start from the first ones, so that points no pass writes (land halos) hold the same values
  const size_t bytes3 = (size_t)G.plane * nz * sizeof(double);
  M6_CUDA(c, cudaMemcpyAsync(hp[1], hp[0], bytes3, cudaMemcpyDeviceToDevice, c->stream));
  M6_CUDA(c, cudaMemcpyAsync(uhr[1], uhr[0], bytes3, cudaMemcpyDeviceToDevice, c->stream));
  M6_CUDA(c, cudaMemcpyAsync(vhr[1], vhr[0], bytes3, cudaMemcpyDeviceToDevice, c->stream));
  for (int m = 0; m < ntr; ++m) M6_CUDA(c, cudaMemcpyAsync(TB[m], TA[m], bytes3, cudaMemcpyDeviceToDevice, c->stream));
  int ihp = 0, iu = 0, iv = 0;  // which copy is current
  AdvPass P = {};
  P.min_h = 0.1 * c->vgrid.Angstrom_H; P.h_neglect = c->vgrid.H_subroundoff; P.H_subroundoff = c->vgrid.H_subroundoff;
  P.areaT = c->grid.areaT; P.limited = d_flags;
  const std::vector<double*>*Tcur = &TA, *Toth = &TB;

  auto pass = [&](bool xdir, int pis, int pie, int pjs, int pje) -> int {
    P.is = pis; P.ie = pie; P.js = pjs; P.je = pje;
    P.maskC = xdir ? c->grid.mask2dCu : c->grid.mask2dCv;
    P.hp_old = hp[ihp]; P.hp_new = hp[1 - ihp];
    P.tr_old = xdir ? uhr[iu] : vhr[iv]; P.tr_new = xdir ? uhr[1 - iu] : vhr[1 - iv];
    if (xdir) {
      P.row_old = d_row[irow]; P.row_new = d_row[1 - irow]; P.nrow = nrow; P.jsd = d.jsd;
      M6_CUDA(c, cudaMemsetAsync(d_row[1 - irow], 0, sizeof(int) * (size_t)nrow * nz, c->stream));
    }
    for (int m0 = 0; m0 < ntr; m0 += NTB) {
      P.nt = std::min(NTB, ntr - m0); P.first = (m0 == 0);
      for (int m = 0; m < NTB; ++m) {
        const int mm = std::min(m0 + m, ntr - 1);
        P.T_old[m] = (*Tcur)[mm]; P.T_new[m] = (*Toth)[mm]; P.scheme[m] = scheme[mm];
        P.underflow[m] = a->conc_underflow ? a->conc_underflow[mm] : 0.0;
      }
      if (xdir) {
        const dim3 grid((pie - pis + 1 + (XB - 2)) / (XB - 1), pje - pjs + 1, nz);
        M6_LAUNCH(c, advect_x_kernel, grid, XB, 0, G, P);
      } else {
        const dim3 grid((pie - pis + 1 + 127) / 128, (pje - pjs + 1 + YSEG - 1) / YSEG, nz);
        M6_LAUNCH(c, advect_y_kernel, grid, 128, 0, G, P);
      }
    }
    ihp = 1 - ihp;
    if (xdir) { iu = 1 - iu; irow = 1 - irow; } else iv = 1 - iv;
    std::swap(Tcur, Toth);
    return 0;
  };

  int itt = 0;
  for (itt = 1; itt <= max_iter; ++itt) {
    {  // do_group_pass(CS%pass_uhr_vhr_t_hprev) :224
      std::vector<double*> f = {uhr[iu], vhr[iv], hp[ihp]};
      std::vector<int> st = {ST_U, ST_V, ST_H};
      for (int m = 0; m < ntr; ++m) { f.push_back((*Tcur)[m]); st.push_back(ST_H); }
      for (size_t f0 = 0; f0 < f.size(); f0 += 8)  // halo groups hold at most 8 fields
        if ((rc = m6_halo_update(c, f.data() + f0, st.data() + f0, (int)std::min<size_t>(8, f.size() - f0), 0, nz))) return rc;
    }
    M6_CUDA(c, cudaMemsetAsync(d_flags, 0, sizeof(int), c->stream));
    if (itt == 1) {  // the first evaluation of domore_u (:226-233), on every row an x pass may visit
      M6_CUDA(c, cudaMemsetAsync(d_row[irow], 0, sizeof(int) * (size_t)nrow * nz, c->stream));
      const int r0 = js - stencil, r1 = je + stencil;
      M6_LAUNCH(c, advect_rowflag_kernel, dim3(r1 - r0 + 1, nz), 128, 0, G, uhr[iu], is, ie, r0, nrow, d.jsd, d_row[irow]);
    }
    if (x_first) {
      if ((rc = pass(true, is, ie, js - stencil, je + stencil)) || (rc = pass(false, is, ie, js, je))) return rc;
    } else {
      if ((rc = pass(false, is - stencil, ie + stencil, js, je)) || (rc = pass(true, is, ie, js, je))) return rc;
    }
    M6_CUDA(c, cudaGetLastError());
    if (itt >= max_iter) break;
    if (itt % nsten_halo == 0) {  // the iterations at which the reference sums domore_k across PEs (:323-333)
      M6_CUDA(c, cudaMemcpyAsync(h_flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      M6_CUDA(c, cudaStreamSynchronize(c->stream));
      int any = h_flags[0];
      if (c->nranks > 1 && (rc = m6_allreduce_max_int(c, &any))) return rc;
      if (any == 0) break;
    }
  }
  // results: the tracers are back in their own arrays after each full iteration (two swaps)
  if (Tcur != &TA) return c->fail(MOM6CU_ERR_CUDA, "advect_tracer: internal ping-pong state error");
  const size_t bytes = bytes3;
  if (d_uo) M6_CUDA(c, cudaMemcpyAsync(d_uo, uhr[iu], bytes, cudaMemcpyDeviceToDevice, c->stream));
  if (d_vo) M6_CUDA(c, cudaMemcpyAsync(d_vo, vhr[iv], bytes, cudaMemcpyDeviceToDevice, c->stream));
  if (d_vol && a->update_vol_prev) M6_CUDA(c, cudaMemcpyAsync(d_vol, hp[ihp], bytes, cudaMemcpyDeviceToDevice, c->stream));
  c->last_iterations = std::min(itt, max_iter);
  return S.finish();
}
