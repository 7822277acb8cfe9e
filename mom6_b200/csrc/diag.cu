// The reference's answer-reproducibility metric on the device: reproducing sums (MOM_coms.F90:101-545),
// bit-count checksums (MOM_checksums.F90 chksum_{h,u,v,B}_{2d,3d}) and write_energy (MOM_sum_output.F90:321-1020), so that an
// ocean.stats line or a DEBUG checksum costs two sweeps over the resident state and a few hundred bytes of D2H traffic
// instead of a copy of the model state to the host.
//
//  * efp_sum_kernel: each thread splits its values into the six 46-bit integers of the extended-fixed-point format (efp.cuh)
//    and adds them up; warps combine with integer shuffles, CTAs with shared memory, and one thread per CTA carries and adds the
//    CTA's six integers to the layer's accumulator with 64-bit integer atomics.  Integer addition is associative, so -- unlike
//    every floating sum of the dycore, which stays k-ordered -- these sums may be tree-reduced and still equal the reference's
//    serial loop bit for bit.  A CTA covers at most 65536 values (< max_count_prec = 131071), so nothing can overflow between
//    carries (:171-183 carries per row for the same reason).
//  * bitcount_kernel / minmax_kernel: popcount sums and extrema of a window, the same reduction shape.
//  * write_energy: pass 1 = layer masses, layer kinetic energies, column heat / salt (k-ordered column sums, then a 2-D EFP
//    sum), maximum CFL numbers; the host turns the layer volumes into the zero-APE depths Z_0APE (:642-665, a search in the
//    sorted depth list); pass 2 = the interface available potential energy (a bottom-up column recursion, :674-685).
#include "ctx.h"
#include "efp.cuh"
#include <cmath>
#include <string>
#include <vector>

using m6::Geom;

namespace {

constexpr int EB = 256;   // threads per CTA
constexpr int ACC = 8;    // words per accumulator: 6 integers, bits of max |r|, flags
constexpr int CTA_VALUES = 65536;

struct Win { int is, ie, js, je; };

__device__ __forceinline__ unsigned long long ordered_bits(double x) {  // monotone map double -> u64
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
inline double from_ordered_bits(unsigned long long o) {
  const unsigned long long u = (o >> 63) ? (o & 0x7fffffffffffffffULL) : ~o;
  double x; memcpy(&x, &u, 8); return x;
}

template <class F>
__global__ void __launch_bounds__(EB) efp_sum_kernel(const Geom G, const Win W, const int rows_per_cta, const int acc_stride, const F f,
                                                     unsigned long long* __restrict__ acc) {
  const int k = blockIdx.y;
  const int j0 = W.js + blockIdx.x * rows_per_cta;
  const int j1 = min(W.je + 1, j0 + rows_per_cta);
  long long s[6] = {0, 0, 0, 0, 0, 0};
  double amax = 0.0;
  int flags = 0;
  for (int j = j0; j < j1; ++j)
    for (int i = W.is + threadIdx.x; i <= W.ie; i += EB) flags |= m6efp::accumulate(f(G, i, j, k), s, amax);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int n = 0; n < 6; ++n) s[n] += __shfl_down_sync(0xffffffffu, s[n], off);
    amax = fmax(amax, __shfl_down_sync(0xffffffffu, amax, off));
    flags |= __shfl_down_sync(0xffffffffu, flags, off);
  }
  __shared__ long long sh[EB / 32][ACC];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
#pragma unroll
    for (int n = 0; n < 6; ++n) sh[w][n] = s[n];
    sh[w][6] = __double_as_longlong(amax);
    sh[w][7] = flags;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < EB / 32; ++q) {
#pragma unroll
      for (int n = 0; n < 6; ++n) s[n] += sh[q][n];
      amax = fmax(amax, __longlong_as_double(sh[q][6]));
      flags |= (int)sh[q][7];
    }
    m6efp::carry_exact(s);
    unsigned long long* a = acc + (size_t)k * acc_stride;
#pragma unroll
    for (int n = 0; n < 6; ++n) if (s[n] != 0) atomicAdd(a + n, (unsigned long long)s[n]);
    atomicMax(a + 6, (unsigned long long)__double_as_longlong(amax));  // amax >= 0: the bit patterns order like the values
    if (flags) atomicOr(a + 7, (unsigned long long)flags);
  }
}

// ---- the summands
struct FieldVal {  // descale * array(i,j,k)
  const double* a; double scale;
  __device__ __forceinline__ double operator()(const Geom& G, int i, int j, int k) const { return scale * __ldg(a + G.idx(i, j) + (long long)k * G.plane); }
};
struct MassVal {  // h * (GV%H_to_RZ * areaTm) (:529), areaTm = mask2dT * areaT (:524)
  const double *h, *mask2dT, *areaT; double H_to_RZ, unscale;
  __device__ __forceinline__ double operator()(const Geom& G, int i, int j, int k) const {
    const long long g = G.idx(i, j);
    const double areaTm = __ldg(mask2dT + g) * __ldg(areaT + g);
    return unscale * (__ldg(h + g + (long long)k * G.plane) * (H_to_RZ * areaTm));
  }
};
struct KEVal {  // :715-718
  const double *u, *v, *h, *mask2dT, *areaT; double H_to_RZ, unscale;
  __device__ __forceinline__ double operator()(const Geom& G, int i, int j, int k) const {
    const long long g = G.idx(i, j), gk = g + (long long)k * G.plane;
    const double areaTm = __ldg(mask2dT + g) * __ldg(areaT + g);
    const double uw = __ldg(u + gk - 1), ue = __ldg(u + gk), vs = __ldg(v + gk - G.pitch), vn = __ldg(v + gk);
    const double t = (0.25 * H_to_RZ * (areaTm * __ldg(h + gk))) * (((uw * uw) + (ue * ue)) + ((vs * vs) + (vn * vn)));
    return unscale * t;
  }
};

// column heat and salt content (:725-728): k-ordered floating sums, one thread per column
__global__ void heat_salt_kernel(const Geom G, const Win W, const int nk, const double* __restrict__ h, const double* __restrict__ T,
                                 const double* __restrict__ S, const double* __restrict__ mask2dT, const double* __restrict__ areaT,
                                 const double H_to_RZ, const double C_p, double* __restrict__ Temp_int, double* __restrict__ Salt_int) {
  const int i = W.is + blockIdx.x * blockDim.x + threadIdx.x, j = W.js + blockIdx.y;
  if (i > W.ie) return;
  const long long g = G.idx(i, j);
  const double areaTm = mask2dT[g] * areaT[g];
  double ti = 0.0, si = 0.0;
  for (int k = 0; k < nk; ++k) {
    const long long gk = g + (long long)k * G.plane;
    const double hk = h[gk];
    si = si + S[gk] * (hk * (H_to_RZ * areaTm));
    ti = ti + (C_p * T[gk]) * (hk * (H_to_RZ * areaTm));
  }
  Temp_int[g] = ti; Salt_int[g] = si;
}

// interface available potential energy, Boussinesq (:674-685): bottom-up column recursion
__global__ void ape_kernel(const Geom G, const Win W, const int nk, const double* __restrict__ h, const double* __restrict__ mask2dT,
                           const double* __restrict__ areaT, const double* __restrict__ bathyT, const double* __restrict__ Z_0APE,
                           const double* __restrict__ g_prime, const double H_to_Z, const double Rho0, const double Z_ref,
                           double* __restrict__ PE_pt) {
  const int i = W.is + blockIdx.x * blockDim.x + threadIdx.x, j = W.js + blockIdx.y;
  if (i > W.ie) return;
  const long long g = G.idx(i, j);
  const double areaTm = mask2dT[g] * areaT[g];
  const double D = bathyT[g] + Z_ref;
  double hbelow = 0.0;
  for (int k = nk - 1; k >= 0; --k) {
    hbelow = hbelow + h[g + (long long)k * G.plane] * H_to_Z;
    const double z0 = Z_0APE[k];
    const double hint = z0 + (hbelow - D);
    double hbot = z0 - D;
    hbot = (hbot + fabs(hbot)) * 0.5;
    PE_pt[g + (long long)k * G.plane] = (0.5 * areaTm) * (Rho0 * g_prime[k]) * (hint * hint - hbot * hbot);
  }
  PE_pt[g + (long long)nk * G.plane] = 0.0;  // PE_pt(:,:,nz+1) stays 0 (:672)
}

// maximum CFL numbers (:746-768): exact maxima, any order
__global__ void __launch_bounds__(EB) cfl_kernel(const Geom G, const int isc, const int iec, const int jsc, const int jec, const int rows_per_cta,
                                                 const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ IareaT,
                                                 const double* __restrict__ dy_Cu, const double* __restrict__ dx_Cv,
                                                 const double* __restrict__ IdxCu, const double* __restrict__ IdyCv, const double dt,
                                                 unsigned long long* __restrict__ out) {
  const int k = blockIdx.y;
  const int j0 = (jsc - 1) + blockIdx.x * rows_per_cta, j1 = min(jec + 1, j0 + rows_per_cta);
  double m0 = 0.0, m1 = 0.0;
  for (int j = j0; j < j1; ++j)
    for (int i = (isc - 1) + threadIdx.x; i <= iec; i += EB) {
      const long long g = G.idx(i, j), gk = g + (long long)k * G.plane;
      if (j >= jsc) {  // u(I,j,k), I = Isq..Ieq
        const double uk = __ldg(u + gk);
        double CFL_Iarea = __ldg(IareaT + g);
        if (uk < 0.0) CFL_Iarea = __ldg(IareaT + g + 1);
        const double CFL_trans = fabs(uk * dt) * (__ldg(dy_Cu + g) * CFL_Iarea);
        const double CFL_lin = fabs(uk * dt) * __ldg(IdxCu + g);
        m0 = (m0 > CFL_trans) ? m0 : CFL_trans; m1 = (m1 > CFL_lin) ? m1 : CFL_lin;
      }
      if (i >= isc) {  // v(i,J,k), J = Jsq..Jeq
        const double vk = __ldg(v + gk);
        double CFL_Iarea = __ldg(IareaT + g);
        if (vk < 0.0) CFL_Iarea = __ldg(IareaT + g + G.pitch);
        const double CFL_trans = fabs(vk * dt) * (__ldg(dx_Cv + g) * CFL_Iarea);
        const double CFL_lin = fabs(vk * dt) * __ldg(IdyCv + g);
        m0 = (m0 > CFL_trans) ? m0 : CFL_trans; m1 = (m1 > CFL_lin) ? m1 : CFL_lin;
      }
    }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double a = __shfl_down_sync(0xffffffffu, m0, off), b = __shfl_down_sync(0xffffffffu, m1, off);
    m0 = (m0 > a) ? m0 : a; m1 = (m1 > b) ? m1 : b;
  }
  if ((threadIdx.x & 31) == 0) {  // non-negative numbers: the bit patterns order like the values (a NaN never wins a '>' above)
    atomicMax(out, (unsigned long long)__double_as_longlong(m0));
    atomicMax(out + 1, (unsigned long long)__double_as_longlong(m1));
  }
}

// subchk (MOM_checksums.F90:520-529): sum of popcnt(transfer(abs(scale*x)))
__global__ void __launch_bounds__(EB) bitcount_kernel(const Geom G, const Win W, const int rows_per_cta, const double* __restrict__ a,
                                                      const double scaling, unsigned long long* __restrict__ out) {
  const int k = blockIdx.y;
  const int j0 = W.js + blockIdx.x * rows_per_cta, j1 = min(W.je + 1, j0 + rows_per_cta);
  unsigned long long s = 0;
  for (int j = j0; j < j1; ++j)
    for (int i = W.is + threadIdx.x; i <= W.ie; i += EB)
      s += __popcll((unsigned long long)__double_as_longlong(fabs(scaling * __ldg(a + G.idx(i, j) + (long long)k * G.plane))));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// subStats extrema (MOM_checksums.F90:531-553)
__global__ void __launch_bounds__(EB) minmax_kernel(const Geom G, const Win W, const int rows_per_cta, const double* __restrict__ a,
                                                    const double scaling, unsigned long long* __restrict__ out) {
  const int k = blockIdx.y;
  const int j0 = W.js + blockIdx.x * rows_per_cta, j1 = min(W.je + 1, j0 + rows_per_cta);
  unsigned long long lo = ~0ULL, hi = 0ULL;
  for (int j = j0; j < j1; ++j)
    for (int i = W.is + threadIdx.x; i <= W.ie; i += EB) {
      const double x = scaling * __ldg(a + G.idx(i, j) + (long long)k * G.plane);
      if (x == x) { const unsigned long long o = ordered_bits(x); lo = o < lo ? o : lo; hi = o > hi ? o : hi; }
    }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long a2 = __shfl_down_sync(0xffffffffu, lo, off), b2 = __shfl_down_sync(0xffffffffu, hi, off);
    lo = a2 < lo ? a2 : lo; hi = b2 > hi ? b2 : hi;
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(out, lo); atomicMax(out + 1, hi); }
}

inline int rows_per_cta(const Win& W) { const int ni = W.ie - W.is + 1; const int r = CTA_VALUES / (ni > 0 ? ni : 1); return r < 1 ? 1 : r; }

template <class F>
int launch_sum(mom6cu_ctx* c, const Win& W, int nlay, int acc_stride, const F& f, unsigned long long* d_acc) {
  if (W.ie < W.is || W.je < W.js || nlay < 1) return 0;
  if (W.ie - W.is + 1 > CTA_VALUES) return c->fail(MOM6CU_ERR_UNSUPPORTED, "reproducing_sum: rows of more than %d values are not supported", CTA_VALUES);
  const int rpc = rows_per_cta(W);
  const dim3 grid((W.je - W.js + 1 + rpc - 1) / rpc, nlay);
  M6_LAUNCH(c, efp_sum_kernel<F>, grid, EB, 0, c->g, W, rpc, acc_stride, f, d_acc);
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

// Host side of a set of device accumulators: the reference's error tests, sum_across_PEs, regularize_ints.
// acc: nacc x ACC words as the device left them.  ints: nacc x 6, regularised on return.
int finish_sums(mom6cu_ctx* c, const unsigned long long* acc, int nacc, bool across_PEs, const char* who, long long* ints) {
  const long long prec_error = 0x7fffffffffffffffLL / (c->nranks > 0 ? c->nranks : 1);  // ((2**62 + (2**62 - 1)) / num_PEs()
  int err = 0;
  for (int a = 0; a < nacc; ++a) {
    const unsigned long long* w = acc + (size_t)a * ACC;
    long long* s = ints + (size_t)a * 6;
    for (int n = 0; n < 6; ++n) s[n] = (long long)w[n];
    double amax; memcpy(&amax, &w[6], 8);
    if (w[7] & m6efp::FLAG_NAN) err = err > 1 ? err : 1;
    else if (amax >= (double)prec_error * m6efp::PR0) err = err > 2 ? err : 2;
    else if (w[7] & m6efp::FLAG_OVERFLOW) err = 3;
    m6efp::carry_exact(s);
    if (m6efp::iabs(s[0]) > prec_error) err = 3;
  }
  if (across_PEs && c->nranks > 1) {
    int rc = m6_allreduce_max_int(c, &err);
    if (rc) return rc;
  }
  if (err == 1) return c->fail(MOM6CU_ERR_BAD_ARG, "NaN in input field of %s.", who);
  if (err == 2) return c->fail(MOM6CU_ERR_BAD_ARG, "Overflow in %s conversion.", who);
  if (err == 3) return c->fail(MOM6CU_ERR_BAD_ARG, "Overflow in %s.", who);
  if (across_PEs && c->nranks > 1) {
    int rc = m6_allreduce_sum_i64(c, ints, nacc * 6);
    if (rc) return rc;
  }
  for (int a = 0; a < nacc; ++a) m6efp::regularize(ints + (size_t)a * 6);
  return 0;
}

// device accumulators + their pinned host mirror
struct AccBuf {
  unsigned long long *d = nullptr, *h = nullptr;
  int n = 0;
  int init(mom6cu_ctx* c, int nacc) {
    n = nacc;
    d = (unsigned long long*)c->buf("diag.acc", (size_t)nacc * ACC);
    h = (unsigned long long*)c->host_scratch("diag.acc", (size_t)nacc * ACC);
    if (!d || !h) return MOM6CU_ERR_CUDA;
    M6_CUDA(c, cudaMemsetAsync(d, 0, sizeof(unsigned long long) * (size_t)nacc * ACC, c->stream));
    return 0;
  }
  int fetch(mom6cu_ctx* c) {
    M6_CUDA(c, cudaMemcpyAsync(h, d, sizeof(unsigned long long) * (size_t)n * ACC, cudaMemcpyDeviceToHost, c->stream));
    M6_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
  }
};

// the layer-sum epilogue of reproducing_sum_3d :453-476, :535-543: sums(k), their k-ordered total, the EFP of the total
void layer_totals(const long long* ints, int nk, double unscale, double* sums, double* total, mom6cu_efp* EFP_sum) {
  double t = 0.0;
  for (int k = 0; k < nk; ++k) {
    const double val = m6efp::to_real(ints + (size_t)k * 6);
    if (sums) sums[k] = val;
    t = t + val;
  }
  if (EFP_sum) {
    long long e[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < nk; ++k) m6efp::increment(e, ints + (size_t)k * 6, -1);
    for (int n = 0; n < 6; ++n) EFP_sum->v[n] = e[n];
  }
  if (unscale != 1.0) {
    double I_unscale = 0.0;
    if (fabs(unscale) > 0.0) I_unscale = 1.0 / unscale;
    t = t * I_unscale;
    if (sums) for (int k = 0; k < nk; ++k) sums[k] = sums[k] * I_unscale;
  }
  *total = t;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ EFP operators (host)
extern "C" void mom6cu_efp_plus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, int* overflow) {
  long long s[6];
  for (int n = 0; n < 6; ++n) s[n] = a->v[n];
  const bool o = m6efp::increment(s, (const long long*)b->v, -1);
  for (int n = 0; n < 6; ++n) out->v[n] = s[n];
  if (overflow) *overflow = o ? 1 : 0;
}
extern "C" void mom6cu_efp_minus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, int* overflow) {
  long long s[6];
  for (int n = 0; n < 6; ++n) s[n] = -1 * b->v[n];
  const bool o = m6efp::increment(s, (const long long*)a->v, -1);
  for (int n = 0; n < 6; ++n) out->v[n] = s[n];
  if (overflow) *overflow = o ? 1 : 0;
}
extern "C" double mom6cu_efp_to_real(mom6cu_efp* a) {
  m6efp::regularize((long long*)a->v);
  return m6efp::to_real((const long long*)a->v);
}
extern "C" int mom6cu_real_to_efp(double val, mom6cu_efp* out) {
  const int fl = m6efp::from_real(val, m6efp::PREC, (long long*)out->v);
  return (fl & m6efp::FLAG_NAN) ? 2 : ((fl & m6efp::FLAG_OVERFLOW) ? 1 : 0);
}
extern "C" double mom6cu_efp_real_diff(const mom6cu_efp* a, const mom6cu_efp* b) {
  mom6cu_efp d;
  mom6cu_efp_minus(a, b, &d, nullptr);
  return mom6cu_efp_to_real(&d);
}

extern "C" int mom6cu_efp_sum_across_pes(mom6cu_ctx* c, mom6cu_efp* EFPs, int nval) {
  if (!c || !EFPs || nval < 0) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  const long long prec_error = 0x7fffffffffffffffLL / (c->nranks > 0 ? c->nranks : 1);
  std::vector<long long> ints((size_t)nval * 6);
  for (int i = 0; i < nval; ++i) for (int n = 0; n < 6; ++n) ints[(size_t)i * 6 + n] = EFPs[i].v[n];
  int rc = m6_allreduce_sum_i64(c, ints.data(), nval * 6);
  if (rc) return rc;
  bool over = false;
  for (int i = 0; i < nval; ++i) {
    over = m6efp::carry_overflow(&ints[(size_t)i * 6], prec_error) || over;
    for (int n = 0; n < 6; ++n) EFPs[i].v[n] = ints[(size_t)i * 6 + n];
  }
  if (over) return c->fail(MOM6CU_ERR_BAD_ARG, "Overflow in EFP_list_sum_across_PEs.");
  return 0;
}

// ------------------------------------------------------------------------------------------------ reproducing_sum
extern "C" int mom6cu_reproducing_sum(mom6cu_ctx* c, const double* array, int stagger, int nk, int isr, int ier, int jsr, int jer,
                                      double unscale, int only_on_PE, double* sum, double* sums, mom6cu_efp* EFP_sum,
                                      mom6cu_efp* EFP_lay_sums) {
  if (!c || !array || !sum || nk < 1 || stagger < 0 || stagger > 3) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  int ilo, ihi, jlo, jhi;
  m6_extent(c, stagger, 0, &ilo, &ihi, &jlo, &jhi);
  int is = 1, ie = ihi - ilo + 1, js = 1, je = jhi - jlo + 1;
  if (isr > 0) { if (isr < is) return c->fail(MOM6CU_ERR_BAD_ARG, "Value of isr too small in reproducing_sum."); is = isr; }
  if (ier > 0) { if (ier > ie) return c->fail(MOM6CU_ERR_BAD_ARG, "Value of ier too large in reproducing_sum."); ie = ier; }
  if (jsr > 0) { if (jsr < js) return c->fail(MOM6CU_ERR_BAD_ARG, "Value of jsr too small in reproducing_sum."); js = jsr; }
  if (jer > 0) { if (jer > je) return c->fail(MOM6CU_ERR_BAD_ARG, "Value of jer too large in reproducing_sum."); je = jer; }
  const Win W = {ilo + is - 1, ilo + ie - 1, jlo + js - 1, jlo + je - 1};
  Stager S(c, "rsum.");
  int rc;
  const double* d_a;
  if ((rc = S.in(array, stagger, 0, nk, "a", &d_a))) return rc;
  AccBuf A;
  if ((rc = A.init(c, nk))) return rc;
  if ((rc = S.begin())) return rc;
  if ((rc = launch_sum(c, W, nk, ACC, FieldVal{d_a, unscale}, A.d))) return rc;
  if ((rc = S.finish()) || (rc = A.fetch(c))) return rc;
  std::vector<long long> ints((size_t)nk * 6);
  const bool layered = (sums != nullptr) || (EFP_lay_sums != nullptr);
  if (layered || nk == 1) {
    if ((rc = finish_sums(c, A.h, nk, !only_on_PE, nk == 1 && !layered ? "reproducing_EFP_sum(_2d)" : "reproducing_sum(_3d)", ints.data()))) return rc;
    if (nk == 1 && !layered) {  // reproducing_sum_2d :286-289
      double I_unscale = 1.0;
      if (unscale != 1.0 && fabs(unscale) > 0.0) I_unscale = 1.0 / unscale;
      *sum = m6efp::to_real(ints.data()) * I_unscale;
      if (EFP_sum) for (int n = 0; n < 6; ++n) EFP_sum->v[n] = ints[n];
      return 0;
    }
    layer_totals(ints.data(), nk, unscale, sums, sum, EFP_sum);
    if (EFP_lay_sums) for (int k = 0; k < nk; ++k) for (int n = 0; n < 6; ++n) EFP_lay_sums[k].v[n] = ints[(size_t)k * 6 + n];
    return 0;
  }
  // one accumulator for all layers (:485-532): fold the layers' integers (exact), then regularise once
  std::vector<unsigned long long> tot(ACC, 0ULL);
  {
    long long s[6] = {0, 0, 0, 0, 0, 0};
    double amax = 0.0; unsigned long long fl = 0;
    for (int k = 0; k < nk; ++k) {
      long long t[6];
      for (int n = 0; n < 6; ++n) t[n] = (long long)A.h[(size_t)k * ACC + n];
      m6efp::carry_exact(t);
      for (int n = 0; n < 6; ++n) s[n] += t[n];
      m6efp::carry_exact(s);
      double am; memcpy(&am, &A.h[(size_t)k * ACC + 6], 8);
      amax = am > amax ? am : amax;
      fl |= A.h[(size_t)k * ACC + 7];
    }
    for (int n = 0; n < 6; ++n) tot[n] = (unsigned long long)s[n];
    memcpy(&tot[6], &amax, 8); tot[7] = fl;
  }
  long long one[6];
  if ((rc = finish_sums(c, tot.data(), 1, !only_on_PE, "reproducing_sum(_3d)", one))) return rc;
  double t = m6efp::to_real(one);
  if (EFP_sum) for (int n = 0; n < 6; ++n) EFP_sum->v[n] = one[n];
  if (unscale != 1.0) { double I_unscale = 0.0; if (fabs(unscale) > 0.0) I_unscale = 1.0 / unscale; t = t * I_unscale; }
  *sum = t;
  return 0;
}

// ------------------------------------------------------------------------------------------------ checksums
extern "C" int mom6cu_chksum(mom6cu_ctx* c, const double* array, int stagger, int nk, int haloshift, int symmetric, int omit_corners,
                             double scale, int* bc, int* kind, double* stats) {
  if (!c || !array || !bc || !kind || nk < 1 || stagger < 0 || stagger > 3) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  const mom6cu_domain& d = c->dom;
  const bool su = (stagger == ST_U || stagger == ST_Q), sv = (stagger == ST_V || stagger == ST_Q);
  const bool sym = symmetric != 0 && stagger != ST_H;
  int hshift = haloshift;
  if (hshift < 0) hshift = d.ied - d.iec;
  if (d.isc - hshift < d.isd || d.iec + hshift > d.ied || d.jsc - hshift < d.jsd || d.jec + hshift > d.jed)
    return c->fail(MOM6CU_ERR_BAD_ARG, "Error in chksum_%s_%dd: haloshift = %d is wider than the halo", stagger == 0 ? "h" : stagger == 1 ? "u" : stagger == 2 ? "v" : "B",
                   nk > 1 ? 3 : 2, hshift);
  // the windows (di, dj) of the case at hand, in the order they are reported
  int win[5][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};
  int nwin = 1;
  const bool plain = (stagger == ST_H) ? (hshift == 0) : ((hshift == 0) && !sym);
  // chksum_B_3d differs from chksum_B_2d (MOM_checksums.F90:1698-1718 against :797-809): its corner windows take the extra row and
  // column with or without `symmetric`, and with omit_corners its S and W windows widen under `symmetric`.  nk > 1 = a rank-3 array.
  const bool q3 = (stagger == ST_Q && nk > 1);
  const int ex = ((sym || q3) && su) ? 1 : 0, ey = ((sym || q3) && sv) ? 1 : 0;
  if (plain) *kind = 1;
  else if (hshift == 0 && stagger == ST_U) { *kind = 4; nwin = 2; win[1][0] = -1; }
  else if (hshift == 0 && stagger == ST_V) { *kind = 5; nwin = 2; win[1][1] = -1; }
  else if (!omit_corners) {
    *kind = 2; nwin = 5;
    win[1][0] = -hshift - ex; win[1][1] = -hshift - ey;  // SW
    win[2][0] = hshift;       win[2][1] = -hshift - ey;  // SE
    win[3][0] = -hshift - ex; win[3][1] = hshift;        // NW
    win[4][0] = hshift;       win[4][1] = hshift;        // NE
  } else {
    *kind = 3; nwin = 5;
    win[1][1] = hshift;                                          // N
    win[2][1] = -hshift - (((stagger == ST_V || q3) && sym) ? 1 : 0);    // S
    win[3][0] = hshift;                                                  // E
    win[4][0] = -hshift - (((stagger == ST_U || q3) && sym) ? 1 : 0);    // W
  }
  Stager S(c, "chk.");
  int rc;
  const double* d_a;
  if ((rc = S.in(array, stagger, 0, nk, "a", &d_a))) return rc;
  // words: [0..4] bit counts, [8], [9] ordered min / max, then one EFP accumulator
  unsigned long long* d_w = (unsigned long long*)c->buf("chk.words", 16 + ACC);
  unsigned long long* h_w = (unsigned long long*)c->host_scratch("chk.words", 16 + ACC);
  if (!d_w || !h_w) return MOM6CU_ERR_CUDA;
  M6_CUDA(c, cudaMemsetAsync(d_w, 0, sizeof(unsigned long long) * (16 + ACC), c->stream));
  M6_CUDA(c, cudaMemsetAsync(d_w + 8, 0xff, sizeof(unsigned long long), c->stream));
  if ((rc = S.begin())) return rc;
  for (int q = 0; q < nwin; ++q) {
    const Win W = {d.isc + win[q][0], d.iec + win[q][0], d.jsc + win[q][1], d.jec + win[q][1]};
    const int rpc = rows_per_cta(W);
    const dim3 grid((W.je - W.js + 1 + rpc - 1) / rpc, nk);
    M6_LAUNCH(c, bitcount_kernel, grid, EB, 0, c->g, W, rpc, d_a, scale, d_w + q);
  }
  if (stats) {
    const bool sym_stats = sym || (haloshift > 0 && stagger != ST_H);
    const Win Wm = {d.isc - ((su && sym_stats) ? 1 : 0), d.iec, d.jsc - ((sv && sym_stats) ? 1 : 0), d.jec};
    const int rpc = rows_per_cta(Wm);
    const dim3 grid((Wm.je - Wm.js + 1 + rpc - 1) / rpc, nk);
    M6_LAUNCH(c, minmax_kernel, grid, EB, 0, c->g, Wm, rpc, d_a, scale, d_w + 8);
    const Win Wh = {d.isc, d.iec, d.jsc, d.jec};
    if ((rc = launch_sum(c, Wh, nk, 0, FieldVal{d_a, scale}, d_w + 16))) return rc;
  }
  M6_CUDA(c, cudaGetLastError());
  M6_CUDA(c, cudaMemcpyAsync(h_w, d_w, sizeof(unsigned long long) * (16 + ACC), cudaMemcpyDeviceToHost, c->stream));
  if ((rc = S.finish())) return rc;
  for (int q = 0; q < 5; ++q) bc[q] = 0;
  for (int q = 0; q < nwin; ++q) {
    // the reference accumulates in a default (32-bit) INTEGER, sums it across PEs, then takes mod(., bc_modulus)
    long long tot = (long long)h_w[q];
    if (c->nranks > 1) { long long v = (long long)(int)(unsigned)tot; if ((rc = m6_allreduce_sum_i64(c, &v, 1))) return rc; tot = v; }
    bc[q] = (int)(unsigned)tot % 1000000000;
  }
  if (stats) {
    double mm[2] = {from_ordered_bits(h_w[8]), from_ordered_bits(h_w[9])};
    long long ints[6];
    if ((rc = finish_sums(c, h_w + 16, 1, true, "reproducing_sum", ints))) return rc;
    long long n = (long long)(d.iec - d.isc + 1) * (d.jec - d.jsc + 1) * nk;
    if (c->nranks > 1) {
      if ((rc = m6_allreduce_sum_i64(c, &n, 1))) return rc;
      double lo = mm[0], hi = mm[1];
      if ((rc = m6_allreduce_min_doubles(c, &lo, 1)) || (rc = m6_allreduce_max_doubles(c, &hi, 1))) return rc;
      mm[0] = lo; mm[1] = hi;
    }
    stats[0] = m6efp::to_real(ints) / (double)n;
    stats[1] = mm[0]; stats[2] = mm[1];
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ write_energy
extern "C" int mom6cu_write_energy(mom6cu_ctx* c, mom6cu_sum_output_cs* CS, const double* u, const double* v, const double* h,
                                   const double* T, const double* Sal, mom6cu_energy_out* out) {
  if (!c || !CS || !u || !v || !h || !out) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (!c->have_grid || !c->have_vgrid) return c->fail(MOM6CU_ERR_BAD_ARG, "write_energy: mom6cu_set_grid / mom6cu_set_vgrid have not been called");
  if (!c->vgrid.Boussinesq) return c->fail(MOM6CU_ERR_UNSUPPORTED, "write_energy: only the Boussinesq branch (:534, :674) is implemented");
  if (CS->use_temperature && (!T || !Sal)) return c->fail(MOM6CU_ERR_BAD_ARG, "write_energy: use_temperature without tv%%T / tv%%S");
  if (CS->do_APE_calc && (!CS->DL_depth || !CS->DL_area || !CS->DL_vol_below || !CS->lH || !CS->g_prime || CS->DL_listsize < 3))
    return c->fail(MOM6CU_ERR_BAD_ARG, "write_energy: do_APE_calc without the depth list (depth_list_setup) or GV%%g_prime");
  const Geom& G = c->g;
  const mom6cu_domain& d = c->dom;
  const mom6cu_vgrid& GV = c->vgrid;
  const int nz = G.nk;
  const Win W = {d.isc, d.iec, d.jsc, d.jec};
  const double RZL4_T2_to_J = CS->RZL2_to_kg * (CS->L_T_to_m_s * CS->L_T_to_m_s);
  const double kg_to_RZL2 = CS->kg_m3_to_R * CS->m_to_Z * (CS->m_to_L * CS->m_to_L);
  const double J_to_QRZL2 = CS->J_kg_to_Q * kg_to_RZL2;
  Stager S(c, "we.");
  int rc;
  const double *d_u, *d_v, *d_h, *d_T = nullptr, *d_S = nullptr;
  if ((rc = S.in3(u, ST_U, "u", &d_u)) || (rc = S.in3(v, ST_V, "v", &d_v)) || (rc = S.in3(h, ST_H, "h", &d_h))) return rc;
  if (CS->use_temperature && ((rc = S.in3(T, ST_H, "T", &d_T)) || (rc = S.in3(Sal, ST_H, "S", &d_S)))) return rc;
  // accumulators: [0,nz) layer mass, [nz,2nz) layer KE, 2nz salt, 2nz+1 heat, then nz+1 interface PE; last: the two CFL maxima
  const int nacc = 2 * nz + 2 + (nz + 1) + 1;
  AccBuf A;
  if ((rc = A.init(c, nacc))) return rc;
  unsigned long long* d_cfl = A.d + (size_t)(nacc - 1) * ACC;
  if ((rc = S.begin())) return rc;
  const GridDev& Gd = c->grid;
  // ---- pass 1
  if ((rc = launch_sum(c, W, nz, ACC, MassVal{d_h, Gd.mask2dT, Gd.areaT, GV.H_to_RZ, CS->RZL2_to_kg}, A.d))) return rc;
  if ((rc = launch_sum(c, W, nz, ACC, KEVal{d_u, d_v, d_h, Gd.mask2dT, Gd.areaT, GV.H_to_RZ, RZL4_T2_to_J}, A.d + (size_t)nz * ACC))) return rc;
  if (CS->use_temperature) {
    double* d_Ti = c->plane2("we.Temp_int");
    double* d_Si = c->plane2("we.Salt_int");
    if (!d_Ti || !d_Si) return MOM6CU_ERR_CUDA;
    const dim3 grid((d.iec - d.isc + 1 + 127) / 128, d.jec - d.jsc + 1);
    M6_LAUNCH(c, heat_salt_kernel, grid, 128, 0, G, W, nz, d_h, d_T, d_S, Gd.mask2dT, Gd.areaT, GV.H_to_RZ, CS->C_p, d_Ti, d_Si);
    if ((rc = launch_sum(c, W, 1, ACC, FieldVal{d_Si, CS->RZL2_to_kg * CS->S_to_ppt}, A.d + (size_t)(2 * nz) * ACC))) return rc;
    if ((rc = launch_sum(c, W, 1, ACC, FieldVal{d_Ti, CS->RZL2_to_kg * CS->Q_to_J_kg}, A.d + (size_t)(2 * nz + 1) * ACC))) return rc;
  }
  {
    const Win Wc = {d.isc - 1, d.iec, d.jsc - 1, d.jec};
    const int rpc = rows_per_cta(Wc);
    const dim3 grid((Wc.je - Wc.js + 1 + rpc - 1) / rpc, nz);
    M6_LAUNCH(c, cfl_kernel, grid, EB, 0, G, d.isc, d.iec, d.jsc, d.jec, rpc, d_u, d_v, Gd.IareaT, Gd.dy_Cu, Gd.dx_Cv, Gd.IdxCu, Gd.IdyCv, CS->dt_in_T, d_cfl);
  }
  M6_CUDA(c, cudaGetLastError());
  if ((rc = A.fetch(c))) return rc;
  std::vector<long long> ints((size_t)nacc * 6);
  std::vector<double> mass_lay(nz), vol_lay(nz), KE(nz), PE(nz + 1, 0.0), Z_0APE(nz + 1, 0.0);
  mom6cu_efp mass_EFP, salt_EFP = {}, heat_EFP = {};
  double mass_tot, KE_tot, PE_tot = 0.0;
  if ((rc = finish_sums(c, A.h, nz, true, "reproducing_sum(_3d)", ints.data()))) return rc;
  layer_totals(ints.data(), nz, CS->RZL2_to_kg, mass_lay.data(), &mass_tot, &mass_EFP);
  for (int k = 0; k < nz; ++k) vol_lay[k] = (1.0 / GV.Rho0) * mass_lay[k];  // :535
  if ((rc = finish_sums(c, A.h + (size_t)nz * ACC, nz, true, "reproducing_sum(_3d)", ints.data()))) return rc;
  layer_totals(ints.data(), nz, RZL4_T2_to_J, KE.data(), &KE_tot, nullptr);

  if (CS->previous_calls == 0) {  // :578-584
    CS->mass_prev_EFP = mass_EFP;
    mom6cu_real_to_efp(0.0, &CS->fresh_water_in_EFP);
    if (CS->use_temperature) { mom6cu_real_to_efp(0.0, &CS->net_salt_in_EFP); mom6cu_real_to_efp(0.0, &CS->net_heat_in_EFP); }
  }

  // ---- pass 2: the zero-APE depths (:642-665, host) and the interface APE
  if (CS->do_APE_calc) {
    const double* DLv = CS->DL_vol_below - 1; const double* DLd = CS->DL_depth - 1; const double* DLa = CS->DL_area - 1;  // 1-based
    int* lH = CS->lH - 1;
    int lbelow = 1, li = 0;
    double volbelow = 0.0;
    for (int k = nz; k >= 1; --k) {
      volbelow = volbelow + vol_lay[k - 1];
      if (lH[k] >= 1 && lH[k] < CS->DL_listsize && (volbelow >= DLv[lH[k]]) && (volbelow < DLv[lH[k] + 1])) li = lH[k];
      else {
        int labove = CS->DL_listsize;
        li = (labove + lbelow) / 2;
        while (li > lbelow) {
          if (volbelow < DLv[li]) labove = li; else lbelow = li;
          li = (labove + lbelow) / 2;
        }
        lH[k] = li;
      }
      lbelow = li;
      Z_0APE[k - 1] = DLd[li] - (volbelow - DLv[li]) / DLa[li];
    }
    Z_0APE[nz] = DLd[2];
    double* d_z0 = c->buf("we.Z_0APE", (size_t)nz + 1);
    double* d_gp = c->buf("we.g_prime", (size_t)nz + 1);
    double* h_z0 = c->host_scratch("we.Z_0APE", 2 * ((size_t)nz + 1));
    double* d_PE = c->plane3k("we.PE_pt", nz + 1);
    if (!d_z0 || !d_gp || !h_z0 || !d_PE) return MOM6CU_ERR_CUDA;
    for (int k = 0; k <= nz; ++k) { h_z0[k] = Z_0APE[k]; h_z0[nz + 1 + k] = CS->g_prime[k]; }
    M6_CUDA(c, cudaMemcpyAsync(d_z0, h_z0, sizeof(double) * (nz + 1), cudaMemcpyHostToDevice, c->stream));
    M6_CUDA(c, cudaMemcpyAsync(d_gp, h_z0 + nz + 1, sizeof(double) * (nz + 1), cudaMemcpyHostToDevice, c->stream));
    const dim3 grid((d.iec - d.isc + 1 + 127) / 128, d.jec - d.jsc + 1);
    M6_LAUNCH(c, ape_kernel, grid, 128, 0, G, W, nz, d_h, Gd.mask2dT, Gd.areaT, Gd.bathyT, d_z0, d_gp, GV.H_to_Z, GV.Rho0, CS->Z_ref, d_PE);
    unsigned long long* d_pe_acc = A.d + (size_t)(2 * nz + 2) * ACC;
    if ((rc = launch_sum(c, W, nz + 1, ACC, FieldVal{d_PE, RZL4_T2_to_J}, d_pe_acc))) return rc;
    M6_CUDA(c, cudaGetLastError());
    if ((rc = A.fetch(c))) return rc;
    if ((rc = finish_sums(c, A.h + (size_t)(2 * nz + 2) * ACC, nz + 1, true, "reproducing_sum(_3d)", ints.data()))) return rc;
    layer_totals(ints.data(), nz + 1, RZL4_T2_to_J, PE.data(), &PE_tot, nullptr);
  }
  if ((rc = S.finish())) return rc;

  const long long prec_error = 0x7fffffffffffffffLL / (c->nranks > 0 ? c->nranks : 1);
  if (CS->use_temperature) {  // :729-744
    long long two[12];
    if ((rc = finish_sums(c, A.h + (size_t)(2 * nz) * ACC, 2, false, "reproducing_EFP_sum(_2d)", two))) return rc;
    for (int n = 0; n < 6; ++n) { salt_EFP.v[n] = two[n]; heat_EFP.v[n] = two[6 + n]; }
    mom6cu_efp list[5] = {salt_EFP, heat_EFP, CS->fresh_water_in_EFP, CS->net_salt_in_EFP, CS->net_heat_in_EFP};
    if ((rc = mom6cu_efp_sum_across_pes(c, list, 5))) return rc;
    salt_EFP = list[0]; heat_EFP = list[1]; CS->fresh_water_in_EFP = list[2]; CS->net_salt_in_EFP = list[3]; CS->net_heat_in_EFP = list[4];
  } else {
    if ((rc = mom6cu_efp_sum_across_pes(c, &CS->fresh_water_in_EFP, 1))) return rc;
  }
  (void)prec_error;

  double max_CFL[2];
  memcpy(max_CFL, A.h + (size_t)(nacc - 1) * ACC, 16);
  if (c->nranks > 1) {  // :770-772
    long long nt = CS->ntrunc;
    if ((rc = m6_allreduce_sum_i64(c, &nt, 1)) || (rc = m6_allreduce_max_doubles(c, max_CFL, 2))) return rc;
    CS->ntrunc = (int)nt;
  }

  double Salt = 0.0, Heat = 0.0, Salt_chg = 0.0, Salt_anom = 0.0, Heat_chg = 0.0, Heat_anom = 0.0;
  if (CS->use_temperature) {  // :774-790
    Salt = kg_to_RZL2 * mom6cu_efp_to_real(&salt_EFP);
    Heat = J_to_QRZL2 * mom6cu_efp_to_real(&heat_EFP);
    if (CS->previous_calls == 0) { CS->salt_prev_EFP = salt_EFP; CS->heat_prev_EFP = heat_EFP; }
    mom6cu_efp chg, anom;
    mom6cu_efp_minus(&salt_EFP, &CS->salt_prev_EFP, &chg, nullptr);
    Salt_chg = kg_to_RZL2 * mom6cu_efp_to_real(&chg);
    mom6cu_efp_minus(&chg, &CS->net_salt_in_EFP, &anom, nullptr);
    Salt_anom = kg_to_RZL2 * mom6cu_efp_to_real(&anom);
    mom6cu_efp_minus(&heat_EFP, &CS->heat_prev_EFP, &chg, nullptr);
    Heat_chg = J_to_QRZL2 * mom6cu_efp_to_real(&chg);
    mom6cu_efp_minus(&chg, &CS->net_heat_in_EFP, &anom, nullptr);
    Heat_anom = J_to_QRZL2 * mom6cu_efp_to_real(&anom);
  }
  mom6cu_efp mass_chg_EFP, mass_anom_EFP;
  mom6cu_efp_minus(&mass_EFP, &CS->mass_prev_EFP, &mass_chg_EFP, nullptr);
  mom6cu_efp_minus(&mass_chg_EFP, &CS->fresh_water_in_EFP, &mass_anom_EFP, nullptr);
  const double mass_anom = kg_to_RZL2 * mom6cu_efp_to_real(&mass_anom_EFP);
  const double mass_chg = kg_to_RZL2 * mom6cu_efp_to_real(&mass_chg_EFP);
  double salin = 0.0, salin_anom = 0.0, temp = 0.0, temp_anom = 0.0;
  if (CS->use_temperature) {
    salin = Salt / mass_tot;
    salin_anom = Salt_anom / mass_tot;
    temp = Heat / (mass_tot * CS->C_p);
    temp_anom = Heat_anom / (mass_tot * CS->C_p);
  }
  const double toten = KE_tot + PE_tot;
  const double En_mass = toten / mass_tot;
  out->En_mass = En_mass; out->toten = toten; out->KE_tot = KE_tot; out->PE_tot = PE_tot; out->mass_tot = mass_tot;
  out->mass_chg = mass_chg; out->mass_anom = mass_anom; out->max_CFL[0] = max_CFL[0]; out->max_CFL[1] = max_CFL[1];
  out->Salt = Salt; out->Salt_chg = Salt_chg; out->Salt_anom = Salt_anom; out->Heat = Heat; out->Heat_chg = Heat_chg;
  out->Heat_anom = Heat_anom; out->salin = salin; out->salin_anom = salin_anom; out->temp = temp; out->temp_anom = temp_anom;
  out->ntrunc = CS->ntrunc;
  if (out->KE) for (int k = 0; k < nz; ++k) out->KE[k] = KE[k];
  if (out->mass_lay) for (int k = 0; k < nz; ++k) out->mass_lay[k] = mass_lay[k];
  if (out->PE) for (int k = 0; k <= nz; ++k) out->PE[k] = PE[k];
  if (out->Z_0APE) for (int k = 0; k <= nz; ++k) out->Z_0APE[k] = Z_0APE[k];
  if (En_mass != En_mass) return c->fail(MOM6CU_ERR_BAD_ARG, "write_energy : NaNs in total model energy forced model termination.");
  CS->ntrunc = 0;  // :1010-1018
  CS->previous_calls = CS->previous_calls + 1;
  CS->mass_prev_EFP = mass_EFP; mom6cu_real_to_efp(0.0, &CS->fresh_water_in_EFP);
  if (CS->use_temperature) {
    CS->salt_prev_EFP = salt_EFP; mom6cu_real_to_efp(0.0, &CS->net_salt_in_EFP);
    CS->heat_prev_EFP = heat_EFP; mom6cu_real_to_efp(0.0, &CS->net_heat_in_EFP);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ the ocean.stats line
namespace {
// Fortran ESw.d / Fw.d / Iw edit descriptors as gfortran prints them
std::string ed_ES(double x, int w, int dgt) {
  char b[80];
  snprintf(b, sizeof b, "%.*E", dgt, x);
  std::string s(b);
  const size_t e = s.find('E');
  if (e != std::string::npos) {
    const int ex = atoi(s.c_str() + e + 1);
    const int ax = ex < 0 ? -ex : ex;
    char eb[16];
    if (ax < 100) snprintf(eb, sizeof eb, "E%c%02d", ex < 0 ? '-' : '+', ax);
    else snprintf(eb, sizeof eb, "%c%03d", ex < 0 ? '-' : '+', ax);
    s = s.substr(0, e) + eb;
  } else {  // NaN / Infinity
    s = (x != x) ? "NaN" : (x > 0 ? "Infinity" : "-Infinity");
  }
  if ((int)s.size() > w) return std::string(w, '*');
  return std::string(w - s.size(), ' ') + s;
}
std::string ed_F(double x, int w, int dgt) {
  char b[400];
  snprintf(b, sizeof b, "%.*f", dgt, x);
  std::string s(b);
  if ((int)s.size() > w) {
    if (s.compare(0, 2, "0.") == 0) s = s.substr(1);
    else if (s.compare(0, 3, "-0.") == 0) s = "-" + s.substr(2);
  }
  if ((int)s.size() > w) return std::string(w, '*');
  return std::string(w - s.size(), ' ') + s;
}
std::string ed_I(long long n, int w) {
  std::string s = std::to_string(n);
  if ((int)s.size() > w) return std::string(w, '*');
  return std::string(w - s.size(), ' ') + s;
}
std::string rtrim(std::string s) { while (!s.empty() && s.back() == ' ') s.pop_back(); return s; }
}  // namespace

extern "C" int mom6cu_ocean_stats_line(const mom6cu_sum_output_cs* CS, const mom6cu_energy_out* e, int n, double reday, char* buf, size_t len) {
  if (!CS || !e || !buf) return MOM6CU_ERR_BAD_ARG;
  // day_str / n_str :826-833
  const std::string day_str = (reday < 1.0e8) ? ed_F(reday, 12, 3) : (reday < 1.0e11) ? ed_F(reday, 15, 3) : ed_ES(reday, 15, 9);
  const std::string n_str = (n < 1000000) ? ed_I(n, 6) : (n < 10000000) ? ed_I(n, 7) : (n < 100000000) ? ed_I(n, 8) : ed_I(n, 10);
  const double vel2 = CS->L_T_to_m_s * CS->L_T_to_m_s;
  const double SL = -CS->Z_to_m * (e->Z_0APE ? e->Z_0APE[0] : 0.0);
  std::string s = rtrim(n_str) + "," + rtrim(day_str) + "," + ed_I(e->ntrunc, 6) + ", En " + ed_ES(vel2 * e->En_mass, 22, 16) + ", CFL " +
                  ed_F(e->max_CFL[0], 8, 5) + ", SL " + ed_ES(SL, 11, 4);
  if (CS->use_temperature)  // :876-889
    s += ", M " + ed_ES(CS->RZL2_to_kg * e->mass_tot, 11, 5) + ", S" + ed_F(e->salin, 8, 4) + ", T" + ed_F(CS->C_to_degC * e->temp, 8, 4) + ", Me " +
         ed_ES(e->mass_anom / e->mass_tot, 9, 2) + ", Se " + ed_ES(e->salin_anom, 9, 2) + ", Te " + ed_ES(CS->C_to_degC * e->temp_anom, 9, 2);
  else  // :891-901
    s += ", Mass " + ed_ES(CS->RZL2_to_kg * e->mass_tot, 11, 5) + ", Me " + ed_ES(e->mass_anom / e->mass_tot, 9, 2);
  if (s.size() + 1 > len) return MOM6CU_ERR_BAD_ARG;
  memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}
