// x**y as the host's libm evaluates it: the double-precision pow of the ARM Optimized Routines (S. Nagy 2018, MIT licence), which every
// glibc >= 2.28 ships as pow(), restated operation for operation as a host/device function with the library's own tables
// (pow_glibc_tables.h, generated from libm.so.6 by tools/gen_pow_tables.py).
//
// Why: bt_rem_u = (SUM frhatu*visc_rem_u)**Instep (MOM_barotropic.F90:1502,1508) is the only operation of the hot path that IEEE 754 does
// not define, so "the reference's answer" is whatever pow() the reference platform links: on Linux, this routine.  Its result is not always
// the correctly rounded one (error bound 0.52 ULP), so a correctly rounded device pow would NOT reproduce it; the same arithmetic does.
// On x86-64 glibc selects at load time the build of this routine compiled with FMA (every CPU since Haswell / Zen) -- the variant FMA=true
// below, where the library calls fma() explicitly (log_inline's r, lo3 and pow's elo) and the compiler contracted nothing else (glibc's libm
// is built with -ffp-contract=off).  FMA=false is the SSE2 build of the same source.  tests/test_pow_glibc.py compiles this header for the
// host and compares both variants with the running libm's pow on >= 1e8 seeded arguments (bt_rem's range and a wide one).
//
// Domain handled: x finite > 0 (subnormals included), y finite with 2^-65 <= |y| < 2^63, |y log x| < 512 -- far more than av_rem**Instep
// needs (0 < x <= ~1, y = 1/nstep).  Outside it *ok is cleared and the caller must not use the value (the kernel reports it).
#pragma once
#include <cstdint>
#include "pow_glibc_tables.h"

#if defined(__CUDACC__)
#define M6POW_HD __host__ __device__ __forceinline__
#define M6POW_CONST __device__ __constant__
#else
#define M6POW_HD inline
#define M6POW_CONST static
#endif

namespace m6pow {

#if defined(__CUDACC__)
__device__ const double d_logtab[384] = M6POW_LOGTAB;
__device__ const uint64_t d_exptab[256] = M6POW_EXPTAB;
#endif
static const double h_logtab[384] = M6POW_LOGTAB;
static const uint64_t h_exptab[256] = M6POW_EXPTAB;

M6POW_HD uint64_t asu(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; __builtin_memcpy(&u, &x, 8); return u;
#endif
}
M6POW_HD double asd(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; __builtin_memcpy(&x, &u, 8); return x;
#endif
}
M6POW_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
// every product and sum below must round on its own (the device library is compiled -fmad=false; the host harness -ffp-contract=off)
M6POW_HD double mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
M6POW_HD double add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}

template <bool FMA>
M6POW_HD double pow_glibc(double x, double y, bool* ok) {
#if defined(__CUDA_ARCH__)
  const double* T = d_logtab; const uint64_t* E = d_exptab;
#else
  const double* T = h_logtab; const uint64_t* E = h_exptab;
#endif
  const double A[7] = M6POW_A;
  const double C[4] = M6POW_C;
  uint64_t ix = asu(x);
  const uint64_t iy = asu(y);
  const uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52);
  // pow(): the argument classes outside the common path (e_pow.c: SmallPowX, ThresPowX, SmallPowY, ThresPowY)
  if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
    if ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu || (ix >> 63) || ix == 0 || topx >= 0x7ff) { *ok = false; return 0.0; }
    // subnormal x: normalise so that the exponent becomes negative
    ix = asu(mul(x, 0x1p52));
    ix &= 0x7fffffffffffffffULL;
    ix -= 52ULL << 52;
  }
  // ---- log_inline: log(x) = k ln2 + log(c) + log1p(z/c - 1) as hi + lo
  const uint64_t OFF = 0x3fe6955500000000ULL;
  const uint64_t tmp = ix - OFF;
  const int i = (int)((tmp >> (52 - 7)) % 128);
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & (0xfffULL << 52));
  const double z = asd(iz);
  const double kd = (double)k;
  const double invc = T[3 * i], logc = T[3 * i + 1], logctail = T[3 * i + 2];
  double r, rhi = 0.0, rlo = 0.0;
  if (FMA) r = fma_(z, invc, -1.0);
  else {
    const double zhi = asd((iz + (1ULL << 31)) & (~0ULL << 32));
    const double zlo = add(z, -zhi);
    rhi = add(mul(zhi, invc), -1.0);
    rlo = mul(zlo, invc);
    r = add(rhi, rlo);
  }
  // FMA build: the compiler (GCC, -ffp-contract=fast, -mfma) fused every product that feeds only a sum; the fusions below are that
  // build's, established by comparing with the library bit for bit (tests/test_pow_glibc.py)
  const double t1 = FMA ? fma_(kd, M6POW_LN2HI, logc) : add(mul(kd, M6POW_LN2HI), logc);
  const double t2 = add(t1, r);
  const double lo1 = FMA ? fma_(kd, M6POW_LN2LO, logctail) : add(mul(kd, M6POW_LN2LO), logctail);
  const double lo2 = add(add(t1, -t2), r);
  const double ar = mul(A[0], r);
  const double ar2 = mul(r, ar);
  const double ar3 = mul(r, ar2);
  double hi, lo3, lo4;
  if (FMA) {
    hi = add(t2, ar2);
    lo3 = fma_(ar, r, -ar2);
    lo4 = add(add(t2, -hi), ar2);
  } else {
    const double arhi = mul(A[0], rhi);
    const double arhi2 = mul(rhi, arhi);
    hi = add(t2, arhi2);
    lo3 = mul(rlo, add(ar, arhi));
    lo4 = add(add(t2, -hi), arhi2);
  }
  // p = ar3 * (A[1] + r*A[2] + ar2*(A[3] + r*A[4] + ar2*(A[5] + r*A[6])))
  const double p = FMA ? mul(ar3, fma_(ar2, fma_(ar2, fma_(r, A[6], A[5]), fma_(r, A[4], A[3])), fma_(r, A[2], A[1])))
                       : mul(ar3, add(add(A[1], mul(r, A[2])), mul(ar2, add(add(A[3], mul(r, A[4])), mul(ar2, add(A[5], mul(r, A[6])))))));
  const double lo = add(add(add(add(lo1, lo2), lo3), lo4), p);
  const double lhi_ = add(hi, lo);
  const double ltail = add(add(hi, -lhi_), lo);
  // ---- ehi + elo = y * log(x)
  double ehi, elo;
  if (FMA) {
    ehi = mul(y, lhi_);
    elo = fma_(y, ltail, fma_(y, lhi_, -ehi));
  } else {
    const double yhi = asd(iy & (~0ULL << 27));
    const double ylo = add(y, -yhi);
    const double lhi = asd(asu(lhi_) & (~0ULL << 27));
    const double llo = add(add(lhi_, -lhi), ltail);
    ehi = mul(yhi, lhi);
    elo = add(mul(ylo, lhi), mul(y, llo));
  }
  // ---- exp_inline(ehi, elo, 0)
  const uint32_t abstop = (uint32_t)(asu(ehi) >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x408u - 0x3c9u) {   // |ehi| < 2^-54 or >= 2^9
    if (abstop - 0x3c9u >= 0x80000000u) return add(1.0, ehi);   // tiny: 1 + x (WANT_ROUNDING)
    *ok = false;   // |y log x| >= 512: the library's overflow / underflow / subnormal-result path, not needed here
    return 0.0;
  }
  double kd2 = FMA ? fma_(M6POW_INVLN2N, ehi, M6POW_SHIFT) : add(mul(M6POW_INVLN2N, ehi), M6POW_SHIFT);
  const uint64_t ki = asu(kd2);
  kd2 = add(kd2, -(M6POW_SHIFT));
  double rr = FMA ? fma_(kd2, M6POW_NEGLN2LON, fma_(kd2, M6POW_NEGLN2HIN, ehi))
                  : add(add(ehi, mul(kd2, M6POW_NEGLN2HIN)), mul(kd2, M6POW_NEGLN2LON));
  rr = add(rr, elo);
  const uint64_t idx = 2 * (ki % 128);
  const uint64_t top = ki << (52 - 7);
  const double tail = asd(E[idx]);
  const uint64_t sbits = E[idx + 1] + top;
  const double r2 = mul(rr, rr);
  // tmp = tail + r + r2*(C2 + r*C3) + r2*r2*(C4 + r*C5)
  const double tm = FMA ? fma_(mul(r2, r2), fma_(rr, C[3], C[2]), fma_(r2, fma_(rr, C[1], C[0]), add(tail, rr)))
                        : add(add(add(tail, rr), mul(r2, add(C[0], mul(rr, C[1])))), mul(mul(r2, r2), add(C[2], mul(rr, C[3]))));
  const double scale = asd(sbits);
  return FMA ? fma_(scale, tm, scale) : add(scale, mul(scale, tm));
}

}  // namespace m6pow
