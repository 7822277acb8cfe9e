// Extended-fixed-point arithmetic of the reference's reproducing sums (src/framework/MOM_coms.F90:30-48, :548-721), written
// once for the device threads (conversion + accumulation) and for the host side of the C ABI (carries, regularisation,
// conversion back).  Host/device code: tests/harness/efp_host.cpp compiles it with g++ so that it is checked against the
// oracle without a GPU.
//
// A real r is split into ni = 6 signed integers of 46 bits: r = sum_i v(i) * pr(i), pr = 2**92, 2**46, 1, 2**-46, 2**-92,
// 2**-138.  Every step of the split (multiply by a power of two, truncate, subtract) is exact, so the integer sums do not
// depend on the order of the additions: the device accumulates them with integer shuffles and atomics in any order and the
// result is bit-identical to the reference's serial loop.
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define M6_HD __host__ __device__ __forceinline__
#else
#define M6_HD inline
#endif

namespace m6efp {

constexpr int NI = 6;
constexpr long long PREC = 1LL << 46;          // prec
constexpr double R_PREC = 70368744177664.0;    // r_prec = 2.0**46
constexpr double I_PREC = 1.0 / R_PREC;        // I_prec
constexpr int MAX_COUNT_PREC = (1 << 17) - 1;  // max_count_prec: values that can be added before a carry is needed
// pr(1:6) and I_pr(1:6) (:41-46)
constexpr double PR0 = R_PREC * R_PREC, PR1 = R_PREC, PR2 = 1.0, PR3 = I_PREC, PR4 = I_PREC * I_PREC, PR5 = I_PREC * I_PREC * I_PREC;
constexpr double IPR0 = I_PREC * I_PREC, IPR1 = I_PREC, IPR2 = 1.0, IPR3 = R_PREC, IPR4 = R_PREC * R_PREC, IPR5 = R_PREC * R_PREC * R_PREC;
constexpr double MAX_EFP_FLOAT = PR0 * (9223372036854775808.0 - 1.);  // max_efp_float (:47)
constexpr int FLAG_NAN = 1, FLAG_OVERFLOW = 2;

M6_HD double pr(int i) { return i == 0 ? PR0 : i == 1 ? PR1 : i == 2 ? PR2 : i == 3 ? PR3 : i == 4 ? PR4 : PR5; }
M6_HD bool is_nan(double r) { return (r >= 1e30) == (r < 1e30); }  // "(r >= 1e30) .eqv. (r < 1e30)"
M6_HD long long iabs(long long v) { return v < 0 ? -v : v; }

// increment_ints_faster :629-658: s(:) += the six integers of r.  Returns FLAG_* (0 = fine); amax tracks max |r| (max_mag_term).
M6_HD int accumulate(double r, long long* s, double& amax) {
  if (is_nan(r)) return FLAG_NAN;
  const long long sgn = (r < 0.0) ? -1 : 1;
  double rs = r < 0.0 ? -r : r;
  if (rs > amax) amax = rs;
  if (rs > MAX_EFP_FLOAT) return FLAG_OVERFLOW;
  long long iv;
  iv = (long long)(rs * IPR0); rs = rs - (double)iv * PR0; s[0] += sgn * iv;
  iv = (long long)(rs * IPR1); rs = rs - (double)iv * PR1; s[1] += sgn * iv;
  iv = (long long)(rs * IPR2); rs = rs - (double)iv * PR2; s[2] += sgn * iv;
  iv = (long long)(rs * IPR3); rs = rs - (double)iv * PR3; s[3] += sgn * iv;
  iv = (long long)(rs * IPR4); rs = rs - (double)iv * PR4; s[4] += sgn * iv;
  iv = (long long)(rs * IPR5); s[5] += sgn * iv;
  return 0;
}

// Value-preserving carry in exact integer arithmetic (what carry_overflow :661-679 does; the reference finds the number of
// carries through a real multiply, which can differ by one for |v| > 2**53 -- the value represented is the same either way, and
// regularize() below maps every representation of a value to the same six integers).
M6_HD void carry_exact(long long* s) {
  for (int i = NI - 1; i >= 1; --i)
    if (iabs(s[i]) >= PREC) {
      const long long c = s[i] / PREC;
      s[i] -= c * PREC;
      s[i - 1] += c;
    }
}

// carry_overflow :661-679, literally (host side of the C ABI)
M6_HD bool carry_overflow(long long* s, long long prec_error) {
  for (int i = NI - 1; i >= 1; --i)
    if (iabs(s[i]) >= PREC) {
      const long long c = (long long)((double)s[i] * I_PREC);
      s[i] = s[i] - c * PREC;
      s[i - 1] = s[i - 1] + c;
    }
  return iabs(s[0]) > prec_error;
}

// regularize_ints :683-721
M6_HD void regularize(long long* s) {
  for (int i = NI - 1; i >= 1; --i)
    if (iabs(s[i]) >= PREC) {
      const long long c = (long long)((double)s[i] * I_PREC);
      s[i] = s[i] - c * PREC;
      s[i - 1] = s[i - 1] + c;
    }
  bool positive = true;
  for (int i = 0; i < NI; ++i)
    if (iabs(s[i]) > 0) { if (s[i] < 0) positive = false; break; }
  if (positive) {
    for (int i = NI - 1; i >= 1; --i) if (s[i] < 0) { s[i] = s[i] + PREC; s[i - 1] = s[i - 1] - 1; }
  } else {
    for (int i = NI - 1; i >= 1; --i) if (s[i] > 0) { s[i] = s[i] - PREC; s[i - 1] = s[i - 1] + 1; }
  }
}

// ints_to_real :587-597
M6_HD double to_real(const long long* s) {
  double r = 0.0;
  for (int i = 0; i < NI; ++i) r = r + pr(i) * (double)s[i];
  return r;
}

// increment_ints :600-626 (prec_error < 0: absent).  Returns true on overflow.
M6_HD bool increment(long long* s, const long long* b, long long prec_error) {
  for (int i = NI - 1; i >= 1; --i) {
    s[i] = s[i] + b[i];
    if (s[i] > PREC) { s[i] = s[i] - PREC; s[i - 1] = s[i - 1] + 1; }
    else if (s[i] < -PREC) { s[i] = s[i] + PREC; s[i - 1] = s[i - 1] - 1; }
  }
  s[0] = s[0] + b[0];
  return iabs(s[0]) > (prec_error >= 0 ? prec_error : PREC);
}

// real_to_ints :548-584 with "overflow" present.  Returns FLAG_*.
M6_HD int from_real(double r, long long prec_err, long long* s) {
  for (int i = 0; i < NI; ++i) s[i] = 0;
  if (is_nan(r)) return FLAG_NAN;
  const long long sgn = (r < 0.0) ? -1 : 1;
  double rs = r < 0.0 ? -r : r;
  const int fl = (!(rs < (double)prec_err * PR0)) ? FLAG_OVERFLOW : 0;
  for (int i = 0; i < NI; ++i) {
    const double ip = i == 0 ? IPR0 : i == 1 ? IPR1 : i == 2 ? IPR2 : i == 3 ? IPR3 : i == 4 ? IPR4 : IPR5;
    const long long iv = (long long)(rs * ip);
    rs = rs - (double)iv * pr(i);
    s[i] = sgn * iv;
  }
  return fl;
}

}  // namespace m6efp
