// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of the order-invariant "extended fixed point" sums and of the bit-count checksums the reference uses as its
// answer-reproducibility metric:
//   /root/reference/src/framework/MOM_coms.F90: parameters :30-48, reproducing_EFP_sum_2d :101-222, reproducing_sum_2d :227-325,
//   reproducing_sum_3d :337-545, real_to_ints :548-584, ints_to_real :587-597, increment_ints :600-626, increment_ints_faster
//   :629-658, carry_overflow :661-679, regularize_ints :683-721, EFP_plus :737, EFP_minus :748, EFP_to_real :775,
//   EFP_real_diff :784, real_to_EFP :797.
//   /root/reference/src/framework/MOM_checksums.F90: chksum_h_2d :387-555, chksum_B_2d :688-875, chksum_u_2d :1005-1205,
//   chksum_v_2d :1209-1409, chksum_h_3d :1413-1583, chksum_B_3d :1586-1779, chksum_u_3d :1782-1983, chksum_v_3d :1986-2187,
//   bitcount :2678-2685.
// PARITY: PINNED for the sums by the reference's own unit test (config_src/drivers/unit_tests/test_reproducing_sum.F90: the
// exact sum of 1..N, order invariance under random swaps, fast == checked conversion, |standard - reproducing| bound), see
// tests/test_oracle_efp.py, and by running MOM_coms.F90 itself (tests/test_reference_f90.py::test_efp_sums_match_the_translated_
// reference).  The checksums: PINNED BY A REFERENCE RUN -- MOM_checksums.F90 executed by oracle/f90run, all four staggers, rank 2 and
// 3, halo shifts, symmetric / omit_corners, scaled or not, statistics (tests/refcases.py "diag/chksum").  That run found that
// chksum_B_3d is not chksum_B_2d with a k loop (see oracle_chksum below).
#include "oracle.h"
#include "efp.hpp"
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

const double efp_pr[EFP_NI] = {EFP_R_PREC * EFP_R_PREC, EFP_R_PREC, 1.0, 1.0 / EFP_R_PREC, (1.0 / EFP_R_PREC) / EFP_R_PREC,
                               ((1.0 / EFP_R_PREC) / EFP_R_PREC) / EFP_R_PREC};
const double efp_I_pr[EFP_NI] = {(1.0 / EFP_R_PREC) / EFP_R_PREC, 1.0 / EFP_R_PREC, 1.0, EFP_R_PREC, EFP_R_PREC * EFP_R_PREC,
                                 EFP_R_PREC * EFP_R_PREC * EFP_R_PREC};

static inline bool is_nan_f(double r) { return (r >= 1e30) == (r < 1e30); }  // "(r >= 1e30) .eqv. (r < 1e30)"
static inline long long iabs(long long v) { return v < 0 ? -v : v; }

// real_to_ints :548-584 with "overflow" present
void real_to_ints(double r, long long prec_err, EfpFlags& F, bool* overflow, long long* ints) {
  for (int i = 0; i < EFP_NI; ++i) ints[i] = 0;
  if (is_nan_f(r)) { F.NaN_error = true; return; }
  const int sgn = (r < 0.0) ? -1 : 1;
  double rs = std::fabs(r);
  if (!(rs < (double)prec_err * efp_pr[0])) *overflow = true;
  for (int i = 0; i < EFP_NI; ++i) {
    const long long ival = (long long)(rs * efp_I_pr[i]);
    rs = rs - (double)ival * efp_pr[i];
    ints[i] = sgn * ival;
  }
}

// ints_to_real :587-597
double ints_to_real(const long long* ints) {
  double r = 0.0;
  for (int i = 0; i < EFP_NI; ++i) r = r + efp_pr[i] * (double)ints[i];
  return r;
}

// increment_ints :600-626 (prec_error < 0: absent)
void increment_ints(long long* int_sum, const long long* int2, long long prec_error, EfpFlags& F) {
  for (int i = EFP_NI - 1; i >= 1; --i) {
    int_sum[i] = int_sum[i] + int2[i];
    if (int_sum[i] > EFP_PREC) { int_sum[i] = int_sum[i] - EFP_PREC; int_sum[i - 1] = int_sum[i - 1] + 1; }
    else if (int_sum[i] < -EFP_PREC) { int_sum[i] = int_sum[i] + EFP_PREC; int_sum[i - 1] = int_sum[i - 1] - 1; }
  }
  int_sum[0] = int_sum[0] + int2[0];
  if (prec_error >= 0) { if (iabs(int_sum[0]) > prec_error) F.overflow_error = true; }
  else if (iabs(int_sum[0]) > EFP_PREC) F.overflow_error = true;
}

// increment_ints_faster :629-658
void increment_ints_faster(long long* int_sum, double r, double& max_mag_term, EfpFlags& F) {
  if (is_nan_f(r)) { F.NaN_error = true; return; }
  const int sgn = (r < 0.0) ? -1 : 1;
  double rs = std::fabs(r);
  if (rs > std::fabs(max_mag_term)) max_mag_term = r;
  const double max_efp_float = efp_pr[0] * (9223372036854775808.0 - 1.);  // pr(1) * (2.**63 - 1.)
  if (rs > max_efp_float) { F.overflow_error = true; return; }
  for (int i = 0; i < EFP_NI; ++i) {
    const long long ival = (long long)(rs * efp_I_pr[i]);
    rs = rs - (double)ival * efp_pr[i];
    int_sum[i] = int_sum[i] + sgn * ival;
  }
}

// carry_overflow :661-679
void carry_overflow(long long* int_sum, long long prec_error, EfpFlags& F) {
  const double I_prec = 1.0 / EFP_R_PREC;
  for (int i = EFP_NI - 1; i >= 1; --i)
    if (iabs(int_sum[i]) >= EFP_PREC) {
      const long long num_carry = (long long)((double)int_sum[i] * I_prec);
      int_sum[i] = int_sum[i] - num_carry * EFP_PREC;
      int_sum[i - 1] = int_sum[i - 1] + num_carry;
    }
  if (iabs(int_sum[0]) > prec_error) F.overflow_error = true;
}

// regularize_ints :683-721
void regularize_ints(long long* int_sum) {
  const double I_prec = 1.0 / EFP_R_PREC;
  for (int i = EFP_NI - 1; i >= 1; --i)
    if (iabs(int_sum[i]) >= EFP_PREC) {
      const long long num_carry = (long long)((double)int_sum[i] * I_prec);
      int_sum[i] = int_sum[i] - num_carry * EFP_PREC;
      int_sum[i - 1] = int_sum[i - 1] + num_carry;
    }
  bool positive = true;
  for (int i = 0; i < EFP_NI; ++i)
    if (iabs(int_sum[i]) > 0) { if (int_sum[i] < 0) positive = false; break; }
  if (positive) {
    for (int i = EFP_NI - 1; i >= 1; --i)
      if (int_sum[i] < 0) { int_sum[i] = int_sum[i] + EFP_PREC; int_sum[i - 1] = int_sum[i - 1] - 1; }
  } else {
    for (int i = EFP_NI - 1; i >= 1; --i)
      if (int_sum[i] > 0) { int_sum[i] = int_sum[i] - EFP_PREC; int_sum[i - 1] = int_sum[i - 1] + 1; }
  }
}

void efp_plus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, EfpFlags& F) {  // :737-745
  long long v[EFP_NI];
  for (int i = 0; i < EFP_NI; ++i) v[i] = a->v[i];
  increment_ints(v, (const long long*)b->v, -1, F);
  for (int i = 0; i < EFP_NI; ++i) out->v[i] = v[i];
}
void efp_minus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, EfpFlags& F) {  // :748-757
  long long v[EFP_NI];
  for (int i = 0; i < EFP_NI; ++i) v[i] = -1 * b->v[i];
  increment_ints(v, (const long long*)a->v, -1, F);
  for (int i = 0; i < EFP_NI; ++i) out->v[i] = v[i];
}
double efp_to_real(mom6cu_efp* a) {  // :775-781
  regularize_ints((long long*)a->v);
  return ints_to_real((const long long*)a->v);
}
double efp_real_diff(const mom6cu_efp* a, const mom6cu_efp* b) {  // :784-794
  EfpFlags F;
  mom6cu_efp d;
  efp_minus(a, b, &d, F);
  return efp_to_real(&d);
}
int real_to_efp(double val, mom6cu_efp* out) {  // :797-815 ; 1 = the FATAL "Overflow in real_to_EFP conversion"
  EfpFlags F;
  bool over = false;
  real_to_ints(val, EFP_PREC, F, &over, (long long*)out->v);
  return over ? 1 : 0;
}

// The accumulation of one layer's window, shared by reproducing_EFP_sum_2d :163-199 and reproducing_sum_3d :421-441, :487-507
// (do_unscale selects "unscale*array" vs "array" in the common branch; the other two branches always use descale*array).
static void sum_window(const double* a, int ni, int is, int ie, int js, int je, bool do_unscale, double unscale, double descale,
                       bool over_check, long long prec_error, long long* ints_sum, double& max_mag_term, EfpFlags& F) {
  auto A = [&](int i, int j) { return a[(size_t)(j - 1) * ni + (i - 1)]; };
  if (over_check) {
    if ((long long)(je + 1 - js) * (ie + 1 - is) < EFP_MAX_COUNT_PREC) {
      if (do_unscale) { for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) increment_ints_faster(ints_sum, unscale * A(i, j), max_mag_term, F); }
      else { for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) increment_ints_faster(ints_sum, A(i, j), max_mag_term, F); }
      carry_overflow(ints_sum, prec_error, F);
    } else if ((ie + 1 - is) < EFP_MAX_COUNT_PREC) {
      for (int j = js; j <= je; ++j) {
        for (int i = is; i <= ie; ++i) increment_ints_faster(ints_sum, descale * A(i, j), max_mag_term, F);
        carry_overflow(ints_sum, prec_error, F);
      }
    } else {
      for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
        long long t[EFP_NI]; bool over = false;
        // real_to_ints without "overflow": its FATAL is reported through overflow_error here
        real_to_ints(descale * A(i, j), prec_error, F, &over, t);
        if (over) F.overflow_error = true;
        increment_ints(ints_sum, t, prec_error, F);
      }
    }
  } else {  // :186-197
    for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      const int sgn = (A(i, j) < 0.0) ? -1 : 1;
      double rs = std::fabs(descale * A(i, j));
      for (int n = 0; n < EFP_NI; ++n) {
        const long long ival = (long long)(rs * efp_I_pr[n]);
        rs = rs - (double)ival * efp_pr[n];
        ints_sum[n] = ints_sum[n] + sgn * ival;
      }
    }
    carry_overflow(ints_sum, prec_error, F);
  }
}

}  // namespace orc

using namespace orc;

static void stag_extent(const mom6cu_domain* d, int stagger, int* ni, int* nj) {
  *ni = d->ied - d->isd + 1 + ((stagger == 1 || stagger == 3) ? 1 : 0);
  *nj = d->jed - d->jsd + 1 + ((stagger == 2 || stagger == 3) ? 1 : 0);
}

// reproducing_sum (2-D when nk == 1 and neither sums nor EFP_lay_sums is given: reproducing_sum_2d :227-325 with
// reproducing_EFP_sum_2d :101-222; otherwise reproducing_sum_3d :337-545).  Single PE.  Returns 0, or the FATALs:
// 21 index range, 22 NaN, 23 conversion overflow, 24 overflow.
extern "C" int oracle_reproducing_sum(const mom6cu_domain* dom, const double* array, int stagger, int nk, int isr, int ier, int jsr,
                                      int jer, double unscale, int reproducing, int overflow_check, double* sum, double* sums,
                                      mom6cu_efp* EFP_sum, mom6cu_efp* EFP_lay_sums) {
  int ni, nj;
  stag_extent(dom, stagger, &ni, &nj);
  const long long prec_error = 0x7fffffffffffffffLL;  // ((2**62 + (2**62 - 1)) / num_PEs()
  int is = 1, ie = ni, js = 1, je = nj;
  if (isr > 0) { if (isr < is) return 21; is = isr; }
  if (ier > 0) { if (ier > ie) return 21; ie = ier; }
  if (jsr > 0) { if (jsr < js) return 21; js = jsr; }
  if (jer > 0) { if (jer > je) return 21; je = jer; }
  const bool do_unscale = (unscale != 1.0);
  const double descale = do_unscale ? unscale : 1.0;
  EfpFlags F;
  double max_mag_term = 0.0;
  const size_t pl = (size_t)ni * nj;
  if (nk == 1 && !sums && !EFP_lay_sums) {
    // ---- reproducing_sum_2d
    double I_unscale = 1.0;
    if (do_unscale && std::fabs(unscale) > 0.0) I_unscale = 1.0 / unscale;
    if (reproducing) {
      long long ints_sum[EFP_NI] = {0, 0, 0, 0, 0, 0};
      sum_window(array, ni, is, ie, js, je, do_unscale, unscale, descale, overflow_check != 0, prec_error, ints_sum, max_mag_term, F);
      if (F.NaN_error) return 22;
      if (std::fabs(max_mag_term) >= (double)prec_error * efp_pr[0]) return 23;
      if (F.overflow_error) return 24;
      regularize_ints(ints_sum);
      *sum = ints_to_real(ints_sum) * I_unscale;
      if (EFP_sum) for (int n = 0; n < EFP_NI; ++n) EFP_sum->v[n] = ints_sum[n];
    } else {
      double rsum = 0.0;
      for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) rsum = rsum + descale * array[(size_t)(j - 1) * ni + (i - 1)];
      *sum = rsum * I_unscale;
      if (EFP_sum) {
        bool over = false;
        real_to_ints(*sum, prec_error, F, &over, (long long*)EFP_sum->v);
        if (over) return 24;
      }
    }
    return 0;
  }
  // ---- reproducing_sum_3d
  double total;
  if (sums || EFP_lay_sums) {
    std::vector<long long> ints_sums((size_t)EFP_NI * nk, 0);
    for (int k = 0; k < nk; ++k)
      sum_window(array + pl * k, ni, is, ie, js, je, do_unscale, unscale, descale, true, prec_error, &ints_sums[(size_t)EFP_NI * k],
                 max_mag_term, F);
    if (F.NaN_error) return 22;
    if (std::fabs(max_mag_term) >= (double)prec_error * efp_pr[0]) return 23;
    if (F.overflow_error) return 24;
    total = 0.0;
    for (int k = 0; k < nk; ++k) {
      regularize_ints(&ints_sums[(size_t)EFP_NI * k]);
      const double val = ints_to_real(&ints_sums[(size_t)EFP_NI * k]);
      if (sums) sums[k] = val;
      total = total + val;
    }
    if (EFP_lay_sums) for (int k = 0; k < nk; ++k) for (int n = 0; n < EFP_NI; ++n) EFP_lay_sums[k].v[n] = ints_sums[(size_t)EFP_NI * k + n];
    if (EFP_sum) {
      long long t[EFP_NI] = {0, 0, 0, 0, 0, 0};
      for (int k = 0; k < nk; ++k) increment_ints(t, &ints_sums[(size_t)EFP_NI * k], -1, F);
      for (int n = 0; n < EFP_NI; ++n) EFP_sum->v[n] = t[n];
    }
  } else {
    long long ints_sum[EFP_NI] = {0, 0, 0, 0, 0, 0};
    // one accumulator for all layers: the three size branches of :487-507 carry per layer / per row / per element
    for (int k = 0; k < nk; ++k)
      sum_window(array + pl * k, ni, is, ie, js, je, do_unscale, unscale, descale, true, prec_error, ints_sum, max_mag_term, F);
    if (F.NaN_error) return 22;
    if (std::fabs(max_mag_term) >= (double)prec_error * efp_pr[0]) return 23;
    if (F.overflow_error) return 24;
    regularize_ints(ints_sum);
    total = ints_to_real(ints_sum);
    if (EFP_sum) for (int n = 0; n < EFP_NI; ++n) EFP_sum->v[n] = ints_sum[n];
  }
  if (do_unscale) {  // :535-543
    double I_unscale = 0.0;
    if (std::fabs(unscale) > 0.0) I_unscale = 1.0 / unscale;
    total = total * I_unscale;
    if (sums) for (int k = 0; k < nk; ++k) sums[k] = sums[k] * I_unscale;
  }
  *sum = total;
  return 0;
}

extern "C" void oracle_efp_plus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, int* overflow) {
  EfpFlags F; efp_plus(a, b, out, F); if (overflow) *overflow = F.overflow_error ? 1 : 0;
}
extern "C" void oracle_efp_minus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, int* overflow) {
  EfpFlags F; efp_minus(a, b, out, F); if (overflow) *overflow = F.overflow_error ? 1 : 0;
}
extern "C" double oracle_efp_to_real(mom6cu_efp* a) { return efp_to_real(a); }
extern "C" int oracle_real_to_efp(double v, mom6cu_efp* out) { return real_to_efp(v, out); }
extern "C" double oracle_efp_real_diff(const mom6cu_efp* a, const mom6cu_efp* b) { return efp_real_diff(a, b); }

// ---------------------------------------------------------------------------------------------- checksums
// bitcount :2678-2685
static inline int bitcount(double x) {
  unsigned long long u;
  std::memcpy(&u, &x, 8);
  return __builtin_popcountll(u);
}

// chksum_{h,u,v,B}_{2d,3d}.  The four staggers share one control flow; what differs is cited at each branch.
// Returns 0, or 31 for the FATAL of the halo-width test.
extern "C" int oracle_chksum(const mom6cu_domain* d, const double* array, int stagger, int nk, int haloshift, int symmetric,
                             int omit_corners, double scale, int* bc, int* kind, double* stats) {
  const bool su = (stagger == 1 || stagger == 3), sv = (stagger == 2 || stagger == 3);  // staggered in i / in j
  const int ilo = d->isd - (su ? 1 : 0), jlo = d->jsd - (sv ? 1 : 0);
  const int ni = d->ied - ilo + 1, nj = d->jed - jlo + 1;
  const size_t pl = (size_t)ni * nj;
  auto A = [&](int i, int j, int k) { return array[pl * k + (size_t)(j - jlo) * ni + (i - ilo)]; };
  const double scaling = scale;
  const bool sym = symmetric != 0;
  bool sym_stats = sym;
  if (haloshift > 0) sym_stats = true;  // (u, v, B forms; the h form has no sym_stats)
  if (stats) {
    // subStats: minima / maxima over the (symmetric) computational domain of the stagger, mean over the h-point one
    const int IsB = d->isc - ((su && sym_stats) ? 1 : 0), JsB = d->jsc - ((sv && sym_stats) ? 1 : 0);
    double aMin = scaling * A(d->isc, d->jsc, 0), aMax = aMin;
    if (scale == 1.0) { aMin = A(d->isc, d->jsc, 0); aMax = aMin; }
    for (int k = 0; k < nk; ++k) for (int j = JsB; j <= d->jec; ++j) for (int i = IsB; i <= d->iec; ++i) {
      const double v = (scale == 1.0) ? A(i, j, k) : scaling * A(i, j, k);
      aMin = fmin2(aMin, v); aMax = fmax2(aMax, v);
    }
    // aMean = reproducing_sum(array(isc:iec,jsc:jec[,:])) of the rescaled array
    long long ints_sum[EFP_NI] = {0, 0, 0, 0, 0, 0};
    EfpFlags F; double mm = 0.0;
    const int wi = d->iec - d->isc + 1, wj = d->jec - d->jsc + 1;
    std::vector<double> w((size_t)wi * wj);
    for (int k = 0; k < nk; ++k) {
      for (int j = d->jsc; j <= d->jec; ++j) for (int i = d->isc; i <= d->iec; ++i)
        w[(size_t)(j - d->jsc) * wi + (i - d->isc)] = (scale == 1.0) ? A(i, j, k) : scaling * A(i, j, k);
      sum_window(w.data(), wi, 1, wi, 1, wj, false, 1.0, 1.0, true, 0x7fffffffffffffffLL, ints_sum, mm, F);
    }
    regularize_ints(ints_sum);
    const long long n = (long long)wi * wj * nk;
    stats[0] = ints_to_real(ints_sum) / (double)n;
    stats[1] = aMin; stats[2] = aMax;
  }
  int hshift = haloshift;
  if (hshift < 0) hshift = d->ied - d->iec;
  const int isc = d->isc - ((stagger == 3) ? 1 : 0), jsc = d->jsc - ((stagger == 3) ? 1 : 0);  // B: iscB..jecB vs isdB..jedB (:775)
  const int isd = d->isd - ((stagger == 3) ? 1 : 0), jsd = d->jsd - ((stagger == 3) ? 1 : 0);
  // (the u form tests the h-point bounds :1901, the v form the B ones in j :1309; all are equivalent for symmetric memory)
  if (isc - hshift < isd || d->iec + hshift > d->ied || jsc - hshift < jsd || d->jec + hshift > d->jed) return 31;
  auto subchk = [&](int di, int dj) {
    int s = 0;  // default INTEGER: wraps like the reference's 32-bit accumulator
    for (int k = 0; k < nk; ++k) for (int j = d->jsc + dj; j <= d->jec + dj; ++j) for (int i = d->isc + di; i <= d->iec + di; ++i)
      s = (int)((unsigned)s + (unsigned)bitcount(std::fabs(scaling * A(i, j, k))));
    return s % 1000000000;  // mod(subchk, bc_modulus)
  };
  bc[0] = subchk(0, 0);
  for (int q = 1; q < 5; ++q) bc[q] = 0;
  const bool plain = (stagger == 0) ? (hshift == 0) : ((hshift == 0) && !sym);
  if (plain) { *kind = 1; return 0; }
  const bool do_corners = !omit_corners;
  // chksum_B_3d is not chksum_B_2d with a k loop: its corner windows take the extra row and column whether or not `symmetric` is
  // set (:1698-1706, both arms of the IF are the same), and with omit_corners its S and W windows widen under `symmetric`
  // (:1712-1718) where the 2-d form's never do (:806-809).  A rank-3 array is what nk > 1 stands for here.
  const bool q3 = (stagger == 3 && nk > 1);
  const int ex = ((sym || q3) && su) ? 1 : 0, ey = ((sym || q3) && sv) ? 1 : 0;  // the extra row / column of the symmetric forms
  if (hshift == 0 && stagger == 1) { bc[1] = subchk(-hshift - 1, 0); *kind = 4; return 0; }  // chksum_u :1916-1918
  if (hshift == 0 && stagger == 2) { bc[1] = subchk(0, -hshift - 1); *kind = 5; return 0; }  // chksum_v :1325-1327
  if (do_corners) {
    bc[1] = subchk(-hshift - ex, -hshift - ey);  // SW
    bc[2] = subchk(hshift, -hshift - ey);        // SE
    bc[3] = subchk(-hshift - ex, hshift);        // NW
    bc[4] = subchk(hshift, hshift);              // NE
    *kind = 2;
  } else {
    const int bS = subchk(0, -hshift - (((stagger == 2 || q3) && sym) ? 1 : 0));  // the v forms widen S (:1340-1344), the u forms W,
    const int bE = subchk(hshift, 0);                                              // the 3-d B form both
    const int bW = subchk(-hshift - (((stagger == 1 || q3) && sym) ? 1 : 0), 0);
    const int bN = subchk(0, hshift);
    bc[1] = bN; bc[2] = bS; bc[3] = bE; bc[4] = bW;
    *kind = 3;
  }
  return 0;
}
