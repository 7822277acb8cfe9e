// TEST HARNESS ONLY: compiles the host/device extended-fixed-point code of mom6_b200/csrc/efp.cuh as plain C++ and replays the
// reduction shape of efp_sum_kernel (per-thread partial sums -> per-CTA sum + exact carry -> accumulator -> host carry,
// regularisation, conversion) so that it can be compared with the oracle without a GPU (tests/test_diag.py).
// Not part of the product: nothing in mom6_b200/ loads this.
#include "../../mom6_b200/csrc/efp.cuh"
#include <vector>
using namespace m6efp;

// values[n]; "threads" threads per CTA, "per_cta" values per CTA, visited in the strided order the kernel uses.
extern "C" int efp_host_sum(long n, const double* values, int threads, int per_cta, long long* ints_out, double* sum_out, double* amax_out) {
  long long acc[NI] = {0, 0, 0, 0, 0, 0};
  double amax_all = 0.0;
  int flags = 0;
  for (long c0 = 0; c0 < n; c0 += per_cta) {
    const long c1 = (c0 + per_cta < n) ? c0 + per_cta : n;
    long long cta[NI] = {0, 0, 0, 0, 0, 0};
    for (int t = threads - 1; t >= 0; --t) {  // any order of the threads gives the same integers
      long long s[NI] = {0, 0, 0, 0, 0, 0};
      double amax = 0.0;
      for (long e = c0 + t; e < c1; e += threads) flags |= accumulate(values[e], s, amax);
      for (int q = 0; q < NI; ++q) cta[q] += s[q];
      if (amax > amax_all) amax_all = amax;
    }
    carry_exact(cta);
    for (int q = 0; q < NI; ++q) acc[q] += cta[q];
  }
  carry_exact(acc);
  regularize(acc);
  for (int q = 0; q < NI; ++q) ints_out[q] = acc[q];
  *sum_out = to_real(acc);
  *amax_out = amax_all;
  return flags;
}
