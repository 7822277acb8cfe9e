// Host harness for mom6_b200/csrc/pow_glibc.cuh: compares the restated pow (both variants) with the running libm's pow() on seeded arguments.
// Built by tests/test_pow_glibc.py with g++ -O2 -ffp-contract=off -mfma.  Not part of the product.
#include "../../mom6_b200/csrc/pow_glibc.cuh"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <omp.h>

// mode 0: x in (0, 1.25], y = 1/n, n = 1..400 (bt_rem's arguments); mode 1: x = 2^U(-60,60), y = U(-8,8); mode 2: x near 1
extern "C" void pow_compare(int mode, unsigned long long seed, long long n, long long* mism_fma, long long* mism_sse, long long* notok) {
  long long m1 = 0, m2 = 0, nk = 0;
#pragma omp parallel reduction(+ : m1, m2, nk)
  {
    std::mt19937_64 g(seed + 7919ULL * omp_get_thread_num());
    std::uniform_real_distribution<double> U(0.0, 1.0);
    const long long per = n / omp_get_num_threads();
    for (long long q = 0; q < per; ++q) {
      double x, y;
      if (mode == 0) { x = 1.25 * U(g); if (U(g) < 0.2) x = std::ldexp(U(g), -(int)(60 * U(g))); y = 1.0 / (double)(1 + (int)(400 * U(g))); }
      else if (mode == 1) { x = std::exp2(120.0 * U(g) - 60.0); y = 16.0 * U(g) - 8.0; }
      else { x = 1.0 + (U(g) - 0.5) * std::ldexp(1.0, -(int)(50 * U(g))); y = 1.0 / (double)(1 + (int)(100 * U(g))); }
      if (!(x > 0.0)) continue;
      const double ref = std::pow(x, y);
      bool ok = true;
      const double a = m6pow::pow_glibc<true>(x, y, &ok);
      if (!ok) { ++nk; continue; }
      const double b = m6pow::pow_glibc<false>(x, y, &ok);
      if (std::memcmp(&a, &ref, 8) != 0) ++m1;
      if (std::memcmp(&b, &ref, 8) != 0) ++m2;
    }
  }
  *mism_fma = m1; *mism_sse = m2; *notok = nk;
}

extern "C" double pow_one(int fma, double x, double y, int* ok) {
  bool o = true;
  const double r = fma ? m6pow::pow_glibc<true>(x, y, &o) : m6pow::pow_glibc<false>(x, y, &o);
  *ok = o ? 1 : 0;
  return r;
}
