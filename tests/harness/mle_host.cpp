// TEST HARNESS ONLY: compiles the host/device code of mom6_b200/csrc/mle_mu.cuh as plain C++ (tests/test_mle.py).
// Not part of the product: nothing in mom6_b200/ loads this.
#include "../../mom6_b200/csrc/mle_mu.cuh"
extern "C" double mle_host_mu(double sigma, double dh) { return m6mle::mu(sigma, dh); }
extern "C" double mle_host_density(int form, double r0, double dT, double dS, double dp, double T, double S, double p) {
  const m6mle::Eos E = {form, r0, dT, dS, dp};
  return m6mle::density(E, T, S, p);
}
