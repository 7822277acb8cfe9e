// TEST HARNESS ONLY: a C++ caller of the host-side mirror include/mom6cu.hpp (tests/test_abi.py builds and runs it).
// Without a CUDA device it must fail loudly at context creation (exit code 3); with one it evaluates the reference's mu unit-test
// values (MOM_mixed_layer_restrat.F90:2022-2041) through the C++ wrapper and exits 0 if they match.
#include <cmath>
#include <cstdio>
#include <cstring>
#include "../../include/mom6cu.hpp"

int main() {
  // host-only entries need no device: the EFP operators (MOM_coms.F90:548-684) and the build target
  if (mom6cu::Context::build_arch() != 100) return 4;
  const mom6cu::EFP a(1.5), b(2.25);
  if ((a + b).to_real() != 3.75 || (a - b).to_real() != -0.75 || b.real_diff(a) != 0.75) return 4;
  mom6cu_domain dom;
  std::memset(&dom, 0, sizeof dom);
  const int halo = 4, ni = 8, nj = 8;
  dom.isc = halo + 1; dom.iec = halo + ni; dom.jsc = halo + 1; dom.jec = halo + nj;
  dom.isd = 1; dom.ied = ni + 2 * halo; dom.jsd = 1; dom.jed = nj + 2 * halo;
  dom.isdw = dom.isd; dom.iedw = dom.ied; dom.jsdw = dom.jsd; dom.jedw = dom.jed;
  dom.nk = 2; dom.cyclic_x = 1; dom.cyclic_y = 0; dom.first_direction = 0; dom.npi = 1; dom.npj = 1; dom.pi = 0; dom.pj = 0;
  try {
    mom6cu::Context ctx(dom, 0);
    const std::vector<double> sigma = {3., 0., -0.25, -0.5, -0.75, -1., -3., -0.5, -1., -1.5}, dh = {0., 0., 0., 0., 0., 0., 0., 0.5, 0.5, 0.5};
    const double want[10] = {0., 0., 0.7946428571428572, 1., 0.7946428571428572, 0., 0., 1., 0.25, 0.};
    const std::vector<double> got = ctx.mu(sigma, dh);
    for (int n = 0; n < 10; ++n)
      if (std::fabs(got[n] - want[n]) > 4.5e-16) { std::printf("mu(%g,%g) = %.17g, expected %.17g\n", sigma[n], dh[n], got[n], want[n]); return 1; }
    // an option outside the frozen set is a FATAL with the reference's message convention
    mom6cu_mle_cs cs;
    std::memset(&cs, 0, sizeof cs);
    cs.use_Bodner = 1;
    double x = 0.0;
    try { ctx.mixedlayer_restrat(cs, &x, &x, &x, &x, &x, &x, 900.0, &x); return 2; }
    catch (const mom6cu::Fatal& e) { std::printf("FATAL as expected: %s\n", e.what()); }
    std::printf("ok: %lld launches\n", ctx.launch_count());
    return 0;
  } catch (const mom6cu::Fatal& e) {
    std::printf("FATAL %d: %s\n", e.code, e.what());
    return 3;
  }
}
