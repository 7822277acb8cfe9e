// TEST HARNESS ONLY: compiles the host/device column code of mom6_b200/csrc/interp_column.cuh as plain C++ (tests/test_ale_chain.py).
// Not part of the product: nothing in mom6_b200/ loads this.
#include "../../mom6_b200/csrc/interp_column.cuh"
extern "C" void interp_host(int nsrc, const double* h_src, const double* u_src, int ndest, const double* h_dest, double* u_dest, int mask_edges) {
  m6interp::interpolate_column(nsrc, [=](int k) { return h_src[k - 1]; }, [=](int k) { return u_src[k - 1]; }, ndest,
                               [=](int k) { return h_dest[k - 1]; }, [=](int k, double v) { u_dest[k - 1] = v; }, mask_edges != 0);
}
