// TEST HARNESS ONLY: compiles the host/device column code of mom6_b200/csrc/remap_column.cuh as plain C++ so that the
// column logic the GPU threads run can be compared with the oracle without a GPU (tests/test_remap_column_host.py).
// Not part of the product: nothing in mom6_b200/ loads this.
#include "../../mom6_b200/csrc/remap_column.cuh"
#include "../../mom6_b200/csrc/remap_stream.cuh"
using namespace m6remap;

template <int KCAP>
static void run(const Params& P, int ncol, int n0, int n1, const double* h0, const double* u0, const double* h1, double* u1) {
  for (int c = 0; c < ncol; ++c) {
    SubGrid<KCAP> S; Recon<KCAP> R; SubVals<KCAP> V;
    S.n0 = n0; S.n1 = n1;
    for (int k = 1; k <= n0; ++k) { S.h0[k] = h0[(long)c * n0 + k - 1]; R.u[k] = u0[(long)c * n0 + k - 1]; }
    for (int k = 1; k <= n1; ++k) S.h1[k] = h1[(long)c * n1 + k - 1];
    intersect<KCAP>(S);
    const int method = build_reconstructions<KCAP>(P, n0, S.h0, R);
    remap_via_sub_cells<KCAP>(P, S, R, method, V, ColOut{u1 + (long)c * n1, 1}, 0.0);
  }
}

extern "C" int remap_host_batch(int scheme, int extrap, int fb_sub, int fb_tgt, int om4, double h_neglect, double h_neglect_edge, int ncol,
                                int n0, const double* h0, const double* u0, int n1, const double* h1, double* u1) {
  const Params P = {scheme, extrap, fb_sub, fb_tgt, om4, h_neglect, h_neglect_edge};
  const int nmax = n0 > n1 ? n0 : n1;
  if (nmax <= 40) run<40>(P, ncol, n0, n1, h0, u0, h1, u1);
  else if (nmax <= 128) run<128>(P, ncol, n0, n1, h0, u0, h1, u1);
  else return 1;
  return 0;
}

// the streaming form (remap_stream.cuh); PPM_IH4 is not covered by it
extern "C" int remap_host_batch_stream(int scheme, int extrap, int fb_sub, int fb_tgt, int om4, double h_neglect, double h_neglect_edge, int ncol,
                                       int n0, const double* h0, const double* u0, int n1, const double* h1, double* u1) {
  if (scheme == SCHEME_PPM_IH4) return 2;
  const Params P = {scheme, extrap, fb_sub, fb_tgt, om4, h_neglect, h_neglect_edge};
  for (int c = 0; c < ncol; ++c) {
    const double *H0 = h0 + (long)c * n0 - 1, *U0 = u0 + (long)c * n0 - 1, *H1 = h1 + (long)c * n1 - 1;
    remap_stream(P, n0, n1, [=](int k) { return H0[k]; }, [=](int k) { return U0[k]; }, [=](int k) { return H1[k]; },
                 ColOut{u1 + (long)c * n1, 1}, 0.0);
  }
  return 0;
}
