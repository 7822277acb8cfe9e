// interpolate_column (src/ALE/MOM_remapping.F90:1247-1314): linear interpolation of interface values from a source to a
// destination column.  Host/device code: tests/harness/interp_host.cpp compiles it with g++ so that the column logic the GPU
// threads run is checked against the reference's own vectors (MOM_remapping.F90:2648-2682) without a GPU.
//
// The reference first records, for every destination interface, the source layer k_src and the fractional position frac_pos,
// then interpolates; here the two loops are fused (the position of interface k_dest depends only on the interfaces above it),
// so no k_src / frac_pos column is stored.  HS(k), US(k): source thickness (k = 1..nsrc) and interface value (1..nsrc+1);
// HD(k): destination thickness (1..ndest); OUT(k, value) receives the interface values 1..ndest+1.
#pragma once
#if defined(__CUDACC__)
#define M6I_HD __host__ __device__ __forceinline__
#else
#define M6I_HD inline
#endif

namespace m6interp {

template <class HS, class US, class HD, class OUT>
M6I_HD void interpolate_column(int nsrc, HS h_src, US u_src, int ndest, HD h_dest, OUT out, bool mask_edges) {
  int ks = 0;
  double dh = 0., x_dest = 0.;
  for (int k_dest = 1; k_dest <= ndest + 1; ++k_dest) {
    while (dh <= x_dest && ks < nsrc) {  // move forward until the interval 0 .. dh spans x_dest
      x_dest = x_dest - dh;
      ks = ks + 1;
      dh = h_src(ks);
    }
    double frac_pos;
    if (dh > 0.) {
      const double q = x_dest / dh;
      const double m = (1. < q) ? 1. : q;      // min(1., x_dest / dh)
      frac_pos = (0. > m) ? 0. : m;            // max(0., .)
    } else {
      frac_pos = 0.5;
    }
    out(k_dest, (1.0 - frac_pos) * u_src(ks) + frac_pos * u_src(ks + 1));
    if (k_dest <= ndest) x_dest = x_dest + h_dest(k_dest);
  }
  if (mask_edges) {  // :1298-1312
    for (int k_dest = 1; k_dest <= ndest; ++k_dest) { if (h_dest(k_dest) > 0.) break; out(k_dest, 0.0); }
    for (int k_dest = ndest; k_dest >= 1; --k_dest) { if (h_dest(k_dest) > 0.) break; out(k_dest + 1, 0.0); }
  }
}

}  // namespace m6interp
