// Shared device-side definitions for the sm_100a dycore kernels.
//
// Device-resident layout ("unified plane"): every 2-D field, whatever its Arakawa-C
// staggering and whether the reference declares it on G's memory domain or on the
// wide-halo barotropic domain, lives in one plane of `rows` x `pitch` doubles indexed
// by the reference's own Fortran indices:
//
//        idx(i,j) = (j - j0) * pitch + (i - i0)
//
// with (i0,j0) the Fortran index of the south-west q-point of the widest memory
// domain.  h(i,j), u(I=i,j), v(i,J=j) and q(I=i,J=j) therefore share one offset, so a
// stencil kernel computes the offset once for all ~50 operands of a point, and `pitch`
// is padded to 16 doubles so that every row starts on a 128-byte line.  3-D fields are
// nk consecutive planes (k slowest), i.e. the reference's (i,j,k) order -- i is the
// coalesced thread index.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace m6 {

struct Geom {
  int i0, j0;      // Fortran index of plane column 0 / row 0
  int nx, ny;      // valid columns / rows (nx = iedw - i0 + 1 ...)
  int pitch;       // doubles per row (multiple of 16)
  int rows;        // ny (allocated rows)
  int nk;
  int isc, iec, jsc, jec;      // computational domain
  int isd, ied, jsd, jed;      // G memory domain
  int isdw, iedw, jsdw, jedw;  // wide memory domain
  long long plane;             // pitch * rows

  __host__ __device__ inline long long idx(int i, int j) const {
    return (long long)(j - j0) * pitch + (i - i0);
  }
  __host__ __device__ inline bool inside(int i, int j) const {
    return (i >= i0) && (i < i0 + nx) && (j >= j0) && (j < j0 + ny);
  }
};

// Fortran intrinsics with the semantics gfortran gives them on x86-64 (no FMA
// contraction anywhere: the library is compiled with -fmad=false).
__device__ __forceinline__ double fmax2(double a, double b) { return (a > b) ? a : b; }
__device__ __forceinline__ double fmin2(double a, double b) { return (a < b) ? a : b; }
__device__ __forceinline__ double fsign(double a, double b) { return copysign(a, b); }

}  // namespace m6
