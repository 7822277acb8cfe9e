// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Fortran-style array views (column-major, arbitrary lower bounds) so that the
// oracle can restate the reference loops index-for-index.
#pragma once
#include <cstddef>
#include <vector>
#include <cmath>

namespace orc {

// 2-D view A(i,j), i fastest, i in [ilo,ihi], j in [jlo,jhi]
struct V2 {
  double* p; int ilo, jlo, ni, nj;
  V2() : p(nullptr), ilo(0), jlo(0), ni(0), nj(0) {}
  V2(double* p_, int ilo_, int ihi_, int jlo_, int jhi_)
      : p(p_), ilo(ilo_), jlo(jlo_), ni(ihi_ - ilo_ + 1), nj(jhi_ - jlo_ + 1) {}
  inline double& operator()(int i, int j) const {
    return p[(size_t)(j - jlo) * ni + (i - ilo)];
  }
  size_t size() const { return (size_t)ni * nj; }
  void fill(double v) const { for (size_t n = 0; n < size(); ++n) p[n] = v; }
};

// 3-D view A(i,j,k), k in [1,nk]
struct V3 {
  double* p; int ilo, jlo, ni, nj, nk;
  V3() : p(nullptr), ilo(0), jlo(0), ni(0), nj(0), nk(0) {}
  V3(double* p_, int ilo_, int ihi_, int jlo_, int jhi_, int nk_)
      : p(p_), ilo(ilo_), jlo(jlo_), ni(ihi_ - ilo_ + 1), nj(jhi_ - jlo_ + 1), nk(nk_) {}
  inline double& operator()(int i, int j, int k) const {
    return p[((size_t)(k - 1) * nj + (j - jlo)) * ni + (i - ilo)];
  }
  size_t size() const { return (size_t)ni * nj * nk; }
  void fill(double v) const { for (size_t n = 0; n < size(); ++n) p[n] = v; }
};

// (m,i,j) view with m in [1,nm] fastest -- e.g. f_4_u(4,I,j) or the array of
// local_BT_cont derived types (10 reals per point).
struct VM2 {
  double* p; int nm, ilo, jlo, ni, nj;
  VM2() : p(nullptr), nm(0), ilo(0), jlo(0), ni(0), nj(0) {}
  VM2(double* p_, int nm_, int ilo_, int ihi_, int jlo_, int jhi_)
      : p(p_), nm(nm_), ilo(ilo_), jlo(jlo_), ni(ihi_ - ilo_ + 1), nj(jhi_ - jlo_ + 1) {}
  inline double& operator()(int m, int i, int j) const {
    return p[((size_t)(j - jlo) * ni + (i - ilo)) * nm + (m - 1)];
  }
  inline double* at(int i, int j) const {
    return p + ((size_t)(j - jlo) * ni + (i - ilo)) * nm;
  }
};

// Owning scratch arrays
struct A2 : V2 {
  std::vector<double> store;
  A2(int ilo_, int ihi_, int jlo_, int jhi_, double init = 0.0)
      : V2(nullptr, ilo_, ihi_, jlo_, jhi_), store((size_t)(ihi_ - ilo_ + 1) * (jhi_ - jlo_ + 1), init) {
    p = store.data();
  }
};
struct A3 : V3 {
  std::vector<double> store;
  A3(int ilo_, int ihi_, int jlo_, int jhi_, int nk_, double init = 0.0)
      : V3(nullptr, ilo_, ihi_, jlo_, jhi_, nk_),
        store((size_t)(ihi_ - ilo_ + 1) * (jhi_ - jlo_ + 1) * nk_, init) {
    p = store.data();
  }
};

// Fortran intrinsics with the semantics gfortran gives them on x86-64.
static inline double fmax2(double a, double b) { return (a > b) ? a : b; }   // MAX(a,b)
static inline double fmin2(double a, double b) { return (a < b) ? a : b; }   // MIN(a,b)
static inline double fsign(double a, double b) { return std::copysign(a, b); } // SIGN(a,b)

}  // namespace orc
