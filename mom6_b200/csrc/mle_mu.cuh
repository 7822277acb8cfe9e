// mu(sigma, dh), the vertical shape of the mixed-layer-eddy streamfunction (src/parameterizations/lateral/
// MOM_mixed_layer_restrat.F90:717-751), and the surface-referenced density of the two equations of state the hot path supports
// (MOM_EOS_linear.F90:60-68, MOM_EOS_Wright.F90:80-97).  Host/device code: tests/harness/mle_host.cpp compiles it with g++ so the
// reference's unit-test values (:2022-2041) are checked on the code the GPU threads run without a GPU.
#pragma once
#include <math.h>
#if defined(__CUDACC__)
#define M6M_HD __host__ __device__ __forceinline__
#else
#define M6M_HD inline
#endif

namespace m6mle {

M6M_HD double fmx(double a, double b) { return (a > b) ? a : b; }  // MAX(a,b)
M6M_HD double fmn(double a, double b) { return (a < b) ? a : b; }  // MIN(a,b)

M6M_HD double mu(double sigma, double dh) {
  const double s1 = 2. * sigma + 1.;
  double m = fmx(0., (1. - s1 * s1) * (1. + (5. / 21.) * (s1 * s1)));
  const double xp = fmx(0., fmn(1., (-sigma - 0.5) * 2. / (1. + 2. * dh)));
  const double base = fmx(1. - (xp * xp) * (3. - 2. * xp), 0.);
  const double e = 1. + 2. * dh;
  // x**1.0 is x exactly in every libm; any other exponent is a real power (the full routine requires MLE_TAIL_DH = 0)
  const double dd = (e == 1.0) ? base : pow(base, e);
  const double bottop = 0.5 * (1. - copysign(1., sigma + 0.5));
  return fmx(m, dd * bottop);
}

struct Eos { int form; double Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp; };
constexpr int EOS_LINEAR = 1, EOS_WRIGHT = 3;  // MOM6CU_EOS_*

M6M_HD double density(const Eos& E, double T, double S, double p) {
  if (E.form == EOS_LINEAR) return E.Rho_T0_S0 + E.dRho_dT * T + E.dRho_dS * S + E.dRho_dp * p;
  // the "buggy" Wright (1997) fit of EOS_WRIGHT, MOM_EOS_Wright.F90:23-37
  const double a0 = 7.057924e-4, a1 = 3.480336e-7, a2 = -1.112733e-7;
  const double b0 = 5.790749e8, b1 = 3.516535e6, b2 = -4.002714e4, b3 = 2.084372e2, b4 = 5.944068e5, b5 = -9.643486e3;
  const double c0 = 1.704853e5, c1 = 7.904722e2, c2 = -7.984422, c3 = 5.140652e-2, c4 = -2.302158e2, c5 = -3.079464;
  const double al0 = (a0 + a1 * T) + a2 * S;
  const double p0 = (b0 + b4 * S) + T * (b1 + T * (b2 + b3 * T) + b5 * S);
  const double lambda = (c0 + c4 * S) + T * (c1 + T * (c2 + c3 * T) + c5 * S);
  return (p + p0) / (lambda + al0 * (p + p0));
}

}  // namespace m6mle
