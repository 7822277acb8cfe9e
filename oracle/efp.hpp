// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/oracle.h).
// Extended-fixed-point arithmetic of src/framework/MOM_coms.F90 (parameters :30-48), shared by efp.cpp and sum_output.cpp.
#pragma once
#include "../include/mom6cu.h"
#include "farray.hpp"

namespace orc {

constexpr int EFP_NI = 6;                              // ni
constexpr long long EFP_PREC = 1LL << 46;              // prec
constexpr double EFP_R_PREC = 70368744177664.0;        // r_prec = 2.0**46
constexpr int EFP_MAX_COUNT_PREC = (1 << (63 - 46)) - 1;  // max_count_prec
extern const double efp_pr[EFP_NI], efp_I_pr[EFP_NI];

struct EfpFlags { bool overflow_error = false, NaN_error = false; };  // the module variables :50-51

void real_to_ints(double r, long long prec_err, EfpFlags& F, bool* overflow, long long* ints);
double ints_to_real(const long long* ints);
void increment_ints(long long* int_sum, const long long* int2, long long prec_error, EfpFlags& F);
void increment_ints_faster(long long* int_sum, double r, double& max_mag_term, EfpFlags& F);
void carry_overflow(long long* int_sum, long long prec_error, EfpFlags& F);
void regularize_ints(long long* int_sum);
void efp_plus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, EfpFlags& F);
void efp_minus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, EfpFlags& F);
double efp_to_real(mom6cu_efp* a);
double efp_real_diff(const mom6cu_efp* a, const mom6cu_efp* b);
int real_to_efp(double val, mom6cu_efp* out);

}  // namespace orc
