// Device planes of the barotropic substep loop (shared by bt_timeloop.cu and btstep.cu).
#pragma once
#include "ctx.h"

struct BtPlanes {
  // ping-pong state
  const double* eta_in; const double* ubt_in; const double* vbt_in;
  double* eta_out; double* ubt_out; double* vbt_out;
  // coefficients
  const double* uhbt0; const double* vhbt0; const double* Datu; const double* Datv;
  const double* bu[10]; const double* bv[10];
  const double* eta_src; const double* eta_PF;
  const double* gtot_E; const double* gtot_W; const double* gtot_N; const double* gtot_S;
  const double* f4u[4]; const double* f4v[4];
  const double* bt_rem_u; const double* bt_rem_v; const double* BT_force_u; const double* BT_force_v;
  const double* Cor_ref_u; const double* Cor_ref_v;
  const double* IareaT; const double* IdxCu; const double* IdyCv;
  // accumulators
  double* u_accel_bt; double* v_accel_bt; double* eta_sum; double* eta_wtd;
  double* ubtav; double* vbtav; double* uhbtav; double* vhbtav; double* ubt_wtd; double* vbt_wtd;
};

struct BtDevice {  // device planes of one timeloop call
  BtPlanes P;
  double* eta[2]; double* ubt[2]; double* vbt[2];
};


// named resident planes "bt.*" of one timeloop call
int m6_bt_alloc(mom6cu_ctx* c, BtDevice& D);
// the substep loop proper on resident planes; `a` supplies only scalars and the host weight arrays
int m6_bt_run(mom6cu_ctx* c, BtDevice& D, const mom6cu_bt_timeloop_args* a, int* final_slot);
