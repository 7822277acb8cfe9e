// Device-resident argument blocks of the dycore stages: every pointer is a unified plane
// (common.cuh).  The C-ABI entries stage host arrays into these; the fused step driver passes its
// own resident planes, so no stage ever touches host memory.
#pragma once
#include "../../include/mom6cu.h"

// ocean_grid_type metrics on the device (src/core/MOM_grid.F90:75-175)
struct GridDev {
  const double *mask2dT, *mask2dCu, *mask2dCv, *mask2dBu;
  const double *dxT, *dyT, *IdxT, *IdyT, *areaT, *IareaT;
  const double *dxCu, *dyCu, *IdxCu, *IdyCu, *dy_Cu, *areaCu, *IareaCu;
  const double *dxCv, *dyCv, *IdxCv, *IdyCv, *dx_Cv, *areaCv, *IareaCv;
  const double *dxBu, *dyBu, *IdxBu, *IdyBu, *areaBu, *IareaBu;
  const double *bathyT, *CoriolisBu, *Coriolis2Bu;
};

// continuity_PPM dummy arguments (MOM_continuity_PPM.F90:86-141); null = absent optional
struct ContinuityDev {
  const double *u, *v, *hin;
  double *h, *uh, *vh;
  double dt;
  const double *por_face_areaU, *por_face_areaV;
  const double *uhbt, *vhbt;
  const double *visc_rem_u, *visc_rem_v;
  double *u_cor, *v_cor, *du_cor, *dv_cor;
  int have_BT_cont;
  double *FA_u_EE, *FA_u_E0, *FA_u_W0, *FA_u_WW, *uBT_WW, *uBT_EE;
  double *FA_v_NN, *FA_v_N0, *FA_v_S0, *FA_v_SS, *vBT_SS, *vBT_NN;
  double *h_u, *h_v;
};

// CorAdCalc dummy arguments (MOM_CoriolisAdv.F90:125-144); null = absent optional / diagnostic not requested
struct CorAdDev {
  const double *u, *v, *h, *uh, *vh;
  double *CAu, *CAv;
  const double *por_face_areaU, *por_face_areaV;
  double *RV, *PV, *gradKEu, *gradKEv;
};

// horizontal_viscosity dummy arguments (MOM_hor_visc.F90:266-305)
struct HorViscDev {
  const double *u, *v, *h, *hu_cont, *hv_cont;
  double *diffu, *diffv;
};

// btstep dummy arguments (MOM_barotropic.F90:455-529); null = absent / not associated
struct BtstepDev {
  double dt;
  const double *U_in, *V_in, *eta_in, *bc_accel_u, *bc_accel_v, *taux, *tauy, *pbce, *eta_PF_in, *U_Cor, *V_Cor;
  const double *visc_rem_u, *visc_rem_v, *taux_bot, *tauy_bot, *uh0, *vh0, *u_uh0, *v_vh0;
  double *accel_layer_u, *accel_layer_v, *eta_out, *uhbtav, *vhbtav, *etaav;
  int have_BT_cont;
  const double *FA_u_EE, *FA_u_E0, *FA_u_W0, *FA_u_WW, *uBT_WW, *uBT_EE;
  const double *FA_v_NN, *FA_v_N0, *FA_v_S0, *FA_v_SS, *vBT_SS, *vBT_NN;
};

// PressureForce dummy arguments (MOM_PressureForce.F90:40-61)
struct PgfDev {
  const double *h, *T, *S, *p_atm;
  double *PFu, *PFv, *pbce, *eta;
};

// vertvisc_coef / vertvisc dummy arguments (MOM_vert_friction.F90:1357, :557)
struct VvCoefDev {
  const double *u, *v, *h, *Kv_bbl_u, *Kv_bbl_v, *bbl_thick_u, *bbl_thick_v, *Kv_shear, *Kv_shear_Bu, *ustar;
  double dt;
};
struct VvDev {
  double *u, *v;
  const double *h, *taux, *tauy, *Ray_u, *Ray_v;
  double dt;
  double *taux_bot, *tauy_bot;
};

struct mom6cu_ctx;
struct Stager;
int m6_vertvisc_coef_run(mom6cu_ctx* c, const VvCoefDev& D);
int m6_vertvisc_run(mom6cu_ctx* c, const VvDev& D);
int m6_vertvisc_remnant_run(mom6cu_ctx* c, const double* Ray_u, const double* Ray_v, double* visc_rem_u, double* visc_rem_v, double dt);
// stage the array members of a barotropic_CS given with host or resident pointers into *CS (device pointers)
int m6_stage_barotropic_cs(mom6cu_ctx* c, Stager& S, const mom6cu_barotropic_cs* CSh, mom6cu_barotropic_cs* CS);
int m6_set_dtbt_run(mom6cu_ctx* c, const mom6cu_set_dtbt_args& a, double* dtbt, double* dtbt_max);
int m6_pressure_force_run(mom6cu_ctx* c, const PgfDev& D);
// CS holds device pointers (resident planes) for every array member
int m6_btstep_run(mom6cu_ctx* c, const mom6cu_barotropic_cs& CS, const BtstepDev& D);
int m6_btcalc_run(mom6cu_ctx* c, const double* h, const double* h_u, const double* h_v, const double* bathyT, int hvel_scheme,
                  int may_use_default, double* frhatu, double* frhatv);
int m6_bt_mass_source_run(mom6cu_ctx* c, const double* h, const double* eta, int set_cor, double* eta_cor);
int m6_hor_visc_run(mom6cu_ctx* c, const HorViscDev& D);
int m6_coradcalc_run(mom6cu_ctx* c, const CorAdDev& D);
int m6_continuity_run(mom6cu_ctx* c, const ContinuityDev& D);
