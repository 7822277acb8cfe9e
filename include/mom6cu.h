/*
 * mom6cu.h -- C ABI of the B200-native split-explicit dycore hot path.
 *
 * Every entry point mirrors one Fortran module procedure of the reference
 * (mom-ocean/MOM6) that the src/core driver calls; the reference routine each
 * one replaces is cited as file:line.  Fortran host code binds these through
 * ISO_C_BINDING (see fortran/mom6cu_interface.F90 and INTEGRATION.md).
 *
 * Conventions
 *  - All reals are IEEE binary64 ("real" in a default MOM6 build, -fdefault-real-8).
 *  - Arrays are contiguous, column-major (i fastest), exactly as the Fortran
 *    dummy arguments are declared.  Extents follow the reference's symmetric
 *    memory macros (src/framework/MOM_memory_macros.h, dynamic branch):
 *        h-points  (isd:ied,   jsd:jed)
 *        u-points  (isd-1:ied, jsd:jed)      SZIB_(G),SZJ_(G)
 *        v-points  (isd:ied,   jsd-1:jed)    SZI_(G),SZJB_(G)
 *        q-points  (isd-1:ied, jsd-1:jed)
 *    "wide" barotropic arrays use (isdw,iedw,jsdw,jedw) the same way
 *    (SZIW_/SZIBW_/SZJW_/SZJBW_, MOM_barotropic.F90:46-51).
 *  - A pointer argument may be a HOST pointer (the Fortran array) or a DEVICE
 *    pointer to a buffer of the same shape; the library detects which with
 *    cudaPointerGetAttributes.  Host arrays are staged to the device-resident
 *    layout and results copied back inside the call.
 *  - Return value: 0 = success; >0 = FATAL (message via mom6cu_last_error),
 *    the Fortran shim turns it into MOM_error(FATAL, msg)
 *    (src/framework/MOM_error_handler.F90); <0 = -(number of WARNINGs).
 *  - There is NO CPU fallback: without a CUDA device every compute entry
 *    returns MOM6CU_ERR_NO_DEVICE.
 */
#ifndef MOM6CU_H
#define MOM6CU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOM6CU_OK 0
#define MOM6CU_ERR_NO_DEVICE 1
#define MOM6CU_ERR_BAD_ARG 2
#define MOM6CU_ERR_UNSUPPORTED 3 /* a run-time option outside the frozen option set */
#define MOM6CU_ERR_CUDA 4
#define MOM6CU_ERR_NCCL 5

typedef struct mom6cu_ctx mom6cu_ctx; /* opaque: one per MPI rank / GPU */

/* Index bounds of one rank's tile: the subset of hor_index_type
 * (src/framework/MOM_hor_index.F90) and of barotropic_CS%isdw.. that the hot
 * path uses.  Fortran (1-based, halo-offset) indices. */
typedef struct mom6cu_domain {
  int isc, iec, jsc, jec;     /* computational domain, tracer points           */
  int isd, ied, jsd, jed;     /* data (memory) domain of G                     */
  int isdw, iedw, jsdw, jedw; /* data domain of the wide-halo barotropic arrays */
  int nk;                     /* GV%ke                                          */
  int cyclic_x, cyclic_y;     /* 1 = reentrant in that direction (MOM_domains.F90 REENTRANT_X/Y) */
  int first_direction;        /* G%first_direction (MOM_grid.F90)               */
  /* layout of tiles over ranks (MOM_domains.F90 LAYOUT): rank = pj*npi + pi    */
  int npi, npj, pi, pj;
} mom6cu_domain;

/* ---------------------------------------------------------------- lifecycle */
/* Replaces nothing in the reference; called once from the Fortran shim's
 * initialize_dyn_split_RK2 wrapper (MOM_dynamics_split_RK2.F90:1338). */
int mom6cu_create(mom6cu_ctx** ctx, const mom6cu_domain* dom, int device);
int mom6cu_destroy(mom6cu_ctx* ctx);
/* Copies the last error/warning message (NUL terminated) into buf. */
int mom6cu_last_error(const mom6cu_ctx* ctx, char* buf, size_t len);
/* Library/build information: returns the sm arch the kernels were built for (100). */
int mom6cu_build_arch(void);
/* sizeof(struct <name>) of any struct declared in this header (-1: unknown name): for bindings to check their mirrors. */
long long mom6cu_sizeof(const char* name);
/* Number of kernel launches issued by this context since creation. */
long long mom6cu_launch_count(const mom6cu_ctx* ctx);
/* Block until all work queued on the context's streams is done. */
int mom6cu_sync(mom6cu_ctx* ctx);
/* Device-time (ms) of the most recent compute entry, measured with CUDA
 * events on the launching stream (excludes host<->device staging). */
double mom6cu_last_kernel_ms(const mom6cu_ctx* ctx);
/* passes made by the most recent iterative entry (mom6cu_advect_tracer: the itt loop, MOM_tracer_advect.F90:222-340) */
int mom6cu_last_iterations(const mom6cu_ctx* ctx);
/* Device time (ms) each stage took INSIDE the most recent mom6cu_step_dyn_split_rk2 call, summed over its calls in the step
 * (CUDA events around the stage's kernels on the compute stream).  Fills ms[0..n): MOM6CU_STAGE_* below; returns the
 * number of entries written. */
#define MOM6CU_STAGE_PRESSURE_FORCE 0
#define MOM6CU_STAGE_CORADCALC 1
#define MOM6CU_STAGE_VERTVISC 2
#define MOM6CU_STAGE_CONTINUITY 3
#define MOM6CU_STAGE_BTCALC 4
#define MOM6CU_STAGE_BTSTEP 5
#define MOM6CU_STAGE_HOR_VISC 6
#define MOM6CU_NSTAGES 7
int mom6cu_last_step_stage_ms(const mom6cu_ctx* ctx, double* ms, int n);
/* Sum of the device times of all repetitions of the most recent *_resident call. */
double mom6cu_total_kernel_ms(const mom6cu_ctx* ctx);

/* --------------------------------------------------------- field residency */
/* The Fortran driver owns the model state; these calls let it keep a field on the device between stages
 * (SURVEY 8b "ownership"): a plane is a device buffer in the library's resident layout.  Passing a plane
 * pointer wherever an entry point takes an array skips the staging copies entirely (inputs are used in place,
 * outputs are left on the device), so a sequence of stages runs with no host traffic; mom6cu_plane_download is
 * the sync point (diagnostics, restarts, halo updates done by un-replaced host code).
 *   stagger 0=h,1=u,2=v,3=q; wide=1 for the barotropic wide-halo arrays; nk = number of levels (1 for 2-D). */
double* mom6cu_plane_alloc(mom6cu_ctx* ctx, const char* name, int nk);
int mom6cu_plane_upload(mom6cu_ctx* ctx, double* plane, const double* host, int stagger, int wide, int nk);
int mom6cu_plane_download(mom6cu_ctx* ctx, const double* plane, double* host, int stagger, int wide, int nk);
/* plane(:,:,1:nk) = 0 on the device (asynchronous on the context's stream). */
int mom6cu_plane_zero(mom6cu_ctx* ctx, double* plane, int nk);

/* ------------------------------------------------------------ grid metrics */
/* The fields of ocean_grid_type (src/core/MOM_grid.F90:75-175) the hot path reads.
 * All are G-sized 2-D arrays of the staggering in the comment; uploaded once and
 * kept resident (the reference computes them once in set_grid_metrics). */
typedef struct mom6cu_grid {
  const double *mask2dT, *mask2dCu, *mask2dCv, *mask2dBu;   /* h,u,v,q */
  const double *dxT, *dyT, *IdxT, *IdyT, *areaT, *IareaT;   /* h */
  const double *dxCu, *dyCu, *IdxCu, *IdyCu, *dy_Cu, *areaCu, *IareaCu; /* u */
  const double *dxCv, *dyCv, *IdxCv, *IdyCv, *dx_Cv, *areaCv, *IareaCv; /* v */
  const double *dxBu, *dyBu, *IdxBu, *IdyBu, *areaBu, *IareaBu;         /* q */
  const double *bathyT;                                     /* h */
  const double *CoriolisBu, *Coriolis2Bu;                   /* q */
} mom6cu_grid;
#define MOM6CU_GRID_NFIELDS 33
int mom6cu_set_grid(mom6cu_ctx* ctx, const mom6cu_grid* G);

/* verticalGrid_type scalars (src/core/MOM_verticalGrid.F90) */
typedef struct mom6cu_vgrid {
  double Angstrom_H, H_subroundoff, Z_to_H, H_to_Z, g_Earth, Rho0, H_to_RZ, RZ_to_H, H_to_m, m_to_H;
  int Boussinesq;
} mom6cu_vgrid;
int mom6cu_set_vgrid(mom6cu_ctx* ctx, const mom6cu_vgrid* GV);

/* ---------------------------------------------------------- continuity_PPM */
/* continuity_PPM_CS, src/core/MOM_continuity_PPM.F90:35-67, resolved values */
typedef struct mom6cu_continuity_cs {
  int upwind_1st, monotonic, simple_2nd, aggress_adjust, vol_CFL, better_iter,
      use_visc_rem_max, marginal_faces;
  double tol_eta, tol_vel, CFL_limit_adjust;
} mom6cu_continuity_cs;

/* BT_cont_type, src/core/MOM_variables.F90:315-350 (G-sized; h_u/h_v 3-D or NULL) */
typedef struct mom6cu_bt_cont {
  double *FA_u_EE, *FA_u_E0, *FA_u_W0, *FA_u_WW, *uBT_WW, *uBT_EE; /* u-points */
  double *FA_v_NN, *FA_v_N0, *FA_v_S0, *FA_v_SS, *vBT_SS, *vBT_NN; /* v-points */
  double *h_u, *h_v;                                               /* 3-D u / v */
} mom6cu_bt_cont;

/* continuity_PPM(u, v, hin, h, uh, vh, dt, G, GV, US, CS, OBC, pbv, uhbt, vhbt, visc_rem_u,
 *                visc_rem_v, u_cor, v_cor, BT_cont, du_cor, dv_cor)
 * src/core/MOM_continuity_PPM.F90:86-194 (aliased `continuity`, MOM_continuity.F90:6).
 * Optional Fortran dummies are NULL when absent (OBC must be absent: rejected).  hin and
 * h may alias (the corrector call, MOM_dynamics_split_RK2.F90:1043). */
typedef struct mom6cu_continuity_args {
  const double *u, *v, *hin; /* 3-D u, v, h */
  double *h, *uh, *vh;       /* 3-D h (inout), u, v (out) */
  double dt;
  const double *por_face_areaU, *por_face_areaV; /* 3-D u, v; NULL = 1 (USE_POROUS_BARRIER=False) */
  const double *uhbt, *vhbt;             /* 2-D u, v, optional */
  const double *visc_rem_u, *visc_rem_v; /* 3-D, optional (both or neither) */
  double *u_cor, *v_cor;                 /* 3-D, optional out */
  mom6cu_bt_cont* BT_cont;               /* optional */
  double *du_cor, *dv_cor;               /* 2-D, optional out */
} mom6cu_continuity_args;

int mom6cu_set_cs_continuity(mom6cu_ctx* ctx, const mom6cu_continuity_cs* CS);
int mom6cu_continuity(mom6cu_ctx* ctx, const mom6cu_continuity_args* a);

/* unit_scale_type factors the hot path uses (src/framework/MOM_unit_scaling.F90); all 1 in an
 * unscaled run.  Defaults to all ones when never set. */
typedef struct mom6cu_unit_scale {
  double m_to_L, L_to_m, m_s_to_L_T, L_T_to_m_s, s_to_T, T_to_s, m_to_Z, Z_to_m, Z_to_L, L_to_Z;
} mom6cu_unit_scale;
int mom6cu_set_unit_scale(mom6cu_ctx* ctx, const mom6cu_unit_scale* US);

/* --------------------------------------------------------------- CorAdCalc */
/* CoriolisAdv_CS, src/core/MOM_CoriolisAdv.F90:30-91, resolved by CoriolisAdv_init :1054.
 * Enumeration values are the reference's (:94-119). */
#define MOM6CU_SADOURNY75_ENERGY 1
#define MOM6CU_ARAKAWA_HSU90 2
#define MOM6CU_ROBUST_ENSTRO 3
#define MOM6CU_SADOURNY75_ENSTRO 4
#define MOM6CU_ARAKAWA_LAMB81 5
#define MOM6CU_AL_BLEND 6
#define MOM6CU_KE_ARAKAWA 10
#define MOM6CU_KE_SIMPLE_GUDONOV 11
#define MOM6CU_KE_GUDONOV 12
#define MOM6CU_PV_ADV_CENTERED 21
#define MOM6CU_PV_ADV_UPWIND1 22
typedef struct mom6cu_coriolisadv_cs {
  int Coriolis_Scheme, KE_Scheme, PV_Adv_Scheme;
  int no_slip, bound_Coriolis, Coriolis_En_Dis;
  double F_eff_max_blend, wt_lin_blend;
} mom6cu_coriolisadv_cs;

/* CorAdCalc(u, v, h, uh, vh, CAu, CAv, OBC, AD, G, GV, US, CS, pbv, Waves)
 * src/core/MOM_CoriolisAdv.F90:125-965 (+ gradKE :969-1051).  OBC and Waves must be absent.
 * Optional diagnostics (NULL = not requested): RV/PV (q-points, 3-D; CS%id_rv/id_PV),
 * gradKEu/gradKEv (AD%gradKEu/v). */
typedef struct mom6cu_coradcalc_args {
  const double *u, *v, *h, *uh, *vh; /* 3-D u, v, h, u, v */
  double *CAu, *CAv;                 /* 3-D u, v (out) */
  const double *por_face_areaU, *por_face_areaV; /* pbv, 3-D u, v; NULL = 1 */
  double *RV, *PV;                   /* 3-D q, optional out */
  double *gradKEu, *gradKEv;         /* 3-D u, v, optional out */
} mom6cu_coradcalc_args;
int mom6cu_set_cs_coriolisadv(mom6cu_ctx* ctx, const mom6cu_coriolisadv_cs* CS);
int mom6cu_coradcalc(mom6cu_ctx* ctx, const mom6cu_coradcalc_args* a);

/* ---------------------------------------------------- horizontal_viscosity */
/* hor_visc_CS, src/parameterizations/lateral/MOM_hor_visc.F90:38-250, as resolved by hor_visc_init
 * (:2322-3302).  The 2-D members are the static arrays hor_visc_init precomputes (:2834-3120), passed as the
 * Fortran arrays they are (h-, q-, u- or v-point, G-sized); an array an option does not use may be NULL.
 * Frozen options: Leith / Leith+E / QG-Leith, GME, MEKE viscosities and backscatter, anisotropic viscosity,
 * ZB2020, resolution-function scaling and OBCs are rejected (MOM6CU_ERR_UNSUPPORTED). */
typedef struct mom6cu_hor_visc_cs {
  int Laplacian, biharmonic, no_slip, bound_Kh, better_bound_Kh, bound_Ah, better_bound_Ah,
      backscatter_underbound, Smagorinsky_Kh, Smagorinsky_Ah, bound_Coriolis, use_land_mask,
      add_LES_viscosity, use_cont_thick, use_cont_thick_bug;
  int unsupported; /* nonzero if any of the rejected options is set in the run's parameters */
  double Kh_bg_min, Re_Ah;
  /* h-points */
  const double *dx2h, *dy2h, *DX_dyT, *DY_dxT, *reduction_xx, *Kh_bg_xx, *Ah_bg_xx, *Kh_Max_xx, *Ah_Max_xx,
      *Laplac2_const_xx, *Biharm_const_xx, *Biharm_const2_xx, *Re_Ah_const_xx;
  /* q-points */
  const double *dx2q, *dy2q, *DX_dyBu, *DY_dxBu, *reduction_xy, *Kh_bg_xy, *Ah_bg_xy, *Kh_Max_xy, *Ah_Max_xy,
      *Laplac2_const_xy, *Biharm_const_xy, *Biharm_const2_xy, *Re_Ah_const_xy;
  /* u-points, v-points */
  const double *Idx2dyCu, *Idxdy2u, *Idx2dyCv, *Idxdy2v;
} mom6cu_hor_visc_cs;
#define MOM6CU_HOR_VISC_NARRAYS 30

/* horizontal_viscosity(u, v, h, uh, vh, diffu, diffv, MEKE, VarMix, G, GV, US, CS, tv, dt, OBC, BT, TD, ADp,
 *                      hu_cont, hv_cont, STOCH)   MOM_hor_visc.F90:266-267.
 * uh, vh feed only the FrictWork diagnostics (not computed here) and may be NULL. */
typedef struct mom6cu_hor_visc_args {
  const double *u, *v, *h, *uh, *vh; /* 3-D u, v, h, u, v */
  double *diffu, *diffv;             /* 3-D u, v (out) */
  const double *hu_cont, *hv_cont;   /* 3-D u, v, optional */
  double dt;
} mom6cu_hor_visc_args;
int mom6cu_set_cs_hor_visc(mom6cu_ctx* ctx, const mom6cu_hor_visc_cs* CS);
int mom6cu_horizontal_viscosity(mom6cu_ctx* ctx, const mom6cu_hor_visc_args* a);

/* ----------------------------------------------------------- PressureForce */
/* PressureForce (src/core/MOM_PressureForce.F90:40-82) with ANALYTIC_FV_PGF=True and GV%Boussinesq, i.e.
 * PressureForce_FV_Bouss (src/core/MOM_PressureForce_FV.F90:947-2017) with the analytic layer integrals
 * int_density_dz (src/core/MOM_density_integrals.F90:42-103 -> src/equation_of_state/MOM_EOS.F90:1384-1499 ->
 * int_density_dz_linear MOM_EOS_linear.F90:275 / int_density_dz_wright MOM_EOS_Wright.F90:389) and
 * Set_pbce_Bouss (src/core/MOM_PressureForce_Montgomery.F90:649-748).
 * PressureForce_FV_CS (:40-107) + the EOS_type / verticalGrid members the routine reads.
 * Frozen options: no tides / SAL, no Stanley SGS term, no RESET_INTXPA_INTEGRAL / CORRECTION_INTXPA, no bulk
 * mixed layer (GV%nk_rho_varies = 0); T,S piecewise constant within layers (analytic integrals, EOS_quadrature off) or,
 * with RECONSTRUCT_FOR_PRESSURE (the reference's default under ALE), PLM / PPM sub-layer profiles integrated by quadrature;
 * EOS forms: none (layered), LINEAR, WRIGHT. */
#define MOM6CU_EOS_NONE 0
#define MOM6CU_EOS_LINEAR 1
#define MOM6CU_EOS_WRIGHT 3
typedef struct mom6cu_pressureforce_cs {
  int EOS_form, MassWghtInterp, use_SSH_in_Z0p, rho_ref_bug, unsupported;
  double rho_ref, GFS_scale, Z_ref, dZ_subroundoff;
  double Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp; /* EOS_LINEAR */
  const double *Rlay, *g_prime;                /* GV%Rlay(1:nk), GV%g_prime(1:nk+1), host arrays (EOS_NONE) */
  /* RECONSTRUCT_FOR_PRESSURE (MOM_PressureForce_FV.F90:2172-2190): sub-layer T,S profiles from the ALE reconstructions
   * (TS_PLM_edge_values / TS_PPM_edge_values, MOM_ALE.F90:1495/1581) integrated by quadrature
   * (int_density_dz_generic_plm / _ppm, MOM_density_integrals.F90:418/874).  reconstruct = CS%reconstruct .and. use_EOS .and.
   * associated(ALE_CSp); Recon_Scheme = PRESSURE_RECONSTRUCTION_SCHEME (1 = PLM, 2 = PPM); boundary_extrap =
   * BOUNDARY_EXTRAPOLATION_PRESSURE; ALE_answer_date = ALE_CS%answer_date (>= 20190101 only). */
  int reconstruct, Recon_Scheme, boundary_extrap, use_inaccurate_pgf_rho_anom, MassWghtInterpVanOnly, ALE_answer_date;
  double h_nonvanished; /* CS%h_nonvanished [H] */
  /* EOS_type unit conversion factors (MOM_EOS.F90:140-150); 0 is read as 1.  The device path requires them to be 1 (the
   * reference's default); the oracle honours them (the dimensional-rescaling test of the pressure force). */
  double kg_m3_to_R, RL2_T2_to_Pa, C_to_degC, S_to_ppt;
} mom6cu_pressureforce_cs;

/* PressureForce(h, tv, PFu, PFv, G, GV, US, CS, ALE_CSp, ADp, p_atm, pbce, eta)  MOM_PressureForce.F90:40 */
typedef struct mom6cu_pressureforce_args {
  const double *h, *T, *S; /* 3-D h; tv%T, tv%S (NULL with EOS_NONE) */
  double *PFu, *PFv;       /* 3-D u, v (out) */
  const double* p_atm;     /* 2-D h, optional */
  double* pbce;            /* 3-D h, optional out */
  double* eta;             /* 2-D h, optional out */
} mom6cu_pressureforce_args;
int mom6cu_set_cs_pressureforce(mom6cu_ctx* ctx, const mom6cu_pressureforce_cs* CS);
int mom6cu_pressure_force(mom6cu_ctx* ctx, const mom6cu_pressureforce_args* a);

/* ---------------------------------------------------------------- ALE remap */
/* remapping_CS, src/ALE/MOM_remapping.F90:37-85, as resolved by initialize_remapping.  Schemes (:89-94):
 * PCM 0, PLM 2, PPM_H4 4, PPM_IH4 5.  Frozen: answer_date >= 20190101, the OM4-era reconstruction functions
 * (not the Recon1d classes), no PQM / hybgen / PPM_CW schemes. */
#define MOM6CU_REMAPPING_PCM 0
#define MOM6CU_REMAPPING_PLM 2
#define MOM6CU_REMAPPING_PPM_H4 4
#define MOM6CU_REMAPPING_PPM_IH4 5
typedef struct mom6cu_remapping_cs {
  int remapping_scheme, boundary_extrapolation, force_bounds_in_subcell, force_bounds_in_target,
      om4_remap_via_sub_cells, answer_date;
  double h_neglect, h_neglect_edge;
} mom6cu_remapping_cs;

/* ALE_remap_tracers(CS, G, GV, h_old, h_new, Reg, ...)  src/ALE/MOM_ALE.F90:760-879: every column of each of the ntr
 * h-point fields (Reg%Tr(m)%t) with mask2dT > 0 is remapped conservatively from h_old to h_new with
 * remapping_core_h (MOM_remapping.F90:234); conc_underflow[m] > 0 flushes tiny values (:822-824). */
int mom6cu_ale_remap_tracers(mom6cu_ctx* ctx, const mom6cu_remapping_cs* CS, const double* h_old, const double* h_new, int ntr,
                             double* const* tr, const double* conc_underflow);
/* ALE_remap_set_h_vel, MOM_ALE.F90:882-925 and ALE_remap_velocities :1089-1300 (vel_remapCS; no KE-conserving
 * correction, no near-bottom masking). */
int mom6cu_ale_remap_set_h_vel(mom6cu_ctx* ctx, const double* h_new, double* h_u, double* h_v);
int mom6cu_ale_remap_velocities(mom6cu_ctx* ctx, const mom6cu_remapping_cs* CS, const double* h_old_u, const double* h_old_v,
                                const double* h_new_u, const double* h_new_v, double* u, double* v);
/* remapping_core_h on a batch of independent columns (the reference's unit-test / diagnostic entry point,
 * MOM_remapping.F90:234): h0,u0 are (ncol,n0), h1,u1 (ncol,n1), row-major, host or device. */
int mom6cu_remapping_core_h(mom6cu_ctx* ctx, const mom6cu_remapping_cs* CS, int ncol, int n0, const double* h0, const double* u0,
                            int n1, const double* h1, double* u1);

/* ---------------------------------------------------------- vertical friction */
/* vertvisc_CS members read by vertvisc_coef / vertvisc / vertvisc_remnant
 * (src/parameterizations/vertical/MOM_vert_friction.F90:48-170).  Frozen: answer_date >= 20190101 (I_amax = 0),
 * Boussinesq (thickness_to_dz: dz = H_to_Z*h, MOM_interface_heights.F90:939; find_ustar = forces%ustar), GV%nkml = 0,
 * DYNAMIC_VISCOUS_ML off, no GL90, no ice shelves, no OBCs, no Stokes mixing / fpmix; the a_u/h_u diagnostics are not
 * produced.  Set `unsupported` when the host configuration enables any of those. */
typedef struct mom6cu_vertvisc_cs {
  int bottomdraglaw, harmonic_visc, direct_stress, fixed_LOTW_ML, apply_LOTW_floor, dynamic_viscous_ML, nkml, answer_date,
      unsupported;
  int CFL_based_trunc; /* CFL_BASED_TRUNCATIONS (default True): see maxvel / CFL_trunc */
  double Hbbl, Kv, Kv_extra_bbl, Kvml_invZ2, Hmix, Hmix_stress, harm_BL_val, vonKar, vel_underflow;
  double dZ_subroundoff; /* GV%dZ_subroundoff */
  /* vertvisc_limit_vel (MOM_vert_friction.F90:2926-3120), the last act of vertvisc: with CFL_based_trunc a velocity whose CFL number
   * exceeds CFL_trunc (CFL_TRUNCATE, default 0.5) is set to the velocity of CFL 0.9*CFL_trunc; otherwise one above maxvel (MAXVEL,
   * default 3e8 m/s) is set to 0.9*maxvel; |u| < vel_underflow is flushed to zero either way.  Truncations in columns thicker than
   * 6 Angstrom are counted (CS%ntrunc: mom6cu_vertvisc_ntrunc).  CFL_trunc <= 0 and maxvel <= 0 switch the respective test off
   * (a zero-filled structure limits nothing).  U_TRUNC_FILE / V_TRUNC_FILE reporting is the host's business. */
  double maxvel, CFL_trunc;
} mom6cu_vertvisc_cs;
int mom6cu_set_cs_vertvisc(mom6cu_ctx* ctx, const mom6cu_vertvisc_cs* CS);
/* CS%ntrunc: the number of velocity truncations vertvisc has made in this context so far (needs h in the vertvisc call);
 * -1 on error (message via mom6cu_last_error) */
long long mom6cu_vertvisc_ntrunc(mom6cu_ctx* ctx);
/* vertvisc_coef(u, v, h, dz, forces, visc, tv, dt, G, GV, US, CS, OBC, VarMix)  MOM_vert_friction.F90:1357: sets the
 * resident CS%a_u, CS%a_v (nk+1 interfaces) and CS%h_u, CS%h_v (nk layers). */
typedef struct mom6cu_vertvisc_coef_args {
  const double *u, *v, *h;                                       /* 3-D u, v, h */
  const double *Kv_bbl_u, *Kv_bbl_v, *bbl_thick_u, *bbl_thick_v; /* visc%, 2-D u / v (bottomdraglaw) */
  const double* Kv_shear;    /* visc%Kv_shear, h-points, nk+1 interfaces; NULL if not associated */
  const double* Kv_shear_Bu; /* visc%Kv_shear_Bu, q-points, nk+1 interfaces; NULL if not associated */
  const double* ustar;       /* forces%ustar, 2-D h (LOTW options only) */
  double dt;
} mom6cu_vertvisc_coef_args;
int mom6cu_vertvisc_coef(mom6cu_ctx* ctx, const mom6cu_vertvisc_coef_args* a);
/* copies of the resident coupling coefficients (tests / diagnostics): a_u, a_v are 3-D u / v with nk+1 levels, h_u, h_v nk */
int mom6cu_vertvisc_get_coef(mom6cu_ctx* ctx, double* a_u, double* a_v, double* h_u, double* h_v);
/* vertvisc(u, v, h, forces, visc, dt, OBC, ADp, CDp, G, GV, US, CS, taux_bot, tauy_bot)  :557 */
typedef struct mom6cu_vertvisc_args {
  double *u, *v;               /* 3-D, in/out */
  const double* h;             /* 3-D (direct_stress only) */
  const double *taux, *tauy;   /* forces%taux, tauy, 2-D u / v */
  const double *Ray_u, *Ray_v; /* visc%Ray_u/v 3-D, NULL if not allocated */
  double dt;
  double *taux_bot, *tauy_bot; /* optional 2-D out */
} mom6cu_vertvisc_args;
int mom6cu_vertvisc(mom6cu_ctx* ctx, const mom6cu_vertvisc_args* a);
/* vertvisc_remnant(visc, visc_rem_u, visc_rem_v, dt, G, GV, US, CS)  :1229 */
int mom6cu_vertvisc_remnant(mom6cu_ctx* ctx, const double* Ray_u, const double* Ray_v, double* visc_rem_u, double* visc_rem_v, double dt);

/* ---------------------------------------------------------------- ALE regrid */
/* regridding_CS members read by the Z* path (src/ALE/MOM_regridding.F90:49-160) with zlike_CS (coord_zlike.F90:12-22).
 * Frozen: REGRIDDING_ZSTAR (regrid_consts.F90:14), CS%nk == GV%ke, no ice shelf (frac_shelf_h absent), Boussinesq
 * (tv%SpV_avg not allocated), no PCM_cell output. */
#define MOM6CU_REGRIDDING_ZSTAR 2
typedef struct mom6cu_regridding_cs {
  int regridding_scheme, nk;
  double min_thickness;                 /* CS%min_thickness == CS%zlike_CS%min_thickness (set_regrid_params :2425-2440) */
  double old_grid_weight, depth_of_time_filter_shallow, depth_of_time_filter_deep;
  double Z_ref;                         /* G%Z_ref */
  const double* coordinateResolution;   /* (nk), host array [Z] */
} mom6cu_regridding_cs;
/* ALE_regrid(G, GV, US, h, h_new, dzRegrid, tv, CS)  src/ALE/MOM_ALE.F90:518-554 -> regridding_main
 * (MOM_regridding.F90:846-972) -> build_zstar_grid :1257 + calc_h_new_by_dz :1008.  h_new is a 3-D h field,
 * dzRegrid a 3-D h field with nk+1 levels; both are set on (isc-1:iec+1, jsc-1:jec+1) as in the reference.
 * A column with a negative thickness, or whose implied new thickness is negative beyond roundoff
 * (adjust_interface_motion :1808-1817), is FATAL. */
int mom6cu_ale_regrid(mom6cu_ctx* ctx, const mom6cu_regridding_cs* CS, const double* h, double* h_new, double* dzRegrid);

/* ------------------------------------------------------------ advect_tracer */
/* tracer_advect_CS, src/tracer/MOM_tracer_advect.F90:32-41; schemes MOM_tracer_advect_schemes.F90:10-12 */
#define MOM6CU_ADVECT_PLM 0
#define MOM6CU_ADVECT_PPMH3 1
#define MOM6CU_ADVECT_PPM 2
typedef struct mom6cu_tracer_advect_cs {
  double dt;                  /* CS%dt, the baroclinic time step (sets max_iter, :176) */
  int default_advect_scheme;  /* TRACER_ADVECTION_SCHEME */
  int useHuynhStencilBug;
} mom6cu_tracer_advect_cs;
/* advect_tracer(h_end, uhtr, vhtr, OBC, dt, G, GV, US, CS, Reg, x_first_in, vol_prev, max_iter_in, update_vol_prev,
 * uhr_out, vhr_out)  MOM_tracer_advect.F90:53-54.  Reg is passed as ntr field pointers with their per-tracer
 * advect_scheme (< 0: CS default) and conc_underflow; OBCs are rejected; the flux diagnostics (ad_x, ad_y, ad2d_*,
 * advection_xy) are not produced. */
typedef struct mom6cu_advect_tracer_args {
  const double *h_end, *uhtr, *vhtr; /* 3-D h, u, v */
  double dt;
  int ntr;
  double* const* tr;            /* ntr 3-D h fields, in/out (halos valid on entry, as in the reference) */
  const int* advect_scheme;     /* ntr, may be NULL (all default) */
  const double* conc_underflow; /* ntr, may be NULL */
  int x_first_in;               /* -1 absent, else 0/1 */
  int max_iter_in;              /* < 0 absent */
  double* vol_prev;             /* optional 3-D h, in(/out) */
  int update_vol_prev;
  double *uhr_out, *vhr_out;    /* optional 3-D u, v out */
} mom6cu_advect_tracer_args;
int mom6cu_advect_tracer(mom6cu_ctx* ctx, const mom6cu_tracer_advect_cs* CS, const mom6cu_advect_tracer_args* a);

/* ------------------------------------------------------- halo communication */
/* The reference's halo API (pass_var / pass_vector / do_group_pass,
 * src/framework/MOM_domains.F90 -> config_src/infra/FMS2/MOM_domain_infra.F90:171-216,
 * :1141-1200) is FMS mpp_update_domains over MPI.  Here each group pass is one pack
 * kernel + ncclSend/ncclRecv to the <=8 neighbour tiles + one unpack kernel.
 * mom6cu_comm_unique_id: rank 0 creates the ncclUniqueId (>=128 bytes buffer), the host
 * (MPI_Bcast in the Fortran shim, torch.distributed in the harness) broadcasts it, every
 * rank calls mom6cu_comm_init. */
int mom6cu_comm_unique_id(void* out, int nbytes);
int mom6cu_comm_init(mom6cu_ctx* ctx, const void* id_bytes, int nbytes, int rank, int nranks);
int mom6cu_comm_destroy(mom6cu_ctx* ctx);
/* pass_var / pass_vector / do_group_pass (src/framework/MOM_domains.F90; the calls that follow thickness_diffuse and mixedlayer_restrat in
 * step_MOM_dynamics, MOM.F90:1396,1427) as an entry of its own, for a caller that chains the entries on resident fields: the halos of
 * nfields fields of nk levels each (host arrays or resident planes; stagger[f] = 0 h, 1 u, 2 v, 3 q) are updated out to the full width
 * of G's memory domain -- reentrant wrap on one tile, NCCL send/recv between tiles, closed edges untouched.  Vector components
 * are passed like scalars (no tripolar fold or other sign-changing boundary is implemented). */
int mom6cu_do_group_pass(mom6cu_ctx* ctx, int nfields, double* const* fields, const int* stagger, int nk);
/* Host-only planning of one neighbour message (no device needed): for direction
 * dir in 0..7 = {E,W,N,S,NE,SW,SE,NW} returns the peer rank (-1: closed edge) and the
 * inclusive Fortran index boxes {i0,i1,j0,j1} of what is sent (from the computational
 * domain) and what is received (into the halo), with the symmetric-memory rule that the
 * shared edge of staggered fields is never overwritten.  stagger 0=h,1=u,2=v,3=q;
 * halo<0 = full halo width of the (wide ? barotropic : G) memory domain. */
int mom6cu_halo_plan(const mom6cu_domain* dom, int stagger, int wide, int halo, int dir,
                     int* send_box, int* recv_box);

/* ------------------------------------------------ barotropic substep loop  */
/* btstep_timeloop, src/core/MOM_barotropic.F90:2175-2832 (private to the
 * module in the reference; exposed here because it is the BASELINE.json
 * "btstep microbench" unit and the inner stage of mom6cu_btstep).
 * Frozen options: no OBCs, no dynamic_psurf, no linear_wave_drag, no
 * clip_velocity, INTEGRAL_BT_CONTINUITY=False, evolving face areas off. */
typedef struct mom6cu_bt_timeloop_args {
  /* state, wide arrays, inout (MOM_barotropic.F90:2188-2193) */
  double* eta; /* h-points wide */
  double* ubt; /* u-points wide */
  double* vbt; /* v-points wide */
  /* transports closure (:2194-2207).  BTCL_* are arrays of the derived types
   * local_BT_cont_u_type / _v_type (:367-416): 10 reals per point in
   * declaration order {FA_EE,FA_E0,FA_W0,FA_WW,uBT_WW,uBT_EE,uh_crvW,uh_crvE,uh_WW,uh_EE}
   * (v: {NN,N0,S0,SS,vBT_SS,vBT_NN,crvS,crvN,vh_SS,vh_NN}); used when use_BT_cont. */
  const double* uhbt0;
  const double* vhbt0;
  const double* Datu; /* used when !use_BT_cont */
  const double* Datv;
  const double* BTCL_u;
  const double* BTCL_v;
  const double* eta_src; /* :2215 */
  const double* eta_PF;  /* :2272 */
  const double* gtot_E;
  const double* gtot_W;
  const double* gtot_N;
  const double* gtot_S;
  const double* f_4_u; /* (4,SZIBW,SZJW) :2232 */
  const double* f_4_v; /* (4,SZIW,SZJBW) :2239 */
  const double* bt_rem_u;
  const double* bt_rem_v;
  const double* BT_force_u;
  const double* BT_force_v;
  const double* Cor_ref_u;
  const double* Cor_ref_v;
  /* control-structure arrays (barotropic_CS, :144-166), wide */
  const double* IareaT_OBCmask; /* h */
  const double* IdxCu;          /* u */
  const double* IdyCv;          /* v */
  /* accumulators */
  double* u_accel_bt; /* wide u, inout :2226 */
  double* v_accel_bt; /* wide v, inout */
  double* eta_sum;    /* wide h, out (only if find_etaav) :2302 */
  double* eta_wtd;    /* wide h, out :2304 */
  double* ubtav;      /* G u-points, out (CS%ubtav :125) */
  double* vbtav;      /* G v-points, out */
  double* uhbtav;     /* G u-points, out :2220 */
  double* vhbtav;     /* G v-points, out */
  double* ubt_wtd;    /* G u-points, out :2306 */
  double* vbt_wtd;    /* G v-points, out */
  /* per-step weights, host arrays (:2331-2345) */
  const double* wt_vel;    /* nstep+nfilter   */
  const double* wt_eta;    /* nstep+nfilter   */
  const double* wt_accel;  /* nstep+nfilter+1 */
  const double* wt_trans;  /* nstep+nfilter+1 */
  const double* wt_accel2; /* nstep+nfilter+1 */
  /* scalars */
  double dtbt, dgeo_de, bebt, vel_underflow;
  int nstep, nfilter;
  int use_BT_cont, find_etaav, BT_project_velocity, use_old_coriolis_bracket_bug;
  int use_wide_halos, min_stencil;
} mom6cu_bt_timeloop_args;

int mom6cu_btstep_timeloop(mom6cu_ctx* ctx, const mom6cu_bt_timeloop_args* a);

/* ------------------------------------------------------------------ btstep */
/* barotropic_CS, src/core/MOM_barotropic.F90:112-364, as resolved by barotropic_init (:5301) and kept current by
 * btcalc / bt_mass_source / set_dtbt.  "wide" members have the BT_Domain extents (isdw..iedw), the others G's.
 * Frozen options (rejected through `unsupported` or at call time): OBCs, tides/SAL, dynamic_psurf,
 * linear_wave_drag, streaming filter / frequency-dependent drag, INTEGRAL_BT_CONTINUITY, nonlinear continuity
 * (USE_BT_CONT_TYPE=False), NONLIN_BT_STRESS, gradual_BT_ICs, eta_PF_start interpolation, answer_date < 20190101. */
typedef struct mom6cu_barotropic_cs {
  int Sadourny, BT_project_velocity, strong_drag, bound_BT_corr, BT_cont_bounds, wt_uv_bug, visc_rem_u_uh0,
      adjust_BT_cont /* refused when set: ADJUST_BT_CONT is outside the frozen option set */, use_wide_halos, min_stencil,
      use_old_coriolis_bracket_bug, unsupported;
  double dtbt, bebt, vel_underflow, maxCFL_BT_cont, G_extra, dt_bt_filter;
  /* wide */
  const double *IareaT, *IareaT_OBCmask, *bathyT, *IdxCu, *IdyCv; /* h, h, h, u, v */
  const double *q_D, *D_u_Cor, *D_v_Cor;                          /* q, u, v (LINEARIZED_BT_CORIOLIS) */
  const double *ua_polarity, *va_polarity;                        /* h */
  const double *OBCmask_u, *OBCmask_v;                            /* u, v (all ones without OBCs) */
  /* G-sized */
  const double *frhatu, *frhatv;   /* 3-D u, v: set by btcalc */
  double *eta_cor;                 /* h, inout: set by bt_mass_source, bounded here (:1552-1585) */
  const double *eta_cor_bound;     /* h, used when bound_BT_corr && !BT_cont_bounds */
  const double *IDatu, *IDatv;     /* u, v */
  double *ubtav, *vbtav;           /* u, v (out) */
} mom6cu_barotropic_cs;

/* btstep(U_in, V_in, eta_in, dt, bc_accel_u, bc_accel_v, forces, pbce, eta_PF_in, U_Cor, V_Cor, accel_layer_u,
 *        accel_layer_v, eta_out, uhbtav, vhbtav, G, GV, US, CS, visc_rem_u, visc_rem_v, SpV_avg, ADp, OBC, BT_cont,
 *        eta_PF_start, taux_bot, tauy_bot, uh0, vh0, u_uh0, v_vh0, etaav)       MOM_barotropic.F90:455-2172
 * Pointer optionals are NULL when not associated / absent; BT_cont is required (USE_BT_CONT_TYPE=True). */
typedef struct mom6cu_btstep_args {
  const double *U_in, *V_in;             /* 3-D u, v */
  const double *eta_in;                  /* 2-D h */
  double dt;
  const double *bc_accel_u, *bc_accel_v; /* 3-D u, v */
  const double *taux, *tauy;             /* forces%taux, forces%tauy: 2-D u, v */
  const double *pbce;                    /* 3-D h */
  const double *eta_PF_in;               /* 2-D h */
  const double *U_Cor, *V_Cor;           /* 3-D u, v */
  double *accel_layer_u, *accel_layer_v; /* 3-D u, v (out) */
  double *eta_out;                       /* 2-D h (out; may alias eta_in) */
  double *uhbtav, *vhbtav;               /* 2-D u, v (out) */
  const double *visc_rem_u, *visc_rem_v; /* 3-D u, v */
  const mom6cu_bt_cont* BT_cont;
  const double *taux_bot, *tauy_bot;     /* 2-D u, v, optional */
  const double *uh0, *vh0, *u_uh0, *v_vh0; /* 3-D, optional (all or none) */
  double *etaav;                         /* 2-D h, optional out */
} mom6cu_btstep_args;
int mom6cu_btstep(mom6cu_ctx* ctx, const mom6cu_barotropic_cs* CS, const mom6cu_btstep_args* a);

/* btcalc(h, G, GV, CS, h_u, h_v, may_use_default, OBC)  MOM_barotropic.F90:4360-4605: the layer weights
 * frhatu/frhatv (3-D u, v; out) from h or, when given, from BT_cont%h_u, h_v.  hvel_scheme: 1 HARMONIC,
 * 2 ARITHMETIC, 3 HYBRID, 4 FROM_BT_CONT (:432-435). */
typedef struct mom6cu_btcalc_args {
  const double *h, *h_u, *h_v; /* 3-D h; optional 3-D u, v */
  double *frhatu, *frhatv;     /* 3-D u, v (out) */
  const double *bathyT;        /* G-sized h (G%bathyT) */
  int hvel_scheme, may_use_default;
} mom6cu_btcalc_args;
int mom6cu_btcalc(mom6cu_ctx* ctx, const mom6cu_btcalc_args* a);

/* bt_mass_source(h, eta, set_cor, G, GV, CS)  MOM_barotropic.F90:5243-5296; eta_cor 2-D h (inout),
 * eta_mass_source: CS%eta_source (2-D h) or NULL. */
int mom6cu_bt_mass_source(mom6cu_ctx* ctx, const double* h, const double* eta, int set_cor, double* eta_cor);

/* set_dtbt(G, GV, US, CS, pbce, gtot_est, BT_cont, eta, SSH_add)  MOM_barotropic.F90:3509-3633: the stable barotropic
 * time step.  The barotropic_CS members it reads are explicit; CS%dy_Cu / dx_Cv are G's.  No SAL (det_de = 0).
 * Returns CS%dtbt and CS%dtbt_max (min_across_PEs included). */
typedef struct mom6cu_set_dtbt_args {
  const double* pbce;            /* 3-D h, or NULL when gtot_est is given */
  double gtot_est;
  int have_gtot_est;
  const mom6cu_bt_cont* BT_cont; /* or NULL */
  const double* eta;             /* 2-D h, or NULL */
  double SSH_add;
  const double *frhatu, *frhatv; /* CS%frhatu, frhatv (3-D u, v) */
  const double* bathyT;          /* CS%bathyT on G's memory domain */
  double bebt, G_extra, dtbt_fraction, BT_Coriolis_scale, Z_ref;
  int Nonlinear_continuity;
} mom6cu_set_dtbt_args;
int mom6cu_set_dtbt(mom6cu_ctx* ctx, const mom6cu_set_dtbt_args* a, double* dtbt, double* dtbt_max);

/* Microbenchmark form: upload once (state + coefficients stay resident in
 * HBM), run the substep loop `reps` times from the same initial state, and
 * return the device time of the LAST repetition through mom6cu_last_kernel_ms.
 * No host<->device traffic happens inside the repetitions. */
int mom6cu_btstep_timeloop_resident(mom6cu_ctx* ctx, const mom6cu_bt_timeloop_args* a,
                                    int reps, int download);

/* ------------------------------------------------- step_MOM_dyn_split_RK2 */
/* MOM_dyn_split_RK2_CS members the step reads / updates (src/core/MOM_dynamics_split_RK2.F90:85-273).  Every array is
 * in/out and may be a resident plane (mom6cu_plane_alloc) -- the intended use: the control structure lives on the
 * device between steps -- or a host array (staged in and out by the call).  The stage control structures are the ones
 * given to mom6cu_set_cs_{continuity,coriolisadv,hor_visc,pressureforce,vertvisc}.
 * Frozen: no OBCs, no Stokes / fpmix, no dynamic surface pressure (p_surf_begin/end absent), BT_USE_LAYER_FLUXES=True,
 * BT_cont associated with h_u/h_v allocated (BT_THICK_SCHEME=FROM_BT_CONT), set_viscous_ML a no-op
 * (DYNAMIC_VISCOUS_ML=False), no diagnostics.  With calc_dtbt the step calls set_dtbt (:665-669) and stores the new
 * barotropic%dtbt. */
typedef struct mom6cu_dyn_split_rk2_cs {
  double be, begw;
  int split_bottom_stress, store_CAu, CAu_pred_stored /* updated */, visc_rem_dt_bug, hvel_scheme /* barotropic CS%hvel_scheme */,
      unsupported;
  int dtbt_use_bt_cont, BT_Nonlinear_continuity;           /* CS%dtbt_use_bt_cont; barotropic CS%Nonlinear_continuity */
  double dtbt_fraction, BT_Coriolis_scale, Z_ref, dtbt_max; /* barotropic CS%dtbt_fraction, BT_Coriolis_scale; G%Z_ref; dtbt_max out */
  double *CAu, *CAv, *CAu_pred, *CAv_pred, *PFu, *PFv, *diffu, *diffv; /* 3-D u / v */
  double *visc_rem_u, *visc_rem_v, *u_accel_bt, *v_accel_bt, *u_av, *v_av; /* 3-D u / v */
  double *h_av, *pbce;                                                  /* 3-D h */
  double *eta, *eta_PF;                                                 /* 2-D h */
  double *uhbt, *vhbt, *taux_bot, *tauy_bot;                            /* 2-D u / v */
  mom6cu_bt_cont* BT_cont;
  mom6cu_barotropic_cs* barotropic; /* dtbt is updated when calc_dtbt */
} mom6cu_dyn_split_rk2_cs;
/* step_MOM_dyn_split_RK2(u_inst, v_inst, h, tv, visc, Time_local, dt, forces, p_surf_begin, p_surf_end, uh, vh, uhtr,
 *   vhtr, eta_av, G, GV, US, CS, calc_dtbt, VarMix, MEKE, thickness_diffuse_CSp, pbv, STOCH, Waves)   :294-296 */
typedef struct mom6cu_step_dyn_args {
  double *u_inst, *v_inst, *h;  /* 3-D, in/out */
  const double *T, *S;          /* tv%T, tv%S (NULL with EOS_NONE) */
  const double *Kv_bbl_u, *Kv_bbl_v, *bbl_thick_u, *bbl_thick_v, *Kv_shear, *Kv_shear_Bu, *Ray_u, *Ray_v; /* visc% */
  const double *taux, *tauy, *ustar, *p_surf;  /* forces% (p_surf optional) */
  double dt;
  double *uh, *vh;      /* 3-D u / v, in/out */
  double *uhtr, *vhtr;  /* 3-D u / v, in/out */
  double *eta_av;       /* 2-D h, out */
  int calc_dtbt;
} mom6cu_step_dyn_args;
int mom6cu_step_dyn_split_rk2(mom6cu_ctx* ctx, mom6cu_dyn_split_rk2_cs* CS, const mom6cu_step_dyn_args* a);
/* remap_dyn_split_RK2_aux_vars(G, GV, CS, h_old_u, h_old_v, h_new_u, h_new_v, ALE_CSp)  MOM_dynamics_split_RK2.F90:1302-1331
 * (CS%remap_aux true): with store_CAu, u_av/v_av and CAu_pred/CAv_pred are remapped and their halos updated; diffu/diffv
 * are remapped.  remapCS is ALE_CSp%vel_remapCS. */
int mom6cu_remap_dyn_split_rk2_aux_vars(mom6cu_ctx* ctx, const mom6cu_remapping_cs* remapCS, const mom6cu_dyn_split_rk2_cs* CS,
                                        const double* h_old_u, const double* h_old_v, const double* h_new_u, const double* h_new_v);

/* ------------------------------------ reproducing sums, checksums, write_energy (the parity metric) */
/* EFP_type, src/framework/MOM_coms.F90:76-78: ni = 6 integers of 46 bits each (prec = 2**46, :30-40). */
typedef struct mom6cu_efp { int64_t v[6]; } mom6cu_efp;
/* EFP_plus :775, EFP_minus :786, EFP_to_real :813 (regularises a in place, as the reference does), real_to_EFP :836,
 * EFP_real_diff :822.  Host arithmetic on 6 integers: no device work. */
void mom6cu_efp_plus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, int* overflow);
void mom6cu_efp_minus(const mom6cu_efp* a, const mom6cu_efp* b, mom6cu_efp* out, int* overflow);
double mom6cu_efp_to_real(mom6cu_efp* a);
int mom6cu_real_to_efp(double val, mom6cu_efp* out); /* returns 1 on overflow, 2 on NaN (the reference's FATALs) */
double mom6cu_efp_real_diff(const mom6cu_efp* a, const mom6cu_efp* b);
/* reproducing_sum(array, isr, ier, jsr, jer, sums, EFP_sum, EFP_lay_sums, err, only_on_PE, unscale)
 *   reproducing_sum_2d :227-325 (nk = 1), reproducing_sum_3d :337-545, reproducing_EFP_sum_2d :101-222.
 * array: a field of the given stagger (nk levels) on G's memory domain, host array or resident plane.  isr..jer are the
 * reference's 1-based positions inside the array (0 = absent: the whole array).  sums / EFP_lay_sums (nk entries) and EFP_sum
 * may be NULL; with sums or EFP_lay_sums the result is the k-ordered floating sum of the layer sums (:470-476), otherwise the
 * conversion of the one extended-fixed-point total (:528-529).  unscale = 1 means absent.  The order-invariant integer sums
 * are accumulated on the device (conversion :640-648 per element, integer atomics), summed across ranks with NCCL
 * (sum_across_PEs) unless only_on_PE, and regularised on the host (:717-754).  NaN or overflow: FATAL as in the reference. */
int mom6cu_reproducing_sum(mom6cu_ctx* ctx, const double* array, int stagger, int nk, int isr, int ier, int jsr, int jer,
                           double unscale, int only_on_PE, double* sum, double* sums, mom6cu_efp* EFP_sum, mom6cu_efp* EFP_lay_sums);
/* EFP_sum_across_PEs(EFPs, nval) :870-916: in place, overflows carried. */
int mom6cu_efp_sum_across_pes(mom6cu_ctx* ctx, mom6cu_efp* EFPs, int nval);

/* hchksum / uvchksum / Bchksum, src/framework/MOM_checksums.F90: chksum_h_2d :387, chksum_h_3d :1413, chksum_u_2d :1005,
 * chksum_u_3d :1782, chksum_v_2d :1209, chksum_v_3d :1986, chksum_B_2d :688, chksum_B_3d :1586; bitcount :2678.
 * stagger 0=h,1=u,2=v,3=q; nk = 1 for the 2-D forms.  haloshift < 0 means the full halo (:470); symmetric / omit_corners as
 * the optional arguments (0 = absent); scale = 1 means absent.  nk > 1 selects the _3d form, which matters at B points:
 * chksum_B_3d shifts its SW / SE / NW windows by haloshift+1 with or without `symmetric` (:1698-1706) and, with omit_corners and
 * `symmetric`, its S and W windows too (:1712-1718); chksum_B_2d does neither (:797-809).  Both are reproduced as they are.
 *   bc[0] = bc0; then, in the order the reference prints them:
 *     kind 1 (haloshift 0, not symmetric): nothing else           (chk_sum_msg1)
 *     kind 2 (corners):  bc[1..4] = SW, SE, NW, NE                 (chk_sum_msg5)
 *     kind 3 (NSEW):     bc[1..4] = N, S, E, W                     (chk_sum_msg_NSEW)
 *     kind 4 (u symmetric, haloshift 0): bc[1] = W                 (chk_sum_msg_W)
 *     kind 5 (v symmetric, haloshift 0): bc[1] = S                 (chk_sum_msg_S)
 *   (at B points `symmetric` with haloshift 0 takes the corner or NSEW form: chk_sum_msg2 is never called by the reference)
 *   *kind receives the case; stats (NULL = calculateStatistics off) receives mean, min, max (subStats). */
int mom6cu_chksum(mom6cu_ctx* ctx, const double* array, int stagger, int nk, int haloshift, int symmetric, int omit_corners,
                  double scale, int* bc, int* kind, double* stats);

/* write_energy, src/diagnostics/MOM_sum_output.F90:321-1020 (Boussinesq; the quantities of one ocean.stats line and of
 * the energy file).  Sum_output_CS members :66-140 as resolved by MOM_sum_output_init :147 and depth_list_setup :1161. */
typedef struct mom6cu_sum_output_cs {
  int do_APE_calc, use_temperature;
  double dt_in_T;                       /* CS%dt_in_T */
  int DL_listsize;                      /* CS%DL%listsize */
  const double *DL_depth, *DL_area, *DL_vol_below; /* CS%DL (host, listsize entries) */
  int* lH;                              /* CS%lH(nk), host, in/out (1-based list positions) */
  const double* g_prime;                /* GV%g_prime(nk+1), host */
  double Z_ref, C_p;                    /* G%Z_ref, tv%C_p */
  /* unit_scale_type factors write_energy uses (all 1 in an unscaled run) */
  double RZL2_to_kg, L_T_to_m_s, Q_to_J_kg, J_kg_to_Q, kg_m3_to_R, m_to_Z, m_to_L, Z_to_m, S_to_ppt, C_to_degC;
  int previous_calls, ntrunc;           /* in/out */
  mom6cu_efp fresh_water_in_EFP, net_salt_in_EFP, net_heat_in_EFP, mass_prev_EFP, salt_prev_EFP, heat_prev_EFP; /* in/out */
} mom6cu_sum_output_cs;
typedef struct mom6cu_energy_out {
  double En_mass, toten, KE_tot, PE_tot, mass_tot, mass_chg, mass_anom, max_CFL[2];
  double Salt, Salt_chg, Salt_anom, Heat, Heat_chg, Heat_anom, salin, salin_anom, temp, temp_anom;
  int ntrunc;
  double *KE, *mass_lay;      /* nk, host, may be NULL */
  double *PE, *Z_0APE;        /* nk+1, host, may be NULL */
} mom6cu_energy_out;
/* u, v, h (3-D), T, S (tv%T, tv%S; NULL unless use_temperature): host arrays or resident planes.  Two passes over the
 * state on the device (mass / KE / heat / salt / CFL, then -- with do_APE_calc -- the interface APE, which needs the
 * layer volumes of the first pass); what returns to the host is 6 integers per sum. */
int mom6cu_write_energy(mom6cu_ctx* ctx, mom6cu_sum_output_cs* CS, const double* u, const double* v, const double* h,
                        const double* T, const double* S, mom6cu_energy_out* out);
/* The line write_energy appends to ocean.stats (:874-902; day-stamped form), NUL terminated, without the newline. */
int mom6cu_ocean_stats_line(const mom6cu_sum_output_cs* CS, const mom6cu_energy_out* e, int n, double reday, char* buf, size_t len);

/* ---------------------------------------------------- ALE_regridding_and_remapping (the thermodynamic-cadence pass) */
/* interpolate_column(nsrc, h_src, u_src, ndest, h_dest, u_dest, mask_edges)  src/ALE/MOM_remapping.F90:1247-1314 for ncol
 * contiguous columns (h_src: ncol x nsrc, u_src: ncol x (nsrc+1), ...): the form the reference's unit tests call. */
int mom6cu_interpolate_column(mom6cu_ctx* ctx, int ncol, int nsrc, const double* h_src, const double* u_src, int ndest,
                              const double* h_dest, double* u_dest, int mask_edges);
/* ALE_remap_interface_vals(CS, G, GV, h_old, h_new, int_val)  src/ALE/MOM_ALE.F90:1303-1339 (int_val: h points, nk+1 levels)
 * and ALE_remap_vertex_vals :1342-1382 (vert_val: q points, nk+1 levels). */
int mom6cu_ale_remap_interface_vals(mom6cu_ctx* ctx, const double* h_old, const double* h_new, double* int_val);
int mom6cu_ale_remap_vertex_vals(mom6cu_ctx* ctx, const double* h_old, const double* h_new, double* vert_val);
/* ALE_CS members the pass uses (src/ALE/MOM_ALE.F90:65-130) and MOM_control_struct%remap_aux_vars. */
typedef struct mom6cu_ale_cs {
  mom6cu_regridding_cs regridCS;          /* old_grid_weight is updated (ALE_update_regrid_weights :1719) */
  mom6cu_remapping_cs remapCS, vel_remapCS;
  double regrid_time_scale;
  int remap_uv_using_old_alg, do_conv_adj, use_hybgen_unmix; /* must be 0 (frozen option set) */
  int remap_aux_vars;
} mom6cu_ale_cs;
typedef struct mom6cu_ale_args {
  double *u, *v, *h;               /* 3-D, in/out */
  int ntr;                         /* CS%tracer_Reg%ntr */
  double* const* tr;               /* Reg%Tr(m)%t, 3-D h, in/out */
  const double* conc_underflow;    /* (ntr) or NULL */
  int iT, iS;                      /* which tracers tv%T / tv%S point at (-1: not associated) */
  double dtdia;
  double *Kd_shear, *Kv_shear;     /* visc%Kd_shear / Kv_shear: h points, nk+1 levels, in/out; NULL = not associated */
  double* Kv_shear_Bu;             /* visc%Kv_shear_Bu: q points, nk+1 levels */
} mom6cu_ale_args;
/* ALE_regridding_and_remapping(CS, G, GV, US, u, v, h, tv, dtdia, Time_end_thermo)  src/core/MOM.F90:1751-1926 without OBCs,
 * ice shelves, particles or diagnostics: halo update of T, S, h; regrid weights; ALE_regrid; ALE_remap_tracers;
 * ALE_remap_set_h_vel on both grids; ALE_remap_velocities; with remap_aux_vars remap_dyn_split_RK2_aux_vars (dynCS) and
 * remap_vertvisc_aux_vars (MOM_set_viscosity.F90:2849) + the halo update of Kv_shear; h = h_new on (isc-1:iec+1, jsc-1:jec+1). */
int mom6cu_ale_regridding_and_remapping(mom6cu_ctx* ctx, mom6cu_ale_cs* CS, const mom6cu_dyn_split_rk2_cs* dynCS,
                                        const mom6cu_ale_args* a);

/* ------------------------------------------------------------- mixedlayer_restrat (first caller) */
/* mixedlayer_restrat_CS, src/parameterizations/lateral/MOM_mixed_layer_restrat.F90:42-115, the members of the Fox-Kemper
 * et al. (2008) general-coordinate path mixedlayer_restrat_OM4 (:189-714), and the equation of state it evaluates at the
 * surface pressure (tv%eqn_of_state: LINEAR or WRIGHT, MOM6CU_EOS_*).  Frozen: Boussinesq, the mixed-layer depth from MLE_USE_PBL_MLD
 * or detected with MLE_DENSITY_DIFF > 0 (detect_mld :1503), no Stanley SGS variance, no Bodner / bulk-mixed-layer variants, constant front length (MLE_FRONT_LENGTH >= 0, not from
 * a file), MLE_TAIL_DH = 0 in the full routine (mu's exponent 1 + 2 dh is a real power otherwise; mom6cu_mle_mu accepts any dh). */
typedef struct mom6cu_mle_cs {
  double ml_restrat_coef, ml_restrat_coef2, front_length, MLE_MLD_decay_time, MLE_MLD_decay_time2, MLE_MLD_stretch, MLE_tail_dh,
      ustar_min, vonKar, MLE_density_diff;
  int MLE_use_PBL_MLD, use_Stanley_ML, use_Bodner, fl_from_file;
  int EOS_form;
  double Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp;
  double *MLD_filtered, *MLD_filtered_slow; /* CS%MLD_filtered, CS%MLD_filtered_slow: 2-D h, in/out (slow may be NULL if decay_time2 <= 0) */
} mom6cu_mle_cs;
/* mixedlayer_restrat(h, uhtr, vhtr, tv, forces, dt, MLD, h_MLD, bflux, VarMix, G, GV, US, CS)  :149-186 -> mixedlayer_restrat_OM4.
 * h, uhtr, vhtr: 3-D in/out; T, S: tv%T, tv%S; ustar: forces%ustar (2-D h); h_MLD: visc%h_ML (2-D h); Rd_dx_h: VarMix%Rd_dx_h
 * (2-D h; NULL allowed when front_length == 0). */
int mom6cu_mixedlayer_restrat(mom6cu_ctx* ctx, mom6cu_mle_cs* CS, double* h, double* uhtr, double* vhtr, const double* T,
                              const double* S, const double* ustar, double dt, const double* h_MLD, const double* Rd_dx_h);
/* mu(sigma, dh) :717-751, the shape function of the restratifying streamfunction, for n values (host or device arrays):
 * the form the reference's unit tests call (mixedlayer_restrat_unit_tests :2014-2058). */
int mom6cu_mle_mu(mom6cu_ctx* ctx, int n, const double* sigma, const double* dh, double* out);

/* ----------------------------------------------------------------- tracer_hordiff (the tracer-step caller) */
/* tracer_hor_diff_CS, src/tracer/MOM_tracer_hor_diff.F90:40-106, and the VarMix switches tracer_hordiff reads (:163-169).  Frozen: the
 * along-surface path (:537-604; USE_NEUTRAL_DIFFUSION, USE_HORIZONTAL_BOUNDARY_DIFFUSION and DIFFUSE_ML_TO_INTERIOR off), online
 * diffusivities (:203-340: constant KhTr, or with VarMix the Eady growth-rate term KhTr_Slope_Cff*L2u*SN_u, MEKE%KhTr_fac*sqrt(Kh Kh) and
 * the KhTr_max / resolution-function / KhTr_min / passivity chain), the MAX_TR_DIFFUSION_CFL limit and the CHECK_DIFFUSIVE_CFL iteration
 * count. */
typedef struct mom6cu_tracer_hor_diff_cs {
  double KhTr, KhTr_min, KhTr_max, KhTr_passivity_coeff, KhTr_passivity_min, KhTr_Slope_Cff, max_diff_CFL;
  int check_diffusive_CFL, use_neutral_diffusion, use_hor_bnd_diffusion, Diffuse_ML_interior;
  int use_variable_mixing, Resoln_scaled_KhTr, use_MEKE_Kh; /* VarMix%use_variable_mixing, VarMix%Resoln_scaled_KhTr, allocated(MEKE%Kh) */
  double MEKE_KhTr_fac;                                     /* MEKE%KhTr_fac */
} mom6cu_tracer_hor_diff_cs;
/* tracer_hordiff(h, dt, MEKE, VarMix, visc, G, GV, US, CS, Reg, tv)  :119: h 3-D; tr: Reg%Tr(m)%t, ntr 3-D h fields, in/out (the
 * routine updates their halos itself); conc_underflow: ntr or NULL; Res_fn_h, Rd_dx_h: VarMix 2-D h fields (NULL unless used);
 * df_x / df_y: Reg%Tr(m)%df_x / df_y, ntr optional 3-D u / v diagnostics (NULL array or NULL entries = not associated). */
typedef struct mom6cu_tracer_hordiff_args {
  const double* h;
  double dt;
  int ntr;
  double* const* tr;
  const double* conc_underflow;
  const double *Res_fn_h, *Rd_dx_h;
  double* const* df_x;
  double* const* df_y;
  const double *L2u, *SN_u, *L2v, *SN_v; /* VarMix%L2u, SN_u (2-D u), L2v, SN_v (2-D v): used when KhTr_Slope_Cff > 0 */
  const double* MEKE_Kh;                 /* MEKE%Kh (2-D h): used when use_MEKE_Kh */
} mom6cu_tracer_hordiff_args;
int mom6cu_tracer_hordiff(mom6cu_ctx* ctx, const mom6cu_tracer_hor_diff_cs* CS, const mom6cu_tracer_hordiff_args* a);

/* -------------------------------------------------------- thickness_diffuse (the last of the three callers) */
/* thickness_diffuse_CS, src/parameterizations/lateral/MOM_thickness_diffuse.F90:40-131, plus the VarMix / MEKE switches the routine reads
 * and the equation of state of tv.  Frozen: the density-gradient path of thickness_diffuse_full (:635-1670) with an equation of state
 * (LINEAR or WRIGHT), slopes computed here (no USE_STORED_SLOPES), the limited streamfunction of :1138-1160 (no
 * KHTH_USE_FGNV_STREAMFUNCTION), constant KHTH with the KHTH_MIN / KHTH_MAX / KHTH_MAX_CFL limits and the VarMix resolution function
 * (no Visbeck / QG-Leith / vertical structure / depth scaling), no interface-height diffusivity (KH_ETA_*), no detangling, no
 * Stanley SGS variance, no GM work diagnostics (MEKE%GM_src, CS%GMwork not allocated), Boussinesq, GV%nkml = 0.  Also supported (the
 * OM4-style selection): USE_STORED_SLOPES (VarMix%slope_x / slope_y as inputs), KHTH_USE_FGNV_STREAMFUNCTION (the elliptic
 * streamfunction of Ferrari et al. 2010, :1103-1122 + streamfn_solver :1674; VarMix%cg1 as input) and the MEKE diffusivity
 * MEKE%KhTh_fac*sqrt(MEKE%Kh(i)*MEKE%Kh(i+1)) (:281-284; MEKE%Kh as input, not MEKE_GEOMETRIC). */
typedef struct mom6cu_thickness_diffuse_cs {
  double Khth, Khth_Min, Khth_Max, max_Khth_CFL, slope_max, kappa_smooth;
  double dZ_subroundoff; /* GV%dZ_subroundoff */
  int thickness_diffuse, read_khth, detangle_interfaces, interface_Kh /* Kh_eta_bg > 0 or Kh_eta_vel > 0 */, use_FGNV_streamfn, use_stanley_gm,
      use_GME_thickness_diffuse, find_work /* allocated(MEKE%GM_src) or allocated(CS%GMwork) or skeb_use_gm */;
  int use_variable_mixing, Resoln_scaled_KhTh, Depth_scaled_KhTh, use_stored_slopes, use_Visbeck, use_QG_Leith_GM, khth_struct, use_MEKE_Kh;
  int EOS_form; /* MOM6CU_EOS_LINEAR | MOM6CU_EOS_WRIGHT */
  double Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp;
  double FGNV_scale, N2_floor; /* FGNV_FILTER_SCALE; (FGNV_STRAT_FLOOR*OMEGA)**2  (:2314-2341) */
  double MEKE_KhTh_fac;        /* MEKE%KhTh_fac */
} mom6cu_thickness_diffuse_cs;
/* thickness_diffuse(h, uhtr, vhtr, tv, dt, G, GV, US, MEKE, VarMix, CDp, CS, STOCH)  :134: h, uhtr, vhtr 3-D in/out (h valid one halo point
 * out); T, S: tv%T, tv%S (3-D h, one halo point); p_surf: tv%p_surf (2-D h) or NULL; Res_fn_u / Res_fn_v: VarMix 2-D u / v (NULL unless
 * Resoln_scaled_KhTh); uhGM / vhGM: CDp%uhGM / vhGM, optional 3-D u / v out. */
typedef struct mom6cu_thickness_diffuse_args {
  double *h, *uhtr, *vhtr;
  const double *T, *S, *p_surf;
  double dt;
  const double *Res_fn_u, *Res_fn_v;
  double *uhGM, *vhGM;
  const double *slope_x, *slope_y; /* VarMix%slope_x / slope_y: 3-D u / v with nk+1 interfaces (use_stored_slopes) */
  const double *cg1;               /* VarMix%cg1: 2-D h (use_FGNV_streamfn) */
  const double *MEKE_Kh;           /* MEKE%Kh: 2-D h (use_MEKE_Kh) */
} mom6cu_thickness_diffuse_args;
int mom6cu_thickness_diffuse(mom6cu_ctx* ctx, const mom6cu_thickness_diffuse_cs* CS, const mom6cu_thickness_diffuse_args* a);

#ifdef __cplusplus
}
#endif
#endif /* MOM6CU_H */
