!> ISO_C_BINDING interfaces to the B200-native dycore library (include/mom6cu.h, libmom6cu.so).
!!
!! This module is what a MOM6 maintainer adds to the source tree, together with the shadow copies of the modules on the hot path that
!! fortran/install_shims.py writes (the reference's own files with a hook at the top of each replaced procedure and the binding procedures
!! of fortran/bodies/*.inc appended), in a directory listed ahead of src/ -- the build-time directory swap MOM6 already uses for
!! config_src/infra/FMS1|FMS2 and config_src/external/* (ac/configure.ac:227-264).  See INTEGRATION.md.
!!
!! NOTE: no Fortran compiler exists in the build container, so this file has not been compiled there.  What a compiler would check is
!! checked by parsers instead: tests/test_abi_layout.py (every bind(C) type below == its struct in include/mom6cu.h, member for member; every
!! C entry has an interface; sizes against mom6cu_sizeof) and tests/test_fortran_shims.py (every member the bindings touch exists).
!! Only standard Fortran 2003 interoperability is used (bind(C), c_ptr, c_loc on contiguous targets).
module mom6cu_interface
  use, intrinsic :: iso_c_binding
  implicit none ; public

  integer(c_int), parameter :: MOM6CU_OK = 0

  !> mom6cu_domain (include/mom6cu.h): the hor_index_type / barotropic_CS bounds the hot path uses
  type, bind(C) :: mom6cu_domain
    integer(c_int) :: isc, iec, jsc, jec
    integer(c_int) :: isd, ied, jsd, jed
    integer(c_int) :: isdw, iedw, jsdw, jedw
    integer(c_int) :: nk
    integer(c_int) :: cyclic_x, cyclic_y
    integer(c_int) :: first_direction
    integer(c_int) :: npi, npj, pi, pj
  end type mom6cu_domain

  !> mom6cu_grid: pointers to the ocean_grid_type metrics (src/core/MOM_grid.F90:75-175)
  type, bind(C) :: mom6cu_grid
    type(c_ptr) :: mask2dT, mask2dCu, mask2dCv, mask2dBu
    type(c_ptr) :: dxT, dyT, IdxT, IdyT, areaT, IareaT
    type(c_ptr) :: dxCu, dyCu, IdxCu, IdyCu, dy_Cu, areaCu, IareaCu
    type(c_ptr) :: dxCv, dyCv, IdxCv, IdyCv, dx_Cv, areaCv, IareaCv
    type(c_ptr) :: dxBu, dyBu, IdxBu, IdyBu, areaBu, IareaBu
    type(c_ptr) :: bathyT, CoriolisBu, Coriolis2Bu
  end type mom6cu_grid

  type, bind(C) :: mom6cu_vgrid
    real(c_double) :: Angstrom_H, H_subroundoff, Z_to_H, H_to_Z, g_Earth, Rho0, H_to_RZ, RZ_to_H, H_to_m, m_to_H
    integer(c_int) :: Boussinesq
  end type mom6cu_vgrid

  !> EFP_type (src/framework/MOM_coms.F90:76-78): the same six 64-bit integers, so an EFP_type argument can be passed as is
  type, bind(C) :: mom6cu_efp
    integer(c_int64_t) :: v(6)
  end type mom6cu_efp

  !> Sum_output_CS members write_energy uses (src/diagnostics/MOM_sum_output.F90:66-140)
  type, bind(C) :: mom6cu_sum_output_cs
    integer(c_int) :: do_APE_calc, use_temperature
    real(c_double) :: dt_in_T
    integer(c_int) :: DL_listsize
    type(c_ptr)    :: DL_depth, DL_area, DL_vol_below, lH, g_prime
    real(c_double) :: Z_ref, C_p
    real(c_double) :: RZL2_to_kg, L_T_to_m_s, Q_to_J_kg, J_kg_to_Q, kg_m3_to_R, m_to_Z, m_to_L, Z_to_m, S_to_ppt, C_to_degC
    integer(c_int) :: previous_calls, ntrunc
    type(mom6cu_efp) :: fresh_water_in_EFP, net_salt_in_EFP, net_heat_in_EFP, mass_prev_EFP, salt_prev_EFP, heat_prev_EFP
  end type mom6cu_sum_output_cs

  !> The quantities write_energy puts on the ocean.stats line and in the energy file
  type, bind(C) :: mom6cu_energy_out
    real(c_double) :: En_mass, toten, KE_tot, PE_tot, mass_tot, mass_chg, mass_anom, max_CFL(2)
    real(c_double) :: Salt, Salt_chg, Salt_anom, Heat, Heat_chg, Heat_anom, salin, salin_anom, temp, temp_anom
    integer(c_int) :: ntrunc
    type(c_ptr)    :: KE, mass_lay, PE, Z_0APE
  end type mom6cu_energy_out

  !> continuity_PPM_CS (src/core/MOM_continuity_PPM.F90:35-67)
  type, bind(C) :: mom6cu_continuity_cs
    integer(c_int) :: upwind_1st, monotonic, simple_2nd, aggress_adjust, vol_CFL, better_iter, &
                      use_visc_rem_max, marginal_faces
    real(c_double) :: tol_eta, tol_vel, CFL_limit_adjust
  end type mom6cu_continuity_cs

  !> BT_cont_type (src/core/MOM_variables.F90:315-350)
  type, bind(C) :: mom6cu_bt_cont
    type(c_ptr) :: FA_u_EE, FA_u_E0, FA_u_W0, FA_u_WW, uBT_WW, uBT_EE
    type(c_ptr) :: FA_v_NN, FA_v_N0, FA_v_S0, FA_v_SS, vBT_SS, vBT_NN
    type(c_ptr) :: h_u, h_v
  end type mom6cu_bt_cont

  !> Arguments of continuity_PPM (src/core/MOM_continuity_PPM.F90:86-141); absent optionals are c_null_ptr
  type, bind(C) :: mom6cu_continuity_args
    type(c_ptr) :: u, v, hin, h, uh, vh
    real(c_double) :: dt
    type(c_ptr) :: por_face_areaU, por_face_areaV, uhbt, vhbt, visc_rem_u, visc_rem_v, u_cor, v_cor
    type(c_ptr) :: BT_cont
    type(c_ptr) :: du_cor, dv_cor
  end type mom6cu_continuity_args

  !> CoriolisAdv_CS (src/core/MOM_CoriolisAdv.F90:30-91)
  type, bind(C) :: mom6cu_coriolisadv_cs
    integer(c_int) :: Coriolis_Scheme, KE_Scheme, PV_Adv_Scheme, no_slip, bound_Coriolis, Coriolis_En_Dis
    real(c_double) :: F_eff_max_blend, wt_lin_blend
  end type mom6cu_coriolisadv_cs

  !> Arguments of CorAdCalc (src/core/MOM_CoriolisAdv.F90:125-144)
  type, bind(C) :: mom6cu_coradcalc_args
    type(c_ptr) :: u, v, h, uh, vh, CAu, CAv, por_face_areaU, por_face_areaV, RV, PV, gradKEu, gradKEv
  end type mom6cu_coradcalc_args

  !> Arguments of horizontal_viscosity (src/parameterizations/lateral/MOM_hor_visc.F90:266-305)
  type, bind(C) :: mom6cu_hor_visc_args
    type(c_ptr) :: u, v, h, uh, vh, diffu, diffv, hu_cont, hv_cont
    real(c_double) :: dt
  end type mom6cu_hor_visc_args

  !> unit_scale_type factors (src/framework/MOM_unit_scaling.F90)
  type, bind(C) :: mom6cu_unit_scale
    real(c_double) :: m_to_L, L_to_m, m_s_to_L_T, L_T_to_m_s, s_to_T, T_to_s, m_to_Z, Z_to_m, Z_to_L, L_to_Z
  end type mom6cu_unit_scale

  !> PressureForce_FV_CS / EOS parameters (src/core/MOM_PressureForce_FV.F90:40-107)
  type, bind(C) :: mom6cu_pressureforce_cs
    integer(c_int) :: EOS_form, MassWghtInterp, use_SSH_in_Z0p, rho_ref_bug, unsupported
    real(c_double) :: rho_ref, GFS_scale, Z_ref, dZ_subroundoff
    real(c_double) :: Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp
    type(c_ptr) :: Rlay, g_prime
    integer(c_int) :: reconstruct, Recon_Scheme, boundary_extrap, use_inaccurate_pgf_rho_anom, MassWghtInterpVanOnly, ALE_answer_date
    real(c_double) :: h_nonvanished
    real(c_double) :: kg_m3_to_R, RL2_T2_to_Pa, C_to_degC, S_to_ppt
  end type mom6cu_pressureforce_cs
  !> Arguments of PressureForce (src/core/MOM_PressureForce.F90:40-61)
  type, bind(C) :: mom6cu_pressureforce_args
    type(c_ptr) :: h, T, S, PFu, PFv, p_atm, pbce, eta
  end type mom6cu_pressureforce_args

  !> vertvisc_CS (src/parameterizations/vertical/MOM_vert_friction.F90:48-170)
  type, bind(C) :: mom6cu_vertvisc_cs
    integer(c_int) :: bottomdraglaw, harmonic_visc, direct_stress, fixed_LOTW_ML, apply_LOTW_floor, dynamic_viscous_ML, &
                      nkml, answer_date, unsupported, CFL_based_trunc
    real(c_double) :: Hbbl, Kv, Kv_extra_bbl, Kvml_invZ2, Hmix, Hmix_stress, harm_BL_val, vonKar, vel_underflow, dZ_subroundoff, &
                      maxvel, CFL_trunc
  end type mom6cu_vertvisc_cs
  type, bind(C) :: mom6cu_vertvisc_coef_args
    type(c_ptr) :: u, v, h, Kv_bbl_u, Kv_bbl_v, bbl_thick_u, bbl_thick_v, Kv_shear, Kv_shear_Bu, ustar
    real(c_double) :: dt
  end type mom6cu_vertvisc_coef_args
  type, bind(C) :: mom6cu_vertvisc_args
    type(c_ptr) :: u, v, h, taux, tauy, Ray_u, Ray_v
    real(c_double) :: dt
    type(c_ptr) :: taux_bot, tauy_bot
  end type mom6cu_vertvisc_args

  !> barotropic_CS members (src/core/MOM_barotropic.F90:112-364) and the arguments of btstep (:455-529)
  type, bind(C) :: mom6cu_barotropic_cs
    integer(c_int) :: Sadourny, BT_project_velocity, strong_drag, bound_BT_corr, BT_cont_bounds, wt_uv_bug, visc_rem_u_uh0, &
                      adjust_BT_cont, use_wide_halos, min_stencil, use_old_coriolis_bracket_bug, unsupported
    real(c_double) :: dtbt, bebt, vel_underflow, maxCFL_BT_cont, G_extra, dt_bt_filter
    type(c_ptr) :: IareaT, IareaT_OBCmask, bathyT, IdxCu, IdyCv, q_D, D_u_Cor, D_v_Cor, ua_polarity, va_polarity, OBCmask_u, OBCmask_v
    type(c_ptr) :: frhatu, frhatv, eta_cor, eta_cor_bound, IDatu, IDatv, ubtav, vbtav
  end type mom6cu_barotropic_cs
  type, bind(C) :: mom6cu_btstep_args
    type(c_ptr) :: U_in, V_in, eta_in
    real(c_double) :: dt
    type(c_ptr) :: bc_accel_u, bc_accel_v, taux, tauy, pbce, eta_PF_in, U_Cor, V_Cor, accel_layer_u, accel_layer_v, eta_out, &
                   uhbtav, vhbtav, visc_rem_u, visc_rem_v, BT_cont, taux_bot, tauy_bot, uh0, vh0, u_uh0, v_vh0, etaav
  end type mom6cu_btstep_args

  !> MOM_dyn_split_RK2_CS members (src/core/MOM_dynamics_split_RK2.F90:85-273) and the arguments of
  !! step_MOM_dyn_split_RK2 (:294-296)
  type, bind(C) :: mom6cu_dyn_split_rk2_cs
    real(c_double) :: be, begw
    integer(c_int) :: split_bottom_stress, store_CAu, CAu_pred_stored, visc_rem_dt_bug, hvel_scheme, unsupported
    integer(c_int) :: dtbt_use_bt_cont, BT_Nonlinear_continuity
    real(c_double) :: dtbt_fraction, BT_Coriolis_scale, Z_ref, dtbt_max
    type(c_ptr) :: CAu, CAv, CAu_pred, CAv_pred, PFu, PFv, diffu, diffv, visc_rem_u, visc_rem_v, u_accel_bt, v_accel_bt, &
                   u_av, v_av, h_av, pbce, eta, eta_PF, uhbt, vhbt, taux_bot, tauy_bot
    type(c_ptr) :: BT_cont, barotropic
  end type mom6cu_dyn_split_rk2_cs
  type, bind(C) :: mom6cu_step_dyn_args
    type(c_ptr) :: u_inst, v_inst, h, T, S, Kv_bbl_u, Kv_bbl_v, bbl_thick_u, bbl_thick_v, Kv_shear, Kv_shear_Bu, Ray_u, Ray_v, &
                   taux, tauy, ustar, p_surf
    real(c_double) :: dt
    type(c_ptr) :: uh, vh, uhtr, vhtr, eta_av
    integer(c_int) :: calc_dtbt
  end type mom6cu_step_dyn_args

  !> tracer_advect_CS and the arguments of advect_tracer (src/tracer/MOM_tracer_advect.F90:32-54)
  type, bind(C) :: mom6cu_tracer_advect_cs
    real(c_double) :: dt
    integer(c_int) :: default_advect_scheme, useHuynhStencilBug
  end type mom6cu_tracer_advect_cs
  type, bind(C) :: mom6cu_advect_tracer_args
    type(c_ptr) :: h_end, uhtr, vhtr
    real(c_double) :: dt
    integer(c_int) :: ntr
    type(c_ptr) :: tr, advect_scheme, conc_underflow   ! tr: array of ntr c_ptr (Reg%Tr(m)%t)
    integer(c_int) :: x_first_in, max_iter_in
    type(c_ptr) :: vol_prev
    integer(c_int) :: update_vol_prev
    type(c_ptr) :: uhr_out, vhr_out
  end type mom6cu_advect_tracer_args

  !> remapping_CS (src/ALE/MOM_remapping.F90:37-85) and regridding_CS (src/ALE/MOM_regridding.F90:49-160)
  !> mixedlayer_restrat_CS (src/parameterizations/lateral/MOM_mixed_layer_restrat.F90:42-115), the mixedlayer_restrat_OM4 members
  type, bind(C) :: mom6cu_mle_cs
    real(c_double) :: ml_restrat_coef, ml_restrat_coef2, front_length, MLE_MLD_decay_time, MLE_MLD_decay_time2, MLE_MLD_stretch, &
                      MLE_tail_dh, ustar_min, vonKar, MLE_density_diff
    integer(c_int) :: MLE_use_PBL_MLD, use_Stanley_ML, use_Bodner, fl_from_file, EOS_form
    real(c_double) :: Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp
    type(c_ptr)    :: MLD_filtered, MLD_filtered_slow
  end type mom6cu_mle_cs
  !> tracer_hor_diff_CS (src/tracer/MOM_tracer_hor_diff.F90:40-106) and the VarMix switches tracer_hordiff reads
  type, bind(C) :: mom6cu_tracer_hor_diff_cs
    real(c_double) :: KhTr, KhTr_min, KhTr_max, KhTr_passivity_coeff, KhTr_passivity_min, KhTr_Slope_Cff, max_diff_CFL
    integer(c_int) :: check_diffusive_CFL, use_neutral_diffusion, use_hor_bnd_diffusion, Diffuse_ML_interior, &
                      use_variable_mixing, Resoln_scaled_KhTr, use_MEKE_Kh
    real(c_double) :: MEKE_KhTr_fac
  end type mom6cu_tracer_hor_diff_cs
  !> the arguments of tracer_hordiff (:119); tr, df_x, df_y are arrays of c_ptr (Reg%Tr(m)%t, %df_x, %df_y)
  type, bind(C) :: mom6cu_tracer_hordiff_args
    type(c_ptr)    :: h
    real(c_double) :: dt
    integer(c_int) :: ntr
    type(c_ptr)    :: tr, conc_underflow, Res_fn_h, Rd_dx_h, df_x, df_y, L2u, SN_u, L2v, SN_v, MEKE_Kh
  end type mom6cu_tracer_hordiff_args
  !> thickness_diffuse_CS (src/parameterizations/lateral/MOM_thickness_diffuse.F90:40-131) + the VarMix / MEKE / EOS switches it reads
  type, bind(C) :: mom6cu_thickness_diffuse_cs
    real(c_double) :: Khth, Khth_Min, Khth_Max, max_Khth_CFL, slope_max, kappa_smooth, dZ_subroundoff
    integer(c_int) :: thickness_diffuse, read_khth, detangle_interfaces, interface_Kh, use_FGNV_streamfn, use_stanley_gm, &
                      use_GME_thickness_diffuse, find_work, use_variable_mixing, Resoln_scaled_KhTh, Depth_scaled_KhTh, &
                      use_stored_slopes, use_Visbeck, use_QG_Leith_GM, khth_struct, use_MEKE_Kh, EOS_form
    real(c_double) :: Rho_T0_S0, dRho_dT, dRho_dS, dRho_dp
    real(c_double) :: FGNV_scale, N2_floor, MEKE_KhTh_fac
  end type mom6cu_thickness_diffuse_cs
  !> the arguments of thickness_diffuse (:134)
  type, bind(C) :: mom6cu_thickness_diffuse_args
    type(c_ptr)    :: h, uhtr, vhtr, T, S, p_surf
    real(c_double) :: dt
    type(c_ptr)    :: Res_fn_u, Res_fn_v, uhGM, vhGM, slope_x, slope_y, cg1, MEKE_Kh
  end type mom6cu_thickness_diffuse_args
  type, bind(C) :: mom6cu_remapping_cs
    integer(c_int) :: remapping_scheme, boundary_extrapolation, force_bounds_in_subcell, force_bounds_in_target, &
                      om4_remap_via_sub_cells, answer_date
    real(c_double) :: h_neglect, h_neglect_edge
  end type mom6cu_remapping_cs
  type, bind(C) :: mom6cu_regridding_cs
    integer(c_int) :: regridding_scheme, nk
    real(c_double) :: min_thickness, old_grid_weight, depth_of_time_filter_shallow, depth_of_time_filter_deep, Z_ref
    type(c_ptr) :: coordinateResolution
  end type mom6cu_regridding_cs

  ! ---- generated from include/mom6cu.h by the same parser tests/test_abi_layout.py checks this file with ----
  !> hor_visc_CS members (src/parameterizations/lateral/MOM_hor_visc.F90:38-250): switches and the precomputed metric / coefficient arrays
  type, bind(C) :: mom6cu_hor_visc_cs
    integer(c_int) :: Laplacian, biharmonic, no_slip, bound_Kh, better_bound_Kh, bound_Ah, better_bound_Ah, &
                      backscatter_underbound, Smagorinsky_Kh, Smagorinsky_Ah, bound_Coriolis, use_land_mask, &
                      add_LES_viscosity, use_cont_thick, use_cont_thick_bug, unsupported
    real(c_double) :: Kh_bg_min, Re_Ah
    type(c_ptr) :: dx2h, dy2h, DX_dyT, DY_dxT, reduction_xx, Kh_bg_xx, Ah_bg_xx, Kh_Max_xx, Ah_Max_xx, &
                   Laplac2_const_xx, Biharm_const_xx, Biharm_const2_xx, Re_Ah_const_xx, dx2q, dy2q, DX_dyBu, DY_dxBu, &
                   reduction_xy, Kh_bg_xy, Ah_bg_xy, Kh_Max_xy, Ah_Max_xy, Laplac2_const_xy, Biharm_const_xy, &
                   Biharm_const2_xy, Re_Ah_const_xy, Idx2dyCu, Idxdy2u, Idx2dyCv, Idxdy2v
  end type mom6cu_hor_visc_cs
  !> btstep_timeloop's arguments (src/core/MOM_barotropic.F90:2175-2260)
  type, bind(C) :: mom6cu_bt_timeloop_args
    type(c_ptr) :: eta, ubt, vbt, uhbt0, vhbt0, Datu, Datv, BTCL_u, BTCL_v, eta_src, eta_PF, gtot_E, gtot_W, gtot_N, &
                   gtot_S, f_4_u, f_4_v, bt_rem_u, bt_rem_v, BT_force_u, BT_force_v, Cor_ref_u, Cor_ref_v, &
                   IareaT_OBCmask, IdxCu, IdyCv, u_accel_bt, v_accel_bt, eta_sum, eta_wtd, ubtav, vbtav, uhbtav, &
                   vhbtav, ubt_wtd, vbt_wtd, wt_vel, wt_eta, wt_accel, wt_trans, wt_accel2
    real(c_double) :: dtbt, dgeo_de, bebt, vel_underflow
    integer(c_int) :: nstep, nfilter, use_BT_cont, find_etaav, BT_project_velocity, use_old_coriolis_bracket_bug, &
                      use_wide_halos, min_stencil
  end type mom6cu_bt_timeloop_args
  !> btcalc's arguments (MOM_barotropic.F90:4360)
  type, bind(C) :: mom6cu_btcalc_args
    type(c_ptr) :: h, h_u, h_v, frhatu, frhatv, bathyT
    integer(c_int) :: hvel_scheme, may_use_default
  end type mom6cu_btcalc_args
  !> set_dtbt's arguments (MOM_barotropic.F90:3509)
  type, bind(C) :: mom6cu_set_dtbt_args
    type(c_ptr) :: pbce
    real(c_double) :: gtot_est
    integer(c_int) :: have_gtot_est
    type(c_ptr) :: BT_cont, eta
    real(c_double) :: SSH_add
    type(c_ptr) :: frhatu, frhatv, bathyT
    real(c_double) :: bebt, G_extra, dtbt_fraction, BT_Coriolis_scale, Z_ref
    integer(c_int) :: Nonlinear_continuity
  end type mom6cu_set_dtbt_args
  !> ALE_CS members and the fields ALE_regridding_and_remapping touches (src/ALE/MOM_ALE.F90:70-150, src/core/MOM.F90:1751-1926)
  type, bind(C) :: mom6cu_ale_cs
    type(mom6cu_regridding_cs) :: regridCS
    type(mom6cu_remapping_cs) :: remapCS, vel_remapCS
    real(c_double) :: regrid_time_scale
    integer(c_int) :: remap_uv_using_old_alg, do_conv_adj, use_hybgen_unmix, remap_aux_vars
  end type mom6cu_ale_cs
  type, bind(C) :: mom6cu_ale_args
    type(c_ptr) :: u, v, h
    integer(c_int) :: ntr
    type(c_ptr) :: tr, conc_underflow
    integer(c_int) :: iT, iS
    real(c_double) :: dtdia
    type(c_ptr) :: Kd_shear, Kv_shear, Kv_shear_Bu
  end type mom6cu_ale_args

  interface
    integer(c_int) function mom6cu_create(ctx, dom, device) bind(C, name="mom6cu_create")
      import :: c_int, c_ptr, mom6cu_domain
      type(c_ptr), intent(out) :: ctx
      type(mom6cu_domain), intent(in) :: dom
      integer(c_int), value :: device
    end function mom6cu_create
    integer(c_int) function mom6cu_destroy(ctx) bind(C, name="mom6cu_destroy")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
    end function mom6cu_destroy
    integer(c_int) function mom6cu_last_error(ctx, buf, len) bind(C, name="mom6cu_last_error")
      import :: c_int, c_ptr, c_char, c_size_t
      type(c_ptr), value :: ctx
      character(kind=c_char), intent(out) :: buf(*)
      integer(c_size_t), value :: len
    end function mom6cu_last_error
    integer(c_int) function mom6cu_set_grid(ctx, G) bind(C, name="mom6cu_set_grid")
      import :: c_int, c_ptr, mom6cu_grid
      type(c_ptr), value :: ctx
      type(mom6cu_grid), intent(in) :: G
    end function mom6cu_set_grid
    integer(c_int) function mom6cu_set_vgrid(ctx, GV) bind(C, name="mom6cu_set_vgrid")
      import :: c_int, c_ptr, mom6cu_vgrid
      type(c_ptr), value :: ctx
      type(mom6cu_vgrid), intent(in) :: GV
    end function mom6cu_set_vgrid
    integer(c_int) function mom6cu_set_cs_continuity(ctx, CS) bind(C, name="mom6cu_set_cs_continuity")
      import :: c_int, c_ptr, mom6cu_continuity_cs
      type(c_ptr), value :: ctx
      type(mom6cu_continuity_cs), intent(in) :: CS
    end function mom6cu_set_cs_continuity
    integer(c_int) function mom6cu_continuity(ctx, a) bind(C, name="mom6cu_continuity")
      import :: c_int, c_ptr, mom6cu_continuity_args
      type(c_ptr), value :: ctx
      type(mom6cu_continuity_args), intent(in) :: a
    end function mom6cu_continuity
    integer(c_int) function mom6cu_set_cs_coriolisadv(ctx, CS) bind(C, name="mom6cu_set_cs_coriolisadv")
      import :: c_int, c_ptr, mom6cu_coriolisadv_cs
      type(c_ptr), value :: ctx
      type(mom6cu_coriolisadv_cs), intent(in) :: CS
    end function mom6cu_set_cs_coriolisadv
    integer(c_int) function mom6cu_coradcalc(ctx, a) bind(C, name="mom6cu_coradcalc")
      import :: c_int, c_ptr, mom6cu_coradcalc_args
      type(c_ptr), value :: ctx
      type(mom6cu_coradcalc_args), intent(in) :: a
    end function mom6cu_coradcalc
    !> mom6cu_set_cs_hor_visc takes the C struct mom6cu_hor_visc_cs (16 ints, 2 doubles, 30 pointers in the order
    !! of include/mom6cu.h); it is passed here as an opaque pointer to a bind(C) copy built by hor_visc_init.
    integer(c_int) function mom6cu_set_cs_hor_visc(ctx, CS) bind(C, name="mom6cu_set_cs_hor_visc")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: CS
    end function mom6cu_set_cs_hor_visc
    integer(c_int) function mom6cu_horizontal_viscosity(ctx, a) bind(C, name="mom6cu_horizontal_viscosity")
      import :: c_int, c_ptr, mom6cu_hor_visc_args
      type(c_ptr), value :: ctx
      type(mom6cu_hor_visc_args), intent(in) :: a
    end function mom6cu_horizontal_viscosity
    !> btstep_timeloop (MOM_barotropic.F90:2175): takes the C struct mom6cu_bt_timeloop_args
    integer(c_int) function mom6cu_btstep_timeloop(ctx, a) bind(C, name="mom6cu_btstep_timeloop")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: a
    end function mom6cu_btstep_timeloop
    integer(c_int) function mom6cu_comm_unique_id(buf, nbytes) bind(C, name="mom6cu_comm_unique_id")
      import :: c_int, c_ptr
      type(c_ptr), value :: buf
      integer(c_int), value :: nbytes
    end function mom6cu_comm_unique_id
    integer(c_int) function mom6cu_comm_init(ctx, id, nbytes, rank, nranks) bind(C, name="mom6cu_comm_init")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, id
      integer(c_int), value :: nbytes, rank, nranks
    end function mom6cu_comm_init
    integer(c_int) function mom6cu_set_unit_scale(ctx, US) bind(C, name="mom6cu_set_unit_scale")
      import ; type(c_ptr), value :: ctx ; type(mom6cu_unit_scale), intent(in) :: US
    end function mom6cu_set_unit_scale
    integer(c_int) function mom6cu_set_cs_pressureforce(ctx, CS) bind(C, name="mom6cu_set_cs_pressureforce")
      import ; type(c_ptr), value :: ctx ; type(mom6cu_pressureforce_cs), intent(in) :: CS
    end function mom6cu_set_cs_pressureforce
    integer(c_int) function mom6cu_pressure_force(ctx, a) bind(C, name="mom6cu_pressure_force")
      import ; type(c_ptr), value :: ctx ; type(mom6cu_pressureforce_args), intent(in) :: a
    end function mom6cu_pressure_force
    integer(c_int) function mom6cu_set_cs_vertvisc(ctx, CS) bind(C, name="mom6cu_set_cs_vertvisc")
      import ; type(c_ptr), value :: ctx ; type(mom6cu_vertvisc_cs), intent(in) :: CS
    end function mom6cu_set_cs_vertvisc
    integer(c_int) function mom6cu_vertvisc_coef(ctx, a) bind(C, name="mom6cu_vertvisc_coef")
      import ; type(c_ptr), value :: ctx ; type(mom6cu_vertvisc_coef_args), intent(in) :: a
    end function mom6cu_vertvisc_coef
    integer(c_int) function mom6cu_vertvisc(ctx, a) bind(C, name="mom6cu_vertvisc")
      import ; type(c_ptr), value :: ctx ; type(mom6cu_vertvisc_args), intent(in) :: a
    end function mom6cu_vertvisc
    integer(c_int) function mom6cu_vertvisc_remnant(ctx, Ray_u, Ray_v, visc_rem_u, visc_rem_v, dt) bind(C, name="mom6cu_vertvisc_remnant")
      import ; type(c_ptr), value :: ctx, Ray_u, Ray_v, visc_rem_u, visc_rem_v ; real(c_double), value :: dt
    end function mom6cu_vertvisc_remnant
    integer(c_int) function mom6cu_btstep(ctx, CS, a) bind(C, name="mom6cu_btstep")
      import ; type(c_ptr), value :: ctx ; type(mom6cu_barotropic_cs), intent(in) :: CS ; type(mom6cu_btstep_args), intent(in) :: a
    end function mom6cu_btstep
    integer(c_int) function mom6cu_bt_mass_source(ctx, h, eta, set_cor, eta_cor) bind(C, name="mom6cu_bt_mass_source")
      import ; type(c_ptr), value :: ctx, h, eta, eta_cor ; integer(c_int), value :: set_cor
    end function mom6cu_bt_mass_source
    integer(c_int) function mom6cu_step_dyn_split_rk2(ctx, CS, a) bind(C, name="mom6cu_step_dyn_split_rk2")
      import ; type(c_ptr), value :: ctx ; type(mom6cu_dyn_split_rk2_cs), intent(inout) :: CS ; type(mom6cu_step_dyn_args), intent(in) :: a
    end function mom6cu_step_dyn_split_rk2
    integer(c_int) function mom6cu_advect_tracer(ctx, CS, a) bind(C, name="mom6cu_advect_tracer")
      import ; type(c_ptr), value :: ctx ; type(mom6cu_tracer_advect_cs), intent(in) :: CS ; type(mom6cu_advect_tracer_args), intent(in) :: a
    end function mom6cu_advect_tracer
    integer(c_int) function mom6cu_ale_regrid(ctx, CS, h, h_new, dzRegrid) bind(C, name="mom6cu_ale_regrid")
      import ; type(c_ptr), value :: ctx, h, h_new, dzRegrid ; type(mom6cu_regridding_cs), intent(in) :: CS
    end function mom6cu_ale_regrid
    integer(c_int) function mom6cu_ale_remap_tracers(ctx, CS, h_old, h_new, ntr, tr, conc_underflow) bind(C, name="mom6cu_ale_remap_tracers")
      import ; type(c_ptr), value :: ctx, h_old, h_new, tr, conc_underflow ; type(mom6cu_remapping_cs), intent(in) :: CS
      integer(c_int), value :: ntr
    end function mom6cu_ale_remap_tracers
    integer(c_int) function mom6cu_ale_remap_set_h_vel(ctx, h_new, h_u, h_v) bind(C, name="mom6cu_ale_remap_set_h_vel")
      import ; type(c_ptr), value :: ctx, h_new, h_u, h_v
    end function mom6cu_ale_remap_set_h_vel
    integer(c_int) function mom6cu_ale_remap_velocities(ctx, CS, h_old_u, h_old_v, h_new_u, h_new_v, u, v) &
        bind(C, name="mom6cu_ale_remap_velocities")
      import ; type(c_ptr), value :: ctx, h_old_u, h_old_v, h_new_u, h_new_v, u, v ; type(mom6cu_remapping_cs), intent(in) :: CS
    end function mom6cu_ale_remap_velocities
    !> reproducing_sum (src/framework/MOM_coms.F90:227, :337); absent optionals are c_null_ptr, isr..jer = 0 when absent
    integer(c_int) function mom6cu_reproducing_sum(ctx, array, stagger, nk, isr, ier, jsr, jer, unscale, only_on_PE, &
                                                   total, sums, EFP_sum, EFP_lay_sums) bind(C, name="mom6cu_reproducing_sum")
      import ; type(c_ptr), value :: ctx, array, sums, EFP_sum, EFP_lay_sums
      integer(c_int), value :: stagger, nk, isr, ier, jsr, jer, only_on_PE
      real(c_double), value :: unscale ; real(c_double), intent(out) :: total
    end function mom6cu_reproducing_sum
    integer(c_int) function mom6cu_efp_sum_across_pes(ctx, EFPs, nval) bind(C, name="mom6cu_efp_sum_across_pes")
      import ; type(c_ptr), value :: ctx ; type(mom6cu_efp), intent(inout) :: EFPs(*) ; integer(c_int), value :: nval
    end function mom6cu_efp_sum_across_pes
    !> hchksum / uvchksum / Bchksum (src/framework/MOM_checksums.F90)
    integer(c_int) function mom6cu_chksum(ctx, array, stagger, nk, haloshift, symmetric, omit_corners, scale, bc, kind, stats) &
        bind(C, name="mom6cu_chksum")
      import ; type(c_ptr), value :: ctx, array, stats
      integer(c_int), value :: stagger, nk, haloshift, symmetric, omit_corners ; real(c_double), value :: scale
      integer(c_int), intent(out) :: bc(5), kind
    end function mom6cu_chksum
    !> write_energy (src/diagnostics/MOM_sum_output.F90:321)
    integer(c_int) function mom6cu_write_energy(ctx, CS, u, v, h, T, S, res) bind(C, name="mom6cu_write_energy")
      import ; type(c_ptr), value :: ctx, u, v, h, T, S
      type(mom6cu_sum_output_cs), intent(inout) :: CS ; type(mom6cu_energy_out), intent(inout) :: res
    end function mom6cu_write_energy
    integer(c_int) function mom6cu_ocean_stats_line(CS, res, n, reday, buf, len) bind(C, name="mom6cu_ocean_stats_line")
      import ; type(mom6cu_sum_output_cs), intent(in) :: CS ; type(mom6cu_energy_out), intent(in) :: res
      integer(c_int), value :: n ; real(c_double), value :: reday
      character(kind=c_char), intent(out) :: buf(*) ; integer(c_size_t), value :: len
    end function mom6cu_ocean_stats_line
    !> mixedlayer_restrat (src/parameterizations/lateral/MOM_mixed_layer_restrat.F90:149), called from step_MOM_dynamics (MOM.F90:1422)
    integer(c_int) function mom6cu_mixedlayer_restrat(ctx, CS, h, uhtr, vhtr, T, S, ustar, dt, h_MLD, Rd_dx_h) &
        bind(C, name="mom6cu_mixedlayer_restrat")
      import ; type(c_ptr), value :: ctx, h, uhtr, vhtr, T, S, ustar, h_MLD, Rd_dx_h
      type(mom6cu_mle_cs), intent(inout) :: CS ; real(c_double), value :: dt
    end function mom6cu_mixedlayer_restrat
    !> mu(sigma, dh) (MOM_mixed_layer_restrat.F90:717) for n values
    integer(c_int) function mom6cu_mle_mu(ctx, n, sigma, dh, res) bind(C, name="mom6cu_mle_mu")
      import ; type(c_ptr), value :: ctx, sigma, dh, res ; integer(c_int), value :: n
    end function mom6cu_mle_mu
    !> tracer_hordiff (src/tracer/MOM_tracer_hor_diff.F90:119), called from step_MOM_tracer_dyn (MOM.F90:1526)
    integer(c_int) function mom6cu_tracer_hordiff(ctx, CS, a) bind(C, name="mom6cu_tracer_hordiff")
      import ; type(c_ptr), value :: ctx
      type(mom6cu_tracer_hor_diff_cs), intent(in) :: CS ; type(mom6cu_tracer_hordiff_args), intent(in) :: a
    end function mom6cu_tracer_hordiff
    !> thickness_diffuse (src/parameterizations/lateral/MOM_thickness_diffuse.F90:134), called from step_MOM_dynamics (MOM.F90:1388)
    integer(c_int) function mom6cu_thickness_diffuse(ctx, CS, a) bind(C, name="mom6cu_thickness_diffuse")
      import ; type(c_ptr), value :: ctx
      type(mom6cu_thickness_diffuse_cs), intent(in) :: CS ; type(mom6cu_thickness_diffuse_args), intent(in) :: a
    end function mom6cu_thickness_diffuse
    !> pass_var / pass_vector / do_group_pass (src/framework/MOM_domains.F90) for fields that live on the device
    integer(c_int) function mom6cu_do_group_pass(ctx, nfields, fields, stagger, nk) bind(C, name="mom6cu_do_group_pass")
      import ; type(c_ptr), value :: ctx ; integer(c_int), value :: nfields, nk
      type(c_ptr), intent(in) :: fields(*) ; integer(c_int), intent(in) :: stagger(*)
    end function mom6cu_do_group_pass
    type(c_ptr) function mom6cu_plane_alloc(ctx, name, nk) bind(C, name="mom6cu_plane_alloc")
      import ; type(c_ptr), value :: ctx ; character(kind=c_char), intent(in) :: name(*) ; integer(c_int), value :: nk
    end function mom6cu_plane_alloc
    ! ---- generated from the prototypes of include/mom6cu.h ----
    integer(c_int) function mom6cu_ale_regridding_and_remapping(ctx, CS, dynCS, a) bind(C, name="mom6cu_ale_regridding_and_remapping")
      import :: c_int, c_ptr, mom6cu_ale_args, mom6cu_ale_cs, mom6cu_dyn_split_rk2_cs
      type(c_ptr), value :: ctx
      type(mom6cu_ale_cs), intent(inout) :: CS
      type(c_ptr), value :: dynCS   ! const mom6cu_dyn_split_rk2_cs*, or c_null_ptr (no auxiliary variables to remap)
      type(mom6cu_ale_args), intent(in) :: a
    end function mom6cu_ale_regridding_and_remapping
    integer(c_int) function mom6cu_ale_remap_interface_vals(ctx, h_old, h_new, int_val) bind(C, name="mom6cu_ale_remap_interface_vals")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: h_old
      type(c_ptr), value :: h_new
      type(c_ptr), value :: int_val
    end function mom6cu_ale_remap_interface_vals
    integer(c_int) function mom6cu_ale_remap_vertex_vals(ctx, h_old, h_new, vert_val) bind(C, name="mom6cu_ale_remap_vertex_vals")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: h_old
      type(c_ptr), value :: h_new
      type(c_ptr), value :: vert_val
    end function mom6cu_ale_remap_vertex_vals
    integer(c_int) function mom6cu_btcalc(ctx, a) bind(C, name="mom6cu_btcalc")
      import :: c_int, c_ptr, mom6cu_btcalc_args
      type(c_ptr), value :: ctx
      type(mom6cu_btcalc_args), intent(in) :: a
    end function mom6cu_btcalc
    integer(c_int) function mom6cu_btstep_timeloop_resident(ctx, a, reps, download) bind(C, name="mom6cu_btstep_timeloop_resident")
      import :: c_int, c_ptr, mom6cu_bt_timeloop_args
      type(c_ptr), value :: ctx
      type(mom6cu_bt_timeloop_args), intent(in) :: a
      integer(c_int), value :: reps
      integer(c_int), value :: download
    end function mom6cu_btstep_timeloop_resident
    integer(c_int) function mom6cu_build_arch() bind(C, name="mom6cu_build_arch")
      import :: c_int
    end function mom6cu_build_arch
    integer(c_int) function mom6cu_comm_destroy(ctx) bind(C, name="mom6cu_comm_destroy")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
    end function mom6cu_comm_destroy
    subroutine mom6cu_efp_minus(a, b, out, overflow) bind(C, name="mom6cu_efp_minus")
      import :: c_ptr, mom6cu_efp
      type(mom6cu_efp), intent(in) :: a
      type(mom6cu_efp), intent(in) :: b
      type(mom6cu_efp), intent(inout) :: out
      type(c_ptr), value :: overflow
    end subroutine mom6cu_efp_minus
    subroutine mom6cu_efp_plus(a, b, out, overflow) bind(C, name="mom6cu_efp_plus")
      import :: c_ptr, mom6cu_efp
      type(mom6cu_efp), intent(in) :: a
      type(mom6cu_efp), intent(in) :: b
      type(mom6cu_efp), intent(inout) :: out
      type(c_ptr), value :: overflow
    end subroutine mom6cu_efp_plus
    real(c_double) function mom6cu_efp_real_diff(a, b) bind(C, name="mom6cu_efp_real_diff")
      import :: c_double, mom6cu_efp
      type(mom6cu_efp), intent(in) :: a
      type(mom6cu_efp), intent(in) :: b
    end function mom6cu_efp_real_diff
    real(c_double) function mom6cu_efp_to_real(a) bind(C, name="mom6cu_efp_to_real")
      import :: c_double, mom6cu_efp
      type(mom6cu_efp), intent(inout) :: a
    end function mom6cu_efp_to_real
    integer(c_int) function mom6cu_halo_plan(dom, stagger, wide, halo, dir, send_box, recv_box) bind(C, name="mom6cu_halo_plan")
      import :: c_int, c_ptr, mom6cu_domain
      type(mom6cu_domain), intent(in) :: dom
      integer(c_int), value :: stagger
      integer(c_int), value :: wide
      integer(c_int), value :: halo
      integer(c_int), value :: dir
      type(c_ptr), value :: send_box
      type(c_ptr), value :: recv_box
    end function mom6cu_halo_plan
    integer(c_int) function mom6cu_interpolate_column(ctx, ncol, nsrc, h_src, u_src, ndest, h_dest, u_dest, mask_edges) bind(C, name="mom6cu_interpolate_column")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: ncol
      integer(c_int), value :: nsrc
      type(c_ptr), value :: h_src
      type(c_ptr), value :: u_src
      integer(c_int), value :: ndest
      type(c_ptr), value :: h_dest
      type(c_ptr), value :: u_dest
      integer(c_int), value :: mask_edges
    end function mom6cu_interpolate_column
    integer(c_int) function mom6cu_last_iterations(ctx) bind(C, name="mom6cu_last_iterations")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
    end function mom6cu_last_iterations
    real(c_double) function mom6cu_last_kernel_ms(ctx) bind(C, name="mom6cu_last_kernel_ms")
      import :: c_double, c_ptr
      type(c_ptr), value :: ctx
    end function mom6cu_last_kernel_ms
    integer(c_int) function mom6cu_last_step_stage_ms(ctx, ms, n) bind(C, name="mom6cu_last_step_stage_ms")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: ms
      integer(c_int), value :: n
    end function mom6cu_last_step_stage_ms
    integer(c_int) function mom6cu_plane_download(ctx, plane, host, stagger, wide, nk) bind(C, name="mom6cu_plane_download")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: plane
      type(c_ptr), value :: host
      integer(c_int), value :: stagger
      integer(c_int), value :: wide
      integer(c_int), value :: nk
    end function mom6cu_plane_download
    integer(c_int) function mom6cu_plane_upload(ctx, plane, host, stagger, wide, nk) bind(C, name="mom6cu_plane_upload")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: plane
      type(c_ptr), value :: host
      integer(c_int), value :: stagger
      integer(c_int), value :: wide
      integer(c_int), value :: nk
    end function mom6cu_plane_upload
    integer(c_int) function mom6cu_real_to_efp(val, out) bind(C, name="mom6cu_real_to_efp")
      import :: c_double, c_int, mom6cu_efp
      real(c_double), value :: val
      type(mom6cu_efp), intent(inout) :: out
    end function mom6cu_real_to_efp
    integer(c_int) function mom6cu_remap_dyn_split_rk2_aux_vars(ctx, remapCS, CS, h_old_u, h_old_v, h_new_u, h_new_v) bind(C, name="mom6cu_remap_dyn_split_rk2_aux_vars")
      import :: c_int, c_ptr, mom6cu_dyn_split_rk2_cs, mom6cu_remapping_cs
      type(c_ptr), value :: ctx
      type(mom6cu_remapping_cs), intent(in) :: remapCS
      type(mom6cu_dyn_split_rk2_cs), intent(in) :: CS
      type(c_ptr), value :: h_old_u
      type(c_ptr), value :: h_old_v
      type(c_ptr), value :: h_new_u
      type(c_ptr), value :: h_new_v
    end function mom6cu_remap_dyn_split_rk2_aux_vars
    integer(c_int) function mom6cu_remapping_core_h(ctx, CS, ncol, n0, h0, u0, n1, h1, u1) bind(C, name="mom6cu_remapping_core_h")
      import :: c_int, c_ptr, mom6cu_remapping_cs
      type(c_ptr), value :: ctx
      type(mom6cu_remapping_cs), intent(in) :: CS
      integer(c_int), value :: ncol
      integer(c_int), value :: n0
      type(c_ptr), value :: h0
      type(c_ptr), value :: u0
      integer(c_int), value :: n1
      type(c_ptr), value :: h1
      type(c_ptr), value :: u1
    end function mom6cu_remapping_core_h
    integer(c_int) function mom6cu_set_dtbt(ctx, a, dtbt, dtbt_max) bind(C, name="mom6cu_set_dtbt")
      import :: c_int, c_ptr, mom6cu_set_dtbt_args
      type(c_ptr), value :: ctx
      type(mom6cu_set_dtbt_args), intent(in) :: a
      type(c_ptr), value :: dtbt
      type(c_ptr), value :: dtbt_max
    end function mom6cu_set_dtbt
    integer(c_int) function mom6cu_sync(ctx) bind(C, name="mom6cu_sync")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
    end function mom6cu_sync
    real(c_double) function mom6cu_total_kernel_ms(ctx) bind(C, name="mom6cu_total_kernel_ms")
      import :: c_double, c_ptr
      type(c_ptr), value :: ctx
    end function mom6cu_total_kernel_ms
    integer(c_int) function mom6cu_vertvisc_get_coef(ctx, a_u, a_v, h_u, h_v) bind(C, name="mom6cu_vertvisc_get_coef")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: a_u
      type(c_ptr), value :: a_v
      type(c_ptr), value :: h_u
      type(c_ptr), value :: h_v
    end function mom6cu_vertvisc_get_coef
    integer(c_long_long) function mom6cu_vertvisc_ntrunc(ctx) bind(C, name="mom6cu_vertvisc_ntrunc")
      import :: c_long_long, c_ptr
      type(c_ptr), value :: ctx
    end function mom6cu_vertvisc_ntrunc
    integer(c_long_long) function mom6cu_launch_count(ctx) bind(C, name="mom6cu_launch_count")
      import :: c_long_long, c_ptr
      type(c_ptr), value :: ctx
    end function mom6cu_launch_count
    integer(c_long_long) function mom6cu_sizeof(name) bind(C, name="mom6cu_sizeof")
      import :: c_long_long, c_char
      character(kind=c_char), intent(in) :: name(*)
    end function mom6cu_sizeof
    integer(c_int) function mom6cu_plane_zero(ctx, plane, nk) bind(C, name="mom6cu_plane_zero")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      type(c_ptr), value :: plane
      integer(c_int), value :: nk
    end function mom6cu_plane_zero
  end interface

  !> The one device context of this PE (one MPI rank = one tile = one GPU)
  type(c_ptr), save :: mom6cu_ctx = c_null_ptr

contains

  !> True once the driver has created the device context (mom6cu_create in initialize_MOM): the hooks the installer adds to the
  !! reference's procedures (fortran/install_shims.py) take the device path only then, so an unmodified run is still possible.
  logical function mom6cu_enabled()
    mom6cu_enabled = c_associated(mom6cu_ctx)
  end function mom6cu_enabled

  !> Turn a nonzero status into MOM_error(FATAL, msg) / MOM_error(WARNING, ...), the reference's error convention
  !! (src/framework/MOM_error_handler.F90).
  subroutine mom6cu_check(rc, where)
    use MOM_error_handler, only : MOM_error, FATAL, WARNING
    integer(c_int),   intent(in) :: rc
    character(len=*), intent(in) :: where
    character(kind=c_char) :: buf(1024)
    character(len=1024) :: msg
    integer :: n, ierr
    if (rc == 0) return
    if (rc < 0) then
      call MOM_error(WARNING, trim(where)//": the device path issued warnings.")
      return
    endif
    ierr = mom6cu_last_error(mom6cu_ctx, buf, int(1024, c_size_t))
    msg = " "
    do n=1,1024 ; if (buf(n) == c_null_char) exit ; msg(n:n) = buf(n) ; enddo
    call MOM_error(FATAL, trim(where)//": "//trim(msg))
  end subroutine mom6cu_check

  !> c_loc of an optional array, or c_null_ptr when it is absent (Fortran optional -> NULL)
  function opt_loc3(a) result(p)
    real(c_double), dimension(:,:,:), optional, target, intent(in) :: a
    type(c_ptr) :: p
    p = c_null_ptr ; if (present(a)) p = c_loc(a)
  end function opt_loc3
  function opt_loc2(a) result(p)
    real(c_double), dimension(:,:), optional, target, intent(in) :: a
    type(c_ptr) :: p
    p = c_null_ptr ; if (present(a)) p = c_loc(a)
  end function opt_loc2

  !> c_loc of an allocatable 2-D array of a control structure, or c_null_ptr when it is not allocated
  function opt_alloc2(x) result(p)
    real(c_double), dimension(:,:), allocatable, target, intent(in) :: x
    type(c_ptr) :: p
    p = c_null_ptr ; if (allocated(x)) p = c_loc(x)
  end function opt_alloc2

end module mom6cu_interface
