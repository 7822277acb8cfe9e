!> Drop-in replacement for the compute entry of src/parameterizations/lateral/MOM_mixed_layer_restrat.F90: same module name and the
!! same dummy argument list for mixedlayer_restrat (:149-186), body forwarded to the sm_100a library through mom6cu_interface.  It is
!! a worked binding of a caller of the dycore (the dycore's own are fortran/bodies/*.inc, installed by fortran/install_shims.py) and shows a caller of step_MOM_dynamics (MOM.F90:1422) whose
!! arguments are updated in place and whose control structure carries model state (CS%MLD_filtered, a restart field).
!! The init / restart-registration routines of the reference module are kept as they are (they read parameters and allocate
!! CS%MLD_filtered, :1618-2009).  The control structure's members are private to the module (:42), so this subroutine is a module
!! procedure of the shadowing copy of MOM_mixed_layer_restrat.F90 -- `contains` it there and have mixedlayer_restrat (:149-186) call it
!! where it calls mixedlayer_restrat_OM4 (:179); it is shown as a file of its own only for readability.  Not compiled in the build
!! container (no Fortran compiler there).
!!
!! subroutine mixedlayer_restrat(h, uhtr, vhtr, tv, forces, dt, MLD, h_MLD, bflux, VarMix, G, GV, US, CS)
!!   ... the reference's declarations (:150-170), arrays given the TARGET attribute ...
subroutine mixedlayer_restrat_mom6cu(h, uhtr, vhtr, tv, forces, dt, h_MLD, VarMix, G, GV, US, CS)
  use, intrinsic :: iso_c_binding
  use mom6cu_interface
  use MOM_error_handler, only : MOM_error, FATAL
  use MOM_forcing_type,  only : mech_forcing
  use MOM_grid,          only : ocean_grid_type
  use MOM_lateral_mixing_coeffs, only : VarMix_CS
  use MOM_unit_scaling,  only : unit_scale_type
  use MOM_variables,     only : thermo_var_ptrs
  use MOM_verticalGrid,  only : verticalGrid_type
  implicit none
#include <MOM_memory.h>
  type(ocean_grid_type),                      intent(inout) :: G
  type(verticalGrid_type),                    intent(in)    :: GV
  type(unit_scale_type),                      intent(in)    :: US
  real, dimension(SZI_(G),SZJ_(G),SZK_(GV)),  target, intent(inout) :: h      !< Layer thickness [H ~> m or kg m-2]
  real, dimension(SZIB_(G),SZJ_(G),SZK_(GV)), target, intent(inout) :: uhtr   !< Accumulated zonal mass flux [H L2 ~> m3 or kg]
  real, dimension(SZI_(G),SZJB_(G),SZK_(GV)), target, intent(inout) :: vhtr   !< Accumulated meridional mass flux [H L2 ~> m3 or kg]
  type(thermo_var_ptrs),                      intent(in)    :: tv
  type(mech_forcing),                         intent(in)    :: forces
  real,                                       intent(in)    :: dt
  real, dimension(:,:),                       pointer       :: h_MLD          !< visc%h_ML [H ~> m or kg m-2]
  type(VarMix_CS),                    target, intent(in)    :: VarMix
  type(mixedlayer_restrat_CS),        target, intent(inout) :: CS

  type(mom6cu_mle_cs) :: c
  type(c_ptr) :: p_MLD, p_Rd

  if (.not.CS%initialized) call MOM_error(FATAL, "mixedlayer_restrat: Module must be initialized before it is used.")
  if (GV%nkml > 0) call MOM_error(FATAL, "mixedlayer_restrat (mom6cu): the bulk-mixed-layer variant is outside the device path's option set.")
  ! the resolved control structure (the members mixedlayer_restrat_OM4 reads, :42-115)
  c%ml_restrat_coef = CS%ml_restrat_coef ; c%ml_restrat_coef2 = CS%ml_restrat_coef2 ; c%front_length = CS%front_length
  c%MLE_MLD_decay_time = CS%MLE_MLD_decay_time ; c%MLE_MLD_decay_time2 = CS%MLE_MLD_decay_time2
  c%MLE_MLD_stretch = CS%MLE_MLD_stretch ; c%MLE_tail_dh = CS%MLE_tail_dh ; c%ustar_min = CS%ustar_min ; c%vonKar = CS%vonKar
  c%MLE_density_diff = CS%MLE_density_diff
  c%MLE_use_PBL_MLD = merge(1, 0, CS%MLE_use_PBL_MLD) ; c%use_Stanley_ML = merge(1, 0, CS%use_Stanley_ML)
  c%use_Bodner = merge(1, 0, CS%use_Bodner) ; c%fl_from_file = merge(1, 0, CS%fl_from_file)
  ! EOS_type keeps its form and coefficients private (MOM_EOS.F90:107-150) and offers no query for them, so the shim's init reads
  ! them once with the same get_param calls EOS_init makes (EQN_OF_STATE, RHO_T0_S0, DRHO_DT, DRHO_DS; MOM_EOS.F90:1535-1580) into
  ! the module variables eos_form (MOM6CU_EOS_*: 1 = LINEAR, 3 = WRIGHT, include/mom6cu.h), lin_Rho_T0_S0, lin_dRho_dT, lin_dRho_dS.
  c%EOS_form = 0
  if (associated(tv%eqn_of_state)) c%EOS_form = eos_form
  c%Rho_T0_S0 = lin_Rho_T0_S0 ; c%dRho_dT = lin_dRho_dT ; c%dRho_dS = lin_dRho_dS ; c%dRho_dp = 0.0
  ! model state held by the control structure: updated in place by the library (restart fields stay on the host side of the ABI)
  c%MLD_filtered = c_null_ptr ; c%MLD_filtered_slow = c_null_ptr
  if (allocated(CS%MLD_filtered))      c%MLD_filtered      = c_loc(CS%MLD_filtered)
  if (allocated(CS%MLD_filtered_slow)) c%MLD_filtered_slow = c_loc(CS%MLD_filtered_slow)
  p_MLD = c_null_ptr ; if (associated(h_MLD)) p_MLD = c_loc(h_MLD)
  p_Rd = c_null_ptr ; if (allocated(VarMix%Rd_dx_h)) p_Rd = c_loc(VarMix%Rd_dx_h)
  ! rc > 0 -> MOM_error(FATAL, message of the library); options outside the frozen set (Bodner, Stanley, front length from a file)
  ! are refused there, so a run never silently diverges from the reference
  call mom6cu_check(mom6cu_mixedlayer_restrat(mom6cu_ctx, c, c_loc(h), c_loc(uhtr), c_loc(vhtr), c_loc(tv%T), c_loc(tv%S), &
                                              c_loc(forces%ustar), dt, p_MLD, p_Rd), "mixedlayer_restrat")
end subroutine mixedlayer_restrat_mom6cu
