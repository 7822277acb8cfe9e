!> Drop-in replacement for src/core/MOM_continuity_PPM.F90: same module name, same public symbols
!! (continuity_PPM, continuity_PPM_init, continuity_PPM_stencil, continuity_PPM_CS, ...; reference public list
!! MOM_continuity_PPM.F90:24-32), bodies forwarded to the sm_100a library through mom6cu_interface.
!! Build-time selection: put this file's directory ahead of src/core in the source list (the same mechanism the
!! reference uses for config_src/external/* stubs, ac/configure.ac:259-264).
!! Not compiled in the build container (no Fortran compiler there).
module MOM_continuity_PPM
  use, intrinsic :: iso_c_binding
  use mom6cu_interface
  use MOM_error_handler, only : MOM_error, FATAL
  use MOM_file_parser,   only : get_param, log_version, param_file_type
  use MOM_grid,          only : ocean_grid_type
  use MOM_open_boundary, only : ocean_OBC_type
  use MOM_unit_scaling,  only : unit_scale_type
  use MOM_variables,     only : BT_cont_type, porous_barrier_type
  use MOM_verticalGrid,  only : verticalGrid_type
  implicit none ; private
#include <MOM_memory.h>

  public continuity_PPM, continuity_PPM_init, continuity_PPM_stencil

  !> Same role as the reference type (MOM_continuity_PPM.F90:35-67); the resolved values live on the device.
  type, public :: continuity_PPM_CS ; private
    logical :: initialized = .false.
    type(mom6cu_continuity_cs) :: c
  end type continuity_PPM_CS

contains

!> continuity_PPM, MOM_continuity_PPM.F90:86-194: identical dummy argument list.
subroutine continuity_PPM(u, v, hin, h, uh, vh, dt, G, GV, US, CS, OBC, pbv, uhbt, vhbt, &
                          visc_rem_u, visc_rem_v, u_cor, v_cor, BT_cont, du_cor, dv_cor)
  type(ocean_grid_type),   intent(inout) :: G
  type(verticalGrid_type), intent(in)    :: GV
  real, dimension(SZIB_(G),SZJ_(G),SZK_(GV)), target, intent(in)    :: u
  real, dimension(SZI_(G),SZJB_(G),SZK_(GV)), target, intent(in)    :: v
  real, dimension(SZI_(G),SZJ_(G),SZK_(GV)),  target, intent(in)    :: hin
  real, dimension(SZI_(G),SZJ_(G),SZK_(GV)),  target, intent(inout) :: h
  real, dimension(SZIB_(G),SZJ_(G),SZK_(GV)), target, intent(out)   :: uh
  real, dimension(SZI_(G),SZJB_(G),SZK_(GV)), target, intent(out)   :: vh
  real,                    intent(in)    :: dt
  type(unit_scale_type),   intent(in)    :: US
  type(continuity_PPM_CS), intent(in)    :: CS
  type(ocean_OBC_type),    pointer       :: OBC
  type(porous_barrier_type), target, intent(in) :: pbv
  real, dimension(SZIB_(G),SZJ_(G)), target, optional, intent(in)  :: uhbt
  real, dimension(SZI_(G),SZJB_(G)), target, optional, intent(in)  :: vhbt
  real, dimension(SZIB_(G),SZJ_(G),SZK_(GV)), target, optional, intent(in)  :: visc_rem_u
  real, dimension(SZI_(G),SZJB_(G),SZK_(GV)), target, optional, intent(in)  :: visc_rem_v
  real, dimension(SZIB_(G),SZJ_(G),SZK_(GV)), target, optional, intent(out) :: u_cor
  real, dimension(SZI_(G),SZJB_(G),SZK_(GV)), target, optional, intent(out) :: v_cor
  type(BT_cont_type),      pointer, optional :: BT_cont
  real, dimension(SZIB_(G),SZJ_(G)), target, optional, intent(out) :: du_cor
  real, dimension(SZI_(G),SZJB_(G)), target, optional, intent(out) :: dv_cor

  type(mom6cu_continuity_args) :: a
  type(mom6cu_bt_cont), target :: b

  if (.not.CS%initialized) call MOM_error(FATAL, &
         "MOM_continuity_PPM: Module must be initialized before it is used.")
  if (associated(OBC)) call MOM_error(FATAL, &
         "MOM_continuity_PPM (mom6cu): open boundary conditions are outside the device path's option set.")

  a%u = c_loc(u) ; a%v = c_loc(v) ; a%hin = c_loc(hin) ; a%h = c_loc(h) ; a%uh = c_loc(uh) ; a%vh = c_loc(vh)
  a%dt = dt
  a%por_face_areaU = c_loc(pbv%por_face_areaU) ; a%por_face_areaV = c_loc(pbv%por_face_areaV)
  a%uhbt = opt_loc2(uhbt) ; a%vhbt = opt_loc2(vhbt)
  a%visc_rem_u = opt_loc3(visc_rem_u) ; a%visc_rem_v = opt_loc3(visc_rem_v)
  a%u_cor = opt_loc3(u_cor) ; a%v_cor = opt_loc3(v_cor)
  a%du_cor = opt_loc2(du_cor) ; a%dv_cor = opt_loc2(dv_cor)
  a%BT_cont = c_null_ptr
  if (present(BT_cont)) then ; if (associated(BT_cont)) then
    b%FA_u_EE = c_loc(BT_cont%FA_u_EE) ; b%FA_u_E0 = c_loc(BT_cont%FA_u_E0)
    b%FA_u_W0 = c_loc(BT_cont%FA_u_W0) ; b%FA_u_WW = c_loc(BT_cont%FA_u_WW)
    b%uBT_WW = c_loc(BT_cont%uBT_WW) ; b%uBT_EE = c_loc(BT_cont%uBT_EE)
    b%FA_v_NN = c_loc(BT_cont%FA_v_NN) ; b%FA_v_N0 = c_loc(BT_cont%FA_v_N0)
    b%FA_v_S0 = c_loc(BT_cont%FA_v_S0) ; b%FA_v_SS = c_loc(BT_cont%FA_v_SS)
    b%vBT_SS = c_loc(BT_cont%vBT_SS) ; b%vBT_NN = c_loc(BT_cont%vBT_NN)
    b%h_u = c_null_ptr ; b%h_v = c_null_ptr
    if (allocated(BT_cont%h_u)) b%h_u = c_loc(BT_cont%h_u)
    if (allocated(BT_cont%h_v)) b%h_v = c_loc(BT_cont%h_v)
    a%BT_cont = c_loc(b)
  endif ; endif
  call mom6cu_check(mom6cu_continuity(mom6cu_ctx, a), "continuity_PPM")
end subroutine continuity_PPM

!> continuity_PPM_init, MOM_continuity_PPM.F90:2674-2754: reads the same parameters with the same defaults,
!! then hands the resolved values to the device context.
subroutine continuity_PPM_init(Time, G, GV, US, param_file, diag, CS)
  use MOM_time_manager,  only : time_type
  use MOM_diag_mediator, only : diag_ctrl
  type(time_type), target, intent(in)    :: Time
  type(ocean_grid_type),   intent(in)    :: G
  type(verticalGrid_type), intent(in)    :: GV
  type(unit_scale_type),   intent(in)    :: US
  type(param_file_type),   intent(in)    :: param_file
  type(diag_ctrl), target, intent(inout) :: diag
  type(continuity_PPM_CS), intent(inout) :: CS
  character(len=40) :: mdl = "MOM_continuity_PPM"
  logical :: l
  real :: tol_eta_m
  CS%initialized = .true.
  call get_param(param_file, mdl, "MONOTONIC_CONTINUITY", l, default=.false.)       ; CS%c%monotonic = merge(1,0,l)
  call get_param(param_file, mdl, "SIMPLE_2ND_PPM_CONTINUITY", l, default=.false.)  ; CS%c%simple_2nd = merge(1,0,l)
  call get_param(param_file, mdl, "UPWIND_1ST_CONTINUITY", l, default=.false.)      ; CS%c%upwind_1st = merge(1,0,l)
  call get_param(param_file, mdl, "ETA_TOLERANCE", CS%c%tol_eta, default=0.5*GV%ke*GV%Angstrom_m, &
                 units="m", scale=GV%m_to_H)
  call get_param(param_file, mdl, "VELOCITY_TOLERANCE", CS%c%tol_vel, default=3.0e8, units="m s-1", scale=US%m_s_to_L_T)
  call get_param(param_file, mdl, "CONT_PPM_AGGRESS_ADJUST", l, default=.false.)    ; CS%c%aggress_adjust = merge(1,0,l)
  call get_param(param_file, mdl, "CONT_PPM_VOLUME_BASED_CFL", l, default=CS%c%aggress_adjust==1) ; CS%c%vol_CFL = merge(1,0,l)
  call get_param(param_file, mdl, "CONTINUITY_CFL_LIMIT", CS%c%CFL_limit_adjust, default=0.5, units="nondim")
  call get_param(param_file, mdl, "CONT_PPM_BETTER_ITER", l, default=.true.)        ; CS%c%better_iter = merge(1,0,l)
  call get_param(param_file, mdl, "CONT_PPM_USE_VISC_REM_MAX", l, default=.true.)   ; CS%c%use_visc_rem_max = merge(1,0,l)
  call get_param(param_file, mdl, "CONT_PPM_MARGINAL_FACE_AREAS", l, default=.true.) ; CS%c%marginal_faces = merge(1,0,l)
  call mom6cu_check(mom6cu_set_cs_continuity(mom6cu_ctx, CS%c), "continuity_PPM_init")
end subroutine continuity_PPM_init

!> continuity_PPM_stencil, MOM_continuity_PPM.F90:2757-2763
function continuity_PPM_stencil(CS) result(stencil)
  type(continuity_PPM_CS), intent(in) :: CS
  integer :: stencil
  stencil = 3 ; if (CS%c%simple_2nd == 1) stencil = 2 ; if (CS%c%upwind_1st == 1) stencil = 1
end function continuity_PPM_stencil

end module MOM_continuity_PPM
