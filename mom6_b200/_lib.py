"""ctypes binding of the C ABI in include/mom6cu.h (libmom6cu.so).

The product path has no CPU fallback: if the CUDA library is missing or no device is
visible, every compute entry raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmom6cu.so")

c_double_p = C.POINTER(C.c_double)


class Domain(C.Structure):
    """mom6cu_domain (include/mom6cu.h); subset of hor_index_type, src/framework/MOM_hor_index.F90."""
    _fields_ = [(n, C.c_int) for n in (
        "isc", "iec", "jsc", "jec", "isd", "ied", "jsd", "jed",
        "isdw", "iedw", "jsdw", "jedw", "nk", "cyclic_x", "cyclic_y", "first_direction",
        "npi", "npj", "pi", "pj")]


_BT_PTRS = [
    "eta", "ubt", "vbt", "uhbt0", "vhbt0", "Datu", "Datv", "BTCL_u", "BTCL_v", "eta_src", "eta_PF",
    "gtot_E", "gtot_W", "gtot_N", "gtot_S", "f_4_u", "f_4_v", "bt_rem_u", "bt_rem_v",
    "BT_force_u", "BT_force_v", "Cor_ref_u", "Cor_ref_v", "IareaT_OBCmask", "IdxCu", "IdyCv",
    "u_accel_bt", "v_accel_bt", "eta_sum", "eta_wtd", "ubtav", "vbtav", "uhbtav", "vhbtav",
    "ubt_wtd", "vbt_wtd", "wt_vel", "wt_eta", "wt_accel", "wt_trans", "wt_accel2"]
_BT_DBL = ["dtbt", "dgeo_de", "bebt", "vel_underflow"]
_BT_INT = ["nstep", "nfilter", "use_BT_cont", "find_etaav", "BT_project_velocity",
           "use_old_coriolis_bracket_bug", "use_wide_halos", "min_stencil"]


class BtTimeloopArgs(C.Structure):
    """mom6cu_bt_timeloop_args: the argument list of btstep_timeloop, MOM_barotropic.F90:2175-2182."""
    _fields_ = ([(n, C.c_void_p) for n in _BT_PTRS] + [(n, C.c_double) for n in _BT_DBL] +
                [(n, C.c_int) for n in _BT_INT])


GRID_FIELDS = [
    ("mask2dT", "h"), ("mask2dCu", "u"), ("mask2dCv", "v"), ("mask2dBu", "q"),
    ("dxT", "h"), ("dyT", "h"), ("IdxT", "h"), ("IdyT", "h"), ("areaT", "h"), ("IareaT", "h"),
    ("dxCu", "u"), ("dyCu", "u"), ("IdxCu", "u"), ("IdyCu", "u"), ("dy_Cu", "u"), ("areaCu", "u"), ("IareaCu", "u"),
    ("dxCv", "v"), ("dyCv", "v"), ("IdxCv", "v"), ("IdyCv", "v"), ("dx_Cv", "v"), ("areaCv", "v"), ("IareaCv", "v"),
    ("dxBu", "q"), ("dyBu", "q"), ("IdxBu", "q"), ("IdyBu", "q"), ("areaBu", "q"), ("IareaBu", "q"),
    ("bathyT", "h"), ("CoriolisBu", "q"), ("Coriolis2Bu", "q")]


class Grid(C.Structure):
    """mom6cu_grid: the ocean_grid_type metrics the hot path reads (src/core/MOM_grid.F90:75-175)."""
    _fields_ = [(n, C.c_void_p) for n, _ in GRID_FIELDS]


class VGrid(C.Structure):
    """mom6cu_vgrid: verticalGrid_type scalars (src/core/MOM_verticalGrid.F90)."""
    _fields_ = [(n, C.c_double) for n in ("Angstrom_H", "H_subroundoff", "Z_to_H", "H_to_Z", "g_Earth", "Rho0",
                                          "H_to_RZ", "RZ_to_H", "H_to_m", "m_to_H")] + [("Boussinesq", C.c_int)]


class ContinuityCS(C.Structure):
    """mom6cu_continuity_cs: continuity_PPM_CS (MOM_continuity_PPM.F90:35-67)."""
    _fields_ = [(n, C.c_int) for n in ("upwind_1st", "monotonic", "simple_2nd", "aggress_adjust", "vol_CFL",
                                       "better_iter", "use_visc_rem_max", "marginal_faces")] + \
               [(n, C.c_double) for n in ("tol_eta", "tol_vel", "CFL_limit_adjust")]


BT_CONT_FIELDS = ["FA_u_EE", "FA_u_E0", "FA_u_W0", "FA_u_WW", "uBT_WW", "uBT_EE",
                  "FA_v_NN", "FA_v_N0", "FA_v_S0", "FA_v_SS", "vBT_SS", "vBT_NN", "h_u", "h_v"]


class BTCont(C.Structure):
    """mom6cu_bt_cont: BT_cont_type (src/core/MOM_variables.F90:315-350)."""
    _fields_ = [(n, C.c_void_p) for n in BT_CONT_FIELDS]


class ContinuityArgs(C.Structure):
    """mom6cu_continuity_args: the dummy arguments of continuity_PPM (MOM_continuity_PPM.F90:86-141)."""
    _fields_ = ([(n, C.c_void_p) for n in ("u", "v", "hin", "h", "uh", "vh")] + [("dt", C.c_double)] +
                [(n, C.c_void_p) for n in ("por_face_areaU", "por_face_areaV", "uhbt", "vhbt", "visc_rem_u",
                                           "visc_rem_v", "u_cor", "v_cor")] +
                [("BT_cont", C.POINTER(BTCont))] + [(n, C.c_void_p) for n in ("du_cor", "dv_cor")])


class UnitScale(C.Structure):
    """mom6cu_unit_scale: unit_scale_type factors (src/framework/MOM_unit_scaling.F90)."""
    _fields_ = [(n, C.c_double) for n in ("m_to_L", "L_to_m", "m_s_to_L_T", "L_T_to_m_s", "s_to_T", "T_to_s",
                                          "m_to_Z", "Z_to_m", "Z_to_L", "L_to_Z")]


# CoriolisAdv enumerations, src/core/MOM_CoriolisAdv.F90:94-119
SADOURNY75_ENERGY, ARAKAWA_HSU90, ROBUST_ENSTRO, SADOURNY75_ENSTRO, ARAKAWA_LAMB81, AL_BLEND = 1, 2, 3, 4, 5, 6
KE_ARAKAWA, KE_SIMPLE_GUDONOV, KE_GUDONOV = 10, 11, 12
PV_ADV_CENTERED, PV_ADV_UPWIND1 = 21, 22


class CoriolisAdvCS(C.Structure):
    """mom6cu_coriolisadv_cs: CoriolisAdv_CS (MOM_CoriolisAdv.F90:30-91)."""
    _fields_ = [(n, C.c_int) for n in ("Coriolis_Scheme", "KE_Scheme", "PV_Adv_Scheme", "no_slip", "bound_Coriolis",
                                       "Coriolis_En_Dis")] + \
               [(n, C.c_double) for n in ("F_eff_max_blend", "wt_lin_blend")]


class CorAdCalcArgs(C.Structure):
    """mom6cu_coradcalc_args: the dummy arguments of CorAdCalc (MOM_CoriolisAdv.F90:125-144)."""
    _fields_ = [(n, C.c_void_p) for n in ("u", "v", "h", "uh", "vh", "CAu", "CAv", "por_face_areaU", "por_face_areaV",
                                          "RV", "PV", "gradKEu", "gradKEv")]


HV_FLAGS = ["Laplacian", "biharmonic", "no_slip", "bound_Kh", "better_bound_Kh", "bound_Ah", "better_bound_Ah",
            "backscatter_underbound", "Smagorinsky_Kh", "Smagorinsky_Ah", "bound_Coriolis", "use_land_mask",
            "add_LES_viscosity", "use_cont_thick", "use_cont_thick_bug", "unsupported"]
HV_H = ["dx2h", "dy2h", "DX_dyT", "DY_dxT", "reduction_xx", "Kh_bg_xx", "Ah_bg_xx", "Kh_Max_xx", "Ah_Max_xx",
        "Laplac2_const_xx", "Biharm_const_xx", "Biharm_const2_xx", "Re_Ah_const_xx"]
HV_Q = ["dx2q", "dy2q", "DX_dyBu", "DY_dxBu", "reduction_xy", "Kh_bg_xy", "Ah_bg_xy", "Kh_Max_xy", "Ah_Max_xy",
        "Laplac2_const_xy", "Biharm_const_xy", "Biharm_const2_xy", "Re_Ah_const_xy"]
HV_UV = ["Idx2dyCu", "Idxdy2u", "Idx2dyCv", "Idxdy2v"]


class HorViscCS(C.Structure):
    """mom6cu_hor_visc_cs: hor_visc_CS (MOM_hor_visc.F90:38-250) as resolved by hor_visc_init (:2322)."""
    _fields_ = ([(n, C.c_int) for n in HV_FLAGS] + [("Kh_bg_min", C.c_double), ("Re_Ah", C.c_double)] +
                [(n, C.c_void_p) for n in HV_H + HV_Q + HV_UV])


class HorViscArgs(C.Structure):
    """mom6cu_hor_visc_args: the dummy arguments of horizontal_viscosity (MOM_hor_visc.F90:266)."""
    _fields_ = [(n, C.c_void_p) for n in ("u", "v", "h", "uh", "vh", "diffu", "diffv", "hu_cont", "hv_cont")] + \
               [("dt", C.c_double)]


BT_CS_INT = ["Sadourny", "BT_project_velocity", "strong_drag", "bound_BT_corr", "BT_cont_bounds", "wt_uv_bug",
             "visc_rem_u_uh0", "adjust_BT_cont", "use_wide_halos", "min_stencil", "use_old_coriolis_bracket_bug",
             "unsupported"]
BT_CS_DBL = ["dtbt", "bebt", "vel_underflow", "maxCFL_BT_cont", "G_extra", "dt_bt_filter"]
BT_CS_PTR = ["IareaT", "IareaT_OBCmask", "bathyT", "IdxCu", "IdyCv", "q_D", "D_u_Cor", "D_v_Cor", "ua_polarity",
             "va_polarity", "OBCmask_u", "OBCmask_v", "frhatu", "frhatv", "eta_cor", "eta_cor_bound", "IDatu", "IDatv",
             "ubtav", "vbtav"]


class BarotropicCS(C.Structure):
    """mom6cu_barotropic_cs: barotropic_CS (MOM_barotropic.F90:112-364)."""
    _fields_ = ([(n, C.c_int) for n in BT_CS_INT] + [(n, C.c_double) for n in BT_CS_DBL] +
                [(n, C.c_void_p) for n in BT_CS_PTR])


class BtstepArgs(C.Structure):
    """mom6cu_btstep_args: the dummy arguments of btstep (MOM_barotropic.F90:455-529)."""
    _fields_ = ([(n, C.c_void_p) for n in ("U_in", "V_in", "eta_in")] + [("dt", C.c_double)] +
                [(n, C.c_void_p) for n in ("bc_accel_u", "bc_accel_v", "taux", "tauy", "pbce", "eta_PF_in", "U_Cor", "V_Cor",
                                           "accel_layer_u", "accel_layer_v", "eta_out", "uhbtav", "vhbtav", "visc_rem_u",
                                           "visc_rem_v")] +
                [("BT_cont", C.POINTER(BTCont))] +
                [(n, C.c_void_p) for n in ("taux_bot", "tauy_bot", "uh0", "vh0", "u_uh0", "v_vh0", "etaav")])


class BtcalcArgs(C.Structure):
    """mom6cu_btcalc_args: btcalc (MOM_barotropic.F90:4360)."""
    _fields_ = [(n, C.c_void_p) for n in ("h", "h_u", "h_v", "frhatu", "frhatv", "bathyT")] + \
               [("hvel_scheme", C.c_int), ("may_use_default", C.c_int)]


EOS_NONE, EOS_LINEAR, EOS_WRIGHT = 0, 1, 3


class PressureForceCS(C.Structure):
    """mom6cu_pressureforce_cs: PressureForce_FV_CS (MOM_PressureForce_FV.F90:40-107) + EOS / vertical-grid members."""
    _fields_ = ([(n, C.c_int) for n in ("EOS_form", "MassWghtInterp", "use_SSH_in_Z0p", "rho_ref_bug", "unsupported")] +
                [(n, C.c_double) for n in ("rho_ref", "GFS_scale", "Z_ref", "dZ_subroundoff", "Rho_T0_S0", "dRho_dT", "dRho_dS",
                                           "dRho_dp")] + [(n, C.c_void_p) for n in ("Rlay", "g_prime")] +
                [(n, C.c_int) for n in ("reconstruct", "Recon_Scheme", "boundary_extrap", "use_inaccurate_pgf_rho_anom",
                                        "MassWghtInterpVanOnly", "ALE_answer_date")] +
                [(n, C.c_double) for n in ("h_nonvanished", "kg_m3_to_R", "RL2_T2_to_Pa", "C_to_degC", "S_to_ppt")])


class PressureForceArgs(C.Structure):
    """mom6cu_pressureforce_args: the dummy arguments of PressureForce (MOM_PressureForce.F90:40)."""
    _fields_ = [(n, C.c_void_p) for n in ("h", "T", "S", "PFu", "PFv", "p_atm", "pbce", "eta")]


REMAPPING_PCM, REMAPPING_PLM, REMAPPING_PPM_H4, REMAPPING_PPM_IH4 = 0, 2, 4, 5


class RemappingCS(C.Structure):
    """mom6cu_remapping_cs: remapping_CS (src/ALE/MOM_remapping.F90:37-85)."""
    _fields_ = [(n, C.c_int) for n in ("remapping_scheme", "boundary_extrapolation", "force_bounds_in_subcell",
                                       "force_bounds_in_target", "om4_remap_via_sub_cells", "answer_date")] + \
               [(n, C.c_double) for n in ("h_neglect", "h_neglect_edge")]


class DynSplitRK2CS(C.Structure):
    """mom6cu_dyn_split_rk2_cs: MOM_dyn_split_RK2_CS members (src/core/MOM_dynamics_split_RK2.F90:85-273)."""
    _fields_ = [("be", C.c_double), ("begw", C.c_double)] + \
               [(n, C.c_int) for n in ("split_bottom_stress", "store_CAu", "CAu_pred_stored", "visc_rem_dt_bug", "hvel_scheme", "unsupported",
                                       "dtbt_use_bt_cont", "BT_Nonlinear_continuity")] + \
               [(n, C.c_double) for n in ("dtbt_fraction", "BT_Coriolis_scale", "Z_ref", "dtbt_max")] + \
               [(n, C.c_void_p) for n in ("CAu", "CAv", "CAu_pred", "CAv_pred", "PFu", "PFv", "diffu", "diffv", "visc_rem_u", "visc_rem_v",
                                          "u_accel_bt", "v_accel_bt", "u_av", "v_av", "h_av", "pbce", "eta", "eta_PF", "uhbt", "vhbt",
                                          "taux_bot", "tauy_bot")] + \
               [("BT_cont", C.POINTER(BTCont)), ("barotropic", C.POINTER(BarotropicCS))]


class SetDtbtArgs(C.Structure):
    """mom6cu_set_dtbt_args: set_dtbt (MOM_barotropic.F90:3509)."""
    _fields_ = [("pbce", C.c_void_p), ("gtot_est", C.c_double), ("have_gtot_est", C.c_int), ("BT_cont", C.POINTER(BTCont)), ("eta", C.c_void_p),
                ("SSH_add", C.c_double), ("frhatu", C.c_void_p), ("frhatv", C.c_void_p), ("bathyT", C.c_void_p)] + \
               [(n, C.c_double) for n in ("bebt", "G_extra", "dtbt_fraction", "BT_Coriolis_scale", "Z_ref")] + [("Nonlinear_continuity", C.c_int)]


class StepDynArgs(C.Structure):
    """mom6cu_step_dyn_args: the arguments of step_MOM_dyn_split_RK2 (:294-296)."""
    _fields_ = [(n, C.c_void_p) for n in ("u_inst", "v_inst", "h", "T", "S", "Kv_bbl_u", "Kv_bbl_v", "bbl_thick_u", "bbl_thick_v", "Kv_shear",
                                          "Kv_shear_Bu", "Ray_u", "Ray_v", "taux", "tauy", "ustar", "p_surf")] + [("dt", C.c_double)] + \
               [(n, C.c_void_p) for n in ("uh", "vh", "uhtr", "vhtr", "eta_av")] + [("calc_dtbt", C.c_int)]


class VertviscCS(C.Structure):
    """mom6cu_vertvisc_cs: vertvisc_CS members (src/parameterizations/vertical/MOM_vert_friction.F90:48-170)."""
    _fields_ = [(n, C.c_int) for n in ("bottomdraglaw", "harmonic_visc", "direct_stress", "fixed_LOTW_ML", "apply_LOTW_floor",
                                       "dynamic_viscous_ML", "nkml", "answer_date", "unsupported", "CFL_based_trunc")] + \
               [(n, C.c_double) for n in ("Hbbl", "Kv", "Kv_extra_bbl", "Kvml_invZ2", "Hmix", "Hmix_stress", "harm_BL_val", "vonKar",
                                          "vel_underflow", "dZ_subroundoff", "maxvel", "CFL_trunc")]


class VertviscCoefArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("u", "v", "h", "Kv_bbl_u", "Kv_bbl_v", "bbl_thick_u", "bbl_thick_v", "Kv_shear", "Kv_shear_Bu",
                                          "ustar")] + [("dt", C.c_double)]


class VertviscArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("u", "v", "h", "taux", "tauy", "Ray_u", "Ray_v")] + [("dt", C.c_double)] + \
               [(n, C.c_void_p) for n in ("taux_bot", "tauy_bot")]


class RegriddingCS(C.Structure):
    """mom6cu_regridding_cs: the regridding_CS members of the Z* path (src/ALE/MOM_regridding.F90:49-160)."""
    _fields_ = [("regridding_scheme", C.c_int), ("nk", C.c_int), ("min_thickness", C.c_double), ("old_grid_weight", C.c_double),
                ("depth_of_time_filter_shallow", C.c_double), ("depth_of_time_filter_deep", C.c_double), ("Z_ref", C.c_double),
                ("coordinateResolution", C.c_void_p)]


class TracerAdvectCS(C.Structure):
    """mom6cu_tracer_advect_cs: tracer_advect_CS (src/tracer/MOM_tracer_advect.F90:32-41)."""
    _fields_ = [("dt", C.c_double), ("default_advect_scheme", C.c_int), ("useHuynhStencilBug", C.c_int)]


class AdvectTracerArgs(C.Structure):
    """mom6cu_advect_tracer_args: the arguments of advect_tracer (MOM_tracer_advect.F90:53-54)."""
    _fields_ = [("h_end", C.c_void_p), ("uhtr", C.c_void_p), ("vhtr", C.c_void_p), ("dt", C.c_double), ("ntr", C.c_int),
                ("tr", C.POINTER(C.c_void_p)), ("advect_scheme", C.POINTER(C.c_int)), ("conc_underflow", C.c_void_p),
                ("x_first_in", C.c_int), ("max_iter_in", C.c_int), ("vol_prev", C.c_void_p), ("update_vol_prev", C.c_int),
                ("uhr_out", C.c_void_p), ("vhr_out", C.c_void_p)]


class AleCS(C.Structure):
    """mom6cu_ale_cs: the ALE_CS members ALE_regridding_and_remapping uses (src/ALE/MOM_ALE.F90:65-130)."""
    _fields_ = [("regridCS", RegriddingCS), ("remapCS", RemappingCS), ("vel_remapCS", RemappingCS), ("regrid_time_scale", C.c_double),
                ("remap_uv_using_old_alg", C.c_int), ("do_conv_adj", C.c_int), ("use_hybgen_unmix", C.c_int), ("remap_aux_vars", C.c_int)]


class AleArgs(C.Structure):
    """mom6cu_ale_args: the state ALE_regridding_and_remapping updates (src/core/MOM.F90:1751)."""
    _fields_ = [("u", C.c_void_p), ("v", C.c_void_p), ("h", C.c_void_p), ("ntr", C.c_int), ("tr", C.POINTER(C.c_void_p)),
                ("conc_underflow", C.c_void_p), ("iT", C.c_int), ("iS", C.c_int), ("dtdia", C.c_double), ("Kd_shear", C.c_void_p),
                ("Kv_shear", C.c_void_p), ("Kv_shear_Bu", C.c_void_p)]


class MleCS(C.Structure):
    """mom6cu_mle_cs: mixedlayer_restrat_CS (src/parameterizations/lateral/MOM_mixed_layer_restrat.F90:42-115), OM4 path."""
    _fields_ = ([(n, C.c_double) for n in ("ml_restrat_coef", "ml_restrat_coef2", "front_length", "MLE_MLD_decay_time",
                                           "MLE_MLD_decay_time2", "MLE_MLD_stretch", "MLE_tail_dh", "ustar_min", "vonKar",
                                           "MLE_density_diff")] +
                [(n, C.c_int) for n in ("MLE_use_PBL_MLD", "use_Stanley_ML", "use_Bodner", "fl_from_file", "EOS_form")] +
                [(n, C.c_double) for n in ("Rho_T0_S0", "dRho_dT", "dRho_dS", "dRho_dp")] +
                [("MLD_filtered", C.c_void_p), ("MLD_filtered_slow", C.c_void_p)])


class TracerHorDiffCS(C.Structure):
    """mom6cu_tracer_hor_diff_cs: tracer_hor_diff_CS (src/tracer/MOM_tracer_hor_diff.F90:40-106) + the VarMix switches it reads."""
    _fields_ = ([(n, C.c_double) for n in ("KhTr", "KhTr_min", "KhTr_max", "KhTr_passivity_coeff", "KhTr_passivity_min", "KhTr_Slope_Cff",
                                           "max_diff_CFL")] +
                [(n, C.c_int) for n in ("check_diffusive_CFL", "use_neutral_diffusion", "use_hor_bnd_diffusion", "Diffuse_ML_interior",
                                        "use_variable_mixing", "Resoln_scaled_KhTr", "use_MEKE_Kh")] + [("MEKE_KhTr_fac", C.c_double)])


class TracerHordiffArgs(C.Structure):
    """mom6cu_tracer_hordiff_args: the arguments of tracer_hordiff (MOM_tracer_hor_diff.F90:119)."""
    _fields_ = [("h", C.c_void_p), ("dt", C.c_double), ("ntr", C.c_int), ("tr", C.POINTER(C.c_void_p)), ("conc_underflow", C.c_void_p),
                ("Res_fn_h", C.c_void_p), ("Rd_dx_h", C.c_void_p), ("df_x", C.POINTER(C.c_void_p)), ("df_y", C.POINTER(C.c_void_p)),
                ("L2u", C.c_void_p), ("SN_u", C.c_void_p), ("L2v", C.c_void_p), ("SN_v", C.c_void_p), ("MEKE_Kh", C.c_void_p)]


class ThicknessDiffuseCS(C.Structure):
    """mom6cu_thickness_diffuse_cs: thickness_diffuse_CS (MOM_thickness_diffuse.F90:40-131) + the VarMix / MEKE / EOS switches it reads."""
    _fields_ = ([(n, C.c_double) for n in ("Khth", "Khth_Min", "Khth_Max", "max_Khth_CFL", "slope_max", "kappa_smooth", "dZ_subroundoff")] +
                [(n, C.c_int) for n in ("thickness_diffuse", "read_khth", "detangle_interfaces", "interface_Kh", "use_FGNV_streamfn", "use_stanley_gm",
                                        "use_GME_thickness_diffuse", "find_work", "use_variable_mixing", "Resoln_scaled_KhTh", "Depth_scaled_KhTh",
                                        "use_stored_slopes", "use_Visbeck", "use_QG_Leith_GM", "khth_struct", "use_MEKE_Kh", "EOS_form")] +
                [(n, C.c_double) for n in ("Rho_T0_S0", "dRho_dT", "dRho_dS", "dRho_dp", "FGNV_scale", "N2_floor", "MEKE_KhTh_fac")])


class ThicknessDiffuseArgs(C.Structure):
    """mom6cu_thickness_diffuse_args: the arguments of thickness_diffuse (MOM_thickness_diffuse.F90:134)."""
    _fields_ = [("h", C.c_void_p), ("uhtr", C.c_void_p), ("vhtr", C.c_void_p), ("T", C.c_void_p), ("S", C.c_void_p), ("p_surf", C.c_void_p),
                ("dt", C.c_double), ("Res_fn_u", C.c_void_p), ("Res_fn_v", C.c_void_p), ("uhGM", C.c_void_p), ("vhGM", C.c_void_p),
                ("slope_x", C.c_void_p), ("slope_y", C.c_void_p), ("cg1", C.c_void_p), ("MEKE_Kh", C.c_void_p)]


class Efp(C.Structure):
    """mom6cu_efp: EFP_type (src/framework/MOM_coms.F90:76-78)."""
    _fields_ = [("v", C.c_int64 * 6)]


_SO_UNITS = ("RZL2_to_kg", "L_T_to_m_s", "Q_to_J_kg", "J_kg_to_Q", "kg_m3_to_R", "m_to_Z", "m_to_L", "Z_to_m", "S_to_ppt", "C_to_degC")
_SO_EFPS = ("fresh_water_in_EFP", "net_salt_in_EFP", "net_heat_in_EFP", "mass_prev_EFP", "salt_prev_EFP", "heat_prev_EFP")


class SumOutputCS(C.Structure):
    """mom6cu_sum_output_cs: Sum_output_CS (src/diagnostics/MOM_sum_output.F90:66-140) as write_energy uses it."""
    _fields_ = ([("do_APE_calc", C.c_int), ("use_temperature", C.c_int), ("dt_in_T", C.c_double), ("DL_listsize", C.c_int),
                 ("DL_depth", C.c_void_p), ("DL_area", C.c_void_p), ("DL_vol_below", C.c_void_p), ("lH", C.c_void_p),
                 ("g_prime", C.c_void_p), ("Z_ref", C.c_double), ("C_p", C.c_double)] +
                [(n, C.c_double) for n in _SO_UNITS] + [("previous_calls", C.c_int), ("ntrunc", C.c_int)] +
                [(n, Efp) for n in _SO_EFPS])


_EO_SCALARS = ("En_mass", "toten", "KE_tot", "PE_tot", "mass_tot", "mass_chg", "mass_anom")
_EO_SCALARS2 = ("Salt", "Salt_chg", "Salt_anom", "Heat", "Heat_chg", "Heat_anom", "salin", "salin_anom", "temp", "temp_anom")


class EnergyOut(C.Structure):
    """mom6cu_energy_out: what write_energy puts on the ocean.stats line and in the energy file."""
    _fields_ = ([(n, C.c_double) for n in _EO_SCALARS] + [("max_CFL", C.c_double * 2)] + [(n, C.c_double) for n in _EO_SCALARS2] +
                [("ntrunc", C.c_int)] + [(n, C.c_void_p) for n in ("KE", "mass_lay", "PE", "Z_0APE")])


def fill_struct(struct, values, keep):
    """Fill a ctypes struct from a dict: numpy arrays / torch tensors -> pointers, scalars as is."""
    for name, ctype in struct._fields_:
        v = values.get(name)
        if ctype is not C.c_void_p and ctype not in (C.c_int, C.c_double):
            continue  # nested struct pointers are filled by the caller
        if ctype is C.c_void_p:
            if v is None:
                setattr(struct, name, None)
            elif isinstance(v, int):               # raw device pointer (resident plane)
                setattr(struct, name, v)
            elif hasattr(v, "ptr") and isinstance(getattr(v, "ptr"), int):   # api.Plane
                setattr(struct, name, v.ptr)
            elif hasattr(v, "data_ptr"):          # torch tensor (device or host)
                keep.append(v)
                setattr(struct, name, v.data_ptr())
            else:                                  # numpy array
                import numpy as np
                if not (isinstance(v, np.ndarray) and v.dtype == np.float64 and v.flags["C_CONTIGUOUS"]):
                    raise TypeError(f"{name}: expected a C-contiguous float64 array")
                keep.append(v)
                setattr(struct, name, v.ctypes.data)
        else:
            if v is None:
                raise KeyError(f"missing scalar argument {name}")
            setattr(struct, name, v)
    return struct


_lib = None


def bind(lib):
    """Attach restype/argtypes to every exported entry point of a loaded library."""
    vp = C.c_void_p
    lib.mom6cu_create.argtypes = [C.POINTER(vp), C.POINTER(Domain), C.c_int]
    lib.mom6cu_destroy.argtypes = [vp]
    lib.mom6cu_last_error.argtypes = [vp, C.c_char_p, C.c_size_t]
    lib.mom6cu_build_arch.argtypes = []
    lib.mom6cu_launch_count.argtypes = [vp]
    lib.mom6cu_launch_count.restype = C.c_longlong
    lib.mom6cu_sync.argtypes = [vp]
    lib.mom6cu_last_kernel_ms.argtypes = [vp]
    lib.mom6cu_last_iterations.argtypes = [vp]
    lib.mom6cu_last_kernel_ms.restype = C.c_double
    lib.mom6cu_last_step_stage_ms.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    lib.mom6cu_last_step_stage_ms.restype = C.c_int
    lib.mom6cu_total_kernel_ms.argtypes = [vp]
    lib.mom6cu_total_kernel_ms.restype = C.c_double
    lib.mom6cu_set_grid.argtypes = [vp, C.POINTER(Grid)]
    lib.mom6cu_set_vgrid.argtypes = [vp, C.POINTER(VGrid)]
    lib.mom6cu_set_cs_continuity.argtypes = [vp, C.POINTER(ContinuityCS)]
    lib.mom6cu_continuity.argtypes = [vp, C.POINTER(ContinuityArgs)]
    lib.mom6cu_set_unit_scale.argtypes = [vp, C.POINTER(UnitScale)]
    lib.mom6cu_set_cs_coriolisadv.argtypes = [vp, C.POINTER(CoriolisAdvCS)]
    lib.mom6cu_coradcalc.argtypes = [vp, C.POINTER(CorAdCalcArgs)]
    lib.mom6cu_set_cs_hor_visc.argtypes = [vp, C.POINTER(HorViscCS)]
    lib.mom6cu_horizontal_viscosity.argtypes = [vp, C.POINTER(HorViscArgs)]
    lib.mom6cu_btstep.argtypes = [vp, C.POINTER(BarotropicCS), C.POINTER(BtstepArgs)]
    lib.mom6cu_btcalc.argtypes = [vp, C.POINTER(BtcalcArgs)]
    lib.mom6cu_bt_mass_source.argtypes = [vp, vp, vp, C.c_int, vp]
    lib.mom6cu_plane_alloc.argtypes = [vp, C.c_char_p, C.c_int]
    lib.mom6cu_plane_alloc.restype = C.c_void_p
    lib.mom6cu_plane_upload.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int]
    lib.mom6cu_plane_download.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int]
    lib.mom6cu_plane_zero.argtypes = [vp, vp, C.c_int]
    lib.mom6cu_sizeof.argtypes = [C.c_char_p]
    lib.mom6cu_sizeof.restype = C.c_longlong
    lib.mom6cu_set_cs_pressureforce.argtypes = [vp, C.POINTER(PressureForceCS)]
    lib.mom6cu_pressure_force.argtypes = [vp, C.POINTER(PressureForceArgs)]
    lib.mom6cu_ale_remap_tracers.argtypes = [vp, C.POINTER(RemappingCS), vp, vp, C.c_int, C.POINTER(vp), vp]
    lib.mom6cu_ale_remap_set_h_vel.argtypes = [vp, vp, vp, vp]
    lib.mom6cu_ale_remap_velocities.argtypes = [vp, C.POINTER(RemappingCS), vp, vp, vp, vp, vp, vp]
    lib.mom6cu_remapping_core_h.argtypes = [vp, C.POINTER(RemappingCS), C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
    lib.mom6cu_set_dtbt.argtypes = [vp, C.POINTER(SetDtbtArgs), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.mom6cu_remap_dyn_split_rk2_aux_vars.argtypes = [vp, C.POINTER(RemappingCS), C.POINTER(DynSplitRK2CS), vp, vp, vp, vp]
    lib.mom6cu_step_dyn_split_rk2.argtypes = [vp, C.POINTER(DynSplitRK2CS), C.POINTER(StepDynArgs)]
    lib.mom6cu_set_cs_vertvisc.argtypes = [vp, C.POINTER(VertviscCS)]
    lib.mom6cu_vertvisc_coef.argtypes = [vp, C.POINTER(VertviscCoefArgs)]
    lib.mom6cu_vertvisc_get_coef.argtypes = [vp, vp, vp, vp, vp]
    lib.mom6cu_vertvisc_ntrunc.argtypes = [vp]
    lib.mom6cu_vertvisc_ntrunc.restype = C.c_longlong
    lib.mom6cu_vertvisc.argtypes = [vp, C.POINTER(VertviscArgs)]
    lib.mom6cu_vertvisc_remnant.argtypes = [vp, vp, vp, vp, vp, C.c_double]
    lib.mom6cu_ale_regrid.argtypes = [vp, C.POINTER(RegriddingCS), vp, vp, vp]
    lib.mom6cu_advect_tracer.argtypes = [vp, C.POINTER(TracerAdvectCS), C.POINTER(AdvectTracerArgs)]
    lib.mom6cu_efp_plus.argtypes = [C.POINTER(Efp), C.POINTER(Efp), C.POINTER(Efp), C.POINTER(C.c_int)]
    lib.mom6cu_efp_plus.restype = None
    lib.mom6cu_efp_minus.argtypes = lib.mom6cu_efp_plus.argtypes
    lib.mom6cu_efp_minus.restype = None
    lib.mom6cu_efp_to_real.argtypes = [C.POINTER(Efp)]
    lib.mom6cu_efp_to_real.restype = C.c_double
    lib.mom6cu_real_to_efp.argtypes = [C.c_double, C.POINTER(Efp)]
    lib.mom6cu_efp_real_diff.argtypes = [C.POINTER(Efp), C.POINTER(Efp)]
    lib.mom6cu_efp_real_diff.restype = C.c_double
    lib.mom6cu_reproducing_sum.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                           C.POINTER(C.c_double), vp, C.POINTER(Efp), C.POINTER(Efp)]
    lib.mom6cu_efp_sum_across_pes.argtypes = [vp, C.POINTER(Efp), C.c_int]
    lib.mom6cu_chksum.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.POINTER(C.c_double)]
    lib.mom6cu_write_energy.argtypes = [vp, C.POINTER(SumOutputCS), vp, vp, vp, vp, vp, C.POINTER(EnergyOut)]
    lib.mom6cu_ocean_stats_line.argtypes = [C.POINTER(SumOutputCS), C.POINTER(EnergyOut), C.c_int, C.c_double, C.c_char_p, C.c_size_t]
    lib.mom6cu_interpolate_column.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int]
    lib.mom6cu_ale_remap_interface_vals.argtypes = [vp, vp, vp, vp]
    lib.mom6cu_ale_remap_vertex_vals.argtypes = [vp, vp, vp, vp]
    lib.mom6cu_ale_regridding_and_remapping.argtypes = [vp, C.POINTER(AleCS), C.POINTER(DynSplitRK2CS), C.POINTER(AleArgs)]
    lib.mom6cu_mixedlayer_restrat.argtypes = [vp, C.POINTER(MleCS), vp, vp, vp, vp, vp, vp, C.c_double, vp, vp]
    lib.mom6cu_mle_mu.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.mom6cu_tracer_hordiff.argtypes = [vp, C.POINTER(TracerHorDiffCS), C.POINTER(TracerHordiffArgs)]
    lib.mom6cu_thickness_diffuse.argtypes = [vp, C.POINTER(ThicknessDiffuseCS), C.POINTER(ThicknessDiffuseArgs)]
    lib.mom6cu_do_group_pass.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    lib.mom6cu_comm_unique_id.argtypes = [C.c_char_p, C.c_int]
    lib.mom6cu_comm_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.mom6cu_comm_destroy.argtypes = [vp]
    lib.mom6cu_halo_plan.argtypes = [C.POINTER(Domain), C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.mom6cu_btstep_timeloop.argtypes = [vp, C.POINTER(BtTimeloopArgs)]
    lib.mom6cu_btstep_timeloop_resident.argtypes = [vp, C.POINTER(BtTimeloopArgs), C.c_int, C.c_int]
    return lib


def _preload_nccl():
    """libmom6cu.so needs libnccl.so.2.  PyTorch bundles a newer NCCL under the same SONAME; whichever is loaded
    first wins for the whole process, so load the bundled (newer) one first when it exists -- otherwise a later
    `import torch` would bind to the older system library and fail on missing symbols."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(loc, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def load():
    """Load libmom6cu.so (built in-tree by __graft_entry__.build()); fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        _preload_nccl()
        _lib = bind(C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL))
    return _lib
