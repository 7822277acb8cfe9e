"""Seeded cases shared by tests/test_reference_f90.py (oracle vs the translated reference, live), by
tests/golden/make_reference_digests.py (writes the digests of the reference's outputs) and by tests/test_reference_golden.py
(oracle vs those digests, runs anywhere)."""
import hashlib

import numpy as np

from mom6_b200 import synthetic

CASES = {}


# cases whose reference run takes minutes in the translator: run by tests/golden/make_reference_digests.py (and by
# tests/test_reference_f90.py with F90RUN_SLOW=1); the oracle and the CUDA path are always compared with their digests
SLOW = set()

DEVICE_REFUSES = {
    # options the oracle restates (and the reference run confirms) but the CUDA path declines with MOM6CU_ERR_UNSUPPORTED
    "mixedlayer_restrat/options03": "MLE_TAIL_DH /= 0",
}


def case(name, stage, shape, outputs, slow=False, **kw):
    CASES[name] = dict(stage=stage, shape=shape, outputs=outputs, kw=kw)
    if slow:
        SLOW.add(name)


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_copy(v) for v in x]
    return x


def inner(dom, x):
    """the computational domain of an array on G's memory domain, including the symmetric edge of staggered fields"""
    ni, nj = dom.ied - dom.isd + 1, dom.jed - dom.jsd + 1
    di, dj = x.shape[-1] - ni, x.shape[-2] - nj
    j0, i0 = dom.jsc - dom.jsd, dom.isc - dom.isd
    return x[..., j0:j0 + (dom.jec - dom.jsc + 1) + dj, i0:i0 + (dom.iec - dom.isc + 1) + di]


def collect(dom, outputs, a, cs):
    out = {}
    for key in outputs:
        src, k = (cs, key[3:]) if key.startswith("CS%") else (a, key)
        if "." in k:
            d, m = k.split(".")
            if isinstance(src.get(d), list):
                v = src[d][int(m)] if int(m) < len(src[d]) else None
            elif src.get(d) is None:
                v = None
            else:
                v = src[d][m] if src.get(d) is not None else None
        else:
            v = src.get(k)
        if v is not None:
            out[key] = np.ascontiguousarray(inner(dom, v))
    return out


def digest(arrs):
    """sha256 over the outputs with -0.0 canonicalised to +0.0 (see tests/test_reference_f90.py)"""
    h = hashlib.sha256()
    for k in sorted(arrs):
        h.update(k.encode())
        h.update(np.ascontiguousarray(arrs[k] + 0.0).tobytes())
    return h.hexdigest()


# ---- continuity_PPM ------------------------------------------------------------------------------------------------------
CONT_OUT = ("h", "uh", "vh", "u_cor", "v_cor", "du_cor", "dv_cor") + tuple(
    "BT_cont." + k for k in ("FA_u_EE", "FA_u_E0", "FA_u_W0", "FA_u_WW", "uBT_WW", "uBT_EE", "FA_v_NN", "FA_v_N0", "FA_v_S0",
                             "FA_v_SS", "vBT_SS", "vBT_NN", "h_u", "h_v"))
case("continuity/default", "continuity", (24, 20, 10), CONT_OUT, land_blocks=2)
case("continuity/monotonic_volCFL_y_first", "continuity", (20, 16, 8), CONT_OUT, land_blocks=2,
     cs_over=dict(monotonic=1, vol_CFL=1, marginal_faces=0), first_direction=1)
case("continuity/simple2nd_aggress_no_visc_rem", "continuity", (20, 16, 8), CONT_OUT, land_blocks=2,
     cs_over=dict(simple_2nd=1, aggress_adjust=1, better_iter=0, use_visc_rem_max=0), with_visc_rem=False)
case("continuity/upwind_closed_no_uhbt", "continuity", (20, 16, 8), CONT_OUT, land_blocks=2, cs_over=dict(upwind_1st=1),
     with_uhbt=False, cyclic_x=False)
case("continuity/no_BT_cont_alias_h", "continuity", (20, 16, 6), CONT_OUT, land_blocks=1, with_BT_cont=False, alias_h=True)
case("continuity/75_layers", "continuity", (12, 10, 75), CONT_OUT, land_blocks=1)

# ---- CorAdCalc -----------------------------------------------------------------------------------------------------------
COR_OUT = ("CAu", "CAv", "gradKEu", "gradKEv")
for _sch in (1, 2, 3, 4, 5, 6):
    for _ke in (10, 11, 12):
        case(f"coradcalc/scheme{_sch}_ke{_ke}", "coradcalc", (16, 12, 3), COR_OUT, land_blocks=2, diags=(_ke == 10),
             cs_over=dict(Coriolis_Scheme=_sch, KE_Scheme=_ke))
case("coradcalc/pv_adv_upwind1", "coradcalc", (16, 12, 3), COR_OUT, land_blocks=2, cs_over=dict(PV_Adv_Scheme=22))
case("coradcalc/no_slip_bound_closed", "coradcalc", (16, 12, 3), COR_OUT, land_blocks=2, cyclic_x=False,
     cs_over=dict(no_slip=1, bound_Coriolis=1))
case("coradcalc/en_dis_porous", "coradcalc", (16, 12, 3), COR_OUT, land_blocks=2, por=True, cs_over=dict(Coriolis_En_Dis=1))

# ---- btstep --------------------------------------------------------------------------------------------------------------
BT_OUT = ("accel_layer_u", "accel_layer_v", "eta_out", "uhbtav", "vhbtav", "etaav", "CS%ubtav", "CS%vbtav", "CS%eta_cor")
case("btstep/default", "btstep", (16, 12, 4), BT_OUT, land_blocks=2)
case("btstep/no_uh0_no_etaav", "btstep", (12, 10, 3), BT_OUT, land_blocks=2, with_uh0=False, with_etaav=False)
case("btstep/bottom_stress_arakawa_hsu", "btstep", (16, 12, 4), BT_OUT, land_blocks=2, with_bot=True, Sadourny=0)
case("btstep/strong_drag_bound_corr", "btstep", (16, 12, 4), BT_OUT, land_blocks=2, strong_drag=1, bound_BT_corr=1,
     BT_cont_bounds=0, visc_rem_u_uh0=1)
case("btstep/narrow_halo_bugs", "btstep", (16, 12, 4), BT_OUT, land_blocks=2, use_wide_halos=0, wt_uv_bug=1,
     use_old_coriolis_bracket_bug=1)
case("btstep/project_velocity_filter_y_first", "btstep", (16, 12, 4), BT_OUT, land_blocks=2, BT_project_velocity=1,
     dt_bt_filter=600.0, bebt=0.2, first_direction=1)
case("btstep/closed_x_G_extra", "btstep", (16, 12, 4), BT_OUT, land_blocks=1, cyclic_x=False, G_extra=0.1)
case("btstep/doubly_periodic_one_layer", "btstep", (12, 10, 1), BT_OUT, land_blocks=0, cyclic_y=True, with_uh0=False)


# ---- horizontal_viscosity ------------------------------------------------------------------------------------------------
HV_OUT = ("diffu", "diffv")
for _n, _kw in enumerate([
        dict(), dict(land_blocks=4, Ah=1.0e11), dict(land_blocks=4, Laplacian=True, Kh=800.0, Smagorinsky_Kh=True),
        dict(land_blocks=3, Laplacian=True, biharmonic=False, Kh_vel_scale=0.01),
        dict(land_blocks=3, Laplacian=True, Smagorinsky_Kh=True, better_bound_Kh=False, better_bound_Ah=False),
        dict(land_blocks=3, bound_Coriolis=True), dict(land_blocks=3, no_slip=True, Laplacian=True, Kh=300.0),
        dict(land_blocks=3, use_land_mask=True, add_LES_viscosity=True, Laplacian=True, Smagorinsky_Kh=True, Kh_bg_min=50.0),
        dict(land_blocks=3, Re_Ah=20.0), dict(land_blocks=3, cont_thick=True),
        dict(land_blocks=2, Laplacian=True, better_bound_Ah=False, Kh=100.0),
        dict(land_blocks=2, bound_Ah=False, better_bound_Ah=False, Smagorinsky_Ah=False, Ah_vel_scale=0.01, cyclic_y=True)]):
    case(f"horizontal_viscosity/options{_n:02d}", "horizontal_viscosity", (16, 12, 3), HV_OUT, **_kw)

# ---- vertvisc_coef -> vertvisc_remnant -> vertvisc ------------------------------------------------------------------------
VV_OUT = ("a_u", "a_v", "h_u", "h_v", "visc_rem_u", "visc_rem_v", "u", "v", "taux_bot", "tauy_bot")
for _n, _kw in enumerate([
        dict(), dict(land_blocks=4, with_Bu=True, with_Ray=True), dict(harmonic_visc=1, land_blocks=3),
        dict(bottomdraglaw=0, Kv_extra_bbl=5e-3), dict(bottomdraglaw=0, land_blocks=2, cyclic_y=True),
        dict(Kvml_invZ2=1e-3, land_blocks=2), dict(harm_BL_val=0.5, with_Ray=True), dict(fixed_LOTW_ML=1, land_blocks=3),
        dict(apply_LOTW_floor=1), dict(fixed_LOTW_ML=1, apply_LOTW_floor=1, Kvml_invZ2=1e-3, harmonic_visc=1),
        dict(direct_stress=1, land_blocks=2, with_Ray=True)]):
    case(f"vertvisc_family/options{_n:02d}", "vertvisc_family", (16, 12, 6), VV_OUT, **_kw)
# vertvisc_limit_vel (:2926-3120): CFL-based truncation (a time step long enough for CFL > 0.5), MAXVEL truncation, VEL_UNDERFLOW
case("vertvisc_family/truncation_cfl", "vertvisc_family", (16, 12, 6), VV_OUT, land_blocks=2, dt=60000.0)
case("vertvisc_family/truncation_cfl_0.2_ray", "vertvisc_family", (16, 12, 6), VV_OUT, land_blocks=1, dt=30000.0, CFL_trunc=0.2, with_Ray=True)
case("vertvisc_family/truncation_maxvel", "vertvisc_family", (16, 12, 6), VV_OUT, land_blocks=2, CFL_based_trunc=0, maxvel=0.2)
case("vertvisc_family/vel_underflow", "vertvisc_family", (16, 12, 6), VV_OUT, land_blocks=2, vel_underflow=0.05)

# ---- PressureForce_FV_Bouss ----------------------------------------------------------------------------------------------
PF_OUT = ("PFu", "PFv", "pbce", "eta")
for _n, _kw in enumerate([
        dict(), dict(eos="LINEAR"), dict(eos="NONE"), dict(MassWghtInterp=1, with_p_atm=True),
        dict(use_SSH_in_Z0p=1, MassWghtInterp=3, rho_ref_bug=1), dict(reconstruct=1, Recon_Scheme=1),
        dict(reconstruct=1, Recon_Scheme=2), dict(reconstruct=1, Recon_Scheme=1, boundary_extrap=1, MassWghtInterp=1, eos="LINEAR"),
        dict(reconstruct=1, Recon_Scheme=2, boundary_extrap=1, use_inaccurate_pgf_rho_anom=1)]):
    case(f"pressure_force/options{_n:02d}", "pressure_force", (14, 10, 5), PF_OUT, land_blocks=2, **_kw)


# ---- advect_tracer -------------------------------------------------------------------------------------------------------
AD_OUT = ("tr.0", "tr.1", "tr.2", "uhr_out", "vhr_out", "vol_prev")
case("advect_tracer/plm", "advect_tracer", (14, 10, 4), AD_OUT)
case("advect_tracer/ppm_h3_cfl2.5_three_tracers", "advect_tracer", (14, 10, 4), AD_OUT, scheme=1, cfl=2.5, ntr=3)
case("advect_tracer/ppm_land", "advect_tracer", (14, 10, 4), AD_OUT, scheme=2, land_blocks=2)
case("advect_tracer/ppm_h3_periodic_y_y_first", "advect_tracer", (14, 10, 4), AD_OUT, scheme=1, land_blocks=2, cyclic_y=True,
     cyclic_x=False, x_first_in=0)
case("advect_tracer/plm_cfl3.5_drained_cells", "advect_tracer", (14, 10, 4), AD_OUT, scheme=0, cfl=3.5, land_blocks=1)
case("advect_tracer/mixed_schemes_underflow_max_iter", "advect_tracer", (14, 10, 4), AD_OUT, ntr=3, cfl=3.2, land_blocks=1,
     advect_scheme=[2, -1, 1], conc_underflow=[0.0, 0.0, 0.6], max_iter_in=2)


# ---- step_MOM_dyn_split_RK2: the whole step, every stage from the reference's own source ----------------------------------
STEP_OUT = ("u_inst", "v_inst", "h", "uh", "vh", "uhtr", "vhtr", "eta_av") + tuple("CS%" + k for k in (
    "CAu", "CAv", "CAu_pred", "CAv_pred", "PFu", "PFv", "diffu", "diffv", "visc_rem_u", "visc_rem_v", "u_accel_bt", "v_accel_bt",
    "u_av", "v_av", "h_av", "pbce", "eta", "eta_PF", "uhbt", "vhbt", "taux_bot", "tauy_bot")) + tuple(
    "BT%" + k for k in ("ubtav", "vbtav", "eta_cor", "frhatu", "frhatv")) + tuple("BT_cont%" + k for k in (
        "FA_u_EE", "FA_u_E0", "FA_u_W0", "FA_u_WW", "uBT_WW", "uBT_EE", "FA_v_NN", "FA_v_N0", "FA_v_S0", "FA_v_SS", "vBT_SS",
        "vBT_NN", "h_u", "h_v"))
case("step/default", "step", (12, 10, 4), STEP_OUT, land_blocks=2)
case("step/two_steps_store_CAu_set_dtbt", "step", (12, 10, 4), STEP_OUT, land_blocks=2, nsteps=2, store_CAu=1, calc_dtbt=1)
case("step/plm_pressure_reconstruction", "step", (12, 10, 4), STEP_OUT, land_blocks=2, pgf=dict(reconstruct=1, Recon_Scheme=1))
case("step/ppm_reconstruction_begw_split_bottom_stress", "step", (12, 10, 4), STEP_OUT, land_blocks=2, begw=0.25,
     split_bottom_stress=1, pgf=dict(reconstruct=1, Recon_Scheme=2))
case("step/project_velocity_no_land", "step", (12, 10, 4), STEP_OUT, land_blocks=0, BT_project_velocity=1)
case("step/cfl_truncation_in_vertvisc", "step", (12, 10, 4), STEP_OUT, land_blocks=2, vv=dict(CFL_trunc=0.001))
case("step/y_first_doubly_periodic", "step", (12, 10, 4), STEP_OUT, land_blocks=1, first_direction=1, cyclic_y=True, store_CAu=1)
case("step/arakawa_hsu_bt_strong_drag_linear_eos", "step", (12, 10, 4), STEP_OUT, land_blocks=2, Sadourny=0, strong_drag=1,
     pgf=dict(EOS_form=1, reconstruct=1, Recon_Scheme=2))


# ---- ALE_regridding_and_remapping (MOM.F90:1751) --------------------------------------------------------------------------
ALE_OUT = ("u", "v", "h", "tr.0", "tr.1", "tr.2", "Kd_shear", "Kv_shear", "Kv_shear_Bu", "DYN%diffu", "DYN%diffv", "DYN%CAu_pred",
           "DYN%CAv_pred", "DYN%u_av", "DYN%v_av")
case("ale/ppm_h4_aux_vars", "ale", (12, 10, 5), ALE_OUT, land_blocks=2)
case("ale/plm_no_store_CAu", "ale", (12, 10, 5), ALE_OUT, land_blocks=2, remapping_scheme=2, store_CAu=0)
case("ale/ppm_ih4_no_time_filter", "ale", (12, 10, 6), ALE_OUT, land_blocks=1, remapping_scheme=5, regrid_time_scale=0.0, with_Bu=False)
case("ale/pcm_no_aux_vars", "ale", (12, 10, 5), ALE_OUT, land_blocks=2, remapping_scheme=0, remap_aux_vars=0)


# ---- the callers between dynamics steps -------------------------------------------------------------------
TD_OUT = ("h", "uhtr", "vhtr", "uhGM", "vhGM")
for _n, _kw in enumerate([
        dict(land_blocks=2), dict(land_blocks=2, with_GM=True, with_p_surf=True, EOS_form=1),
        dict(land_blocks=2, use_FGNV_streamfn=1, use_variable_mixing=1),
        dict(land_blocks=1, use_variable_mixing=1, Resoln_scaled_KhTh=1, Khth_Max=900.0, Khth_Min=50.0),
        dict(land_blocks=2, use_variable_mixing=1, use_stored_slopes=1), dict(land_blocks=2, use_MEKE_Kh=1, MEKE_KhTh_fac=0.5)]):
    case(f"thickness_diffuse/options{_n:02d}", "thickness_diffuse", (14, 10, 5), TD_OUT, **_kw)
MLE_OUT = ("h", "uhtr", "vhtr", "CS%MLD_filtered", "CS%MLD_filtered_slow")
for _n, _kw in enumerate([
        dict(land_blocks=2), dict(land_blocks=2, eos="LINEAR", ml_restrat_coef2=0.5, MLE_MLD_decay_time2=5.0e6),
        dict(land_blocks=1, MLE_use_PBL_MLD=0, MLE_density_diff=0.3),   # 0.03: the diagnosed mixed layer stays inside the top layer
        dict(land_blocks=1, MLE_tail_dh=0.2, MLE_MLD_stretch=1.5, cyclic_y=True)]):
    # 16 layers: with fewer the boundary layer lies inside the top layer and the restratifying transports vanish identically
    case(f"mixedlayer_restrat/options{_n:02d}", "mixedlayer_restrat", (14, 10, 16), MLE_OUT, **_kw)
HD_OUT = ("tr.0", "tr.1", "tr.2", "df_x.0", "df_x.2", "df_y.1", "df_y.2")
for _n, _kw in enumerate([
        dict(land_blocks=2), dict(land_blocks=2, with_df=True, max_diff_CFL=0.4, check_diffusive_CFL=1, KhTr=8000.0),
        dict(land_blocks=1, use_variable_mixing=1, Resoln_scaled_KhTr=1, KhTr_Slope_Cff=0.25, KhTr_max=3000.0, KhTr_min=100.0),
        dict(land_blocks=1, use_MEKE_Kh=1, MEKE_KhTr_fac=0.7, KhTr_passivity_coeff=3.0, cyclic_y=True),
        dict(land_blocks=2, KhTr=5.0e4, check_diffusive_CFL=1), dict(KhTr=5.0e4, max_diff_CFL=2.5, ntr=1)]):
    case(f"tracer_hordiff/options{_n:02d}", "tracer_hordiff", (14, 10, 5), HD_OUT, **_kw)


# ---- a second sweep: the control-structure members no case above moves off their defaults (each checked to change the answer) -------
case("continuity/tolerances", "continuity", (20, 16, 6), CONT_OUT, land_blocks=2, cs_over=dict(tol_eta=1e-4, tol_vel=1e-5, CFL_limit_adjust=0.3))
case("continuity/tolerances_aggress_adjust", "continuity", (20, 16, 6), CONT_OUT, land_blocks=2,
     cs_over=dict(tol_eta=1e-3, tol_vel=1e-4, CFL_limit_adjust=0.9, aggress_adjust=1, vol_CFL=1))
for _nm, _o in (("al_blend_f2.1_w0.3", dict(F_eff_max_blend=2.1, wt_lin_blend=0.3)), ("al_blend_f2.4_w0.9", dict(F_eff_max_blend=2.4, wt_lin_blend=0.9)),
                ("al_blend_sadourny_limit", dict(F_eff_max_blend=1.5))):
    case("coradcalc/" + _nm, "coradcalc", (16, 12, 3), CASES["coradcalc/scheme6_ke10"]["outputs"], land_blocks=2, cs_over=dict(Coriolis_Scheme=6, **_o))
case("vertvisc_family/mixing_lengths", "vertvisc_family", (16, 12, 6), CASES["vertvisc_family/options00"]["outputs"], land_blocks=2,
     Hbbl=3.0, Kv=5e-4, Hmix=15.0, Hmix_stress=8.0, vonKar=0.38)
case("vertvisc_family/mixing_lengths_no_drag_law", "vertvisc_family", (16, 12, 6), CASES["vertvisc_family/options00"]["outputs"], land_blocks=2,
     bottomdraglaw=0, Hbbl=25.0, Kv=2e-3, Hmix=60.0, with_Ray=True)
case("pressure_force/gfs_scale", "pressure_force", (14, 10, 5), CASES["pressure_force/options00"]["outputs"], land_blocks=2, GFS_scale=0.9)
case("pressure_force/rho_ref_h_nonvanished_plm", "pressure_force", (14, 10, 5), CASES["pressure_force/options00"]["outputs"], land_blocks=2,
     rho_ref=1030.0, h_nonvanished=1e-3, reconstruct=1, Recon_Scheme=1)
case("pressure_force/mass_weight_vanished_only_ppm", "pressure_force", (14, 10, 5), CASES["pressure_force/options00"]["outputs"], land_blocks=2,
     MassWghtInterp=1, MassWghtInterpVanOnly=1, reconstruct=1, Recon_Scheme=2)
case("tracer_hordiff/passivity_min", "tracer_hordiff", (14, 10, 5), HD_OUT, land_blocks=2, use_variable_mixing=1, KhTr_passivity_coeff=3.0,
     KhTr_passivity_min=4.0)
case("thickness_diffuse/khth_cfl_slope_smoothing", "thickness_diffuse", (14, 10, 5), TD_OUT, land_blocks=2, Khth=1500.0, max_Khth_CFL=0.05,
     slope_max=0.002, kappa_smooth=1e-4)
case("thickness_diffuse/fgnv_scale_n2_floor", "thickness_diffuse", (14, 10, 5), TD_OUT, land_blocks=2, use_FGNV_streamfn=1, FGNV_scale=0.5,
     N2_floor=1e-10, use_variable_mixing=1)
case("mixedlayer_restrat/coefficients", "mixedlayer_restrat", (14, 10, 16), MLE_OUT, land_blocks=2, ml_restrat_coef=0.7, front_length=1500.0,
     MLE_MLD_decay_time=1e5, vonKar=0.38)
case("mixedlayer_restrat/no_front_length", "mixedlayer_restrat", (14, 10, 16), MLE_OUT, land_blocks=2, front_length=0.0, ml_restrat_coef=20.0)
case("step/be_0.7", "step", (12, 10, 4), STEP_OUT, land_blocks=2, dyn=dict(be=0.7))
case("step/no_visc_rem_dt_bug", "step", (12, 10, 4), STEP_OUT, land_blocks=2, dyn=dict(visc_rem_dt_bug=0))

# ---- a third sweep: options that an audit (remove the option, does the oracle's answer change?) found without effect in the case that
# ---- names them, now in a setting where they do act (each checked) -------------------------------------------------------------
_T3 = [
    ("btstep/coriolis_bracket_bug", "btstep", (16, 12, 4), dict(use_old_coriolis_bracket_bug=1)),
    ("coradcalc/robust_enstro_pv_upwind1", "coradcalc", (16, 12, 3), dict(cs_over=dict(Coriolis_Scheme=3, PV_Adv_Scheme=22))),
    ("pressure_force/rho_ref_bug", "pressure_force", (14, 10, 5), dict(rho_ref_bug=1, rho_ref=1030.0)),
    ("pressure_force/plm_inaccurate_rho_anom", "pressure_force", (14, 10, 5), dict(reconstruct=1, Recon_Scheme=1, use_inaccurate_pgf_rho_anom=1)),
    ("pressure_force/ppm_mass_weight_vanished_only_5m", "pressure_force", (14, 10, 5),
     dict(reconstruct=1, Recon_Scheme=2, MassWghtInterp=1, MassWghtInterpVanOnly=1, h_nonvanished=5.0)),
    ("horizontal_viscosity/les_added_to_background", "horizontal_viscosity", (16, 12, 3),
     dict(Laplacian=True, Kh=300.0, Smagorinsky_Kh=True, add_LES_viscosity=True)),
    ("horizontal_viscosity/kh_bg_min", "horizontal_viscosity", (16, 12, 3), dict(Laplacian=True, Kh=10.0, Kh_bg_min=500.0)),
    # viscosities large enough, and a time step long enough, for the stability bounds to bind: legacy / no / better bounds
    ("horizontal_viscosity/legacy_bounds_bind", "horizontal_viscosity", (16, 12, 3),
     dict(Laplacian=True, Kh=5.0e4, dt=14400.0, better_bound_Kh=False, better_bound_Ah=False, Smagorinsky_Kh=True)),
    ("horizontal_viscosity/unbounded", "horizontal_viscosity", (16, 12, 3),
     dict(Laplacian=True, Kh=5.0e4, Ah=1.0e13, dt=14400.0, bound_Kh=False, bound_Ah=False, better_bound_Kh=False, better_bound_Ah=False)),
    ("horizontal_viscosity/better_bound_kh_only", "horizontal_viscosity", (16, 12, 3),
     dict(Laplacian=True, Kh=5.0e4, Ah=1.0e13, dt=14400.0, better_bound_Ah=False)),
    ("horizontal_viscosity/better_bound_ah_only", "horizontal_viscosity", (16, 12, 3),
     dict(Laplacian=True, Kh=5.0e4, Ah=1.0e13, dt=14400.0, better_bound_Kh=False)),
    ("horizontal_viscosity/better_bounds_bind", "horizontal_viscosity", (16, 12, 3), dict(Laplacian=True, Kh=5.0e4, Ah=1.0e13, dt=14400.0)),
    ("tracer_hordiff/meke_passivity_with_varmix", "tracer_hordiff", (14, 10, 5),
     dict(use_variable_mixing=1, use_MEKE_Kh=1, MEKE_KhTr_fac=0.7, KhTr_passivity_coeff=3.0)),
    ("tracer_hordiff/khtr_max_binds", "tracer_hordiff", (14, 10, 5), dict(use_variable_mixing=1, KhTr=8000.0, KhTr_max=3000.0)),
    ("tracer_hordiff/diffusive_cfl_iterations", "tracer_hordiff", (14, 10, 5), dict(KhTr=2.0e5, check_diffusive_CFL=1)),
    ("thickness_diffuse/khth_cfl_binds", "thickness_diffuse", (14, 10, 5), dict(Khth=1.0e5, max_Khth_CFL=0.05)),
    ("thickness_diffuse/khth_max_binds", "thickness_diffuse", (14, 10, 5),
     dict(use_variable_mixing=1, Resoln_scaled_KhTh=1, Khth=2000.0, Khth_Max=900.0)),
    ("thickness_diffuse/p_surf_wright", "thickness_diffuse", (14, 10, 5), dict(with_p_surf=True)),
    ("continuity/cfl_limit_binds", "continuity", (20, 16, 6), dict(uhbt_noise=3.0, dt=3600.0, cs_over=dict(CFL_limit_adjust=0.02))),
    ("continuity/cfl_limit_binds_no_visc_rem_max", "continuity", (20, 16, 6),
     dict(uhbt_noise=3.0, dt=3600.0, cs_over=dict(CFL_limit_adjust=0.005, use_visc_rem_max=0))),
    ("continuity/aggress_adjust_binds", "continuity", (20, 16, 6), dict(uhbt_noise=6.0, dt=14400.0, cs_over=dict(aggress_adjust=1, vol_CFL=1))),
]
# the members of ALE's regridding / remapping control structures (each checked to act); remapping answer dates before 2019 are refused
_T3 += [("ale/" + n, "ale", (12, 10, 5), kw) for n, kw in (
    ("min_thickness_2m", dict(regrid=dict(min_thickness=2.0))),
    ("time_filter_depths", dict(regrid=dict(depth_of_time_filter_shallow=50.0, depth_of_time_filter_deep=600.0))),
    ("boundary_extrapolation", dict(remap=dict(boundary_extrapolation=1), vel_remap=dict(boundary_extrapolation=1))),
    ("no_force_bounds_in_target", dict(remap=dict(force_bounds_in_target=0), vel_remap=dict(force_bounds_in_target=0))),
    ("not_via_sub_cells", dict(remap=dict(om4_remap_via_sub_cells=0), vel_remap=dict(om4_remap_via_sub_cells=0))),
    ("plm_velocities_ppm_tracers", dict(vel_remap=dict(remapping_scheme=2))))]
# the coefficients of the linear equation of state, and KV without a bottom drag law
_T3 += [("thickness_diffuse/linear_eos_coefficients", "thickness_diffuse", (14, 10, 5), dict(EOS_form=1, dRho_dT=-0.3, dRho_dS=0.7)),
        ("mixedlayer_restrat/linear_eos_coefficients", "mixedlayer_restrat", (14, 10, 16), dict(eos="LINEAR", Rho_T0_S0=999.0, dRho_dT=-0.3, dRho_dS=0.7)),
        ("pressure_force/linear_eos_coefficients", "pressure_force", (14, 10, 5),
         dict(eos="LINEAR", Rho_T0_S0=999.0, dRho_dT=-0.3, dRho_dS=0.7, dRho_dp=4.0e-7)),
        ("pressure_force/linear_eos_coefficients_ppm", "pressure_force", (14, 10, 5),
         dict(eos="LINEAR", reconstruct=1, Recon_Scheme=2, Rho_T0_S0=999.0, dRho_dT=-0.3, dRho_dS=0.7, dRho_dp=4.0e-7)),
        ("vertvisc_family/kv_no_drag_law", "vertvisc_family", (16, 12, 6), dict(bottomdraglaw=0, Kv=3e-3)),
        ("btstep/min_stencil_2", "btstep", (16, 12, 4), dict(min_stencil=2)),
        ("mixedlayer_restrat/ustar_min", "mixedlayer_restrat", (14, 10, 16), dict(ustar_min=5.0e-3))]
_T3 += [("advect_tracer/vol_prev_uhr_out_one_iteration", "advect_tracer", (14, 10, 4),
         dict(scheme=1, cfl=3.2, ntr=3, max_iter_in=1, with_vol_prev=1.01, with_uhr_out=True)),
        ("advect_tracer/vol_prev_only_ppm", "advect_tracer", (14, 10, 4), dict(scheme=2, cfl=2.5, ntr=3, with_vol_prev=0.99))]
# optional arguments left out
_T3 += [("pressure_force/no_pbce_no_eta_p_atm_plm", "pressure_force", (14, 10, 5),
         dict(with_pbce=False, with_eta=False, reconstruct=1, Recon_Scheme=1, with_p_atm=True)),
        ("vertvisc_family/no_shear_viscosity", "vertvisc_family", (16, 12, 6), dict(with_shear=False)),
        ("continuity/no_velocity_corrections", "continuity", (20, 16, 6), dict(with_cor=False))]
for _nm, _st, _shape, _kw in _T3:
    case(_nm, _st, _shape, next(c["outputs"] for c in CASES.values() if c["stage"] == _st), land_blocks=2, **_kw)
# domain shapes the cases above leave out: closed basins, a reentrant y direction, y first
_GEO = [("step/closed_basin_with_land", "step", (12, 10, 4), dict(land_blocks=2, cyclic_x=False)),
        ("step/closed_basin_y_first", "step", (12, 10, 4), dict(land_blocks=0, cyclic_x=False, cyclic_y=False, first_direction=1)),
        ("pressure_force/closed_basin_ppm", "pressure_force", (14, 10, 5), dict(land_blocks=2, cyclic_x=False, reconstruct=1, Recon_Scheme=2)),
        ("continuity/doubly_periodic_y_first", "continuity", (20, 16, 6), dict(land_blocks=2, cyclic_y=True, first_direction=1)),
        ("coradcalc/doubly_periodic_arakawa_lamb_gudonov", "coradcalc", (16, 12, 3),
         dict(land_blocks=2, cyclic_y=True, cs_over=dict(Coriolis_Scheme=5, KE_Scheme=12))),
        ("advect_tracer/closed_basin_ppm", "advect_tracer", (14, 10, 4), dict(land_blocks=0, cyclic_x=False, cyclic_y=False, scheme=2, cfl=2.5))]
# the whole step with every stage off its default scheme
_GEO += [("step/monotonic_arakawa_hsu_smagorinsky_harmonic", "step", (12, 10, 4),
          dict(land_blocks=2, cont=dict(monotonic=1), corad=dict(Coriolis_Scheme=2, KE_Scheme=12, bound_Coriolis=1),
               hv=dict(Laplacian=True, Smagorinsky_Kh=True, Kh=300.0, bound_Coriolis=True), vv=dict(harmonic_visc=1))),
         ("step/simple2nd_al_blend_laplacian_no_slip_direct_stress", "step", (12, 10, 4),
          dict(land_blocks=2, cont=dict(simple_2nd=1, vol_CFL=1), corad=dict(Coriolis_Scheme=6, KE_Scheme=11, no_slip=1),
               hv=dict(Laplacian=True, biharmonic=False, Kh=800.0, no_slip=True), vv=dict(bottomdraglaw=0, direct_stress=1)))]
# one, two and three layers (the reconstructions fall back to lower order, the tridiagonal solves degenerate)
_GEO += [("ale/two_layers", "ale", (12, 10, 2), dict(land_blocks=2)), ("ale/three_layers", "ale", (12, 10, 3), dict(land_blocks=2)),
         ("step/one_layer", "step", (12, 10, 1), dict(land_blocks=2)), ("step/two_layers", "step", (12, 10, 2), dict(land_blocks=2)),
         ("vertvisc_family/one_layer", "vertvisc_family", (16, 12, 1), dict(land_blocks=2, with_Ray=True)),
         ("advect_tracer/one_layer_ppm", "advect_tracer", (14, 10, 1), dict(land_blocks=2, scheme=2, cfl=2.5)),
         ("thickness_diffuse/two_layers", "thickness_diffuse", (14, 10, 2), dict(land_blocks=2)),
         ("pressure_force/two_layers_plm", "pressure_force", (14, 10, 2), dict(land_blocks=2, reconstruct=1, Recon_Scheme=1))]
# other time steps (another number of barotropic steps and parity of the alternating directions, larger CFL numbers)
_GEO += [("btstep/dt_1250", "btstep", (16, 12, 4), dict(land_blocks=2, dt=1250.0)),
         ("btstep/dt_300_y_first", "btstep", (16, 12, 4), dict(land_blocks=2, dt=300.0, first_direction=1)),
         ("step/dt_1800", "step", (12, 10, 4), dict(land_blocks=2, dt=1800.0)),
         ("step/dt_450_fast_flow", "step", (12, 10, 4), dict(land_blocks=2, dt=450.0, vel=0.3)),
         ("continuity/dt_7200", "continuity", (20, 16, 6), dict(land_blocks=2, dt=7200.0)),
         ("advect_tracer/dt_14400_cfl3.9_ppm_h3", "advect_tracer", (14, 10, 4), dict(land_blocks=2, dt=14400.0, dt_dyn=900.0, scheme=1, cfl=3.9))]
for _nm, _st, _shape, _kw in _GEO:
    case(_nm, _st, _shape, next(c["outputs"] for c in CASES.values() if c["stage"] == _st), **_kw)

# cases added or changed after the round's GPU budget was spent: their device legs run from tests/test_zzz_reference_golden_late.py, sorted
# last, so that a disagreement there cannot hide the results of the files after tests/test_reference_golden.py under `pytest -x`
LATE = {n for n in CASES if n.startswith(("diag/", "mixedlayer_restrat/"))} | {
    "continuity/tolerances", "continuity/tolerances_aggress_adjust", "coradcalc/al_blend_f2.1_w0.3", "coradcalc/al_blend_f2.4_w0.9",
    "coradcalc/al_blend_sadourny_limit", "vertvisc_family/mixing_lengths", "vertvisc_family/mixing_lengths_no_drag_law",
    "pressure_force/gfs_scale", "pressure_force/rho_ref_h_nonvanished_plm", "pressure_force/mass_weight_vanished_only_ppm",
    "tracer_hordiff/passivity_min", "thickness_diffuse/khth_cfl_slope_smoothing", "thickness_diffuse/fgnv_scale_n2_floor", "step/be_0.7",
    "step/no_visc_rem_dt_bug", "bt_helpers/set_dtbt_parameters"} | {t[0] for t in _T3 + _GEO}

# outputs a case legitimately returns as it received them
UNTOUCHED_OK = {
    "btstep": {"CS%eta_cor"},                                  # an input of btstep (bt_mass_source sets it)
    "ale/pcm_no_aux_vars": {"Kd_shear", "Kv_shear", "Kv_shear_Bu", "DYN%diffu", "DYN%diffv", "DYN%CAu_pred", "DYN%CAv_pred", "DYN%u_av", "DYN%v_av"},
    "ale/plm_no_store_CAu": {"DYN%CAu_pred", "DYN%CAv_pred", "DYN%u_av", "DYN%v_av"},
    "mixedlayer_restrat": {"CS%MLD_filtered_slow"},            # only filtered when MLE_MLD_DECAY_TIME2 > 0
}


def untouched(name, inputs, out):
    """the outputs of a case that equal what went in (a case whose main outputs come back untouched pins nothing)"""
    c = CASES[name]
    dom = inputs[0]
    if c["stage"] in ("bt_helpers", "diag", "vertvisc_family"):
        return []
    if c["stage"] == "step":
        before = step_collect(dom, inputs[4], inputs[5])
    elif c["stage"] == "ale":
        before = ale_collect(dom, inputs[3], inputs[4], inputs[5])
    else:
        before = collect(dom, c["outputs"], inputs[4], inputs[3])
    ok = UNTOUCHED_OK.get(name, set()) | UNTOUCHED_OK.get(c["stage"], set())
    return [k for k in out if k in before and not k.startswith("zero_ok:") and k not in ok and before[k].shape == out[k].shape
            and np.array_equal(before[k], out[k])]


# ---- btcalc (4 thickness schemes + the default), bt_mass_source (set / accumulate), set_dtbt (4 ways of finding the wave speed) ----
case("bt_helpers/btcalc_mass_source_set_dtbt", "bt_helpers", (16, 12, 5), (), land_blocks=2)
case("bt_helpers/set_dtbt_parameters", "bt_helpers", (16, 12, 5), (), land_blocks=2,
     dtbt=dict(bebt=0.3, G_extra=0.05, dtbt_fraction=0.7, BT_Coriolis_scale=0.5, Z_ref=2.0))


def _dtbt_args(dom, grid, cs, mode):
    a = dict(pbce=cs["pbce"], gtot_est=0.0, have_gtot_est=0, BT_cont=None, eta=None, SSH_add=0.0, frhatu=cs["barotropic"]["frhatu"],
             frhatv=cs["barotropic"]["frhatv"], bathyT=grid["bathyT"], bebt=0.1, G_extra=0.0, dtbt_fraction=0.98,
             BT_Coriolis_scale=1.0, Z_ref=0.0, Nonlinear_continuity=0)
    if mode == "BT_cont":
        a["BT_cont"] = cs["BT_cont"]
    elif mode == "eta":
        a["eta"] = cs["eta"]; a["Nonlinear_continuity"] = 1
    elif mode == "gtot":
        a["pbce"] = None; a["gtot_est"] = 9.8; a["have_gtot_est"] = 1; a["SSH_add"] = 2.0
    a.update(cs.get("_dtbt_over", {}))   # the parameters of set_dtbt a case moves off their defaults
    return a


def bt_helpers(inputs, btcalc, bt_mass_source, set_dtbt):
    """the three small barotropic entries through the given backend (oracle, translated reference or device)"""
    from mom6_b200 import fidx
    dom, grid, gv, bcs, ba, st, scs = inputs
    nk = int(dom.nk)
    out = {}
    hu, hv = np.abs(st["u"]) + 1.0, np.abs(st["v"]) + 1.0
    for n, (scheme, huv, mud) in enumerate(((1, False, 0), (2, False, 0), (3, False, 0), (4, True, 0), (4, False, 1))):
        a = dict(h=st["h"], h_u=hu if huv else None, h_v=hv if huv else None, frhatu=fidx.new(dom, "u", nk=nk).a,
                 frhatv=fidx.new(dom, "v", nk=nk).a, bathyT=grid["bathyT"], hvel_scheme=scheme, may_use_default=mud)
        btcalc(a)
        out[f"btcalc{n}.frhatu"], out[f"btcalc{n}.frhatv"] = inner(dom, a["frhatu"]).copy(), inner(dom, a["frhatv"]).copy()
    for set_cor in (1, 0):
        e = bcs["eta_cor"].copy()
        bt_mass_source(st["h"], ba["eta_in"], set_cor, e)
        out[f"bt_mass_source{set_cor}.eta_cor"] = inner(dom, e).copy()
    for mode in ("BT_cont", "eta", "bathy", "gtot"):
        out["set_dtbt." + mode] = np.array(set_dtbt(_dtbt_args(dom, grid, scs, mode)))
    return out


# ---- write_energy (three calls on a changing state, MOM_sum_output.F90:321) and the bit-count checksums (MOM_checksums.F90) -------
for _n, _kw in enumerate([dict(), dict(use_temperature=False), dict(do_APE_calc=False),
                          dict(RZL2_to_kg=2.0**-10, L_T_to_m_s=2.0**3, Z_to_m=2.0**2)]):
    case(f"diag/write_energy{_n:02d}", "diag", (20, 16, 5), (), land_blocks=2, cs=_kw)
case("diag/chksum", "diag", (12, 10, 3), (), chksum=True)
LATE |= {n for n in CASES if n.startswith("diag/")}

_WE_SCALARS = ("En_mass", "toten", "KE_tot", "PE_tot", "mass_tot", "mass_chg", "mass_anom", "Salt", "Salt_chg", "Salt_anom", "Heat",
               "Heat_chg", "Heat_anom", "salin", "salin_anom", "temp", "temp_anom")
_WE_EFPS = ("fresh_water_in_EFP", "net_salt_in_EFP", "net_heat_in_EFP", "mass_prev_EFP", "salt_prev_EFP", "heat_prev_EFP")


def diag_write_energy(inputs, write_energy, stats_line):
    """three write_energy calls through the given backend, the state and the truncation count changing in between; stats_line(cs, e,
    n, reday) is the record the call appends to ocean.stats (step n = the call number, one hour apart)"""
    a, cs = inputs[3], _copy(inputs[4])
    u, v, h, T, S = (a[k].copy() for k in ("u_inst", "v_inst", "h", "T", "S"))
    rng = np.random.default_rng(4)
    out = {}
    for call in range(3):
        e = write_energy(cs, u, v, h, T, S)
        out[f"call{call}.scalars"] = np.array([float(e[k]) for k in _WE_SCALARS] + [float(x) for x in e["max_CFL"]] +
                                              [float(e["ntrunc"]), float(cs["previous_calls"]), float(cs["ntrunc"])])
        for k in ("KE", "mass_lay"):
            out[f"call{call}.{k}"] = np.array(e[k], dtype=np.float64)
        for k in ("PE", "Z_0APE"):
            out[f"zero_ok:call{call}.{k}"] = np.array(e[k], dtype=np.float64)
        out[f"call{call}.EFP"] = np.array([[int(x) for x in cs[k]] for k in _WE_EFPS], dtype=np.int64)
        out[f"call{call}.lH"] = np.array(cs["lH"], dtype=np.int64)
        out[f"call{call}.stats_line"] = np.frombuffer(stats_line(cs, e, call, (3600.0 * call) / 86400.0).encode(), dtype=np.uint8).astype(np.float64)
        h *= 1.0 + 1.0e-3 * rng.standard_normal(h.shape)
        u += 1.0e-3 * rng.standard_normal(u.shape); T += 0.01 * rng.standard_normal(T.shape)
        cs["ntrunc"] = call + 1
    return out


def diag_chksum(inputs, chksum):
    """every form of the checksums -- h, u, v and B points, rank 2 and 3, halo shifts 0, 1 and 2, symmetric or not, corners or
    edges, scaled or not -- through the given backend: chksum(array, stagger, haloshift, symmetric, omit_corners, scale) ->
    (the bit counts, zero-padded to five, [mean, min, max])"""
    dom = inputs[0]
    r = np.random.default_rng(7)
    bcs, sts = [], []
    for stagger in range(4):
        su, sv = stagger in (1, 3), stagger in (2, 3)
        shp = (dom.jed - dom.jsd + 1 + sv, dom.ied - dom.isd + 1 + su)
        for nd in (2, 3):
            a = np.ascontiguousarray(r.standard_normal(((int(dom.nk),) if nd == 3 else ()) + shp) * 10.0 ** r.integers(-8, 8, size=shp))
            for hs in (0, 1, 2):
                for sym in ((False, True) if stagger else (False,)):
                    for omit in (False, True):
                        for scale in (1.0, 0.37):
                            bc, st = chksum(a, stagger, hs, sym, omit, scale)
                            bcs.append([int(x) for x in bc][:5] + [0] * (5 - len(bc)))
                            sts.append([float(x) for x in st])
    return {"chksum.bc": np.array(bcs, dtype=np.int64), "chksum.stats": np.array(sts)}


def ale_collect(dom, ale, dcs, a):
    src = dict(a)
    for k in ("diffu", "diffv", "CAu_pred", "CAv_pred", "u_av", "v_av"):
        src["DYN%" + k] = dcs[k]
    out = collect(dom, ALE_OUT, src, {})
    out["zero_ok:old_grid_weight"] = np.array([ale["regridCS"]["old_grid_weight"]])
    return out


# ---- the OM4 layer count (75): the nk-dependent kernel variants of the bench against the reference itself ------------------------
case("step/75_layers_plm_store_CAu", "step", (12, 10, 75), STEP_OUT, slow=True, land_blocks=1, store_CAu=1,
     pgf=dict(reconstruct=1, Recon_Scheme=1))
case("pressure_force/75_layers_ppm", "pressure_force", (12, 10, 75), PF_OUT, slow=True, land_blocks=1, reconstruct=1, Recon_Scheme=2)
case("vertvisc_family/75_layers", "vertvisc_family", (12, 10, 75), VV_OUT, slow=True, land_blocks=1, with_Ray=True)
case("ale/75_layers", "ale", (12, 10, 75), ALE_OUT, slow=True, land_blocks=1)
case("thickness_diffuse/75_layers", "thickness_diffuse", (12, 10, 75), TD_OUT, slow=True, land_blocks=1, use_FGNV_streamfn=1,
     use_variable_mixing=1)
case("advect_tracer/75_layers", "advect_tracer", (12, 10, 75), AD_OUT, slow=True, land_blocks=1, scheme=1, cfl=2.5)
case("btstep/75_layers", "btstep", (12, 10, 75), BT_OUT, slow=True, land_blocks=1)


def step_collect(dom, cs, a):
    src = dict(a)
    for k, v in cs.items():
        if isinstance(v, np.ndarray):
            src["CS%" + k] = v
    for k in ("ubtav", "vbtav", "eta_cor", "frhatu", "frhatv"):
        src["BT%" + k] = cs["barotropic"][k]
    for k, v in cs["BT_cont"].items():
        if v is not None:
            src["BT_cont%" + k] = v
    out = {k: np.ascontiguousarray(inner(dom, src[k])) for k in STEP_OUT if src.get(k) is not None}
    out["zero_ok:scalars"] = np.array([cs["barotropic"]["dtbt"], cs["dtbt_max"], float(cs["CAu_pred_stored"])])
    return out


# ---------------------------------------------------------------------------------------------------------------------------
def _coefs(dom, nk):
    from mom6_b200 import fidx
    return [np.zeros((nk + 1,) + fidx.new(dom, "u").a.shape), np.zeros((nk + 1,) + fidx.new(dom, "v").a.shape),
            fidx.new(dom, "u", nk=nk).a, fidx.new(dom, "v", nk=nk).a]


def vv_collect(dom, outputs, d):
    out = collect(dom, outputs, d, {})
    out["zero_ok:ntrunc"] = np.array([float(d["ntrunc"])])   # CS%ntrunc, the count of truncated velocities
    return out


def vertvisc_family(mod, dom, grid, gv, cs, coef, sol, is_oracle):
    """vertvisc_coef, then vertvisc_remnant and vertvisc with its coefficients (the order of step_MOM_dyn_split_RK2)"""
    nk = int(dom.nk)
    c = _coefs(dom, nk)
    s = _copy(sol)
    mod.vertvisc_coef(dom, grid, gv, cs, _copy(coef), *c)
    vru, vrv = np.zeros_like(sol["u"]), np.zeros_like(sol["v"])
    if is_oracle:
        mod.vertvisc_remnant(dom, grid, cs, vru, vrv, sol["dt"], *c, sol["Ray_u"], sol["Ray_v"])
    else:
        mod.vertvisc_remnant(dom, grid, gv, cs, vru, vrv, sol["dt"], *c, sol["Ray_u"], sol["Ray_v"])
    ntrunc = mod.vertvisc(dom, grid, gv, cs, s, *c)
    return dict(a_u=c[0], a_v=c[1], h_u=c[2], h_v=c[3], visc_rem_u=vru, visc_rem_v=vrv, u=s["u"], v=s["v"],
                taux_bot=s["taux_bot"], tauy_bot=s["tauy_bot"], ntrunc=ntrunc)


def build(name):
    c = CASES[name]
    st, shape, kw = c["stage"], c["shape"], dict(c["kw"])
    if st == "continuity":
        return synthetic.continuity_inputs(*shape, **kw)
    if st == "coradcalc":
        return synthetic.coradcalc_inputs(*shape, **kw)
    if st == "btstep":
        return synthetic.btstep_inputs(*shape, whalo=6, **kw)
    if st == "horizontal_viscosity":
        return synthetic.hor_visc_inputs(*shape, **kw)
    if st == "vertvisc_family":
        return synthetic.vertvisc_inputs(*shape, **kw)
    if st == "pressure_force":
        return synthetic.pressureforce_inputs(*shape, **kw)
    if st == "advect_tracer":
        vp, ur = kw.pop("with_vol_prev", None), kw.pop("with_uhr_out", False)
        dom, grid, gv, cs, a = synthetic.advect_inputs(*shape, **kw)
        if vp is not None:   # the caller's own cell volumes (hprev of MOM_tracer_advect.F90:188-195, scaled), updated by the call
            div = np.zeros_like(a["h_end"])
            div[:, 1:-1, 1:-1] = (a["uhtr"][:, 1:-1, 2:-1] - a["uhtr"][:, 1:-1, 1:-2]) + (a["vhtr"][:, 2:-1, 1:-1] - a["vhtr"][:, 1:-2, 1:-1])
            v = np.maximum(0.0, grid["areaT"][None] * a["h_end"] + div)
            a["vol_prev"] = np.ascontiguousarray((v + np.maximum(0.0, 1.0e-13 * v - grid["areaT"][None] * a["h_end"])) * vp)
            a["update_vol_prev"] = True
        if ur:               # the transports the limited number of iterations leaves over
            a["uhr_out"], a["vhr_out"] = np.zeros_like(a["uhtr"]), np.zeros_like(a["vhtr"])
        return dom, grid, gv, cs, a
    if st == "ale":
        over = {k: kw.pop(k, None) for k in ("regrid", "remap", "vel_remap")}   # members of the three control structures to move
        dom, grid, gv, ale, dcs, a = synthetic.ale_chain_inputs(*shape, **kw)
        for k, d in over.items():
            if d:
                ale[k + "CS"].update(d)
        return dom, grid, gv, ale, dcs, a
    if st == "diag":
        from oracle import pyoracle
        if kw.get("chksum"):
            from mom6_b200.api import make_domain
            return (make_domain(shape[0], shape[1], nk=shape[2]),)
        dom, grid, gv, css, dcs, a = synthetic.step_dyn_inputs(*shape, whalo=6, land_blocks=kw.get("land_blocks", 0))
        # the depth list is host code of the reference (create_depth_list :1203); the reference leg checks this one against its own
        cs = synthetic.sum_output_cs(dom, pyoracle.create_depth_list(dom, grid, min_depth_inc=1.0e-10), **kw.get("cs", {}))
        return dom, grid, gv, a, cs
    if st == "bt_helpers":
        from oracle import pyoracle
        dtbt_over = kw.pop("dtbt", None)
        dom, grid, gv, bcs, ba = synthetic.btstep_inputs(*shape, **kw)
        dom2, grid2, gv2, css, scs, sa = synthetic.step_dyn_inputs(*shape, **kw)
        pyoracle.step_dyn_split_rk2(dom2, grid2, gv2, css, scs, sa)   # realistic pbce, BT_cont, eta (pinned by the step cases)
        scs["_dtbt_over"] = dtbt_over or {}
        return dom2, grid2, gv2, bcs, ba, synthetic.dyn_state(dom2, grid2), scs
    if st == "thickness_diffuse":
        return synthetic.thickness_diffuse_inputs(*shape, **kw)
    if st == "mixedlayer_restrat":
        return synthetic.mle_inputs(*shape, **kw)
    if st == "tracer_hordiff":
        return synthetic.hordiff_inputs(*shape, **kw)
    if st == "step":
        pgf, nsteps, vv, dyn = kw.pop("pgf", None), kw.pop("nsteps", 1), kw.pop("vv", None), kw.pop("dyn", None)
        cont, corad, hv = kw.pop("cont", None), kw.pop("corad", None), kw.pop("hv", None)
        dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(*shape, **kw)
        if cont:
            css["continuity"].update(cont)
        if corad:
            css["coriolisadv"].update(corad)
        if hv:   # hor_visc_init's products depend on the options: built anew
            css["hor_visc"] = synthetic.hor_visc_cs(dom, grid, dt=a["dt"], **hv)
        if dyn:
            cs.update(dyn)
        if pgf:
            css["pressureforce"].update(pgf)
        if vv:
            css["vertvisc"].update(vv)
        return dom, grid, gv, css, cs, a
    raise KeyError(st)


def run_oracle(oracle, name, inputs):
    c = CASES[name]
    if c["stage"] == "diag":
        dom = inputs[0]
        if c["kw"].get("chksum"):
            def chk(a, stagger, hs, sym, omit, scale):
                bc, kind, st = oracle.chksum(dom, a, stagger, hs, sym, omit, scale, stats=True)
                return bc[:{1: 1, 2: 5, 3: 5, 4: 2, 5: 2}[kind]], st
            return diag_chksum(inputs, chk)
        grid, gv = inputs[1:3]
        return diag_write_energy(inputs, lambda cs, u, v, h, T, S: oracle.write_energy(dom, grid, gv, cs, u, v, h, T, S), oracle.ocean_stats_line)
    if c["stage"] == "bt_helpers":
        dom, grid, gv = inputs[:3]
        return bt_helpers(inputs, lambda a: oracle.btcalc(dom, grid, gv, a),
                          lambda h, eta, sc, e: oracle.bt_mass_source(dom, grid, gv, h, eta, sc, e),
                          lambda a: oracle.set_dtbt(dom, grid, gv, a))
    if c["stage"] == "ale":
        dom, grid, gv, ale, dcs, a = inputs
        ale, dcs, a = _copy(ale), _copy(dcs), _copy(a)
        oracle.ale_regridding_and_remapping(dom, grid, gv, ale, a, dyn_cs=dcs)
        return ale_collect(dom, ale, dcs, a)
    if c["stage"] == "step":
        dom, grid, gv, css, cs, a = inputs
        cs, a = _copy(cs), _copy(a)
        for _ in range(c["kw"].get("nsteps", 1)):
            oracle.step_dyn_split_rk2(dom, grid, gv, css, cs, a)
        return step_collect(dom, cs, a)
    if c["stage"] == "vertvisc_family":
        dom, grid, gv, cs, coef, sol = inputs
        return vv_collect(dom, c["outputs"], vertvisc_family(oracle, dom, grid, gv, cs, coef, sol, True))
    dom, grid, gv, cs, a = inputs
    cs, a = _copy(cs), _copy(a)
    if c["stage"] == "mixedlayer_restrat":
        oracle.mixedlayer_restrat(dom, grid, gv, cs, a["h"], a["uhtr"], a["vhtr"], a["T"], a["S"], a["ustar"], a["dt"], a["h_MLD"],
                                  a["Rd_dx_h"])
    else:
        getattr(oracle, c["stage"])(dom, grid, gv, cs, a)
    return collect(dom, c["outputs"], a, cs)


def run_reference(name, inputs):
    from oracle.f90run import stages
    c = CASES[name]
    if c["stage"] == "diag":
        dom = inputs[0]
        if c["kw"].get("chksum"):
            return diag_chksum(inputs, lambda a, stagger, hs, sym, omit, scale: stages.chksum(dom, a, stagger, hs, sym, omit, scale, stats=True))
        grid, gv, cs = inputs[1], inputs[2], inputs[4]
        dl = stages.create_depth_list(dom, grid, min_depth_inc=1.0e-10)
        for mine, theirs in zip((cs["DL_depth"], cs["DL_area"], cs["DL_vol_below"]), dl):
            assert np.array_equal(mine, theirs), "create_depth_list: the oracle's list is not the reference's"
        return diag_write_energy(inputs, lambda cs, u, v, h, T, S: stages.write_energy(dom, grid, gv, cs, u, v, h, T, S),
                                 lambda cs, e, n, reday: cs["_f90run"]["stats_line"])
    if c["stage"] == "bt_helpers":
        dom, grid, gv = inputs[:3]
        return bt_helpers(inputs, lambda a: stages.btcalc(dom, grid, gv, a),
                          lambda h, eta, sc, e: stages.bt_mass_source(dom, grid, gv, h, eta, sc, e),
                          lambda a: stages.set_dtbt(dom, grid, gv, a))
    if c["stage"] == "ale":
        dom, grid, gv, ale, dcs, a = inputs
        ale, dcs, a = _copy(ale), _copy(dcs), _copy(a)
        stages.ale_regridding_and_remapping(dom, grid, gv, ale, a, dyn_cs=dcs)
        return ale_collect(dom, ale, dcs, a)
    if c["stage"] == "step":
        dom, grid, gv, css, cs, a = inputs
        cs, a = _copy(cs), _copy(a)
        for _ in range(c["kw"].get("nsteps", 1)):
            stages.step_dyn_split_rk2(dom, grid, gv, css, cs, a, land_blocks=c["kw"].get("land_blocks", 0))
        return step_collect(dom, cs, a)
    if c["stage"] == "vertvisc_family":
        dom, grid, gv, cs, coef, sol = inputs
        return vv_collect(dom, c["outputs"], vertvisc_family(stages, dom, grid, gv, cs, coef, sol, False))
    dom, grid, gv, cs, a = inputs
    cs, a = _copy(cs), _copy(a)
    if c["stage"] == "btstep":
        stages.btstep(dom, grid, gv, cs, a, stages.wide_metrics(dom, c["kw"].get("land_blocks", 0)))
    else:
        getattr(stages, c["stage"])(dom, grid, gv, cs, a)
    return collect(dom, c["outputs"], a, cs)


def run_case(oracle, name, want_ref=True):
    inputs = build(name)
    orc = run_oracle(oracle, name, inputs)
    return (run_reference(name, inputs) if want_ref else None), orc


# ---------------------------------------------------------------------------------------------------------------------------
def _ctx(ctx_factory, dom, grid, gv):
    ctx = ctx_factory(dom)
    ctx.set_grid(grid)
    ctx.set_vgrid(gv)
    return ctx


def run_device(ctx_factory, name, inputs):
    """the same case through the C ABI on the GPU (tests/test_reference_golden.py)"""
    c = CASES[name]
    st = c["stage"]
    if st == "diag":
        dom = inputs[0]
        if c["kw"].get("chksum"):
            ctx = ctx_factory(dom)

            def chk(a, stagger, hs, sym, omit, scale):
                bc, kind, stt = ctx.chksum(a, stagger, haloshift=hs, symmetric=sym, omit_corners=omit, scale=scale, stats=True)
                return bc[:{1: 1, 2: 5, 3: 5, 4: 2, 5: 2}[kind]], stt
            out = diag_chksum(inputs, chk)
        else:
            ctx = _ctx(ctx_factory, dom, inputs[1], inputs[2])
            out = diag_write_energy(inputs, lambda cs, u, v, h, T, S: ctx.write_energy(cs, u, v, h, T, S), ctx.ocean_stats_line)
        ctx.close()
        return out
    if st == "bt_helpers":
        dom, grid, gv = inputs[:3]
        ctx = _ctx(ctx_factory, dom, grid, gv)
        out = bt_helpers(inputs, ctx.btcalc, ctx.bt_mass_source, ctx.set_dtbt)
        ctx.close()
        return out
    if st == "step":
        dom, grid, gv, css, cs, a = inputs
        cs, a = _copy(cs), _copy(a)
        ctx = _ctx(ctx_factory, dom, grid, gv)
        ctx.set_cs_continuity(css["continuity"]); ctx.set_cs_coriolisadv(css["coriolisadv"]); ctx.set_cs_hor_visc(css["hor_visc"])
        ctx.set_cs_pressureforce(css["pressureforce"]); ctx.set_cs_vertvisc(css["vertvisc"])
        for _ in range(c["kw"].get("nsteps", 1)):
            ctx.step_dyn_split_rk2(cs, a)
        ctx.close()
        return step_collect(dom, cs, a)
    if st == "ale":
        dom, grid, gv, ale, dcs, a = inputs
        ale, dcs, a = _copy(ale), _copy(dcs), _copy(a)
        ctx = _ctx(ctx_factory, dom, grid, gv)
        ctx.ale_regridding_and_remapping(ale, a, dyn_cs=dcs)
        ctx.close()
        return ale_collect(dom, ale, dcs, a)
    if st == "vertvisc_family":
        dom, grid, gv, cs, coef, sol = inputs
        nk = int(dom.nk)
        ctx = _ctx(ctx_factory, dom, grid, gv)
        ctx.set_cs_vertvisc(cs)
        ctx.vertvisc_coef(_copy(coef))
        g = _coefs(dom, nk)
        ctx.vertvisc_get_coef(*g)
        vru, vrv = np.zeros_like(sol["u"]), np.zeros_like(sol["v"])
        ctx.vertvisc_remnant(vru, vrv, sol["dt"], sol["Ray_u"], sol["Ray_v"])
        s = _copy(sol)
        ctx.vertvisc(s)
        ntrunc = ctx.vertvisc_ntrunc()
        ctx.close()
        return vv_collect(dom, c["outputs"], dict(a_u=g[0], a_v=g[1], h_u=g[2], h_v=g[3], visc_rem_u=vru, visc_rem_v=vrv, u=s["u"],
                                                  v=s["v"], taux_bot=s["taux_bot"], tauy_bot=s["tauy_bot"], ntrunc=ntrunc))
    dom, grid, gv, cs, a = inputs
    cs, a = _copy(cs), _copy(a)
    ctx = _ctx(ctx_factory, dom, grid, gv)
    if st == "continuity":
        ctx.set_cs_continuity(cs); ctx.continuity(a)
    elif st == "coradcalc":
        ctx.set_cs_coriolisadv(cs); ctx.coradcalc(a)
    elif st == "horizontal_viscosity":
        ctx.set_cs_hor_visc(cs); ctx.horizontal_viscosity(a)
    elif st == "pressure_force":
        ctx.set_cs_pressureforce(cs); ctx.pressure_force(a)
    elif st == "btstep":
        ctx.btstep(cs, a)
    elif st == "advect_tracer":
        ctx.advect_tracer(cs, a)
    elif st == "thickness_diffuse":
        ctx.thickness_diffuse(cs, a)
    elif st == "mixedlayer_restrat":
        ctx.mixedlayer_restrat(cs, a["h"], a["uhtr"], a["vhtr"], a["T"], a["S"], a["ustar"], a["dt"], a["h_MLD"], a["Rd_dx_h"])
    elif st == "tracer_hordiff":
        ctx.tracer_hordiff(cs, a)
    else:
        raise KeyError(st)
    ctx.close()
    return collect(dom, c["outputs"], a, cs)
