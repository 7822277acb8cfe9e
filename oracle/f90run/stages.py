"""The hot-path stages run from the REFERENCE'S OWN SOURCE (translated by oracle/f90run), with the calling convention of
oracle/pyoracle.py: (dom, grid, gv, cs, args) dictionaries in, results written into args / cs in place.  TEST INFRASTRUCTURE ONLY.

tests/test_reference_f90.py calls a stage here and the same stage of the C++ oracle on copies of the same seeded inputs and
compares every output bit for bit over the computational domain.  Each function cites the reference entry it executes."""
import numpy as np

from . import adapt, halo, load, new, rt
from .rt import FArray, NS

_REF = {}
# tests/test_fortran_shims_executed.py: run the same stages with the shadow copies fortran/install_shims.py makes of some modules
SHADOW = dict(files={}, extra_files=[], stubs={}, post_load=None, tag=None)


def ref(*paths):
    key = tuple(paths) + (SHADOW["tag"],)
    if key not in _REF:
        files = [SHADOW["files"].get(p, p) for p in paths]
        files += [f for f in SHADOW["extra_files"] if f not in files]
        for p, q in SHADOW["files"].items():   # a shadowed module another shadowed module uses (its accessors)
            if q not in files and SHADOW["tag"] is not None:
                files.append(q)
        stubs = dict(halo.STUBS)
        stubs.update(SHADOW["stubs"])
        _REF[key] = load(files, extra_stubs=stubs)
        if SHADOW["post_load"] is not None:
            SHADOW["post_load"](_REF[key])
    return _REF[key]


def _types(dom, grid, gv, us=None):
    G = adapt.grid_type(dom, grid)
    G.domain = halo.domain_type(dom)
    return G, adapt.vgrid_type(dom, gv), adapt.unit_scale_type(us)


def _fa(dom, a, skip=()):
    return {k: adapt.farr(dom, v) for k, v in a.items() if isinstance(v, np.ndarray) and k not in skip and not k.startswith("_")}


def _back(fa, a, keys):
    for k in keys:
        if k in fa and isinstance(a.get(k), np.ndarray):
            adapt.back(fa[k], a[k])


def _set(CS, cs, logical=()):
    for k, v in cs.items():
        if isinstance(v, np.ndarray) or v is None:
            continue
        setattr(CS, k.lower(), bool(v) if k in logical else v)
    return CS


def _pbv(dom, a, ushape, vshape):
    return NS(por_face_areau=adapt.farr(dom, a.get("por_face_areaU", np.ones(ushape))),
              por_face_areav=adapt.farr(dom, a.get("por_face_areaV", np.ones(vshape))),
              por_layer_widthu=None, por_layer_widthv=None)


# ---------------------------------------------------------------------------------------------------------------------------
CONT_LOGICAL = ("upwind_1st", "monotonic", "simple_2nd", "aggress_adjust", "vol_CFL", "better_iter", "use_visc_rem_max",
                "marginal_faces")


def continuity(dom, grid, gv, cs, a):
    """continuity_PPM, src/core/MOM_continuity_PPM.F90:86-194 and everything it calls in that module"""
    R = ref("src/core/MOM_continuity_PPM.F90")
    F = R["mom_continuity_ppm"]
    G, GV, US = _types(dom, grid, gv)
    CS = _set(new(R, "mom_continuity_ppm", "continuity_ppm_cs", initialized=True), cs, CONT_LOGICAL)
    fa = _fa(dom, a)
    if a["h"] is a["hin"]:
        fa["h"] = fa["hin"]
    BT = None
    if a.get("BT_cont") is not None:
        BT = NS(**{k: adapt.farr(dom, v) for k, v in a["BT_cont"].items() if v is not None})
    F["continuity_ppm"](fa["u"], fa["v"], fa["hin"], fa["h"], fa["uh"], fa["vh"], a["dt"], G, GV, US, CS, None,
                        _pbv(dom, a, a["uh"].shape, a["vh"].shape), uhbt=fa.get("uhbt"), vhbt=fa.get("vhbt"),
                        visc_rem_u=fa.get("visc_rem_u"), visc_rem_v=fa.get("visc_rem_v"), u_cor=fa.get("u_cor"),
                        v_cor=fa.get("v_cor"), bt_cont=BT, du_cor=fa.get("du_cor"), dv_cor=fa.get("dv_cor"))
    _back(fa, a, ("h", "uh", "vh", "u_cor", "v_cor", "du_cor", "dv_cor"))
    if BT is not None:
        for k, v in a["BT_cont"].items():
            if v is not None:
                adapt.back(getattr(BT, k.lower()), v)


def coradcalc(dom, grid, gv, cs, a):
    """CorAdCalc, src/core/MOM_CoriolisAdv.F90:125-965, and gradKE :969-1060"""
    R = ref("src/core/MOM_CoriolisAdv.F90")
    F = R["mom_coriolisadv"]
    G, GV, US = _types(dom, grid, gv)
    CS = _set(new(R, "mom_coriolisadv", "coriolisadv_cs", initialized=True), cs, ("no_slip", "bound_Coriolis", "Coriolis_En_Dis"))
    fa = _fa(dom, a)
    AD = NS(gradkeu=fa.get("gradKEu"), gradkev=fa.get("gradKEv"), rv_x_u=None, rv_x_v=None)
    if "RV" in a:
        CS.id_rv, CS.id_pv = 1, 1   # the diagnostics are filled only when registered; post_data is a stub
    F["coradcalc"](fa["u"], fa["v"], fa["h"], fa["uh"], fa["vh"], fa["CAu"], fa["CAv"], None, AD, G, GV, US, CS,
                   _pbv(dom, a, a["u"].shape, a["v"].shape))
    _back(fa, a, ("CAu", "CAv", "gradKEu", "gradKEv"))


# ---------------------------------------------------------------------------------------------------------------------------
BT_LOGICAL = ("Sadourny", "BT_project_velocity", "strong_drag", "bound_BT_corr", "BT_cont_bounds", "wt_uv_bug", "visc_rem_u_uh0",
              "adjust_BT_cont", "use_wide_halos", "use_old_coriolis_bracket_bug")
BT_WIDE = ("IareaT", "IareaT_OBCmask", "bathyT", "IdxCu", "IdyCv", "q_D", "D_u_Cor", "D_v_Cor", "ua_polarity", "va_polarity",
           "OBCmask_u", "OBCmask_v")
BT_GSIZED = ("frhatu", "frhatv", "eta_cor", "eta_cor_bound", "IDatu", "IDatv", "ubtav", "vbtav")


def wide_farr(dom, a):
    """a numpy array on the wide-halo barotropic memory domain -> FArray with bounds (isdw[-1]:iedw, jsdw[-1]:jedw)"""
    ni, nj = dom.iedw - dom.isdw + 1, dom.jedw - dom.jsdw + 1
    sh = a.shape[-2:]
    return FArray.from_numpy(a, (dom.isdw - (sh[1] - ni), dom.jsdw - (sh[0] - nj)))


def barotropic_cs(R, dom, grid, gv, cs, wide_metrics):
    """barotropic_CS (src/core/MOM_barotropic.F90:110-366) with the members barotropic_init (:5301-6190) would have set for
    the frozen option set; wide_metrics = (dy_Cu, dx_Cv) on the wide memory domain"""
    CS = _set(new(R, "mom_barotropic", "barotropic_cs", module_is_initialized=True), cs, BT_LOGICAL)
    for k in BT_WIDE:
        if cs.get(k) is not None:
            setattr(CS, k.lower(), wide_farr(dom, cs[k]))
    for k in BT_GSIZED:
        if cs.get(k) is not None:
            setattr(CS, k.lower(), adapt.farr(dom, cs[k]))
    CS.dy_cu, CS.dx_cv = wide_farr(dom, wide_metrics[0]), wide_farr(dom, wide_metrics[1])
    CS.isdw, CS.iedw, CS.jsdw, CS.jedw = int(dom.isdw), int(dom.iedw), int(dom.jsdw), int(dom.jedw)
    CS.bt_domain = halo.domain_type(dom)
    CS.answer_date = 99991231
    CS.split = True
    CS.linearized_bt_pv = True
    CS.bt_coriolis_scale = 1.0
    CS.dtbt_fraction = 0.98
    CS.nonlin_cont_update_period = 1
    CS.rho_bt_lin = gv["Rho0"]
    CS.hvel_scheme = 4
    CS.dtbt_max = 0.0
    for k in ("integral_bt_cont", "integral_obcs", "nonlinear_continuity", "gradual_bt_ics", "nonlin_stress", "clip_velocity",
              "dynamic_psurf", "calculate_sal", "linear_wave_drag", "use_filter", "linear_freq_drag", "debug", "debug_bt",
              "tidal_sal_flather", "tidal_sal_bug", "debug_wide_halos"):
        setattr(CS, k, False)
    return CS


def wide_metrics(dom, land_blocks, seed=None):
    from mom6_b200 import synthetic
    from mom6_b200.api import make_domain
    ni, nj = dom.iec - dom.isc + 1, dom.jec - dom.jsc + 1
    wh = dom.isc - dom.isdw
    domw = make_domain(ni, nj, nk=int(dom.nk), halo=wh, whalo=wh, cyclic_x=bool(dom.cyclic_x), cyclic_y=bool(dom.cyclic_y))
    gw = synthetic.make_grid(domw, land_blocks, synthetic.SEED if seed is None else seed)
    return gw["dy_Cu"], gw["dx_Cv"]


def bt_cont_type(dom, b):
    BT = NS(**{k: adapt.farr(dom, v) for k, v in b.items() if v is not None})
    BT.pass_polarity_bt, BT.pass_fa_uv = NS(), NS()
    return BT


def btstep(dom, grid, gv, cs, a, wide):
    """btstep, src/core/MOM_barotropic.F90:455-2172, with btstep_timeloop :2175-2832 and every helper it calls"""
    R = ref("src/core/MOM_barotropic.F90")
    F = R["mom_barotropic"]
    G, GV, US = _types(dom, grid, gv)
    CS = barotropic_cs(R, dom, grid, gv, cs, wide)
    fa = _fa(dom, a)
    forces = NS(taux=fa["taux"], tauy=fa["tauy"])
    BT = bt_cont_type(dom, a["BT_cont"])
    F["btstep"](fa["U_in"], fa["V_in"], fa["eta_in"], a["dt"], fa["bc_accel_u"], fa["bc_accel_v"], forces, fa["pbce"],
                fa["eta_PF_in"], fa["U_Cor"], fa["V_Cor"], fa["accel_layer_u"], fa["accel_layer_v"], fa["eta_out"], fa["uhbtav"],
                fa["vhbtav"], G, GV, US, CS, fa["visc_rem_u"], fa["visc_rem_v"], None, NS(), None, BT, None, fa.get("taux_bot"),
                fa.get("tauy_bot"), fa.get("uh0"), fa.get("vh0"), fa.get("u_uh0"), fa.get("v_vh0"), fa.get("etaav"))
    _back(fa, a, ("accel_layer_u", "accel_layer_v", "eta_out", "uhbtav", "vhbtav", "etaav"))
    for k in ("ubtav", "vbtav", "eta_cor"):
        adapt.back(getattr(CS, k), cs[k])
    return CS.nstep_last


# ---------------------------------------------------------------------------------------------------------------------------
HV_LOGICAL = ("Laplacian", "biharmonic", "no_slip", "bound_Kh", "better_bound_Kh", "bound_Ah", "better_bound_Ah",
              "backscatter_underbound", "Smagorinsky_Kh", "Smagorinsky_Ah", "bound_Coriolis", "use_land_mask", "add_LES_viscosity",
              "use_cont_thick", "use_cont_thick_bug")


def horizontal_viscosity(dom, grid, gv, cs, a):
    """horizontal_viscosity, src/parameterizations/lateral/MOM_hor_visc.F90:266-2290"""
    R = ref("src/parameterizations/lateral/MOM_hor_visc.F90")
    F = R["mom_hor_visc"]
    G, GV, US = _types(dom, grid, gv)
    CS = _set(new(R, "mom_hor_visc", "hor_visc_cs", initialized=True), cs, HV_LOGICAL)
    for k, v in cs.items():
        if isinstance(v, np.ndarray):
            setattr(CS, k.lower(), adapt.farr(dom, v))
    CS.answer_date = 99991231
    for k in ("anisotropic", "leith_kh", "leith_ah", "use_leithy", "modified_leith", "use_qg_leith_visc", "use_beta_in_leith",
              "use_gme", "use_zb2020", "smooth_ah", "debug", "res_scale_meke", "frictwork_bug", "ey24_ebt_bs"):
        setattr(CS, k, False)
    fa = _fa(dom, a)
    uh = fa.get("uh") or adapt.farr(dom, np.zeros_like(a["u"]))
    vh = fa.get("vh") or adapt.farr(dom, np.zeros_like(a["v"]))
    MEKE, VarMix, tv = NS(), NS(use_variable_mixing=False, resoln_scaled_kh=False, resoln_scaled_khth=False), NS()
    F["horizontal_viscosity"](fa["u"], fa["v"], fa["h"], uh, vh, fa["diffu"], fa["diffv"], MEKE, VarMix, G, GV, US, CS, tv,
                              a["dt"], hu_cont=fa.get("hu_cont"), hv_cont=fa.get("hv_cont"))
    _back(fa, a, ("diffu", "diffv"))


# ---------------------------------------------------------------------------------------------------------------------------
VV_LOGICAL = ("bottomdraglaw", "harmonic_visc", "direct_stress", "fixed_LOTW_ML", "apply_LOTW_floor", "dynamic_viscous_ML")


def _vertvisc_setup(dom, grid, gv, cs, a_u, a_v, h_u, h_v):
    R = ref("src/parameterizations/vertical/MOM_vert_friction.F90", "src/core/MOM_forcing_type.F90")
    G, GV, US = _types(dom, grid, gv)
    CS = _set(new(R, "mom_vert_friction", "vertvisc_cs", initialized=True), cs, VV_LOGICAL)
    GV.nkml = int(cs.get("nkml", 0))
    GV.dz_subroundoff = float(cs.get("dZ_subroundoff", 1.0e-30))
    nk = int(dom.nk)
    lbu, lbv = adapt.stagger_lb(dom, "u"), adapt.stagger_lb(dom, "v")
    CS.a_u, CS.a_v = FArray.from_numpy(a_u, lbu + (1,)), FArray.from_numpy(a_v, lbv + (1,))
    CS.h_u, CS.h_v = FArray.from_numpy(h_u, lbu + (1,)), FArray.from_numpy(h_v, lbv + (1,))
    for k in ("debug", "use_gl90_in_ssw", "stokesmixing"):
        setattr(CS, k, False)
    CS.pass_ke_uv = NS()
    # vertvisc_limit_vel (:2926-3120) with its defaults (vertvisc_init :3389-3402): truncation of velocities whose CFL exceeds 0.5
    CS.maxvel, CS.cfl_based_trunc = float(cs.get("maxvel", 3.0e8)), bool(cs.get("CFL_based_trunc", 1))
    CS.cfl_trunc = float(cs.get("CFL_trunc", 0.5))
    CS.cfl_report, CS.truncramptime = CS.cfl_trunc, 0.0
    CS.u_trunc_file, CS.v_trunc_file, CS.ntrunc = "", "", 0
    return R["mom_vert_friction"], G, GV, US, CS


def _interfaces(dom, a, st):
    """(nk+1, nj, ni) interface array at h- or q-points -> FArray (i, j, K=1:nk+1)"""
    if a is None:
        return None
    return FArray.from_numpy(a, adapt.stagger_lb(dom, st) + (1,))


def vertvisc_coef(dom, grid, gv, cs, a, a_u, a_v, h_u, h_v):
    """vertvisc_coef, src/parameterizations/vertical/MOM_vert_friction.F90:1357-1840 with find_coupling_coef :2314-2760"""
    F, G, GV, US, CS = _vertvisc_setup(dom, grid, gv, cs, a_u, a_v, h_u, h_v)
    f = lambda k: adapt.farr(dom, a[k])  # noqa: E731
    dz = adapt.farr(dom, gv["H_to_Z"] * a["h"])   # thickness_to_dz, Boussinesq: src/core/MOM_interface_heights.F90:790-801
    forces = NS(ustar=f("ustar"), tau_mag=None, frac_shelf_u=None, frac_shelf_v=None)
    visc = NS(kv_bbl_u=f("Kv_bbl_u"), kv_bbl_v=f("Kv_bbl_v"), bbl_thick_u=f("bbl_thick_u"), bbl_thick_v=f("bbl_thick_v"),
              kv_shear=_interfaces(dom, a.get("Kv_shear"), "h"), kv_shear_bu=_interfaces(dom, a.get("Kv_shear_Bu"), "q"))
    F["vertvisc_coef"](f("u"), f("v"), f("h"), dz, forces, visc, NS(), a["dt"], G, GV, US, CS, None, NS())
    for fa, out in ((CS.a_u, a_u), (CS.a_v, a_v), (CS.h_u, h_u), (CS.h_v, h_v)):
        adapt.back(fa, out)


def vertvisc_remnant(dom, grid, gv, cs, visc_rem_u, visc_rem_v, dt, a_u, a_v, h_u, h_v, Ray_u=None, Ray_v=None):
    """vertvisc_remnant, MOM_vert_friction.F90:1229-1330"""
    F, G, GV, US, CS = _vertvisc_setup(dom, grid, gv, cs, a_u, a_v, h_u, h_v)
    visc = NS(ray_u=adapt.farr(dom, Ray_u), ray_v=adapt.farr(dom, Ray_v))
    fu, fv = adapt.farr(dom, visc_rem_u), adapt.farr(dom, visc_rem_v)
    F["vertvisc_remnant"](visc, fu, fv, dt, G, GV, US, CS)
    adapt.back(fu, visc_rem_u); adapt.back(fv, visc_rem_v)


def vertvisc(dom, grid, gv, cs, a, a_u, a_v, h_u, h_v):
    """vertvisc, MOM_vert_friction.F90:557-1010"""
    F, G, GV, US, CS = _vertvisc_setup(dom, grid, gv, cs, a_u, a_v, h_u, h_v)
    fa = _fa(dom, a)
    forces = NS(taux=fa["taux"], tauy=fa["tauy"])
    visc = NS(ray_u=fa.get("Ray_u"), ray_v=fa.get("Ray_v"))
    F["vertvisc"](fa["u"], fa["v"], fa["h"], forces, visc, a["dt"], None, NS(), NS(), G, GV, US, CS,
                  fa.get("taux_bot"), fa.get("tauy_bot"))
    _back(fa, a, ("u", "v", "taux_bot", "tauy_bot"))
    return int(CS.ntrunc)


# ---------------------------------------------------------------------------------------------------------------------------
EOS_FILES = ("src/equation_of_state/MOM_EOS.F90", "src/equation_of_state/MOM_EOS_base_type.F90",
             "src/equation_of_state/MOM_EOS_Wright.F90", "src/equation_of_state/MOM_EOS_linear.F90")
PGF_FILES = ("src/core/MOM_PressureForce_FV.F90", "src/core/MOM_PressureForce_Montgomery.F90", "src/core/MOM_density_integrals.F90",
             "src/ALE/MOM_ALE.F90", "src/ALE/regrid_solvers.F90", "src/ALE/PLM_functions.F90", "src/ALE/PPM_functions.F90", "src/ALE/regrid_edge_values.F90") + EOS_FILES


def eos_type(R, cs):
    """EOS_type (src/equation_of_state/MOM_EOS.F90:107-152) the way EOS_init (:1540-1790) leaves it for LINEAR / WRIGHT"""
    form = int(cs["EOS_form"])
    if form == 0:
        return None
    E = new(R, "mom_eos", "eos_type")
    E.form_of_eos = form
    E.eos_quadrature = False
    E.compressible = True
    for k in ("Rho_T0_S0", "dRho_dT", "dRho_dS", "dRho_dp"):
        setattr(E, k.lower(), float(cs.get(k, 0.0)))
    one = lambda k: float(cs.get(k, 0.0)) or 1.0  # noqa: E731
    E.kg_m3_to_r, E.rl2_t2_to_pa, E.c_to_degc, E.s_to_ppt = one("kg_m3_to_R"), one("RL2_T2_to_Pa"), one("C_to_degC"), one("S_to_ppt")
    E.r_to_kg_m3, E.degc_to_c, E.ppt_to_s = 1.0 / E.kg_m3_to_r, 1.0 / E.c_to_degc, 1.0 / E.s_to_ppt
    if form == 1:
        E.type_ = new(R, "mom_eos_linear", "linear_eos")
        E.type_.rho_t0_s0, E.type_.drho_dt, E.type_.drho_ds = E.rho_t0_s0, E.drho_dt, E.drho_ds
        E.type_.drho_dp = E.drho_dp
    elif form == 3:
        E.type_ = new(R, "mom_eos_wright", "buggy_wright_eos")
    else:
        raise ValueError("EOS form outside the frozen option set")
    return E


def pressure_force(dom, grid, gv, cs, a, us=None):
    """PressureForce_FV_Bouss, src/core/MOM_PressureForce_FV.F90:947-2017, with int_density_dz / int_density_dz_generic_plm /
    _ppm (MOM_density_integrals.F90), the EOS modules, TS_PLM/PPM_edge_values (MOM_ALE.F90:1495-1660) and Set_pbce_Bouss"""
    R = ref(*PGF_FILES)
    F = R["mom_pressureforce_fv"]
    G, GV, US = _types(dom, grid, gv, us)
    G.z_ref = float(cs.get("Z_ref", 0.0))
    GV.dz_subroundoff = float(cs.get("dZ_subroundoff", 1.0e-30))
    GV.nk_rho_varies = 0
    if cs.get("Rlay") is not None:
        GV.rlay = FArray.from_numpy(np.asarray(cs["Rlay"], dtype=np.float64), (1,))
        GV.g_prime = FArray.from_numpy(np.asarray(cs["g_prime"], dtype=np.float64), (1,))
    CS = new(R, "mom_pressureforce_fv", "pressureforce_fv_cs", initialized=True)
    CS.masswghtinterp = int(cs["MassWghtInterp"])
    CS.use_ssh_in_z0p, CS.rho_ref_bug = bool(cs["use_SSH_in_Z0p"]), bool(cs["rho_ref_bug"])
    CS.rho_ref, CS.gfs_scale, CS.rho0 = float(cs["rho_ref"]), float(cs["GFS_scale"]), float(gv["Rho0"])
    CS.reconstruct = bool(cs.get("reconstruct", 0))
    CS.recon_scheme = int(cs.get("Recon_Scheme", 0)) or 1
    CS.boundary_extrap = bool(cs.get("boundary_extrap", 0))
    CS.use_inaccurate_pgf_rho_anom = bool(cs.get("use_inaccurate_pgf_rho_anom", 0))
    CS.masswghtinterpvanonly = bool(cs.get("MassWghtInterpVanOnly", 0))
    CS.h_nonvanished = float(cs.get("h_nonvanished", 0.0))
    for k in ("calculate_sal", "tides", "sal_use_bpa", "bq_sal_tides", "use_stanley_pgf", "reset_intxpa_integral", "debug",
              "correction_intxpa", "reset_intxpa_flattest"):
        setattr(CS, k, False)
    E = eos_type(R, cs)
    tv = NS(eqn_of_state=E, t=adapt.farr(dom, a.get("T")), s=adapt.farr(dom, a.get("S")), p_ref=0.0, vart=None)
    ALE = None
    if CS.reconstruct and E is not None:
        ALE = new(R, "mom_ale", "ale_cs")
        ALE.answer_date = int(cs.get("ALE_answer_date", 0)) or 99991231
        ALE.nk = int(dom.nk)
    fa = _fa(dom, a, skip=("T", "S"))
    F["pressureforce_fv_bouss"](fa["h"], tv, fa["PFu"], fa["PFv"], G, GV, US, CS, ALE, None, fa.get("p_atm"),
                                pbce=fa.get("pbce"), eta=fa.get("eta"))
    _back(fa, a, ("PFu", "PFv", "pbce", "eta"))


# ---------------------------------------------------------------------------------------------------------------------------
def advect_tracer(dom, grid, gv, cs, a):
    """advect_tracer, src/tracer/MOM_tracer_advect.F90:53-365, with advect_x :370-740 and advect_y :745-1130"""
    R = ref("src/tracer/MOM_tracer_advect.F90", "src/tracer/MOM_tracer_advect_schemes.F90")
    F = R["mom_tracer_advect"]
    G, GV, US = _types(dom, grid, gv)
    CS = new(R, "mom_tracer_advect", "tracer_advect_cs")
    CS.dt, CS.default_advect_scheme = float(cs["dt"]), int(cs["default_advect_scheme"])
    CS.usehuynhstencilbug = bool(cs.get("useHuynhStencilBug", 0))
    CS.debug = False
    CS.pass_uhr_vhr_t_hprev = NS()
    ntr = len(a["tr"])
    Tr = FArray.alloc("o", [(1, ntr)])
    fts = [adapt.farr(dom, t) for t in a["tr"]]
    sch = a.get("advect_scheme") if a.get("advect_scheme") is not None else [-1] * ntr
    cu = a.get("conc_underflow") if a.get("conc_underflow") is not None else [0.0] * ntr
    for m, ft in enumerate(fts):
        Tr.v[m] = NS(t=ft, advect_scheme=int(sch[m]), conc_underflow=float(cu[m]), ntr_index=m + 1, advection_xy=None, ad_x=None,
                     ad_y=None, ad2d_x=None, ad2d_y=None, tres=None)
    Reg = NS(ntr=ntr, tr=Tr)
    fa = _fa(dom, a)
    F["advect_tracer"](fa["h_end"], fa["uhtr"], fa["vhtr"], None, a["dt"], G, GV, US, CS, Reg,
                       x_first_in=(None if a.get("x_first_in") is None else bool(a["x_first_in"])),
                       max_iter_in=a.get("max_iter_in"), vol_prev=fa.get("vol_prev"),
                       update_vol_prev=(bool(a["update_vol_prev"]) if a.get("update_vol_prev") is not None else None),
                       uhr_out=fa.get("uhr_out"), vhr_out=fa.get("vhr_out"))
    _back(fa, a, ("vol_prev", "uhr_out", "vhr_out"))
    for ft, t in zip(fts, a["tr"]):
        adapt.back(ft, t)


# ---------------------------------------------------------------------------------------------------------------------------
STEP_FILES = ("src/core/MOM_dynamics_split_RK2.F90", "src/core/MOM_continuity_PPM.F90", "src/core/MOM_continuity.F90",
              "src/core/MOM_CoriolisAdv.F90", "src/core/MOM_PressureForce.F90", "src/parameterizations/lateral/MOM_hor_visc.F90",
              "src/parameterizations/vertical/MOM_vert_friction.F90", "src/parameterizations/vertical/MOM_set_viscosity.F90",
              "src/core/MOM_barotropic.F90", "src/core/MOM_forcing_type.F90", "src/core/MOM_interface_heights.F90") + PGF_FILES
STEP_CS_ARRAYS = ("CAu", "CAv", "CAu_pred", "CAv_pred", "PFu", "PFv", "diffu", "diffv", "visc_rem_u", "visc_rem_v", "u_accel_bt",
                  "v_accel_bt", "u_av", "v_av", "h_av", "pbce", "eta", "eta_PF", "uhbt", "vhbt", "taux_bot", "tauy_bot")
STEP_STATE = ("u_inst", "v_inst", "h", "uh", "vh", "uhtr", "vhtr", "eta_av")


def step_dyn_split_rk2(dom, grid, gv, css, cs, a, land_blocks=0):
    """step_MOM_dyn_split_RK2, src/core/MOM_dynamics_split_RK2.F90:294-1205, calling the reference's own continuity, CorAdCalc,
    PressureForce, horizontal_viscosity, vertvisc*, btcalc, bt_mass_source, set_dtbt and btstep"""
    R = ref(*STEP_FILES)
    F = R["mom_dynamics_split_rk2"]
    G, GV, US = _types(dom, grid, gv)
    nk = int(dom.nk)
    pf = css["pressureforce"]
    G.z_ref = float(pf.get("Z_ref", 0.0))
    GV.dz_subroundoff = float(pf.get("dZ_subroundoff", 1.0e-30))
    GV.nk_rho_varies, GV.nkml = 0, int(css["vertvisc"].get("nkml", 0))
    if pf.get("Rlay") is not None:
        GV.rlay = FArray.from_numpy(np.asarray(pf["Rlay"], dtype=np.float64), (1,))
        GV.g_prime = FArray.from_numpy(np.asarray(pf["g_prime"], dtype=np.float64), (1,))
    # ---- the stage control structures
    cont = _set(new(R, "mom_continuity_ppm", "continuity_ppm_cs", initialized=True), css["continuity"], CONT_LOGICAL)
    cor = _set(new(R, "mom_coriolisadv", "coriolisadv_cs", initialized=True), css["coriolisadv"],
               ("no_slip", "bound_Coriolis", "Coriolis_En_Dis"))
    hv = _set(new(R, "mom_hor_visc", "hor_visc_cs", initialized=True), css["hor_visc"], HV_LOGICAL)
    for k, v in css["hor_visc"].items():
        if isinstance(v, np.ndarray):
            setattr(hv, k.lower(), adapt.farr(dom, v))
    hv.answer_date = 99991231
    for k in ("anisotropic", "leith_kh", "leith_ah", "use_leithy", "modified_leith", "use_qg_leith_visc", "use_beta_in_leith",
              "use_gme", "use_zb2020", "smooth_ah", "debug", "res_scale_meke", "frictwork_bug", "ey24_ebt_bs"):
        setattr(hv, k, False)
    pfv = new(R, "mom_pressureforce_fv", "pressureforce_fv_cs", initialized=True)
    pfv.masswghtinterp = int(pf["MassWghtInterp"])
    pfv.use_ssh_in_z0p, pfv.rho_ref_bug = bool(pf["use_SSH_in_Z0p"]), bool(pf["rho_ref_bug"])
    pfv.rho_ref, pfv.gfs_scale, pfv.rho0 = float(pf["rho_ref"]), float(pf["GFS_scale"]), float(gv["Rho0"])
    pfv.reconstruct = bool(pf.get("reconstruct", 0))
    pfv.recon_scheme = int(pf.get("Recon_Scheme", 0)) or 1
    pfv.boundary_extrap = bool(pf.get("boundary_extrap", 0))
    pfv.use_inaccurate_pgf_rho_anom = bool(pf.get("use_inaccurate_pgf_rho_anom", 0))
    pfv.masswghtinterpvanonly = bool(pf.get("MassWghtInterpVanOnly", 0))
    pfv.h_nonvanished = float(pf.get("h_nonvanished", 0.0))
    for k in ("calculate_sal", "tides", "sal_use_bpa", "bq_sal_tides", "use_stanley_pgf", "reset_intxpa_integral", "debug",
              "correction_intxpa", "reset_intxpa_flattest"):
        setattr(pfv, k, False)
    pgf = NS(analytic_fv_pgf=True, pressureforce_fv=pfv)
    E = eos_type(R, pf)
    ALE = None
    if pfv.reconstruct and E is not None:
        ALE = new(R, "mom_ale", "ale_cs")
        ALE.answer_date = int(pf.get("ALE_answer_date", 0)) or 99991231
        ALE.nk = nk
    vv = _set(new(R, "mom_vert_friction", "vertvisc_cs", initialized=True), css["vertvisc"], VV_LOGICAL)
    lbu, lbv = adapt.stagger_lb(dom, "u"), adapt.stagger_lb(dom, "v")
    shu, shv = a["u_inst"].shape, a["v_inst"].shape
    vv.a_u = FArray.from_numpy(np.zeros((nk + 1,) + shu[1:]), lbu + (1,))
    vv.a_v = FArray.from_numpy(np.zeros((nk + 1,) + shv[1:]), lbv + (1,))
    vv.h_u, vv.h_v = FArray.from_numpy(np.zeros(shu), lbu + (1,)), FArray.from_numpy(np.zeros(shv), lbv + (1,))
    for k in ("debug", "use_gl90_in_ssw", "stokesmixing"):
        setattr(vv, k, False)
    vv.pass_ke_uv = NS()
    vvd = css["vertvisc"]
    vv.maxvel, vv.cfl_based_trunc = float(vvd.get("maxvel", 3.0e8)), bool(vvd.get("CFL_based_trunc", 1))
    vv.cfl_trunc = float(vvd.get("CFL_trunc", 0.5))
    vv.cfl_report, vv.truncramptime = vv.cfl_trunc, 0.0
    vv.u_trunc_file, vv.v_trunc_file, vv.ntrunc = "", "", 0
    GV.dz_subroundoff = float(css["vertvisc"].get("dZ_subroundoff", GV.dz_subroundoff))
    bt = barotropic_cs(R, dom, grid, gv, cs["barotropic"], wide_metrics(dom, land_blocks))
    bt.hvel_scheme = int(cs.get("hvel_scheme", 4))
    bt.dtbt_fraction = float(cs.get("dtbt_fraction", 0.98))
    bt.bt_coriolis_scale = float(cs.get("BT_Coriolis_scale", 1.0))
    bt.nonlinear_continuity = bool(cs.get("BT_Nonlinear_continuity", 0))
    bt.dtbt_max = float(cs.get("dtbt_max", 0.0))
    setv = new(R, "mom_set_visc", "set_visc_cs", initialized=True)
    setv.dynamic_viscous_ml = bool(css["vertvisc"].get("dynamic_viscous_ML", 0))
    # ---- MOM_dyn_split_RK2_CS
    CS = new(R, "mom_dynamics_split_rk2", "mom_dyn_split_rk2_cs", module_is_initialized=True)
    fcs = {k: adapt.farr(dom, cs[k]) for k in STEP_CS_ARRAYS}
    for k, v in fcs.items():
        setattr(CS, k.lower(), v)
    BT = bt_cont_type(dom, cs["BT_cont"])
    CS.bt_cont = BT
    CS.be, CS.begw = float(cs["be"]), float(cs["begw"])
    CS.split_bottom_stress, CS.store_cau = bool(cs["split_bottom_stress"]), bool(cs["store_CAu"])
    CS.cau_pred_stored, CS.visc_rem_dt_bug = bool(cs["CAu_pred_stored"]), bool(cs["visc_rem_dt_bug"])
    CS.dtbt_use_bt_cont = bool(cs.get("dtbt_use_bt_cont", 0))
    CS.bt_use_layer_fluxes = True
    for k in ("debug", "debug_obc", "fpmix", "calculate_sal", "use_tides", "remap_aux"):
        setattr(CS, k, False)
    CS.continuity_csp, CS.coriolisadv, CS.hor_visc, CS.pressureforce_csp = cont, cor, hv, pgf
    CS.vertvisc_csp, CS.barotropic_csp, CS.set_visc_csp, CS.ale_csp = vv, bt, setv, ALE
    CS.adp, CS.cdp, CS.ad_pred = NS(), NS(), NS()
    for k in ("pass_eta", "pass_visc_rem", "pass_uvp", "pass_hp_uv", "pass_uv", "pass_h", "pass_av_uvh"):
        setattr(CS, k, NS())
    # ---- arguments
    fa = {k: adapt.farr(dom, a[k]) for k in STEP_STATE}
    f2 = lambda k: adapt.farr(dom, a.get(k))  # noqa: E731
    tv = NS(eqn_of_state=E, t=f2("T"), s=f2("S"), p_ref=0.0, vart=None, spv_avg=None, valid_spv_halo=-1)
    visc = NS(kv_bbl_u=f2("Kv_bbl_u"), kv_bbl_v=f2("Kv_bbl_v"), bbl_thick_u=f2("bbl_thick_u"), bbl_thick_v=f2("bbl_thick_v"),
              kv_shear=_interfaces(dom, a.get("Kv_shear"), "h"), kv_shear_bu=_interfaces(dom, a.get("Kv_shear_Bu"), "q"),
              ray_u=f2("Ray_u"), ray_v=f2("Ray_v"))
    forces = NS(taux=f2("taux"), tauy=f2("tauy"), ustar=f2("ustar"), tau_mag=None, p_surf=f2("p_surf"))
    pbv = NS(por_face_areau=adapt.farr(dom, np.ones(shu)), por_face_areav=adapt.farr(dom, np.ones(shv)))
    VarMix = NS(use_variable_mixing=False, resoln_scaled_kh=False, resoln_scaled_khth=False)
    F["step_mom_dyn_split_rk2"](fa["u_inst"], fa["v_inst"], fa["h"], tv, visc, 0.0, a["dt"], forces, forces.p_surf, forces.p_surf,
                                fa["uh"], fa["vh"], fa["uhtr"], fa["vhtr"], fa["eta_av"], G, GV, US, CS, bool(a.get("calc_dtbt", 0)),
                                VarMix, NS(), None, pbv)
    _back(fa, a, STEP_STATE)
    for k, v in fcs.items():
        adapt.back(v, cs[k])
    for k, v in cs["BT_cont"].items():
        if v is not None:
            adapt.back(getattr(BT, k.lower()), v)
    for k in ("ubtav", "vbtav", "eta_cor", "frhatu", "frhatv"):
        adapt.back(getattr(bt, k), cs["barotropic"][k])
    cs["CAu_pred_stored"] = int(bool(CS.cau_pred_stored))
    cs["dtbt_max"] = float(bt.dtbt_max)
    cs["barotropic"]["dtbt"] = float(bt.dtbt)


# ---------------------------------------------------------------------------------------------------------------------------
ALE_FILES = ("src/core/MOM.F90", "src/ALE/MOM_ALE.F90", "src/ALE/MOM_regridding.F90", "src/ALE/MOM_remapping.F90",
             "src/ALE/coord_zlike.F90", "src/ALE/regrid_consts.F90", "src/ALE/regrid_interp.F90", "src/ALE/regrid_edge_values.F90",
             "src/ALE/regrid_solvers.F90", "src/ALE/PCM_functions.F90", "src/ALE/PLM_functions.F90", "src/ALE/PPM_functions.F90",
             "src/ALE/PQM_functions.F90", "src/ALE/P1M_functions.F90", "src/ALE/P3M_functions.F90", "src/ALE/polynomial_functions.F90",
             "src/tracer/MOM_tracer_registry.F90", "src/tracer/MOM_tracer_types.F90", "src/parameterizations/vertical/MOM_set_viscosity.F90",
             "src/core/MOM_dynamics_split_RK2.F90", "src/core/MOM_interface_heights.F90")
REMAP_LOGICAL = ("boundary_extrapolation", "force_bounds_in_subcell", "force_bounds_in_target", "om4_remap_via_sub_cells")


def remapping_cs(R, d):
    """remapping_CS (src/ALE/MOM_remapping.F90:42-80) as initialize_remapping (:1640-1700) resolves it"""
    C = _set(new(R, "mom_remapping", "remapping_cs"), d, REMAP_LOGICAL)
    C.degree = {0: 0, 2: 1, 4: 2, 5: 2}[int(d["remapping_scheme"])]
    return C


def regridding_cs(R, d):
    """regridding_CS (src/ALE/MOM_regridding.F90:47-130) for the Z* coordinate, as initialize_regridding (:200-800) and
    set_regrid_params (:2380-2470) leave it; zlike_CS from init_coord_zlike (src/ALE/coord_zlike.F90:37-50)"""
    C = new(R, "mom_regridding", "regridding_cs")
    nk = int(d["nk"])
    C.nk, C.regridding_scheme = nk, int(d["regridding_scheme"])
    C.min_thickness = float(d["min_thickness"])
    C.old_grid_weight = float(d["old_grid_weight"])
    C.depth_of_time_filter_shallow = float(d["depth_of_time_filter_shallow"])
    C.depth_of_time_filter_deep = float(d["depth_of_time_filter_deep"])
    res = np.asarray(d["coordinateResolution"], dtype=np.float64)
    C.coordinateresolution = FArray.from_numpy(res.copy(), (1,))
    Z = new(R, "coord_zlike", "zlike_cs")
    Z.nk, Z.min_thickness = nk, C.min_thickness
    Z.coordinateresolution = FArray.from_numpy(res.copy(), (1,))
    C.zlike_cs = Z
    return C


def ale_regridding_and_remapping(dom, grid, gv, ale, a, dyn_cs=None):
    """ALE_regridding_and_remapping, src/core/MOM.F90:1751-1926, and everything it calls in MOM_ALE / MOM_regridding /
    MOM_remapping / coord_zlike, remap_dyn_split_RK2_aux_vars (MOM_dynamics_split_RK2.F90:1211-1240) and
    remap_vertvisc_aux_vars (MOM_set_viscosity.F90:2849-2873)"""
    R = ref(*ALE_FILES)
    F = R["mom"]
    G, GV, US = _types(dom, grid, gv)
    G.z_ref = float(ale["regridCS"].get("Z_ref", 0.0))
    nk = int(dom.nk)
    A = new(R, "mom_ale", "ale_cs")
    A.regridcs, A.remapcs, A.vel_remapcs = regridding_cs(R, ale["regridCS"]), remapping_cs(R, ale["remapCS"]), \
        remapping_cs(R, ale["vel_remapCS"])
    A.regrid_time_scale = float(ale["regrid_time_scale"])
    A.nk, A.answer_date = nk, 99991231
    A.bbl_h_vel_mask, A.h_vel_mask = 0.0, 0.0
    for k in ("remap_uv_using_old_alg", "partial_cell_vel_remap", "use_hybgen_unmix", "do_conv_adj", "conserve_ke", "debug",
              "show_call_tree", "remap_after_initialization"):
        setattr(A, k, False)
    ntr = len(a["tr"])
    fts = [adapt.farr(dom, t) for t in a["tr"]]
    Tr = FArray.alloc("o", [(1, ntr)])
    cu = a.get("conc_underflow") if a.get("conc_underflow") is not None else [0.0] * ntr
    for m, ft in enumerate(fts):
        T = new(R, "mom_tracer_types", "tracer_type")
        T.t, T.conc_underflow, T.remap_tr, T.ntr_index = ft, float(cu[m]), True, m + 1
        Tr.v[m] = T
    Reg = NS(ntr=ntr, tr=Tr)
    A.do_tendency_diag = FArray.alloc("l", [(1, ntr)])
    tv = NS(t=(fts[a["iT"]] if a.get("iT", -1) >= 0 else None), s=(fts[a["iS"]] if a.get("iS", -1) >= 0 else None),
            spv_avg=None, valid_spv_halo=-1, eqn_of_state=None)
    visc = NS(kd_shear=_interfaces(dom, a.get("Kd_shear"), "h"), kv_shear=_interfaces(dom, a.get("Kv_shear"), "h"),
              kv_shear_bu=_interfaces(dom, a.get("Kv_shear_Bu"), "q"))
    D = None
    fd = {}
    if dyn_cs is not None:
        D = NS(remap_aux=True, store_cau=bool(dyn_cs.get("store_CAu", 0)), cau_pred_stored=bool(dyn_cs.get("CAu_pred_stored", 0)),
               be=float(dyn_cs.get("be", 0.6)), begw=float(dyn_cs.get("begw", 0.0)))
        for k in ("diffu", "diffv", "CAu_pred", "CAv_pred", "u_av", "v_av"):
            fd[k] = adapt.farr(dom, dyn_cs[k])
            setattr(D, k.lower(), fd[k])
    CS = NS(ale_csp=A, debug=False, dyn_split_rk2_csp=D, split=True, use_alt_split=False, tracer_reg=Reg,
            remap_aux_vars=bool(ale.get("remap_aux_vars", 0)) and D is not None, remap_uv_using_old_alg=False, use_particles=False,
            visc=visc, obc=None, tv=tv, use_ale_algorithm=True, frac_shelf_h=None, diag=None)
    fu, fv, fh = adapt.farr(dom, a["u"]), adapt.farr(dom, a["v"]), adapt.farr(dom, a["h"])
    F["ale_regridding_and_remapping"](CS, G, GV, US, fu, fv, fh, tv, float(a["dtdia"]), 0.0)
    adapt.back(fu, a["u"]); adapt.back(fv, a["v"]); adapt.back(fh, a["h"])
    for ft, t in zip(fts, a["tr"]):
        adapt.back(ft, t)
    for k in ("Kd_shear", "Kv_shear", "Kv_shear_Bu"):
        if a.get(k) is not None:
            adapt.back(getattr(visc, k.lower()), a[k])
    for k, v in fd.items():
        adapt.back(v, dyn_cs[k])
    ale["regridCS"]["old_grid_weight"] = float(A.regridcs.old_grid_weight)


# ---------------------------------------------------------------------------------------------------------------------------
TD_FILES = ("src/parameterizations/lateral/MOM_thickness_diffuse.F90", "src/core/MOM_isopycnal_slopes.F90",
            "src/core/MOM_interface_heights.F90") + EOS_FILES


def thickness_diffuse(dom, grid, gv, cs, a):
    """thickness_diffuse, src/parameterizations/lateral/MOM_thickness_diffuse.F90:134-630, thickness_diffuse_full :635-1530,
    streamfn_solver :1535-1570, vert_fill_TS (MOM_isopycnal_slopes.F90:560-640) and the EOS derivative routines"""
    R = ref(*TD_FILES)
    F = R["mom_thickness_diffuse"]
    G, GV, US = _types(dom, grid, gv)
    G.obcmaskcu, G.obcmaskcv = G.mask2dcu, G.mask2dcv   # no open boundaries: MOM_grid.F90 sets OBCmaskCu = mask2dCu
    GV.dz_subroundoff = float(cs["dZ_subroundoff"])
    GV.nkml, GV.semi_boussinesq = 0, False
    CS = new(R, "mom_thickness_diffuse", "thickness_diffuse_cs", initialized=True)
    for k in ("Khth", "Khth_Min", "Khth_Max", "max_Khth_CFL", "slope_max", "kappa_smooth", "FGNV_scale", "N2_floor"):
        setattr(CS, k.lower(), float(cs[k]))
    for k in ("thickness_diffuse", "read_khth", "detangle_interfaces", "use_FGNV_streamfn", "use_stanley_gm",
              "use_GME_thickness_diffuse"):
        setattr(CS, k.lower(), bool(cs[k]))
    CS.kh_eta_bg, CS.kh_eta_vel, CS.khth_slope_cff = 0.0, 0.0, 0.0
    for k in ("debug", "meke_geometric", "use_kh_in_meke", "gm_src_alt", "full_depth_khth_min", "use_gm_work_bug", "meke_src_slope_bug"):
        setattr(CS, k, False)
    CS.meke_src_answer_date, CS.meke_geom_answer_date = 99991231, 99991231
    f2 = lambda k: adapt.farr(dom, a.get(k))  # noqa: E731
    V = NS(use_variable_mixing=bool(cs["use_variable_mixing"]), resoln_scaled_khth=bool(cs["Resoln_scaled_KhTh"]),
           depth_scaled_khth=bool(cs["Depth_scaled_KhTh"]), use_stored_slopes=bool(cs["use_stored_slopes"]),
           use_visbeck=bool(cs["use_Visbeck"]), use_qg_leith_gm=bool(cs["use_QG_Leith_GM"]), khth_struct=None,
           res_fn_u=f2("Res_fn_u"), res_fn_v=f2("Res_fn_v"), slope_x=_interfaces(dom, a.get("slope_x"), "u"),
           slope_y=_interfaces(dom, a.get("slope_y"), "v"), cg1=f2("cg1"))
    M = NS(kh=(f2("MEKE_Kh") if cs["use_MEKE_Kh"] else None), khth_fac=float(cs["MEKE_KhTh_fac"]), gm_src=None, meke=None)
    fa = {k: adapt.farr(dom, a[k]) for k in ("h", "uhtr", "vhtr")}
    tv = NS(eqn_of_state=eos_type(R, cs), t=f2("T"), s=f2("S"), p_surf=f2("p_surf"), spv_avg=None, vart=None)
    CDp = NS(uhgm=f2("uhGM"), vhgm=f2("vhGM"))
    F["thickness_diffuse"](fa["h"], fa["uhtr"], fa["vhtr"], tv, float(a["dt"]), G, GV, US, M, V, CDp, CS, NS())
    _back(fa, a, ("h", "uhtr", "vhtr"))
    for k, m in (("uhGM", CDp.uhgm), ("vhGM", CDp.vhgm)):
        if a.get(k) is not None:
            adapt.back(m, a[k])


# ---------------------------------------------------------------------------------------------------------------------------
MLE_FILES = ("src/parameterizations/lateral/MOM_mixed_layer_restrat.F90", "src/core/MOM_forcing_type.F90",
             "src/core/MOM_interface_heights.F90") + EOS_FILES


def mixedlayer_restrat(dom, grid, gv, cs, a):
    """mixedlayer_restrat, src/parameterizations/lateral/MOM_mixed_layer_restrat.F90:149-186 -> mixedlayer_restrat_OM4 :189-714
    (with mu :1545-1575 and rmean2ts :1100-1130)"""
    R = ref(*MLE_FILES)
    F = R["mom_mixed_layer_restrat"]
    G, GV, US = _types(dom, grid, gv)
    G.obcmaskcu, G.obcmaskcv = G.mask2dcu, G.mask2dcv
    GV.nkml, GV.semi_boussinesq = 0, False
    CS = new(R, "mom_mixed_layer_restrat", "mixedlayer_restrat_cs", initialized=True)
    for k in ("ml_restrat_coef", "ml_restrat_coef2", "front_length", "MLE_MLD_decay_time", "MLE_MLD_decay_time2", "MLE_MLD_stretch",
              "MLE_tail_dh", "ustar_min", "vonKar", "MLE_density_diff"):
        setattr(CS, k.lower(), float(cs[k]))
    for k in ("MLE_use_PBL_MLD", "use_Stanley_ML", "use_Bodner", "fl_from_file"):
        setattr(CS, k.lower(), bool(cs[k]))
    CS.debug = False
    fml, fmls = adapt.farr(dom, cs["MLD_filtered"]), adapt.farr(dom, cs["MLD_filtered_slow"])
    CS.mld_filtered, CS.mld_filtered_slow = fml, fmls
    f2 = lambda k: adapt.farr(dom, a.get(k))  # noqa: E731
    fa = {k: adapt.farr(dom, a[k]) for k in ("h", "uhtr", "vhtr")}
    tv = NS(eqn_of_state=eos_type(R, cs), t=f2("T"), s=f2("S"), vart=None, spv_avg=None)
    forces = NS(ustar=f2("ustar"), tau_mag=None)
    V = NS(rd_dx_h=f2("Rd_dx_h"))
    F["mixedlayer_restrat"](fa["h"], fa["uhtr"], fa["vhtr"], tv, forces, float(a["dt"]), None, f2("h_MLD"), None, V, G, GV, US, CS)
    _back(fa, a, ("h", "uhtr", "vhtr"))
    adapt.back(fml, cs["MLD_filtered"]); adapt.back(fmls, cs["MLD_filtered_slow"])


# ---------------------------------------------------------------------------------------------------------------------------
def tracer_hordiff(dom, grid, gv, cs, a):
    """tracer_hordiff, src/tracer/MOM_tracer_hor_diff.F90:119-690 (the along-layer path: no neutral / boundary diffusion)"""
    R = ref("src/tracer/MOM_tracer_hor_diff.F90", "src/tracer/MOM_tracer_types.F90")
    F = R["mom_tracer_hor_diff"]
    G, GV, US = _types(dom, grid, gv)
    GV.nkml, GV.nk_rho_varies = 0, 0
    CS = new(R, "mom_tracer_hor_diff", "tracer_hor_diff_cs")
    for k in ("KhTr", "KhTr_min", "KhTr_max", "KhTr_passivity_coeff", "KhTr_passivity_min", "KhTr_Slope_Cff", "max_diff_CFL"):
        setattr(CS, k.lower(), float(cs[k]))
    for k in ("check_diffusive_CFL", "use_neutral_diffusion", "use_hor_bnd_diffusion", "Diffuse_ML_interior"):
        setattr(CS, k.lower(), bool(cs[k]))
    for k in ("debug", "show_call_tree", "first_call", "full_depth_khtr_min", "khtr_use_vert_struct", "recalc_neutral_surf"):
        setattr(CS, k, False)
    CS.ml_khtr_scale = 1.0
    CS.pass_t = NS()
    ntr = len(a["tr"])
    fts = [adapt.farr(dom, t) for t in a["tr"]]
    Tr = FArray.alloc("o", [(1, ntr)])
    cu = a.get("conc_underflow") if a.get("conc_underflow") is not None else [0.0] * ntr
    dfx = a.get("df_x") or [None] * ntr
    dfy = a.get("df_y") or [None] * ntr
    fdx, fdy = [adapt.farr(dom, x) for x in dfx], [adapt.farr(dom, x) for x in dfy]
    for m, ft in enumerate(fts):
        T = new(R, "mom_tracer_types", "tracer_type")
        T.t, T.conc_underflow, T.df_x, T.df_y, T.name = ft, float(cu[m]), fdx[m], fdy[m], f"tr{m}"
        Tr.v[m] = T
    Reg = NS(ntr=ntr, tr=Tr)
    f2 = lambda k: adapt.farr(dom, a.get(k))  # noqa: E731
    V = NS(use_variable_mixing=bool(cs["use_variable_mixing"]), resoln_scaled_khtr=bool(cs["Resoln_scaled_KhTr"]),
           res_fn_h=f2("Res_fn_h"), rd_dx_h=f2("Rd_dx_h"), l2u=f2("L2u"), l2v=f2("L2v"), sn_u=f2("SN_u"), sn_v=f2("SN_v"),
           khtr_struct=None, ebt_struct=None)
    M = NS(kh=(f2("MEKE_Kh") if cs["use_MEKE_Kh"] else None), khtr_fac=float(cs["MEKE_KhTr_fac"]))
    F["tracer_hordiff"](adapt.farr(dom, a["h"]), float(a["dt"]), M, V, NS(), G, GV, US, CS, Reg, NS(t=None, s=None, p_surf=None))
    for ft, t in zip(fts, a["tr"]):
        adapt.back(ft, t)
    for fl, ol in ((fdx, dfx), (fdy, dfy)):
        for fx, ox in zip(fl, ol):
            if ox is not None:
                adapt.back(fx, ox)


# ---------------------------------------------------------------------------------------------------------------------------
def _bt_small_cs(R, dom, gv):
    """barotropic_CS with only what btcalc / bt_mass_source / set_dtbt read"""
    CS = new(R, "mom_barotropic", "barotropic_cs", module_is_initialized=True)
    CS.split, CS.debug, CS.calculate_sal, CS.tidal_sal_bug = True, False, False, False
    CS.rho_bt_lin = float(gv["Rho0"])
    CS.isdw, CS.iedw, CS.jsdw, CS.jedw = int(dom.isdw), int(dom.iedw), int(dom.jsdw), int(dom.jedw)
    return CS


def _wide_from_g(dom, a):
    """a G-sized h-point array placed in a wide-halo array (zero outside G's memory domain)"""
    wi, wj = dom.isd - dom.isdw, dom.jsd - dom.jsdw
    w = np.zeros((dom.jedw - dom.jsdw + 1, dom.iedw - dom.isdw + 1))
    w[wj:wj + a.shape[0], wi:wi + a.shape[1]] = a
    return FArray.from_numpy(w, (int(dom.isdw), int(dom.jsdw)))


def btcalc(dom, grid, gv, a):
    """btcalc, src/core/MOM_barotropic.F90:4360-4605"""
    R = ref("src/core/MOM_barotropic.F90")
    G, GV, US = _types(dom, grid, gv)
    CS = _bt_small_cs(R, dom, gv)
    CS.hvel_scheme = int(a["hvel_scheme"])
    fu, fv = adapt.farr(dom, a["frhatu"]), adapt.farr(dom, a["frhatv"])
    CS.frhatu, CS.frhatv = fu, fv
    CS.bathyt = _wide_from_g(dom, a["bathyT"])
    R["mom_barotropic"]["btcalc"](adapt.farr(dom, a["h"]), G, GV, CS, adapt.farr(dom, a.get("h_u")), adapt.farr(dom, a.get("h_v")),
                                 bool(a.get("may_use_default", 0)), None)
    adapt.back(fu, a["frhatu"]); adapt.back(fv, a["frhatv"])


def bt_mass_source(dom, grid, gv, h, eta, set_cor, eta_cor):
    """bt_mass_source, src/core/MOM_barotropic.F90:5243-5296"""
    R = ref("src/core/MOM_barotropic.F90")
    G, GV, US = _types(dom, grid, gv)
    CS = _bt_small_cs(R, dom, gv)
    fe = adapt.farr(dom, eta_cor)
    CS.eta_cor = fe
    R["mom_barotropic"]["bt_mass_source"](adapt.farr(dom, h), adapt.farr(dom, eta), bool(set_cor), G, GV, CS)
    adapt.back(fe, eta_cor)


def set_dtbt(dom, grid, gv, a):
    """set_dtbt, src/core/MOM_barotropic.F90:3509-3633, with find_face_areas :5146 / BT_cont_to_face_areas :5107; -> (dtbt, dtbt_max)"""
    R = ref("src/core/MOM_barotropic.F90")
    G, GV, US = _types(dom, grid, gv)
    G.z_ref = float(a.get("Z_ref", 0.0))
    CS = _bt_small_cs(R, dom, gv)
    CS.frhatu, CS.frhatv = adapt.farr(dom, a["frhatu"]), adapt.farr(dom, a["frhatv"])
    CS.bathyt = _wide_from_g(dom, a["bathyT"])
    CS.bebt, CS.g_extra, CS.dtbt_fraction = float(a["bebt"]), float(a["G_extra"]), float(a["dtbt_fraction"])
    CS.bt_coriolis_scale, CS.nonlinear_continuity = float(a["BT_Coriolis_scale"]), bool(a["Nonlinear_continuity"])
    CS.dtbt, CS.dtbt_max = 0.0, 0.0
    CS.dy_cu, CS.dx_cv = _wide_u(dom, grid["dy_Cu"]), _wide_v(dom, grid["dx_Cv"])
    BT = bt_cont_type(dom, a["BT_cont"]) if a.get("BT_cont") is not None else None
    kw = {}
    if a.get("pbce") is not None:
        kw["pbce"] = adapt.farr(dom, a["pbce"])
    if a.get("have_gtot_est"):
        kw["gtot_est"] = float(a["gtot_est"])
    if BT is not None:
        kw["bt_cont"] = BT
    if a.get("eta") is not None:
        kw["eta"] = adapt.farr(dom, a["eta"])
    if a.get("SSH_add"):
        kw["ssh_add"] = float(a["SSH_add"])
    R["mom_barotropic"]["set_dtbt"](G, GV, US, CS, **kw)
    return float(CS.dtbt), float(CS.dtbt_max)


def _wide_u(dom, a):
    wi, wj = dom.isd - dom.isdw, dom.jsd - dom.jsdw
    w = np.zeros((dom.jedw - dom.jsdw + 1, dom.iedw - dom.isdw + 2))
    w[wj:wj + a.shape[0], wi:wi + a.shape[1]] = a
    return FArray.from_numpy(w, (int(dom.isdw) - 1, int(dom.jsdw)))


def _wide_v(dom, a):
    wi, wj = dom.isd - dom.isdw, dom.jsd - dom.jsdw
    w = np.zeros((dom.jedw - dom.jsdw + 2, dom.iedw - dom.isdw + 1))
    w[wj:wj + a.shape[0], wi:wi + a.shape[1]] = a
    return FArray.from_numpy(w, (int(dom.isdw), int(dom.jsdw) - 1))


class _Time:
    """time_type (config_src/infra/FMS*/MOM_time_manager.F90, an FMS type) as write_energy uses it: a whole number of seconds with
    +, -, comparisons, integer * time, time / integer and the integer quotient time / time."""
    __slots__ = ("s",)

    def __init__(self, s=0):
        self.s = int(s)

    def __add__(self, o): return _Time(self.s + o.s)
    def __sub__(self, o): return _Time(abs(self.s - o.s))   # FMS time differences are magnitudes
    def __mul__(self, n): return _Time(self.s * int(n))
    __rmul__ = __mul__
    def __truediv__(self, o): return self.s // o.s if isinstance(o, _Time) else _Time(round(self.s / o))
    def __lt__(self, o): return self.s < o.s
    def __le__(self, o): return self.s <= o.s
    def __gt__(self, o): return self.s > o.s
    def __ge__(self, o): return self.s >= o.s
    def __eq__(self, o): return self.s == o.s
    def __hash__(self): return hash(self.s)


class _EnergyFile:
    """stands in for the MOM_netcdf_file the energies are written to: keeps what write_energy hands to write_field, by variable name"""

    def __init__(self):
        self.names, self.rows = [], []

    def write_field(self, field, value, reday=None):
        row = self.rows[-1]
        row[field.name] = value.to_numpy().copy() if isinstance(value, FArray) else float(value)

    def flush(self):
        pass


def _sum_output_stubs():
    def var_desc(name, units=None, longname=None, hor_grid=None, z_grid=None, **kw):
        return NS(name=name)

    def open_file(handle, path, vars, novars, fields, *a, **k):
        for m in range(1, int(novars) + 1):
            fields.s1(m, NS(name=vars.g1(m).name))

    return dict(set_time=lambda seconds=0, days=0, **k: _Time(int(seconds) + 86400 * int(days)), _new_time_type=lambda: _Time(0),
                var_desc=var_desc, create_mom_file=open_file, reopen_mom_file=open_file, open_ascii_file=lambda *a, **k: None,
                call_tracer_stocks=lambda *a, **k: None, array_global_min_max=lambda *a, **k: None, get_time=lambda t, *a, **k: (t.s % 86400, t.s // 86400),
                get_date=lambda *a, **k: None, get_calendar_type=lambda: 0, no_calendar=0, flush_file=lambda *a, **k: None,
                append_file=1, writeonly_file=2, single_file=1, stdout=6, max_across_pes=lambda *a, **k: None,
                efp_sum_across_pes=lambda *a, **k: None, sum_across_pes=lambda *a, **k: None, find_eta=None, is_nan=lambda x: x != x)


SUM_OUTPUT_FILES = ["src/diagnostics/MOM_sum_output.F90", "src/framework/MOM_coms.F90"]


def _sum_output_ref():
    key = ("sum_output",)
    if key not in _REF:
        _REF[key] = load(SUM_OUTPUT_FILES, extra_stubs=_sum_output_stubs(), expose=("write_energy",))
    return _REF[key]


def create_depth_list(dom, grid, Z_ref=0.0, min_depth_inc=1.0e-10):
    """create_depth_list, src/diagnostics/MOM_sum_output.F90:1203-1299 -> (depth, area, vol_below)"""
    R = _sum_output_ref()
    G = _types(dom, grid, {})[0]
    G.z_ref = float(Z_ref)
    G.isg, G.jsg = G.isc, G.jsc
    G.domain.niglobal, G.domain.njglobal = G.iec - G.isc + 1, G.jec - G.jsc + 1
    DL = R["mom_sum_output"]["_new_depth_list"]()
    R["mom_sum_output"]["create_depth_list"](G, DL, float(min_depth_inc))
    return tuple(np.array(x.tolist()) for x in (DL.depth, DL.area, DL.vol_below))


def write_energy(dom, grid, gv, cs, u, v, h, T=None, S=None):
    """write_energy, src/diagnostics/MOM_sum_output.F90:321-1030, on the single-PE case.  cs is the dict of
    mom6_b200.synthetic.sum_output_cs and is updated the way oracle.pyoracle.write_energy updates it (previous_calls, ntrunc, lH and
    the six EFP members); the reference's own Sum_output_CS lives on in cs["_f90run"] from one call to the next.  Returns the
    oracle's result dict: what the reference hands to write_field (:970-994) and, for the numbers it only prints on the ocean.stats
    line (En_mass, KE_tot, PE_tot, salin, temp and their anomalies), its local variables at the final RETURN."""
    R = _sum_output_ref()
    M = R["mom_sum_output"]
    gvd = dict(gv)
    gvd["g_prime"] = np.asarray(cs["g_prime"], dtype=np.float64)
    G, GV, US = _types(dom, grid, gvd, {k: cs.get(k, 1.0) for k in ("RZL2_to_kg", "L_T_to_m_s", "Q_to_J_kg", "J_kg_to_Q", "kg_m3_to_R",
                                                                  "m_to_Z", "m_to_L", "Z_to_m", "S_to_ppt", "C_to_degC")})
    G.z_ref = float(cs.get("Z_ref", 0.0))
    state = cs.setdefault("_f90run", {})
    CS = state.get("CS")
    if CS is None:
        CS = M["_new_sum_output_cs"]()
        CS.initialized = True
        CS.do_ape_calc, CS.use_temperature = bool(cs["do_APE_calc"]), bool(cs["use_temperature"])
        CS.dt_in_t = float(cs["dt_in_T"])
        CS.dl = M["_new_depth_list"]()
        CS.dl.listsize = int(cs["DL_listsize"])
        for n in ("depth", "area", "vol_below"):
            setattr(CS.dl, n, FArray.from_numpy(np.asarray(cs["DL_" + n], dtype=np.float64), (1,)))
        CS.lh = FArray.alloc("i", [(1, int(dom.nk))])
        CS.lh.v[:] = [int(x) for x in cs["lH"]]
        CS.energysavedays, CS.energysavedays_geometric, CS.energysave_geometric = _Time(3600), _Time(0), False
        CS.start_time, CS.write_energy_time, CS.geometric_end_time = _Time(0), _Time(0), _Time(0)
        CS.timeunit, CS.date_stamped_output, CS.iso_date_stamped_output = 86400.0, False, False
        CS.max_energy, CS.maxtrunc = 1.0e30, 1 << 30
        CS.write_stocks, CS.write_min_max, CS.write_min_max_loc = False, False, False
        CS.previous_calls = int(cs.get("previous_calls", 0))
        CS.fileenergy_nc, CS.fileenergy_ascii, CS.energyfile = _EnergyFile(), 17, "ocean.stats"
        CS.fields = FArray.alloc("o", [(1, 17 + 50)])   # NUM_FIELDS + MAX_FIELDS_
        state["CS"], state["n"] = CS, 0
    CS.ntrunc = int(cs.get("ntrunc", 0))
    f = CS.fileenergy_nc
    f.rows.append({})
    tv = NS(c_p=float(cs.get("C_p", 3991.86795711963)))
    if T is not None:
        tv.t, tv.s = adapt.farr(dom, T), adapt.farr(dom, S)
    n = state["n"]
    rt.UNITS[CS.fileenergy_ascii] = []
    M["write_energy"](adapt.farr(dom, u), adapt.farr(dom, v), adapt.farr(dom, h), tv, _Time(3600 * n), n, G, GV, US, CS)
    state["n"] = n + 1
    state["stats_line"] = rt.UNITS.pop(CS.fileenergy_ascii)[-1]   # the record appended to ocean.stats (:880-905); n, then day n/24
    row, loc = f.rows[-1], M["_SAVE"]["write_energy.__locals__"]
    names = {"En": "toten", "APE": "PE", "KE": "KE", "H0": "Z_0APE", "Mass_lay": "mass_lay", "Mass": "mass_tot", "Mass_chg": "mass_chg",
             "Mass_anom": "mass_anom", "Salt": "Salt", "Salt_chg": "Salt_chg", "Salt_anom": "Salt_anom", "Heat": "Heat",
             "Heat_chg": "Heat_chg", "Heat_anom": "Heat_anom"}
    out = {names[k]: (x if isinstance(x, np.ndarray) else float(x)) for k, x in row.items() if k in names}
    out["max_CFL"] = np.array([row["max_CFL_trans"], row["max_CFL_lin"]])
    out["ntrunc"] = int(row["Ntrunc"])
    for k in ("En_mass", "KE_tot", "PE_tot"):
        out[k] = float(loc[k.lower()])
    for k in ("Salt", "Salt_chg", "Salt_anom", "Heat", "Heat_chg", "Heat_anom", "salin", "salin_anom", "temp", "temp_anom"):
        if k not in out:   # not written without temperature; the oracle reports them as zero
            out[k] = float(loc[k.lower()]) if bool(cs["use_temperature"]) or k in ("Salt", "Heat") else 0.0
    cs["previous_calls"], cs["ntrunc"] = int(CS.previous_calls), int(CS.ntrunc)
    cs["lH"][...] = np.array(CS.lh.tolist(), dtype=np.int32)
    for k in ("fresh_water_in_EFP", "net_salt_in_EFP", "net_heat_in_EFP", "mass_prev_EFP", "salt_prev_EFP", "heat_prev_EFP"):
        cs[k] = np.array(getattr(CS, k.lower()).v.tolist(), dtype=np.int64)
    return out


CHKSUM_FILES = ["src/framework/MOM_checksums.F90", "src/framework/MOM_coms.F90"]


def chksum(dom, array, stagger=0, haloshift=0, symmetric=False, omit_corners=False, scale=1.0, stats=False):
    """hchksum / uchksum / vchksum / Bchksum, src/framework/MOM_checksums.F90 (chksum_h_2d :387, chksum_B_2d :688, chksum_u_2d :1005,
    chksum_v_2d :1209 and the _3d forms :1413-2188) -> (the bit counts in the order the reference hands them to chk_sum_msg,
    [mean, min, max] or None).  The messages themselves (formatted writes) are not reproduced: chk_sum_msg is replaced by a recorder."""
    key = ("chksum",)
    if key not in _REF:
        no = lambda *a, **k: None  # noqa: E731
        _REF[key] = load(CHKSUM_FILES, extra_stubs=dict(sum_across_pes=no, min_across_pes=no, max_across_pes=no, error_unit=0,
                                                        is_root_pe=lambda: True))
    m = _REF[key]["mom_checksums"]
    seen = {"bc": [], "stats": None}

    def record(fmsg, *a):
        vals = a[:-2]   # ..., mesg, iounit
        if len(vals) == 3 and all(type(x) is float for x in vals):
            seen["stats"] = list(vals)
        else:
            seen["bc"] += [int(x) for x in vals]

    for n in ("chk_sum_msg", "chk_sum_msg_nsew", "chk_sum_msg_s", "chk_sum_msg_w"):
        m[n] = record
    m["calculatestatistics"], m["writechksums"], m["checkfornans"], m["writehash"] = bool(stats), True, False, False
    G = adapt.grid_type(dom, {})
    HI = G.hi
    HI.turns = 0
    st = "huvq"[stagger]
    fa = adapt.farr(dom, array, st)
    name = "chksum_" + {"h": "h", "u": "u", "v": "v", "q": "b"}[st] + ("_2d" if array.ndim == 2 else "_3d")
    kw = dict(haloshift=int(haloshift), omit_corners=bool(omit_corners))
    if st != "h":
        kw["symmetric"] = bool(symmetric)
    if scale != 1.0:
        kw["unscale"] = float(scale)
    m[name](fa, "x", HI, **kw)
    return seen["bc"], seen["stats"]
