"""The hot-path stages run from the REFERENCE'S OWN SOURCE (translated by oracle/f90run), with the calling convention of
oracle/pyoracle.py: (dom, grid, gv, cs, args) dictionaries in, results written into args / cs in place.  TEST INFRASTRUCTURE ONLY.

tests/test_reference_f90.py calls a stage here and the same stage of the C++ oracle on copies of the same seeded inputs and
compares every output bit for bit over the computational domain.  Each function cites the reference entry it executes."""
import numpy as np

from . import adapt, halo, load, new
from .rt import FArray, NS

_REF = {}


def ref(*paths):
    key = tuple(paths)
    if key not in _REF:
        _REF[key] = load(list(paths), extra_stubs=halo.STUBS)
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
