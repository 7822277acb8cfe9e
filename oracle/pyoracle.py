"""TEST INFRASTRUCTURE ONLY (see oracle/oracle.h)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """Compile the C++ restatement with the committed recipe (oracle/Makefile)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


def btstep_timeloop(dom, args, nthreads=1):
    """oracle_btstep_timeloop: MOM_barotropic.F90:2175-2832 on host arrays, in place."""
    from mom6_b200._lib import BtTimeloopArgs, fill_struct
    lib = load()
    keep = []
    st = fill_struct(BtTimeloopArgs(), args, keep)
    lib.oracle_btstep_timeloop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = lib.oracle_btstep_timeloop(C.byref(dom), C.byref(st), None, None, nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_btstep_timeloop rc={rc}")
    return rc


def continuity(dom, grid, gv, cs, args, nthreads=1):
    """oracle_continuity: continuity_PPM, MOM_continuity_PPM.F90:86-194, on host arrays."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep)
    v = marshal.vgrid(gv)
    c = marshal.continuity_cs(cs)
    a = marshal.continuity_args(args, keep)
    lib.oracle_continuity.argtypes = [C.c_void_p] * 5 + [C.c_int]
    rc = lib.oracle_continuity(C.byref(dom), C.byref(g), C.byref(v), C.byref(c), C.byref(a), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_continuity rc={rc}")
    return rc


def coradcalc(dom, grid, gv, cs, args, us=None, nthreads=1):
    """oracle_coradcalc: CorAdCalc, MOM_CoriolisAdv.F90:125-965, on host arrays."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep)
    v = marshal.vgrid(gv)
    u = marshal.unit_scale(us)
    c = marshal.coriolisadv_cs(cs)
    a = marshal.coradcalc_args(args, keep)
    lib.oracle_coradcalc.argtypes = [C.c_void_p] * 6 + [C.c_int]
    rc = lib.oracle_coradcalc(C.byref(dom), C.byref(g), C.byref(v), C.byref(u), C.byref(c), C.byref(a), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_coradcalc rc={rc}")
    return rc


def horizontal_viscosity(dom, grid, gv, cs, args, nthreads=1):
    """oracle_horizontal_viscosity: MOM_hor_visc.F90:266-2317 (frozen option set) on host arrays."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep)
    v = marshal.vgrid(gv)
    c = marshal.hor_visc_cs(cs, keep)
    a = marshal.hor_visc_args(args, keep)
    lib.oracle_horizontal_viscosity.argtypes = [C.c_void_p] * 5 + [C.c_int]
    rc = lib.oracle_horizontal_viscosity(C.byref(dom), C.byref(g), C.byref(v), C.byref(c), C.byref(a), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_horizontal_viscosity rc={rc}")
    return rc


def btstep(dom, grid, gv, cs, args, nthreads=1):
    """oracle_btstep: btstep, MOM_barotropic.F90:455-2172 (frozen option set) on host arrays."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep)
    v = marshal.vgrid(gv)
    c = marshal.barotropic_cs(cs, keep)
    a = marshal.btstep_args(args, keep)
    lib.oracle_btstep.argtypes = [C.c_void_p] * 5 + [C.c_int]
    rc = lib.oracle_btstep(C.byref(dom), C.byref(g), C.byref(v), C.byref(c), C.byref(a), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_btstep rc={rc}")
    return rc


def btcalc(dom, grid, gv, args, nthreads=1):
    """oracle_btcalc: btcalc, MOM_barotropic.F90:4360-4605."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep)
    v = marshal.vgrid(gv)
    a = marshal.btcalc_args(args, keep)
    lib.oracle_btcalc.argtypes = [C.c_void_p] * 4 + [C.c_int]
    rc = lib.oracle_btcalc(C.byref(dom), C.byref(g), C.byref(v), C.byref(a), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_btcalc rc={rc}")
    return rc


def bt_mass_source(dom, grid, gv, h, eta, set_cor, eta_cor):
    """oracle_bt_mass_source: MOM_barotropic.F90:5243-5296."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep)
    v = marshal.vgrid(gv)
    lib.oracle_bt_mass_source.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
    return lib.oracle_bt_mass_source(C.byref(dom), C.byref(g), C.byref(v), h.ctypes.data, eta.ctypes.data, int(set_cor),
                                     eta_cor.ctypes.data)


def pressure_force(dom, grid, gv, cs, args, nthreads=1):
    """oracle_pressure_force: PressureForce_FV_Bouss, MOM_PressureForce_FV.F90:947-2017 (frozen option set)."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep)
    v = marshal.vgrid(gv)
    c = marshal.pressureforce_cs(cs, keep)
    a = marshal.pressureforce_args(args, keep)
    lib.oracle_pressure_force.argtypes = [C.c_void_p] * 5 + [C.c_int]
    rc = lib.oracle_pressure_force(C.byref(dom), C.byref(g), C.byref(v), C.byref(c), C.byref(a), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_pressure_force rc={rc}")
    return rc


# ---- ALE remapping (oracle/remap.cpp)
def _dp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def remapping_core_h(cs, h0, u0, h1):
    """oracle_remapping_core_h: remapping_core_h (MOM_remapping.F90:234) for one column; returns (u1, net_err)."""
    import numpy as np
    from mom6_b200 import marshal
    lib = load()
    h0, u0, h1 = (np.ascontiguousarray(x, dtype=np.float64) for x in (h0, u0, h1))
    u1 = np.zeros(len(h1)); err = C.c_double(0.0)
    lib.oracle_remapping_core_h.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    c = marshal.remapping_cs(cs)
    rc = lib.oracle_remapping_core_h(C.byref(c), len(h0), _dp(h0), _dp(u0), len(h1), _dp(h1), _dp(u1), C.byref(err))
    if rc != 0:
        raise RuntimeError(f"oracle_remapping_core_h rc={rc}")
    return u1, err.value


def remap_intersect(h0, h1):
    """intersect_src_tgt_grids (MOM_remapping.F90:642): returns a dict of its 8 outputs."""
    import numpy as np
    lib = load()
    h0, h1 = (np.ascontiguousarray(x, dtype=np.float64) for x in (h0, h1))
    n0, n1 = len(h0), len(h1)
    o = dict(h_sub=np.zeros(n0 + n1 + 1), h0_eff=np.zeros(n0), isrc_start=np.zeros(n0, np.int32), isrc_end=np.zeros(n0, np.int32),
             isrc_max=np.zeros(n0, np.int32), itgt_start=np.zeros(n1, np.int32), itgt_end=np.zeros(n1, np.int32),
             isub_src=np.zeros(n0 + n1 + 1, np.int32))
    lib.oracle_remap_intersect.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_void_p] * 8
    lib.oracle_remap_intersect(n0, _dp(h0), n1, _dp(h1), *[_dp(o[k]) for k in ("h_sub", "h0_eff", "isrc_start", "isrc_end", "isrc_max",
                                                                               "itgt_start", "itgt_end", "isub_src")])
    return o


RECON = dict(PCM=0, PLM=1, PLM_extrap=2, edge_h4=3, PPM=4, edge_ih4=5)


def remap_reconstruct(which, h, u, h_neglect=1.0e-30, E=None):
    """One of the reconstruction routines (see oracle_remap_reconstruct); returns (E[2,N], coefs[3,N])."""
    import numpy as np
    lib = load()
    h, u = (np.ascontiguousarray(x, dtype=np.float64) for x in (h, u))
    N = len(h)
    Ea = np.zeros((2, N)) if E is None else np.ascontiguousarray(np.array(E, dtype=np.float64).reshape(2, N))
    co = np.zeros((3, N))
    lib.oracle_remap_reconstruct.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    rc = lib.oracle_remap_reconstruct(RECON[which], N, _dp(h), _dp(u), h_neglect, _dp(Ea), _dp(co))
    assert rc == 0
    return Ea, co


def remap_plm_sub(om4, h0, u0, h1, h_neglect=1.0e-30):
    """PLM + boundary extrapolation, remap_src_to_sub_grid[_om4], remap_sub_to_tgt_grid_om4; returns (u_sub, u1)."""
    import numpy as np
    lib = load()
    h0, u0, h1 = (np.ascontiguousarray(x, dtype=np.float64) for x in (h0, u0, h1))
    us = np.zeros(len(h0) + len(h1) + 1); u1 = np.zeros(len(h1))
    lib.oracle_remap_plm_sub.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    lib.oracle_remap_plm_sub(int(om4), len(h0), _dp(h0), _dp(u0), len(h1), _dp(h1), h_neglect, _dp(us), _dp(u1))
    return us, u1


def ale_remap_scalar(dom, grid, cs, h_old, h_new, field, conc_underflow=0.0, nthreads=1):
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); c = marshal.remapping_cs(cs)
    lib.oracle_ale_remap_scalar.argtypes = [C.c_void_p] * 6 + [C.c_double, C.c_int]
    return lib.oracle_ale_remap_scalar(C.byref(dom), C.byref(g), C.byref(c), _dp(h_old), _dp(h_new), _dp(field), conc_underflow, nthreads)


def ale_remap_set_h_vel(dom, grid, h_new, h_u, h_v):
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep)
    lib.oracle_ale_remap_set_h_vel.argtypes = [C.c_void_p] * 5
    return lib.oracle_ale_remap_set_h_vel(C.byref(dom), C.byref(g), _dp(h_new), _dp(h_u), _dp(h_v))


def ale_remap_velocities(dom, grid, cs, h_old_u, h_old_v, h_new_u, h_new_v, u, v, nthreads=1):
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); c = marshal.remapping_cs(cs)
    lib.oracle_ale_remap_velocities.argtypes = [C.c_void_p] * 9 + [C.c_int]
    return lib.oracle_ale_remap_velocities(C.byref(dom), C.byref(g), C.byref(c), _dp(h_old_u), _dp(h_old_v), _dp(h_new_u), _dp(h_new_v),
                                           _dp(u), _dp(v), nthreads)


def advect_tracer(dom, grid, gv, cs, a):
    """oracle_advect_tracer: advect_tracer (MOM_tracer_advect.F90:53); returns the number of passes made."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); v = marshal.vgrid(gv); c = marshal.tracer_advect_cs(cs); st = marshal.advect_tracer_args(a, keep)
    it = C.c_int(0)
    lib.oracle_advect_tracer.argtypes = [C.c_void_p] * 6
    rc = lib.oracle_advect_tracer(C.byref(dom), C.byref(g), C.byref(v), C.byref(c), C.byref(st), C.byref(it))
    if rc:
        raise RuntimeError(f"oracle_advect_tracer rc={rc}")
    return it.value


US_ONE = dict(m_to_L=1., L_to_m=1., m_s_to_L_T=1., L_T_to_m_s=1., s_to_T=1., T_to_s=1., m_to_Z=1., Z_to_m=1., Z_to_L=1., L_to_Z=1.)


def ale_regrid(dom, grid, gv, cs, h, h_new, dzRegrid, us=None):
    """oracle_ale_regrid: ALE_regrid (MOM_ALE.F90:518), Z* coordinate; returns 0 or 10 + the code of the first FATAL."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); v = marshal.vgrid(gv); u = marshal.unit_scale(us or US_ONE); c = marshal.regridding_cs(cs, keep)
    lib.oracle_ale_regrid.argtypes = [C.c_void_p] * 8
    return lib.oracle_ale_regrid(C.byref(dom), C.byref(g), C.byref(v), C.byref(u), C.byref(c), _dp(h), _dp(h_new), _dp(dzRegrid))


def vertvisc_coef(dom, grid, gv, cs, a, a_u, a_v, h_u, h_v, us=None):
    """oracle_vertvisc_coef: vertvisc_coef (MOM_vert_friction.F90:1357); the CS%a_u.. arrays are explicit."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); v = marshal.vgrid(gv); u = marshal.unit_scale(us or US_ONE); c = marshal.vertvisc_cs(cs)
    st = marshal.vertvisc_coef_args(a, keep)
    lib.oracle_vertvisc_coef.argtypes = [C.c_void_p] * 10
    rc = lib.oracle_vertvisc_coef(C.byref(dom), C.byref(g), C.byref(v), C.byref(u), C.byref(c), C.byref(st), _dp(a_u), _dp(a_v), _dp(h_u), _dp(h_v))
    if rc:
        raise RuntimeError(f"oracle_vertvisc_coef rc={rc}")


def vertvisc(dom, grid, gv, cs, a, a_u, a_v, h_u, h_v):
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); v = marshal.vgrid(gv); c = marshal.vertvisc_cs(cs); st = marshal.vertvisc_args(a, keep)
    lib.oracle_vertvisc.argtypes = [C.c_void_p] * 9
    lib.oracle_vertvisc_ntrunc.restype = C.c_longlong
    lib.oracle_vertvisc_ntrunc.argtypes = [C.c_int]
    lib.oracle_vertvisc_ntrunc(1)
    rc = lib.oracle_vertvisc(C.byref(dom), C.byref(g), C.byref(v), C.byref(c), C.byref(st), _dp(a_u), _dp(a_v), _dp(h_u), _dp(h_v))
    if rc:
        raise RuntimeError(f"oracle_vertvisc rc={rc}")
    return int(lib.oracle_vertvisc_ntrunc(1))   # CS%ntrunc of this call (vertvisc_limit_vel)


def vertvisc_remnant(dom, grid, cs, visc_rem_u, visc_rem_v, dt, a_u, a_v, h_u, h_v, Ray_u=None, Ray_v=None):
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); c = marshal.vertvisc_cs(cs)
    lib.oracle_vertvisc_remnant.argtypes = [C.c_void_p] * 7 + [C.c_double] + [C.c_void_p] * 4
    rc = lib.oracle_vertvisc_remnant(C.byref(dom), C.byref(g), C.byref(c), None if Ray_u is None else _dp(Ray_u),
                                     None if Ray_v is None else _dp(Ray_v), _dp(visc_rem_u), _dp(visc_rem_v), float(dt), _dp(a_u), _dp(a_v),
                                     _dp(h_u), _dp(h_v))
    if rc:
        raise RuntimeError(f"oracle_vertvisc_remnant rc={rc}")


def step_dyn_split_rk2(dom, grid, gv, css, cs, args, us=None, nthreads=1):
    """oracle_step_dyn_split_rk2: step_MOM_dyn_split_RK2 (MOM_dynamics_split_RK2.F90:294).  css: dict of the stage control
    structures (continuity, coriolisadv, hor_visc, pressureforce, vertvisc); cs["CAu_pred_stored"] is updated."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); v = marshal.vgrid(gv); u = marshal.unit_scale(us or US_ONE)
    cc = marshal.continuity_cs(css["continuity"]); ca = marshal.coriolisadv_cs(css["coriolisadv"]); hv = marshal.hor_visc_cs(css["hor_visc"], keep)
    pg = marshal.pressureforce_cs(css["pressureforce"], keep); vv = marshal.vertvisc_cs(css["vertvisc"])
    st = marshal.dyn_split_rk2_cs(cs, keep); a = marshal.step_dyn_args(args, keep)
    lib.oracle_step_dyn_split_rk2.argtypes = [C.c_void_p] * 11 + [C.c_int]
    rc = lib.oracle_step_dyn_split_rk2(C.byref(dom), C.byref(g), C.byref(v), C.byref(u), C.byref(cc), C.byref(ca), C.byref(hv), C.byref(pg),
                                       C.byref(vv), C.byref(st), C.byref(a), nthreads)
    if rc:
        raise RuntimeError(f"oracle_step_dyn_split_rk2 rc={rc}")
    cs["CAu_pred_stored"] = int(st.CAu_pred_stored)
    cs["dtbt_max"] = float(st.dtbt_max)
    cs["barotropic"]["dtbt"] = float(st.barotropic.contents.dtbt)


def set_dtbt(dom, grid, gv, args, us=None):
    """oracle_set_dtbt: set_dtbt (MOM_barotropic.F90:3509); returns (dtbt, dtbt_max)."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); v = marshal.vgrid(gv); u = marshal.unit_scale(us or US_ONE); a = marshal.set_dtbt_args(args, keep)
    dtbt, dmax = C.c_double(0.0), C.c_double(0.0)
    lib.oracle_set_dtbt.argtypes = [C.c_void_p] * 7
    rc = lib.oracle_set_dtbt(C.byref(dom), C.byref(g), C.byref(v), C.byref(u), C.byref(a), C.byref(dtbt), C.byref(dmax))
    if rc:
        raise RuntimeError(f"oracle_set_dtbt rc={rc}")
    return dtbt.value, dmax.value


def remap_dyn_split_rk2_aux_vars(dom, grid, remap_cs, cs, h_old_u, h_old_v, h_new_u, h_new_v, nthreads=1):
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); r = marshal.remapping_cs(remap_cs); st = marshal.dyn_split_rk2_cs(cs, keep)
    lib.oracle_remap_dyn_split_rk2_aux_vars.argtypes = [C.c_void_p] * 8 + [C.c_int]
    return lib.oracle_remap_dyn_split_rk2_aux_vars(C.byref(dom), C.byref(g), C.byref(r), C.byref(st), _dp(h_old_u), _dp(h_old_v), _dp(h_new_u),
                                                   _dp(h_new_v), nthreads)


# ---- reproducing sums, checksums, write_energy (efp.cpp, sum_output.cpp)
def reproducing_sum(dom, array, stagger=0, isr=0, ier=0, jsr=0, jer=0, unscale=1.0, reproducing=True, overflow_check=True,
                    want_sums=False, want_efp=False, want_lay_efp=False):
    """oracle_reproducing_sum: reproducing_sum_2d / _3d (MOM_coms.F90:227, :337).  array is (nk, nj, ni) or (nj, ni).
    Returns a dict: sum, and sums / EFP_sum / EFP_lay_sums (int64 arrays) when requested."""
    import numpy as np
    from mom6_b200 import marshal
    from mom6_b200._lib import Efp
    lib = load()
    nk = 1 if array.ndim == 2 else array.shape[0]
    total = C.c_double(0.0)
    sums = np.zeros(nk) if want_sums else None
    e = Efp() if want_efp else None
    le = (Efp * nk)() if want_lay_efp else None
    lib.oracle_reproducing_sum.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                                                    C.c_void_p, C.c_void_p]
    rc = lib.oracle_reproducing_sum(C.byref(dom), _dp(array), stagger, nk, isr, ier, jsr, jer, float(unscale), int(reproducing),
                                    int(overflow_check), C.byref(total), _dp(sums), C.byref(e) if e is not None else None,
                                    C.cast(le, C.c_void_p) if le is not None else None)
    if rc != 0:
        raise RuntimeError(f"oracle_reproducing_sum: FATAL {rc}")
    r = {"sum": total.value}
    if want_sums:
        r["sums"] = sums
    if want_efp:
        r["EFP_sum"] = marshal.efp_back(e)
    if want_lay_efp:
        r["EFP_lay_sums"] = np.array([marshal.efp_back(le[k]) for k in range(nk)])
    return r


def efp_op(op, a, b=None):
    """EFP_plus / EFP_minus / EFP_to_real / real_to_EFP / EFP_real_diff on int64[6] arrays."""
    from mom6_b200 import marshal
    from mom6_b200._lib import Efp
    lib = load()
    lib.oracle_efp_to_real.restype = C.c_double
    lib.oracle_efp_real_diff.restype = C.c_double
    lib.oracle_real_to_efp.argtypes = [C.c_double, C.c_void_p]
    if op == "from_real":
        out = Efp()
        rc = lib.oracle_real_to_efp(float(a), C.byref(out))
        if rc:
            raise OverflowError("Overflow in real_to_EFP conversion")
        return marshal.efp_back(out)
    ea = marshal.efp(a)
    if op == "to_real":
        return float(lib.oracle_efp_to_real(C.byref(ea)))
    eb = marshal.efp(b)
    if op == "diff":
        return float(lib.oracle_efp_real_diff(C.byref(ea), C.byref(eb)))
    out, ov = Efp(), C.c_int(0)
    getattr(lib, "oracle_efp_" + op)(C.byref(ea), C.byref(eb), C.byref(out), C.byref(ov))
    return marshal.efp_back(out)


def chksum(dom, array, stagger=0, haloshift=0, symmetric=False, omit_corners=False, scale=1.0, stats=False):
    """oracle_chksum: hchksum / uchksum / vchksum / Bchksum (MOM_checksums.F90).  Returns (bc[5], kind, stats or None)."""
    import numpy as np
    lib = load()
    nk = 1 if array.ndim == 2 else array.shape[0]
    bc = (C.c_int * 5)()
    kind = C.c_int(0)
    st = (C.c_double * 3)()
    lib.oracle_chksum.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib.oracle_chksum(C.byref(dom), _dp(array), stagger, nk, haloshift, int(symmetric), int(omit_corners), float(scale), bc,
                           C.byref(kind), st if stats else None)
    if rc != 0:
        raise RuntimeError(f"oracle_chksum: FATAL {rc}")
    return np.array(list(bc), dtype=np.int64), kind.value, (np.array(list(st)) if stats else None)


def create_depth_list(dom, grid, Z_ref=0.0, min_depth_inc=1.0e-10):
    """oracle_create_depth_list: create_depth_list (MOM_sum_output.F90:1203); returns (depth, area, vol_below)."""
    import numpy as np
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep)
    mls = (dom.iec - dom.isc + 1) * (dom.jec - dom.jsc + 1)
    depth, area, vol = np.zeros(mls + 2), np.zeros(mls + 2), np.zeros(mls + 2)
    n = C.c_int(0)
    lib.oracle_create_depth_list.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib.oracle_create_depth_list(C.byref(dom), C.byref(g), float(Z_ref), float(min_depth_inc), C.byref(n), _dp(depth), _dp(area), _dp(vol))
    assert rc == 0
    return depth[:n.value].copy(), area[:n.value].copy(), vol[:n.value].copy()


def write_energy(dom, grid, gv, cs, u, v, h, T=None, S=None):
    """oracle_write_energy: write_energy (MOM_sum_output.F90:321); cs (dict) is updated like the reference's CS."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); vg = marshal.vgrid(gv)
    st = marshal.sum_output_cs(cs, keep)
    out, arrs = marshal.energy_out(dom.nk, keep)
    lib.oracle_write_energy.argtypes = [C.c_void_p] * 10
    rc = lib.oracle_write_energy(C.byref(dom), C.byref(g), C.byref(vg), C.byref(st), _dp(u), _dp(v), _dp(h), _dp(T), _dp(S), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle_write_energy: FATAL {rc}")
    marshal.sum_output_cs_back(st, cs)
    return marshal.energy_out_back(out, arrs)


def ocean_stats_line(cs, e, n, reday):
    """oracle_ocean_stats_line: the line write_energy appends to ocean.stats (MOM_sum_output.F90:874-902)."""
    import numpy as np
    from mom6_b200 import marshal
    from mom6_b200._lib import EnergyOut, _EO_SCALARS, _EO_SCALARS2
    lib = load()
    keep = []
    st = marshal.sum_output_cs(cs, keep)
    out = EnergyOut()
    for k in _EO_SCALARS + _EO_SCALARS2:
        setattr(out, k, float(e[k]))
    out.max_CFL[0], out.max_CFL[1] = float(e["max_CFL"][0]), float(e["max_CFL"][1])
    out.ntrunc = int(e["ntrunc"])
    z = np.ascontiguousarray(e["Z_0APE"], dtype=np.float64)
    out.Z_0APE = z.ctypes.data
    buf = C.create_string_buffer(512)
    lib.oracle_ocean_stats_line.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_char_p, C.c_size_t]
    rc = lib.oracle_ocean_stats_line(C.byref(st), C.byref(out), int(n), float(reday), buf, 512)
    assert rc == 0
    return buf.value.decode()


# ---- the thermodynamic-cadence ALE pass (ale_chain.cpp)
def interpolate_column(h_src, u_src, h_dest, mask_edges=False):
    """oracle_interpolate_column: interpolate_column (MOM_remapping.F90:1247) for one column."""
    import numpy as np
    lib = load()
    h_src, u_src, h_dest = (np.ascontiguousarray(x, dtype=np.float64) for x in (h_src, u_src, h_dest))
    u_dest = np.zeros(len(h_dest) + 1)
    lib.oracle_interpolate_column.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.oracle_interpolate_column.restype = None
    lib.oracle_interpolate_column(len(h_src), _dp(h_src), _dp(u_src), len(h_dest), _dp(h_dest), _dp(u_dest), int(mask_edges))
    return u_dest


def ale_remap_vals(dom, grid, h_old, h_new, val, vertex=False):
    """oracle_ale_remap_interface_vals / _vertex_vals (MOM_ALE.F90:1303 / :1342), in place."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep)
    fn = lib.oracle_ale_remap_vertex_vals if vertex else lib.oracle_ale_remap_interface_vals
    fn.argtypes = [C.c_void_p] * 5
    return fn(C.byref(dom), C.byref(g), _dp(h_old), _dp(h_new), _dp(val))


def ale_regridding_and_remapping(dom, grid, gv, cs, args, dyn_cs=None, us=None, nthreads=1):
    """oracle_ale_regridding_and_remapping: ALE_regridding_and_remapping (MOM.F90:1751-1926) on one tile."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); v = marshal.vgrid(gv); u = marshal.unit_scale(us or US_ONE)
    st = marshal.ale_cs(cs, keep)
    dyn = marshal.dyn_split_rk2_cs(dyn_cs, keep) if dyn_cs is not None else None
    lib.oracle_ale_regridding_and_remapping.argtypes = [C.c_void_p] * 7 + [C.c_int]
    rc = lib.oracle_ale_regridding_and_remapping(C.byref(dom), C.byref(g), C.byref(v), C.byref(u), C.byref(st),
                                                 C.byref(dyn) if dyn is not None else None, C.byref(marshal.ale_args(args, keep)), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_ale_regridding_and_remapping: FATAL {rc}")
    cs["regridCS"]["old_grid_weight"] = float(st.regridCS.old_grid_weight)
    return rc


# ---- mixedlayer_restrat (mle.cpp)
def mle_mu(sigma, dh):
    """oracle_mle_mu: mu(sigma, dh), MOM_mixed_layer_restrat.F90:717-751."""
    lib = load()
    lib.oracle_mle_mu.argtypes = [C.c_double, C.c_double]
    lib.oracle_mle_mu.restype = C.c_double
    return lib.oracle_mle_mu(float(sigma), float(dh))


def mixedlayer_restrat(dom, grid, gv, cs, h, uhtr, vhtr, T, S, ustar, dt, h_MLD, Rd_dx_h=None):
    """oracle_mixedlayer_restrat: mixedlayer_restrat_OM4 (MOM_mixed_layer_restrat.F90:189-714), in place."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); v = marshal.vgrid(gv)
    st = marshal.mle_cs(cs, keep)
    lib.oracle_mixedlayer_restrat.argtypes = [C.c_void_p] * 10 + [C.c_double] + [C.c_void_p] * 2
    rc = lib.oracle_mixedlayer_restrat(C.byref(dom), C.byref(g), C.byref(v), C.byref(st), _dp(h), _dp(uhtr), _dp(vhtr), _dp(T), _dp(S),
                                       _dp(ustar), float(dt), _dp(h_MLD), _dp(Rd_dx_h))
    if rc != 0:
        raise RuntimeError(f"oracle_mixedlayer_restrat: FATAL {rc}")
    return rc


# ---- tracer_hordiff (hordiff.cpp)
def tracer_hordiff(dom, grid, gv, cs, a):
    """oracle_tracer_hordiff: tracer_hordiff (MOM_tracer_hor_diff.F90:119), the along-surface path; returns the iterations made."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); v = marshal.vgrid(gv); c = marshal.tracer_hor_diff_cs(cs); st = marshal.tracer_hordiff_args(a, keep)
    it = C.c_int(0)
    lib.oracle_tracer_hordiff.argtypes = [C.c_void_p] * 6
    rc = lib.oracle_tracer_hordiff(C.byref(dom), C.byref(g), C.byref(v), C.byref(c), C.byref(st), C.byref(it))
    if rc:
        raise RuntimeError(f"oracle_tracer_hordiff: FATAL {rc}")
    return it.value


# ---- thickness_diffuse (thickdiff.cpp)
def thickness_diffuse(dom, grid, gv, cs, a, us=None):
    """oracle_thickness_diffuse: thickness_diffuse (MOM_thickness_diffuse.F90:134) -> thickness_diffuse_full :635, in place."""
    from mom6_b200 import marshal
    lib = load()
    keep = []
    g = marshal.grid(grid, keep); v = marshal.vgrid(gv); u = marshal.unit_scale(us or US_ONE)
    c = marshal.thickness_diffuse_cs(cs); st = marshal.thickness_diffuse_args(a, keep)
    lib.oracle_thickness_diffuse.argtypes = [C.c_void_p] * 6
    rc = lib.oracle_thickness_diffuse(C.byref(dom), C.byref(g), C.byref(v), C.byref(u), C.byref(c), C.byref(st))
    if rc:
        raise RuntimeError(f"oracle_thickness_diffuse: FATAL {rc}")
    return rc


def eos_eval(which, form, T, S, p, rho_ref=0.0, lin4=None, scales=None, impl="pgf"):
    """Equation-of-state elements: which = "rho" | "anom" | "drho_dT" | "drho_dS"; impl = "pgf" (oracle/eos.hpp), "mle", "thickdiff"."""
    lib = load()
    l4 = (C.c_double * 4)(*lin4) if lin4 is not None else None
    if impl == "mle":
        lib.oracle_mle_eos_density.restype = C.c_double
        lib.oracle_mle_eos_density.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double]
        return lib.oracle_mle_eos_density(form, l4, T, S, p)
    if impl == "thickdiff":
        a, b = C.c_double(), C.c_double()
        lib.oracle_thickdiff_eos_derivs.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        lib.oracle_thickdiff_eos_derivs(form, l4, T, S, p, C.byref(a), C.byref(b))
        return a.value if which == "drho_dT" else b.value
    sc = (C.c_double * 4)(*scales) if scales is not None else None
    lib.oracle_eos_eval.restype = C.c_double
    lib.oracle_eos_eval.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
    return lib.oracle_eos_eval(dict(rho=0, anom=1, drho_dT=2, drho_dS=3)[which], form, l4, sc, T, S, p, rho_ref)


def ale_edge_values(scheme, h, Q, bdry_extrap=False, h_neglect=1.0e-30):
    """oracle_ale_edge_values: the top / bottom edge values the pressure force reconstructs T and S with (scheme 1 = PLM, 2 = PPM)."""
    import numpy as np
    lib = load()
    h = np.ascontiguousarray(h, dtype=np.float64); Q = np.ascontiguousarray(Q, dtype=np.float64)
    qt, qb = np.zeros_like(Q), np.zeros_like(Q)
    lib.oracle_ale_edge_values.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    rc = lib.oracle_ale_edge_values(scheme, len(h), h.ctypes.data, Q.ctypes.data, int(bdry_extrap), h_neglect, qt.ctypes.data, qb.ctypes.data)
    if rc:
        raise RuntimeError(f"oracle_ale_edge_values rc={rc}")
    return qt, qb
