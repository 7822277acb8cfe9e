"""dict -> ctypes struct marshalling shared by the product binding (api.py) and the oracle binding."""
import ctypes as C

import numpy as np

from ._lib import (Grid, VGrid, ContinuityCS, ContinuityArgs, BTCont, UnitScale, CoriolisAdvCS, CorAdCalcArgs,
                   HorViscCS, HorViscArgs, BarotropicCS, BtstepArgs, BtcalcArgs, PressureForceCS,
                   PressureForceArgs, RemappingCS, TracerAdvectCS, AdvectTracerArgs, RegriddingCS, VertviscCS, VertviscCoefArgs, VertviscArgs, DynSplitRK2CS, StepDynArgs, SetDtbtArgs, AleCS, AleArgs, MleCS, TracerHorDiffCS, TracerHordiffArgs, ThicknessDiffuseCS, ThicknessDiffuseArgs, Efp, SumOutputCS, EnergyOut, _SO_UNITS, _SO_EFPS, _EO_SCALARS, _EO_SCALARS2, fill_struct)


def _scalars(struct, d):
    for name, _ in struct._fields_:
        setattr(struct, name, d[name])
    return struct


def grid(g, keep):
    return fill_struct(Grid(), g, keep)


def vgrid(d):
    return _scalars(VGrid(), d)


def continuity_cs(d):
    return _scalars(ContinuityCS(), d)


def continuity_args(a, keep):
    st = fill_struct(ContinuityArgs(), a, keep)
    b = a.get("BT_cont")
    if b is not None:
        bs = fill_struct(BTCont(), b, keep)
        keep.append(bs)
        st.BT_cont = C.pointer(bs)
    return st


def unit_scale(d=None):
    st = UnitScale()
    for name, _ in UnitScale._fields_:
        setattr(st, name, (d or {}).get(name, 1.0))
    return st


def coriolisadv_cs(d):
    return _scalars(CoriolisAdvCS(), d)


def coradcalc_args(a, keep):
    return fill_struct(CorAdCalcArgs(), a, keep)


def hor_visc_cs(d, keep):
    return fill_struct(HorViscCS(), d, keep)


def hor_visc_args(a, keep):
    return fill_struct(HorViscArgs(), a, keep)


def barotropic_cs(d, keep):
    return fill_struct(BarotropicCS(), d, keep)


def btstep_args(a, keep):
    st = fill_struct(BtstepArgs(), a, keep)
    bs = fill_struct(BTCont(), a["BT_cont"], keep)
    keep.append(bs)
    st.BT_cont = C.pointer(bs)
    return st


def btcalc_args(a, keep):
    return fill_struct(BtcalcArgs(), a, keep)


PGF_RECON_DEFAULTS = dict(reconstruct=0, Recon_Scheme=1, boundary_extrap=0, use_inaccurate_pgf_rho_anom=0, MassWghtInterpVanOnly=0,
                          ALE_answer_date=99991231, h_nonvanished=0.0, kg_m3_to_R=1.0, RL2_T2_to_Pa=1.0, C_to_degC=1.0, S_to_ppt=1.0)


def pressureforce_cs(d, keep):
    """RECONSTRUCT_FOR_PRESSURE members default to off / unscaled when the dict does not carry them."""
    dd = dict(PGF_RECON_DEFAULTS)
    dd.update(d)
    return fill_struct(PressureForceCS(), dd, keep)


def pressureforce_args(a, keep):
    return fill_struct(PressureForceArgs(), a, keep)


def remapping_cs(d):
    return _scalars(RemappingCS(), d)


def _addr(x):
    if x is None:
        return None
    if hasattr(x, "ptr") and isinstance(getattr(x, "ptr"), int):   # api.Plane
        return x.ptr
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"] or x.dtype != np.float64:
            raise ValueError("array arguments must be C-contiguous float64")
        return x.ctypes.data
    return int(x)


def tracer_advect_cs(d):
    return _scalars(TracerAdvectCS(), d)


def advect_tracer_args(a, keep):
    """a: dict(h_end, uhtr, vhtr, dt, tr=[...], advect_scheme=None, conc_underflow=None, x_first_in=None, max_iter_in=None,
    vol_prev=None, update_vol_prev=False, uhr_out=None, vhr_out=None)."""
    s = AdvectTracerArgs()
    s.h_end, s.uhtr, s.vhtr = _addr(a["h_end"]), _addr(a["uhtr"]), _addr(a["vhtr"])
    s.dt = float(a["dt"])
    tr = a["tr"]
    s.ntr = len(tr)
    ptrs = (C.c_void_p * max(len(tr), 1))(*[_addr(t) for t in tr])
    keep.append(ptrs)
    s.tr = C.cast(ptrs, C.POINTER(C.c_void_p))
    if a.get("advect_scheme") is not None:
        sch = (C.c_int * max(len(tr), 1))(*[int(x) for x in a["advect_scheme"]])
        keep.append(sch)
        s.advect_scheme = C.cast(sch, C.POINTER(C.c_int))
    if a.get("conc_underflow") is not None:
        cu = np.ascontiguousarray(a["conc_underflow"], dtype=np.float64)
        keep.append(cu)
        s.conc_underflow = cu.ctypes.data
    s.x_first_in = -1 if a.get("x_first_in") is None else int(bool(a["x_first_in"]))
    s.max_iter_in = -1 if a.get("max_iter_in") is None else int(a["max_iter_in"])
    s.vol_prev = _addr(a.get("vol_prev"))
    s.update_vol_prev = int(bool(a.get("update_vol_prev", False)))
    s.uhr_out, s.vhr_out = _addr(a.get("uhr_out")), _addr(a.get("vhr_out"))
    return s


def regridding_cs(d, keep):
    return fill_struct(RegriddingCS(), d, keep)


def vertvisc_cs(d):
    return _scalars(VertviscCS(), d)


def vertvisc_coef_args(a, keep):
    return fill_struct(VertviscCoefArgs(), a, keep)


def vertvisc_args(a, keep):
    return fill_struct(VertviscArgs(), a, keep)


def dyn_split_rk2_cs(d, keep):
    """d: the scalar members, the arrays, BT_cont (dict) and barotropic (dict of barotropic_CS members)."""
    st = fill_struct(DynSplitRK2CS(), d, keep)
    bs = fill_struct(BTCont(), d["BT_cont"], keep)
    bt = fill_struct(BarotropicCS(), d["barotropic"], keep)
    keep += [bs, bt]
    st.BT_cont = C.pointer(bs)
    st.barotropic = C.pointer(bt)
    return st


def step_dyn_args(a, keep):
    return fill_struct(StepDynArgs(), a, keep)


def set_dtbt_args(a, keep):
    st = fill_struct(SetDtbtArgs(), a, keep)
    if a.get("BT_cont") is not None:
        bs = fill_struct(BTCont(), a["BT_cont"], keep)
        keep.append(bs)
        st.BT_cont = C.pointer(bs)
    return st


def efp(v=None):
    """numpy int64[6] (or None = zero) -> mom6cu_efp"""
    e = Efp()
    if v is not None:
        for n in range(6):
            e.v[n] = int(v[n])
    return e


def efp_back(e):
    return np.array([int(e.v[n]) for n in range(6)], dtype=np.int64)


def sum_output_cs(d, keep):
    """Sum_output_CS as write_energy uses it.  Arrays: DL_depth / DL_area / DL_vol_below / g_prime float64, lH int32 (updated in
    place); the six EFP members are int64[6] arrays (absent = zero)."""
    st = SumOutputCS()
    st.do_APE_calc, st.use_temperature = int(d["do_APE_calc"]), int(d["use_temperature"])
    st.dt_in_T = float(d["dt_in_T"])
    st.Z_ref, st.C_p = float(d.get("Z_ref", 0.0)), float(d.get("C_p", 3991.86795711963))
    for n in _SO_UNITS:
        setattr(st, n, float(d.get(n, 1.0)))
    st.previous_calls, st.ntrunc = int(d.get("previous_calls", 0)), int(d.get("ntrunc", 0))
    st.DL_listsize = int(d.get("DL_listsize", 0))
    for n in ("DL_depth", "DL_area", "DL_vol_below", "g_prime"):
        a = d.get(n)
        if a is not None:
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
            keep.append(a)
            setattr(st, n, a.ctypes.data)
    if d.get("lH") is not None:
        assert d["lH"].dtype == np.int32
        keep.append(d["lH"])
        st.lH = d["lH"].ctypes.data
    for n in _SO_EFPS:
        setattr(st, n, efp(d.get(n)))
    return st


def sum_output_cs_back(st, d):
    d["previous_calls"], d["ntrunc"] = int(st.previous_calls), int(st.ntrunc)
    for n in _SO_EFPS:
        d[n] = efp_back(getattr(st, n))


def energy_out(nk, keep):
    st = EnergyOut()
    arrs = {"KE": np.zeros(nk), "mass_lay": np.zeros(nk), "PE": np.zeros(nk + 1), "Z_0APE": np.zeros(nk + 1)}
    for n, a in arrs.items():
        keep.append(a)
        setattr(st, n, a.ctypes.data)
    return st, arrs


def energy_out_back(st, arrs):
    r = {n: float(getattr(st, n)) for n in _EO_SCALARS + _EO_SCALARS2}
    r["max_CFL"] = np.array([st.max_CFL[0], st.max_CFL[1]])
    r["ntrunc"] = int(st.ntrunc)
    r.update(arrs)
    return r


def ale_cs(d, keep):
    """d: regridCS / remapCS / vel_remapCS (dicts), regrid_time_scale, remap_aux_vars."""
    st = AleCS()
    st.regridCS = fill_struct(RegriddingCS(), d["regridCS"], keep)
    st.remapCS = _scalars(RemappingCS(), d["remapCS"])
    st.vel_remapCS = _scalars(RemappingCS(), d["vel_remapCS"])
    st.regrid_time_scale = float(d.get("regrid_time_scale", 0.0))
    st.remap_uv_using_old_alg = int(d.get("remap_uv_using_old_alg", 0))
    st.do_conv_adj, st.use_hybgen_unmix = int(d.get("do_conv_adj", 0)), int(d.get("use_hybgen_unmix", 0))
    st.remap_aux_vars = int(d.get("remap_aux_vars", 0))
    return st


def ale_args(a, keep):
    st = AleArgs()
    st.u, st.v, st.h = _addr(a["u"]), _addr(a["v"]), _addr(a["h"])
    tr = a.get("tr") or []
    st.ntr = len(tr)
    arr = (C.c_void_p * max(len(tr), 1))(*[_addr(t) for t in tr])
    keep.append(arr)
    st.tr = arr
    cu = a.get("conc_underflow")
    if cu is not None:
        cu = np.ascontiguousarray(cu, dtype=np.float64)
        keep.append(cu)
        st.conc_underflow = cu.ctypes.data
    st.iT, st.iS = int(a.get("iT", -1)), int(a.get("iS", -1))
    st.dtdia = float(a["dtdia"])
    st.Kd_shear, st.Kv_shear, st.Kv_shear_Bu = _addr(a.get("Kd_shear")), _addr(a.get("Kv_shear")), _addr(a.get("Kv_shear_Bu"))
    return st


def mle_cs(d, keep):
    return fill_struct(MleCS(), d, keep)


def tracer_hor_diff_cs(d):
    return _scalars(TracerHorDiffCS(), d)


def tracer_hordiff_args(a, keep):
    """a: dict(h, dt, tr=[...], conc_underflow=None, Res_fn_h=None, Rd_dx_h=None, df_x=None, df_y=None); df_x / df_y are lists with
    None for the tracers whose diagnostic is not associated."""
    s = TracerHordiffArgs()
    s.h, s.dt = _addr(a["h"]), float(a["dt"])
    tr = a["tr"]
    s.ntr = len(tr)
    ptrs = (C.c_void_p * max(len(tr), 1))(*[_addr(t) for t in tr])
    keep.append(ptrs)
    s.tr = C.cast(ptrs, C.POINTER(C.c_void_p))
    if a.get("conc_underflow") is not None:
        cu = np.ascontiguousarray(a["conc_underflow"], dtype=np.float64)
        keep.append(cu)
        s.conc_underflow = cu.ctypes.data
    s.Res_fn_h, s.Rd_dx_h = _addr(a.get("Res_fn_h")), _addr(a.get("Rd_dx_h"))
    for key in ("L2u", "SN_u", "L2v", "SN_v", "MEKE_Kh"):
        setattr(s, key, _addr(a.get(key)))
    for key in ("df_x", "df_y"):
        if a.get(key) is not None:
            p = (C.c_void_p * max(len(tr), 1))(*[_addr(t) for t in a[key]])
            keep.append(p)
            setattr(s, key, C.cast(p, C.POINTER(C.c_void_p)))
    return s


def thickness_diffuse_cs(d):
    return _scalars(ThicknessDiffuseCS(), d)


def thickness_diffuse_args(a, keep):
    d = {k: a.get(k) for k, _ in ThicknessDiffuseArgs._fields_}
    return fill_struct(ThicknessDiffuseArgs(), d, keep)
