"""dict -> ctypes struct marshalling shared by the product binding (api.py) and the oracle binding."""
import ctypes as C

from ._lib import (Grid, VGrid, ContinuityCS, ContinuityArgs, BTCont, UnitScale, CoriolisAdvCS, CorAdCalcArgs,
                   HorViscCS, HorViscArgs, BarotropicCS, BtstepArgs, BtcalcArgs, PressureForceCS,
                   PressureForceArgs, RemappingCS, fill_struct)


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


def pressureforce_cs(d, keep):
    return fill_struct(PressureForceCS(), d, keep)


def pressureforce_args(a, keep):
    return fill_struct(PressureForceArgs(), a, keep)


def remapping_cs(d):
    return _scalars(RemappingCS(), d)
