"""dict -> ctypes struct marshalling shared by the product binding (api.py) and the oracle binding."""
import ctypes as C

from ._lib import Grid, VGrid, ContinuityCS, ContinuityArgs, BTCont, fill_struct


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
