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
