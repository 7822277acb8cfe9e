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


def fill_struct(struct, values, keep):
    """Fill a ctypes struct from a dict: numpy arrays / torch tensors -> pointers, scalars as is."""
    for name, ctype in struct._fields_:
        v = values.get(name)
        if ctype is C.c_void_p:
            if v is None:
                setattr(struct, name, None)
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
    lib.mom6cu_last_kernel_ms.restype = C.c_double
    lib.mom6cu_total_kernel_ms.argtypes = [vp]
    lib.mom6cu_total_kernel_ms.restype = C.c_double
    lib.mom6cu_comm_unique_id.argtypes = [C.c_char_p, C.c_int]
    lib.mom6cu_comm_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.mom6cu_comm_destroy.argtypes = [vp]
    lib.mom6cu_halo_plan.argtypes = [C.POINTER(Domain), C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.mom6cu_btstep_timeloop.argtypes = [vp, C.POINTER(BtTimeloopArgs)]
    lib.mom6cu_btstep_timeloop_resident.argtypes = [vp, C.POINTER(BtTimeloopArgs), C.c_int, C.c_int]
    return lib


def load():
    """Load libmom6cu.so (built in-tree by __graft_entry__.build()); fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        _lib = bind(C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL))
    return _lib
