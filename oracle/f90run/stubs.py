"""Stand-ins for the framework modules the translated reference routines `use` but that are outside the hot path
(error handler, clocks, diagnostics, checksums, halo updates).  TEST INFRASTRUCTURE ONLY.

Halo updates are supplied by the test (extra_stubs of f90run.load) when a routine needs them; the default raises, so that a
translated routine can never silently skip an exchange."""
from . import rt

FATAL, WARNING, NOTE = 3, 2, 1
WARNINGS = []


def mom_error(level, message="", all_print=None):
    if level == FATAL:
        raise rt.FortranStop("FATAL: " + str(message))
    WARNINGS.append(str(message))


def _noop(*a, **k):
    return None


def _false(*a, **k):
    return False


def _need(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name} was called: the test must supply it through extra_stubs")
    return f


NAMES = {
    "fatal": FATAL, "warning": WARNING, "note": NOTE,
    "mom_error": mom_error, "mom_mesg": _noop, "is_root_pe": lambda: True, "calltree_enter": _noop, "calltree_leave": _noop,
    "calltree_waypoint": _noop, "calltree_showquery": _false, "assert_": _noop,
    "cpu_clock_begin": _noop, "cpu_clock_end": _noop, "cpu_clock_id": lambda *a, **k: 0,
    "clock_module": 0, "clock_routine": 0, "clock_module_driver": 0, "clock_component": 0,
    "post_data": _noop, "query_averaging_enabled": _false, "enable_averages": _noop, "disable_averaging": _noop,
    "register_diag_field": lambda *a, **k: -1, "post_product_u": _noop, "post_product_v": _noop,
    "post_product_sum_u": _noop, "post_product_sum_v": _noop,
    "hchksum": _noop, "uvchksum": _noop, "bchksum": _noop, "qchksum": _noop, "chksum": _noop, "mom_state_chksum": _noop,
    "mom_accel_chksum": _noop, "check_redundant": _noop, "mom_thermo_chksum": _noop, "vec_chksum": _noop,
    "hchksum_pair": _noop, "uvchksum_pair": _noop,
    "pass_var": _need("pass_var"), "pass_vector": _need("pass_vector"), "do_group_pass": _need("do_group_pass"),
    "start_group_pass": _need("start_group_pass"), "complete_group_pass": _need("complete_group_pass"),
    "create_group_pass": _noop, "pass_var_start": _need("pass_var_start"), "pass_var_complete": _need("pass_var_complete"),
    "pass_vector_start": _need("pass_vector_start"), "pass_vector_complete": _need("pass_vector_complete"),
    "to_all": 1, "to_north": 2, "to_south": 4, "to_east": 8, "to_west": 16, "omit_corners": 32,
    "scalar_pair": 64, "agrid": 1, "bgrid_ne": 2, "cgrid_ne": 3, "corner": 4, "center": 0,
    "obc_none": 0, "obc_direction_n": 100, "obc_direction_s": 200, "obc_direction_e": 300, "obc_direction_w": 400,
    "max_across_pes": _noop, "min_across_pes": _noop, "sum_across_pes": _noop,
    "time_type_to_real": lambda t: float(t), "real_to_time": lambda x: x,
    "ns": rt.NS,
    "diag_update_remap_grids": _noop, "safe_alloc_ptr": _noop, "safe_alloc_alloc": _noop, "query_debugging_checks": _noop,
    "diag_save_grids": _noop, "diag_restore_grids": _noop, "diag_copy_diag_to_storage": _noop,
    "time_type": rt.NS, "get_diag_time_end": lambda *a, **k: 0.0,
    "int64": 8, "int32": 4, "real64": 8, "real32": 4,
    "num_pes": lambda: 1, "pe_here": lambda: 0, "root_pe": lambda: 0, "sync_pes": _noop, "broadcast": _noop,
    "any_across_pes": lambda x: bool(x), "all_across_pes": lambda x: bool(x),
    "uppercase": lambda s_: s_.upper(), "lowercase": lambda s_: s_.lower(), "stdout": 6, "stderr": 0,
}
