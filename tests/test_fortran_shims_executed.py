"""The Fortran bindings, EXECUTED.

fortran/install_shims.py makes shadow copies of the reference's modules with the hooks and the binding procedures of
fortran/bodies/*.inc inserted.  There is no Fortran compiler here, so these tests run the shadow copies through oracle/f90run: the
hooked reference procedure is called with the reference's own derived types, takes the device branch (mom6cu_enabled()), the
binding fills the bind(C) structures of fortran/mom6cu_interface.F90 and calls the mom6cu_* entry points -- which land in
oracle/f90run/cabi.py, a stand-in that rebuilds the C structures field by field and calls the C++ oracle (no GPU here).  The
results must equal, bit for bit, those of the UNHOOKED reference procedure on the same inputs.  So a member that the binding
forgets, mis-converts or puts in the wrong field, or a control structure that never reaches the library, fails here.

(The whole-step case found one: with step_MOM_dyn_split_RK2 on the device the stage entries are never reached, so their control
structures were never sent; every shadow module now exports a *_send_cs_mom6cu routine that the step binding calls.)"""
import os

import numpy as np
import pytest

from oracle import f90run

pytestmark = pytest.mark.skipif(not f90run.available(), reason="the reference tree is not present")

import refcases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


_LOADED = {}   # the translated shadow modules are kept across cases; only the C-ABI stand-in and the SAVEd flags are renewed


@pytest.fixture()
def shadow(oracle):
    """installs the shims in a temporary directory and points oracle/f90run/stages.py at the shadow copies"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "fortran"))
    import install_shims
    from oracle.f90run import cabi, stages
    if "installed" not in _LOADED:
        # the shadow copies contain the reference's text: they live in a temporary directory outside the repository, removed at exit
        import atexit, shutil, tempfile
        _LOADED["dir"] = tempfile.mkdtemp(prefix="mom6cu_shims_")
        atexit.register(shutil.rmtree, _LOADED["dir"], ignore_errors=True)
        install_shims.install(f90run.REFERENCE_ROOT, _LOADED["dir"])
        _LOADED["installed"] = True
        _LOADED["current"] = {}
        names = cabi.Abi(oracle, None, None, None).stubs()
        # every stub forwards to the stand-in of the current case
        _LOADED["stubs"] = {k: ((lambda *a, _k=k: _LOADED["current"]["stubs"][_k](*a)) if callable(v) else v) for k, v in names.items()}
        _LOADED["plain"] = dict(stages._REF)
        _LOADED["shadow"] = {}
    out = _LOADED["dir"]
    plain = dict(stages._REF)
    _LOADED["plain"].update(plain)

    def activate(dom, grid, gv):
        abi = cabi.Abi(oracle, dom, grid, gv)
        _LOADED["current"]["stubs"] = abi.stubs()
        _LOADED["plain"].update(stages._REF)
        stages._REF.clear()
        stages._REF.update(_LOADED["shadow"])
        stages.SHADOW.update(files={rel: os.path.join(out, os.path.basename(rel)) for rel in install_shims.SHIMS},
                             extra_files=[os.path.join(out, "mom6cu_interface.F90")], stubs=_LOADED["stubs"], tag="shadow",
                             post_load=lambda R: [ns.__setitem__("mom6cu_ctx", "ctx") for ns in R.values() if "mom6cu_ctx" in ns])
        for R in stages._REF.values():      # `logical, save :: cs_sent`: a new run of the model
            for ns in R.values():
                ns["_SAVE"].clear()
        return abi
    yield activate
    _LOADED["shadow"].update({k: v for k, v in stages._REF.items() if k[-1] == "shadow"})
    stages.SHADOW.update(files={}, extra_files=[], stubs={}, post_load=None, tag=None)
    stages._REF.clear()
    stages._REF.update(_LOADED["plain"])


def _same(a, b):
    bad = [k for k in a if not np.array_equal((a[k] + 0.0).view(np.int64), (b[k] + 0.0).view(np.int64))]
    assert sorted(a) == sorted(b) and not bad, bad


@pytest.mark.parametrize("name", ["step/default", "step/two_steps_store_CAu_set_dtbt", "step/plm_pressure_reconstruction",
                                  "step/cfl_truncation_in_vertvisc"])
def test_hooked_step_equals_the_reference_step(shadow, name):
    inputs = refcases.build(name)
    want = refcases.run_reference(name, inputs)          # the reference's own step_MOM_dyn_split_RK2
    abi = shadow(*inputs[:3])
    got = refcases.run_reference(name, inputs)           # the shadow copy: hook -> binding -> C ABI stand-in -> oracle
    n = refcases.CASES[name]["kw"].get("nsteps", 1)
    assert abi.calls.count("step_dyn_split_rk2") == n, abi.calls
    for k in ("continuity", "coriolisadv", "hor_visc", "pressureforce", "vertvisc"):
        assert abi.calls.count("set_cs_" + k) == 1, (k, abi.calls)       # every stage control structure arrives, once
    _same(want, got)


@pytest.mark.parametrize("name", ["continuity/default", "continuity/monotonic_volCFL_y_first", "continuity/upwind_closed_no_uhbt",
                                  "coradcalc/scheme1_ke10", "coradcalc/scheme6_ke12", "coradcalc/no_slip_bound_closed",
                                  "horizontal_viscosity/options00", "horizontal_viscosity/options02", "horizontal_viscosity/options09",
                                  "pressure_force/options00", "pressure_force/options01", "pressure_force/options03",
                                  "pressure_force/options05", "pressure_force/options08", "btstep/default",
                                  "btstep/strong_drag_bound_corr", "btstep/project_velocity_filter_y_first",
                                  "advect_tracer/ppm_h3_cfl2.5_three_tracers", "advect_tracer/mixed_schemes_underflow_max_iter",
                                  "advect_tracer/ppm_h3_periodic_y_y_first", "mixedlayer_restrat/options00",
                                  "mixedlayer_restrat/options01", "thickness_diffuse/options01", "thickness_diffuse/options02",
                                  "thickness_diffuse/options05", "tracer_hordiff/options01", "tracer_hordiff/options02",
                                  "tracer_hordiff/options03", "ale/ppm_h4_aux_vars", "ale/plm_no_store_CAu",
                                  "ale/ppm_ih4_no_time_filter", "ale/pcm_no_aux_vars", "vertvisc_family/options00",
                                  "vertvisc_family/options01", "vertvisc_family/options09", "vertvisc_family/truncation_cfl_0.2_ray"])
def test_hooked_stage_equals_the_reference_stage(shadow, name):
    inputs = refcases.build(name)
    want = refcases.run_reference(name, inputs)
    abi = shadow(*inputs[:3])
    got = refcases.run_reference(name, inputs)
    entry = {"ale": "ale_regridding_and_remapping", "vertvisc_family": "vertvisc"}.get(refcases.CASES[name]["stage"],
                                                                                       refcases.CASES[name]["stage"])
    assert abi.calls.count(entry) == 1, abi.calls
    _same(want, got)
