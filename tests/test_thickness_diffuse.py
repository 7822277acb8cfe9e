"""thickness_diffuse -> thickness_diffuse_full (src/parameterizations/lateral/MOM_thickness_diffuse.F90:134-1670), SURVEY 8f row 2 (the
isopycnal-height diffusion of step_MOM_dynamics, MOM.F90:1388).  The reference holds no vector for this routine (parity unpinned): CPU
tests check what the algorithm guarantees on the oracle restatement (no net transport through a face, column thickness kept, thickness
floor, flat isopycnals at rest, slopes flattened), that the host build of the code the GPU threads run (csrc/thickdiff_column.cuh)
equals the oracle bit for bit, and tests/test_rotation.py / test_rescaling.py hold its invariances.  GPU: C ABI == oracle, bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mom6_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    return x


def _inner(dom, x):
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]


def test_oracle_properties(oracle):
    dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(30, 22, 12, land_blocks=3, with_GM=True)
    ref = _copy(a)
    assert oracle.thickness_diffuse(dom, grid, gv, cs, ref) == 0
    dt = a["dt"]
    uhD, vhD = (ref["uhtr"] - a["uhtr"]) / dt, (ref["vhtr"] - a["vhtr"]) / dt
    sel = ref["uhGM"] != 7.0
    assert np.allclose(uhD[sel], ref["uhGM"][sel], rtol=1e-9, atol=1e-6 * np.abs(ref["uhGM"][sel]).max()) and sel.any()
    gm_u, gm_v = np.where(ref["uhGM"] != 7.0, ref["uhGM"], 0.0), np.where(ref["vhGM"] != 7.0, ref["vhGM"], 0.0)
    scale = np.abs(gm_u).sum(axis=0).max()
    assert scale > 0 and np.abs(gm_u.sum(axis=0)).max() < 1e-12 * scale and np.abs(gm_v.sum(axis=0)).max() < 1e-12 * scale   # Sfn(z=0) = 0
    assert (gm_u[:, grid["mask2dCu"] == 0] == 0).all() and (gm_v[:, grid["mask2dCv"] == 0] == 0).all()                       # nothing through land
    hi, ho = _inner(dom, a["h"]), _inner(dom, ref["h"])
    m = _inner(dom, grid["mask2dT"]) > 0
    assert np.allclose(ho.sum(axis=0)[m], hi.sum(axis=0)[m], rtol=1e-13) and np.abs(ho - hi).max() > 1e-3
    assert ho.min() >= gv["Angstrom_H"]
    # h = h - dt*IareaT*div(uhD, vhD) with the diagnosed transports, in the reference's operation order (:611-615)
    j0, i0 = dom.jsc - dom.jsd, dom.isc - dom.isd
    nj, ni = dom.jec - dom.jsc + 1, dom.iec - dom.isc + 1
    div = ((gm_u[:, j0:j0 + nj, i0 + 1:i0 + 1 + ni] - gm_u[:, j0:j0 + nj, i0:i0 + ni]) + (gm_v[:, j0 + 1:j0 + 1 + nj, i0:i0 + ni] - gm_v[:, j0:j0 + nj, i0:i0 + ni]))
    want = np.maximum(hi - dt * _inner(dom, grid["IareaT"])[None] * div, gv["Angstrom_H"])
    assert np.array_equal(want, ho)


def test_oracle_flat_stratification_is_at_rest_and_slopes_are_flattened(oracle):
    # level interfaces and horizontally uniform T, S: no slope, no transport
    dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(20, 16, 8, front=0.0)
    grid["bathyT"][...] = 4000.0
    a["h"][...] = 500.0
    a["T"][...] = (20.0 - 2.0 * np.arange(8))[:, None, None]
    a["S"][...] = 35.0
    ref = _copy(a)
    oracle.thickness_diffuse(dom, grid, gv, cs, ref)
    assert np.array_equal(ref["h"], a["h"]) and np.array_equal(ref["uhtr"], a["uhtr"])
    # a tilted interface between two uniform water masses relaxes: the transport above it is directed towards the thin side of the upper layer
    a["h"][3] += 40.0 * np.sin(2 * np.pi * np.arange(a["h"].shape[-1]) / 20.0)[None, :]
    a["h"][4] -= 40.0 * np.sin(2 * np.pi * np.arange(a["h"].shape[-1]) / 20.0)[None, :]
    ref = _copy(a)
    oracle.thickness_diffuse(dom, grid, gv, cs, ref)
    var0 = _inner(dom, a["h"][3]).var(); var1 = _inner(dom, ref["h"][3]).var()
    assert 0 < var1 < var0


def test_oracle_rejects_options_outside_the_frozen_set(oracle):
    dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(12, 10, 4)
    for bad in (dict(use_Visbeck=1), dict(interface_Kh=1), dict(khth_struct=1), dict(detangle_interfaces=1), dict(EOS_form=0), dict(find_work=1)):
        with pytest.raises(RuntimeError):
            oracle.thickness_diffuse(dom, grid, gv, dict(cs, **bad), _copy(a))
    with pytest.raises(RuntimeError):                                              # "cg1 must be associated when using FGNV streamfunction." (:858)
        oracle.thickness_diffuse(dom, grid, gv, dict(cs, use_FGNV_streamfn=1), dict(_copy(a), cg1=None))
    h0 = a["h"].copy()
    assert oracle.thickness_diffuse(dom, grid, gv, dict(cs, Khth=0.0), a) == 0 and np.array_equal(h0, a["h"])      # nothing to do (:196)


@pytest.fixture(scope="module")
def td_host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("td") / "libtd_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "harness", "thickdiff_host.cpp")])
    lib = C.CDLL(so)
    lib.td_host_run.restype = None
    lib.td_host_run.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong] + [C.c_void_p] * 26
    return lib


def _unified(dom, x, st):
    nj, ni = dom.jed - dom.jsd + 2, dom.ied - dom.isd + 2
    out = np.zeros(x.shape[:-2] + (nj, ni))
    out[..., (0 if st in "vq" else 1):, (0 if st in "uq" else 1):] = x
    return out


def _from_unified(x, st):
    return np.ascontiguousarray(x[..., (0 if st in "vq" else 1):, (0 if st in "uq" else 1):])


def _run_device_code_on_host(lib, dom, grid, gv, cs, a):
    """thickness_diffuse through the host build of csrc/thickdiff_column.cuh, parameters set as mom6cu_thickness_diffuse sets them."""
    p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)   # noqa: E731
    nk, dt = dom.nk, a["dt"]
    res = bool(cs["use_variable_mixing"] and cs["Resoln_scaled_KhTh"])
    kap = (2.0 * (cs["kappa_smooth"] * dt)) * (1.0 * gv["m_to_H"])
    par = np.array([nk, cs["EOS_form"], int(res), int(a["p_surf"] is not None), dt, 0.25 / dt, gv["Angstrom_H"], gv["H_subroundoff"],
                    gv["H_subroundoff"] * gv["H_subroundoff"], cs["dZ_subroundoff"], gv["H_to_Z"], gv["Z_to_H"], gv["g_Earth"] * gv["H_to_RZ"], 1.0,
                    cs["Khth"], cs["Khth_Min"], cs["Khth_Max"], cs["max_Khth_CFL"], 1.0 / (cs["slope_max"] * cs["slope_max"]), kap,
                    1.0e-16 * np.sqrt(0.5 * kap), cs["dRho_dT"], cs["dRho_dS"],
                    int(cs["use_stored_slopes"]), int(cs["use_FGNV_streamfn"]), int(cs["use_MEKE_Kh"]), gv["g_Earth"] / gv["Rho0"],
                    cs["dZ_subroundoff"] * cs["dZ_subroundoff"], cs["N2_floor"], cs["FGNV_scale"], cs["MEKE_KhTh_fac"]], dtype=np.float64)
    box = np.array([dom.isc, dom.iec, dom.jsc, dom.jec, dom.isd - 1, dom.jsd - 1], dtype=np.int32)
    F = {k: (None if a[k] is None else _unified(dom, a[k], st)) for k, st in (("h", "h"), ("uhtr", "u"), ("vhtr", "v"), ("T", "h"), ("S", "h"), ("p_surf", "h"),
                                                                           ("Res_fn_u", "u"), ("Res_fn_v", "v"), ("uhGM", "u"), ("vhGM", "v"),
                                                                           ("slope_x", "u"), ("slope_y", "v"), ("cg1", "h"), ("MEKE_Kh", "h"))}
    Gd = {k: _unified(dom, grid[k], st) for k, st in (("areaT", "h"), ("IareaT", "h"), ("bathyT", "h"), ("IdxCu", "u"), ("IdyCu", "u"), ("dy_Cu", "u"),
                                                      ("IdxCv", "v"), ("IdyCv", "v"), ("dx_Cv", "v"), ("mask2dCu", "u"), ("mask2dCv", "v"))}
    nj, ni = Gd["areaT"].shape
    scratch = np.zeros((8 * (nk + 1) + 6 * nk, nj, ni))
    lib.td_host_run(p(par), p(box), ni, ni * nj, p(F["h"]), p(F["uhtr"]), p(F["vhtr"]), p(F["T"]), p(F["S"]), p(F["p_surf"]), p(F["Res_fn_u"]),
                    p(F["Res_fn_v"]), p(F["uhGM"]), p(F["vhGM"]), p(Gd["areaT"]), p(Gd["IareaT"]), p(Gd["bathyT"]), p(Gd["IdxCu"]), p(Gd["IdyCu"]),
                    p(Gd["dy_Cu"]), p(Gd["IdxCv"]), p(Gd["IdyCv"]), p(Gd["dx_Cv"]), p(scratch), p(Gd["mask2dCu"]), p(Gd["mask2dCv"]), p(F["slope_x"]),
                    p(F["slope_y"]), p(F["cg1"]), p(F["MEKE_Kh"]))
    return {k: _from_unified(F[k], st) for k, st in (("h", "h"), ("uhtr", "u"), ("vhtr", "v"), ("uhGM", "u"), ("vhGM", "v")) if F[k] is not None}


CASES = [dict(), dict(land_blocks=4, EOS_form=1, with_GM=True), dict(land_blocks=2, cyclic_y=True, with_p_surf=True, Khth=3000.0, max_Khth_CFL=0.2),
         dict(use_variable_mixing=1, Resoln_scaled_KhTh=1, Khth_Max=400.0, Khth_Min=50.0, land_blocks=2), dict(kappa_smooth=0.0, slope_max=0.001, front=6.0),
         dict(dt=7200.0, Khth=2000.0, kappa_smooth=1.0e-4, with_GM=True)]
# the OM4-style selection: stored slopes, the FGNV streamfunction, the MEKE diffusivity (m6td::face_ext / td_face_ext_kernel)
EXT_CASES = [dict(use_stored_slopes=1, land_blocks=2), dict(use_FGNV_streamfn=1, land_blocks=3, with_GM=True), dict(use_MEKE_Kh=1, Khth=0.0, use_variable_mixing=1),
         dict(use_FGNV_streamfn=1, use_stored_slopes=1, use_MEKE_Kh=1, Khth=0.0, use_variable_mixing=1, Resoln_scaled_KhTh=1, FGNV_scale=0.1, cyclic_y=True)]


def _assert_same(dom, want, got, kw):
    assert np.array_equal(_inner(dom, want["h"]).view(np.int64), _inner(dom, got["h"]).view(np.int64)), kw
    for k in ("uhtr", "vhtr", "uhGM", "vhGM"):
        if want.get(k) is not None:
            assert np.array_equal(want[k].view(np.int64), got[k].view(np.int64)), (k, kw)


@pytest.mark.parametrize("kw", CASES + EXT_CASES)
def test_device_column_code_equals_oracle_on_the_host(oracle, td_host, kw):
    for (ni, nj, nk) in ((28, 20, 12), (9, 31, 2), (14, 12, 40)):
        dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(ni, nj, nk, **kw)
        ref = _copy(a)
        oracle.thickness_diffuse(dom, grid, gv, cs, ref)
        got = _run_device_code_on_host(td_host, dom, grid, gv, cs, a)
        _assert_same(dom, ref, got, kw)
        assert not np.array_equal(ref["uhtr"], a["uhtr"])


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_thickness_diffuse_bitwise(oracle, ctx_factory, kw):
    for (ni, nj, nk) in ((44, 40, 20), (131, 9, 2), (30, 22, 75)):
        dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(ni, nj, nk, **kw)
        ref = _copy(a)
        oracle.thickness_diffuse(dom, grid, gv, cs, ref)
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv)
        n0 = ctx.launches
        ctx.thickness_diffuse(cs, a)
        assert ctx.launches - n0 >= 4
        _assert_same(dom, ref, a, kw)


@pytest.mark.gpu
def test_thickness_diffuse_errors(ctx_factory):
    from mom6_b200.api import Mom6cuError
    dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(16, 12, 5)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    for bad in (dict(use_Visbeck=1), dict(interface_Kh=1), dict(khth_struct=1), dict(EOS_form=0), dict(find_work=1)):
        with pytest.raises(Mom6cuError):
            ctx.thickness_diffuse(dict(cs, **bad), a)
