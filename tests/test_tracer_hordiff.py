"""tracer_hordiff, the along-surface path (src/tracer/MOM_tracer_hor_diff.F90:119-640), SURVEY 8f row 2 (the tracer-step caller,
MOM.F90:1526).  The reference holds no vector for this routine (parity unpinned): CPU tests check what the algorithm guarantees on
the oracle restatement (tracer inventory, maximum principle, uniform fields, the CFL iteration count, the flux diagnostics), that
the host build of the code the GPU threads run (csrc/hordiff_cell.cuh) equals the oracle bit for bit, and tests/test_rotation.py holds
its index-rotation invariance.  GPU: C ABI == oracle, bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mom6_b200 import fidx, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_copy(v) for v in x]
    return x


def _inner(dom, x):
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]


def test_oracle_conserves_the_inventory_and_obeys_the_maximum_principle(oracle):
    dom, grid, gv, cs, a = synthetic.hordiff_inputs(30, 22, 8, land_blocks=3, ntr=2, KhTr=3000.0)
    a["conc_underflow"] = None
    ref = _copy(a)
    assert oracle.tracer_hordiff(dom, grid, gv, cs, ref) == 1
    vol = _inner(dom, a["h"] * grid["areaT"][None])
    for m in range(2):
        t0, t1 = _inner(dom, a["tr"][m]), _inner(dom, ref["tr"][m])
        assert np.abs(t1 - t0).max() > 1e-3
        inv0, inv1 = (vol * t0).sum(axis=(1, 2)), (vol * t1).sum(axis=(1, 2))         # closed / reentrant tile: every layer conserves
        assert np.allclose(inv1, inv0, rtol=1e-12)
        # CFL = 0.27 < 1: the update is a convex combination of a cell and its four neighbours (halos included)
        f = fidx.FA(dom.isd, dom.ied, dom.jsd, dom.jed, nk=dom.nk); f.a[...] = a["tr"][m]; fidx.fill_halo(dom, f, "h")
        j0, i0 = dom.jsc - dom.jsd, dom.isc - dom.isd
        nj, ni = dom.jec - dom.jsc + 1, dom.iec - dom.isc + 1
        nb = np.stack([f.a[:, j0 + dj:j0 + dj + nj, i0 + di:i0 + di + ni] for dj, di in ((0, 0), (1, 0), (-1, 0), (0, 1), (0, -1))])
        assert (t1 <= nb.max(axis=0) + 1e-12).all() and (t1 >= nb.min(axis=0) - 1e-12).all()


def test_oracle_keeps_a_uniform_tracer_uniform_and_counts_iterations(oracle):
    dom, grid, gv, cs, a = synthetic.hordiff_inputs(24, 18, 5, land_blocks=2, ntr=1, KhTr=5.0e4, check_diffusive_CFL=1)
    a["tr"][0][...] = 3.25
    n = oracle.tracer_hordiff(dom, grid, gv, cs, a)
    area_min = _inner(dom, grid["areaT"]).min()
    assert n == int(np.ceil(2.0 * 4 * 5.0e4 * 7200.0 / area_min)) or n >= 4        # CFL = 2*sum(khdt)/area ~ 4.6 on the 25 km mesh
    assert (_inner(dom, a["tr"][0]) == 3.25).all()
    # MAX_TR_DIFFUSION_CFL caps khdt, and alone (no CFL check) sets the iteration count (:367-369)
    dom, grid, gv, cs, a = synthetic.hordiff_inputs(24, 18, 5, ntr=1, KhTr=5.0e4, max_diff_CFL=2.5)
    assert oracle.tracer_hordiff(dom, grid, gv, cs, a) == 3
    # nothing to do: KHTR = 0 without variable mixing (:153)
    dom, grid, gv, cs, a = synthetic.hordiff_inputs(12, 10, 3, ntr=1, KhTr=0.0)
    t0 = a["tr"][0].copy()
    assert oracle.tracer_hordiff(dom, grid, gv, cs, a) == 0 and np.array_equal(t0, a["tr"][0])


def test_oracle_flux_diagnostics_close_the_tendency(oracle):
    dom, grid, gv, cs, a = synthetic.hordiff_inputs(20, 16, 4, land_blocks=1, ntr=3, with_df=True, KhTr=2.0e4, check_diffusive_CFL=1)
    a["conc_underflow"] = None
    ref = _copy(a)
    n = oracle.tracer_hordiff(dom, grid, gv, cs, ref)
    assert n >= 2 and ref["df_x"][1] is None and ref["df_y"][0] is None
    m = 2                                                                            # both diagnostics associated
    fx, fy = ref["df_x"][m], ref["df_y"][m]                                            # [conc H L2 T-1], accumulated over the iterations
    j0, i0 = dom.jsc - dom.jsd, dom.isc - dom.isd
    nj, ni = dom.jec - dom.jsc + 1, dom.iec - dom.isc + 1
    div = ((fx[:, j0:j0 + nj, i0 + 1:i0 + 1 + ni] - fx[:, j0:j0 + nj, i0:i0 + ni]) + (fy[:, j0 + 1:j0 + 1 + nj, i0:i0 + ni] - fy[:, j0:j0 + nj, i0:i0 + ni]))
    dT = _inner(dom, ref["tr"][m] - a["tr"][m])
    want = -a["dt"] * div * _inner(dom, grid["IareaT"])[None] / (_inner(dom, a["h"]) + gv["H_subroundoff"])
    assert np.allclose(dT, want, rtol=1e-9, atol=1e-12 * np.abs(dT).max())
    assert (fx[:, :j0, :] == 7.0).all()                                              # only the computational faces are zeroed (:374-381)


@pytest.fixture(scope="module")
def hd_host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hd") / "libhd_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "harness", "hordiff_host.cpp")])
    lib = C.CDLL(so)
    lib.hd_host_khdt.restype = C.c_double
    lib.hd_host_khdt.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong] + [C.c_void_p] * 15
    lib.hd_host_sweep.restype = None
    lib.hd_host_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_double, C.c_double] + [C.c_void_p] * 8
    return lib


def _unified(dom, x, st):
    nj, ni = dom.jed - dom.jsd + 2, dom.ied - dom.isd + 2
    out = np.zeros(x.shape[:-2] + (nj, ni))
    out[..., (0 if st in "vq" else 1):, (0 if st in "uq" else 1):] = x
    return out


def _from_unified(x, st):
    return np.ascontiguousarray(x[..., (0 if st in "vq" else 1):, (0 if st in "uq" else 1):])


def _run_device_code_on_host(lib, dom, grid, gv, cs, a):
    """tracer_hordiff through the host build of csrc/hordiff_cell.cuh, driven the way mom6cu_tracer_hordiff drives the kernels."""
    p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)   # noqa: E731
    dt, nk = a["dt"], dom.nk
    vm = bool(cs["use_variable_mixing"])
    par = np.array([dt, 1.0 / dt, gv["H_subroundoff"], cs["KhTr"], cs["KhTr_min"], cs["KhTr_max"], cs["KhTr_passivity_coeff"],
                    cs["KhTr_passivity_min"], cs["max_diff_CFL"], int(vm), int(vm and cs["Resoln_scaled_KhTr"]),
                    int(vm and cs["KhTr_Slope_Cff"] > 0.0), int(vm and cs["use_MEKE_Kh"]), cs["KhTr_Slope_Cff"], cs["MEKE_KhTr_fac"]], dtype=np.float64)
    box = np.array([dom.isc, dom.iec, dom.jsc, dom.jec, dom.isd - 1, dom.jsd - 1], dtype=np.int32)
    Gd = {k: _unified(dom, grid[k], st) for k, st in (("dy_Cu", "u"), ("IdxCu", "u"), ("dx_Cv", "v"), ("IdyCv", "v"), ("areaT", "h"), ("IareaT", "h"))}
    res, rd = _unified(dom, a["Res_fn_h"], "h"), _unified(dom, a["Rd_dx_h"], "h")
    h = _unified(dom, a["h"], "h")
    nj, ni = res.shape
    khx, khy = np.zeros((nj, ni)), np.zeros((nj, ni))
    V = {k: _unified(dom, a[k], st) for k, st in (("L2u", "u"), ("SN_u", "u"), ("L2v", "v"), ("SN_v", "v"), ("MEKE_Kh", "h"))}
    max_cfl = lib.hd_host_khdt(p(par), p(box), ni, p(Gd["dy_Cu"]), p(Gd["IdxCu"]), p(Gd["dx_Cv"]), p(Gd["IdyCv"]), p(Gd["areaT"]), p(Gd["IareaT"]),
                               p(res), p(rd), p(khx), p(khy), p(V["L2u"]), p(V["SN_u"]), p(V["L2v"]), p(V["SN_v"]), p(V["MEKE_Kh"]))
    eps = np.finfo(np.float64).eps
    if cs["check_diffusive_CFL"]:
        n = max(1, int(np.ceil(max_cfl - 4.0 * eps)))
    elif cs["max_diff_CFL"] > 0.0:
        n = max(1, int(np.ceil(cs["max_diff_CFL"] - 4.0 * eps)))
    else:
        n = 1
    scale = 1.0 / float(n)
    out = _copy(a)
    for key, st in (("df_x", "u"), ("df_y", "v")):
        for m, f in enumerate(out.get(key) or []):
            if f is not None:                                              # hd_zero_faces_kernel
                j0, i0 = dom.jsc - dom.jsd, dom.isc - dom.isd
                nj_, ni_ = dom.jec - dom.jsc + 1, dom.iec - dom.isc + 1
                if st == "u":
                    f[:, j0:j0 + nj_, i0:i0 + ni_ + 1] = 0.0
                else:
                    f[:, j0:j0 + nj_ + 1, i0:i0 + ni_] = 0.0
    for itt in range(n):
        for m in range(len(out["tr"])):
            f = fidx.FA(dom.isd, dom.ied, dom.jsd, dom.jed, nk=nk); f.a[...] = out["tr"][m]; fidx.fill_halo(dom, f, "h")   # m6_halo_update
            T = _unified(dom, f.a, "h")
            Tn = T.copy()
            dfx = _unified(dom, out["df_x"][m], "u") if out.get("df_x") and out["df_x"][m] is not None else None
            dfy = _unified(dom, out["df_y"][m], "v") if out.get("df_y") and out["df_y"][m] is not None else None
            uf = 0.0 if a.get("conc_underflow") is None else float(a["conc_underflow"][m])
            lib.hd_host_sweep(p(par), p(box), ni, ni * nj, nk, scale, uf, p(h), p(T), p(khx), p(khy), p(Gd["IareaT"]), p(Tn), p(dfx), p(dfy))
            out["tr"][m] = _from_unified(Tn, "h")
            if dfx is not None:
                out["df_x"][m] = _from_unified(dfx, "u")
            if dfy is not None:
                out["df_y"][m] = _from_unified(dfy, "v")
    return n, out


CASES = [dict(), dict(land_blocks=3, KhTr=5.0e4, check_diffusive_CFL=1, with_df=True),
         dict(land_blocks=2, cyclic_y=True, KhTr=8.0e4, max_diff_CFL=2.5, check_diffusive_CFL=1),
         dict(use_variable_mixing=1, Resoln_scaled_KhTr=1, KhTr_max=1500.0, KhTr_min=100.0, KhTr_passivity_coeff=2.0, land_blocks=2),
         dict(use_variable_mixing=1, KhTr=0.0, KhTr_min=800.0, max_diff_CFL=0.1, with_df=True), dict(KhTr=2.0e4, max_diff_CFL=1.5, ntr=9)]


# the VarMix / MEKE diffusivity terms (:208-210): KHTR_SLOPE_CFF * L2u * SN_u and MEKE%KhTr_fac * sqrt(Kh Kh)
EXT_CASES = [dict(use_variable_mixing=1, KhTr_Slope_Cff=0.1, KhTr=10.0, KhTr_max=900.0, land_blocks=2),
             dict(use_variable_mixing=1, use_MEKE_Kh=1, KhTr=0.0, MEKE_KhTr_fac=0.5, check_diffusive_CFL=1, with_df=True),
             dict(use_variable_mixing=1, KhTr_Slope_Cff=0.05, use_MEKE_Kh=1, Resoln_scaled_KhTr=1, KhTr_passivity_coeff=2.0, KhTr_min=50.0, cyclic_y=True)]


def _assert_same(dom, want, got, kw):
    for m in range(len(want["tr"])):
        assert np.array_equal(_inner(dom, want["tr"][m]).view(np.int64), _inner(dom, got["tr"][m]).view(np.int64)), (m, kw)
    for key in ("df_x", "df_y"):
        for m, f in enumerate(want.get(key) or []):
            if f is not None:
                assert np.array_equal(f.view(np.int64), got[key][m].view(np.int64)), (key, m, kw)


@pytest.mark.parametrize("kw", CASES + EXT_CASES)
def test_device_cell_code_equals_oracle_on_the_host(oracle, hd_host, kw):
    for (ni, nj, nk) in ((28, 20, 6), (9, 31, 3)):
        dom, grid, gv, cs, a = synthetic.hordiff_inputs(ni, nj, nk, **kw)
        ref = _copy(a)
        n_ref = oracle.tracer_hordiff(dom, grid, gv, cs, ref)
        n, got = _run_device_code_on_host(hd_host, dom, grid, gv, cs, a)
        assert n == n_ref, kw
        _assert_same(dom, ref, got, kw)
        assert not np.array_equal(ref["tr"][0], a["tr"][0])


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_tracer_hordiff_bitwise(oracle, ctx_factory, kw):
    for (ni, nj, nk) in ((44, 40, 20), (131, 9, 3), (30, 22, 75)):
        dom, grid, gv, cs, a = synthetic.hordiff_inputs(ni, nj, nk, **kw)
        ref = _copy(a)
        n_ref = oracle.tracer_hordiff(dom, grid, gv, cs, ref)
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv)
        n0 = ctx.launches
        assert ctx.tracer_hordiff(cs, a) == n_ref
        assert ctx.launches - n0 >= 2 + 2 * len(a["tr"])
        _assert_same(dom, ref, a, kw)


@pytest.mark.gpu
def test_tracer_hordiff_on_resident_planes_and_errors(oracle, ctx_factory):
    from mom6_b200.api import Mom6cuError
    dom, grid, gv, cs, a = synthetic.hordiff_inputs(36, 28, 10, land_blocks=2, KhTr=3.0e4, check_diffusive_CFL=1)
    ref = _copy(a)
    n_ref = oracle.tracer_hordiff(dom, grid, gv, cs, ref)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    ra = dict(a, h=ctx.plane("hd.h", a["h"], "h", False, dom.nk), tr=[ctx.plane(f"hd.t{m}", t, "h", False, dom.nk) for m, t in enumerate(a["tr"])])
    assert ctx.tracer_hordiff(cs, ra) == n_ref
    for m, pl in enumerate(ra["tr"]):
        got = np.zeros_like(a["tr"][m]); pl.download(got)
        assert np.array_equal(_inner(dom, ref["tr"][m]).view(np.int64), _inner(dom, got).view(np.int64)), m
    for bad in (dict(use_neutral_diffusion=1), dict(Diffuse_ML_interior=1), dict(use_hor_bnd_diffusion=1)):
        with pytest.raises(Mom6cuError):
            ctx.tracer_hordiff(dict(cs, **bad), a)
