"""mixedlayer_restrat -> mixedlayer_restrat_OM4 (src/parameterizations/lateral/MOM_mixed_layer_restrat.F90:149-714), SURVEY 8f row 2.
CPU: mu(sigma, dh) against the reference's own unit-test values (mixedlayer_restrat_unit_tests :2022-2041) -- on the oracle and on
the host build of the code the GPU threads run (csrc/mle_mu.cuh) -- and properties of the oracle restatement of the routine (the
routine as a whole has no vector in the reference: parity unpinned).  GPU: C ABI == oracle, bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mom6_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (sigma, dh, expected, tolerance) -- MOM_mixed_layer_restrat.F90:2022-2041; tol = epsilon(1.) where the reference gives one
MU_VECTORS = [(3.0, 0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 0.0), (-0.25, 0.0, 0.7946428571428572, np.finfo(np.float64).eps),
              (-0.5, 0.0, 1.0, 0.0), (-0.75, 0.0, 0.7946428571428572, np.finfo(np.float64).eps), (-1.0, 0.0, 0.0, 0.0),
              (-3.0, 0.0, 0.0, 0.0), (-0.5, 0.5, 1.0, 0.0), (-1.0, 0.5, 0.25, 0.0), (-1.5, 0.5, 0.0, 0.0)]


@pytest.fixture(scope="module")
def mle_host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("mle") / "libmle_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "harness", "mle_host.cpp")])
    lib = C.CDLL(so)
    lib.mle_host_mu.argtypes = [C.c_double, C.c_double]; lib.mle_host_mu.restype = C.c_double
    lib.mle_host_density.argtypes = [C.c_int] + [C.c_double] * 7; lib.mle_host_density.restype = C.c_double
    return lib


def test_mu_reference_unit_test_vectors(oracle, mle_host):
    for sigma, dh, want, tol in MU_VECTORS:
        assert abs(oracle.mle_mu(sigma, dh) - want) <= tol, (sigma, dh)
        assert abs(mle_host.mle_host_mu(sigma, dh) - want) <= tol, (sigma, dh)


def test_mu_column_code_equals_oracle(oracle, mle_host):
    r = synthetic.rng(77)
    for s in np.concatenate((r.uniform(-1.6, 0.1, size=4000), [-0.5, -1.0, 0.0, -0.0, -0.49999999999, -1.0000000001])):
        assert mle_host.mle_host_mu(s, 0.0) == oracle.mle_mu(s, 0.0)            # the exponent-1 power is exact
    for s in r.uniform(-2.0, 0.1, size=500):                                     # a real power: same libm here, to an ulp anywhere
        a, b = mle_host.mle_host_mu(s, 0.25), oracle.mle_mu(s, 0.25)
        assert a == b or abs(a - b) <= 4 * np.finfo(np.float64).eps


CASES = [dict(), dict(land_blocks=4, eos="LINEAR"), dict(front_length=0.0, ml_restrat_coef=60.0, cyclic_y=True),
         dict(MLE_MLD_decay_time2=7.776e6, ml_restrat_coef2=0.5, land_blocks=2), dict(MLE_MLD_decay_time=0.0, MLE_MLD_stretch=1.5),
         dict(dt=7200.0, front=6.0, ml_restrat_coef=20.0, MLE_MLD_decay_time2=7.776e6, ml_restrat_coef2=5.0)]
# the mixed-layer depth detected from the density profile (MLE_DENSITY_DIFF > 0, detect_mld :1503) instead of the boundary-layer depth
EXT_CASES = [dict(MLE_density_diff=0.03, MLE_use_PBL_MLD=0, land_blocks=2), dict(MLE_density_diff=0.3, MLE_MLD_stretch=1.5, eos="LINEAR", cyclic_y=True),
             dict(MLE_density_diff=1.0e-4, MLE_MLD_decay_time2=7.776e6, ml_restrat_coef2=0.5)]


def _unified(dom, x, st):
    """A field in the library's unified plane layout (csrc/common.cuh): one offset for every staggering."""
    nj, ni = dom.jed - dom.jsd + 2, dom.ied - dom.isd + 2
    out = np.zeros(x.shape[:-2] + (nj, ni))
    out[..., (0 if st in "vq" else 1):, (0 if st in "uq" else 1):] = x
    return out


def _from_unified(x, st):
    return np.ascontiguousarray(x[..., (0 if st in "vq" else 1):, (0 if st in "uq" else 1):])


def _run_device_code_on_host(lib, dom, grid, gv, cs, a):
    """mixedlayer_restrat through the host build of csrc/mle_column.cuh, parameters set as mom6cu_mixedlayer_restrat sets them."""
    nk, dt = dom.nk, a["dt"]
    f1, f2 = cs["MLE_MLD_decay_time"] > 0.0, cs["MLE_MLD_decay_time2"] > 0.0
    par = [nk, dt, gv["Z_to_H"], gv["Angstrom_H"], gv["H_subroundoff"], gv["H_to_Z"] * gv["g_Earth"] / gv["Rho0"], 0.25 / dt,
           0.5 * gv["Angstrom_H"], cs["vonKar"] * 9.8696, cs["ustar_min"], cs["ml_restrat_coef"], cs["ml_restrat_coef2"], cs["front_length"],
           cs["MLE_MLD_stretch"], cs["MLE_tail_dh"],
           cs["MLE_MLD_decay_time"] / (dt + cs["MLE_MLD_decay_time"]) if f1 else 0.0, dt / (dt + cs["MLE_MLD_decay_time"]) if f1 else 0.0,
           cs["MLE_MLD_decay_time2"] / (dt + cs["MLE_MLD_decay_time2"]) if f2 else 0.0, dt / (dt + cs["MLE_MLD_decay_time2"]) if f2 else 0.0,
           int(f1), int(f2), int(cs["front_length"] > 0.0), cs["EOS_form"], cs["Rho_T0_S0"], cs["dRho_dT"], cs["dRho_dS"], cs["dRho_dp"],
           int(cs["MLE_density_diff"] > 0.0), cs["MLE_density_diff"]]
    par = np.array(par, dtype=np.float64)
    box = np.array([dom.isc, dom.iec, dom.jsc, dom.jec, dom.isd - 1, dom.jsd - 1], dtype=np.int32)
    F = {k: _unified(dom, a[k], st) for k, st in (("h", "h"), ("uhtr", "u"), ("vhtr", "v"), ("T", "h"), ("S", "h"), ("ustar", "h"),
                                                  ("h_MLD", "h"), ("Rd_dx_h", "h"))}
    F["MLD_filtered"], F["MLD_filtered_slow"] = _unified(dom, cs["MLD_filtered"], "h"), _unified(dom, cs["MLD_filtered_slow"], "h")
    Gd = {k: _unified(dom, grid[k], st) for k, st in (("areaT", "h"), ("IareaT", "h"), ("CoriolisBu", "q"), ("mask2dCu", "u"), ("mask2dCv", "v"),
                                                      ("dxCu", "u"), ("dyCu", "u"), ("dxCv", "v"), ("dyCv", "v"), ("IdxCu", "u"), ("IdyCv", "v"))}
    nj, ni = F["ustar"].shape
    scratch = np.zeros((4 + 2 * nk, nj, ni))
    p = lambda x: x.ctypes.data_as(C.c_void_p)   # noqa: E731
    lib.mle_host_run.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong] + [C.c_void_p] * 22
    lib.mle_host_run.restype = None
    lib.mle_host_run(p(par), p(box), ni, ni * nj, p(F["h"]), p(F["uhtr"]), p(F["vhtr"]), p(F["T"]), p(F["S"]), p(F["ustar"]), p(F["h_MLD"]),
                     p(F["Rd_dx_h"]), p(F["MLD_filtered"]), p(F["MLD_filtered_slow"]), p(Gd["areaT"]), p(Gd["IareaT"]), p(Gd["CoriolisBu"]),
                     p(Gd["mask2dCu"]), p(Gd["mask2dCv"]), p(Gd["dxCu"]), p(Gd["dyCu"]), p(Gd["dxCv"]), p(Gd["dyCv"]), p(Gd["IdxCu"]),
                     p(Gd["IdyCv"]), p(scratch))
    out = {k: _from_unified(F[k], st) for k, st in (("h", "h"), ("uhtr", "u"), ("vhtr", "v"), ("MLD_filtered", "h"), ("MLD_filtered_slow", "h"))}
    return out


def _run_oracle(oracle, dom, grid, gv, cs, a):
    o = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    c = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in cs.items()}
    assert oracle.mixedlayer_restrat(dom, grid, gv, c, o["h"], o["uhtr"], o["vhtr"], o["T"], o["S"], o["ustar"], o["dt"], o["h_MLD"],
                                     o["Rd_dx_h"]) == 0
    return c, o


def _inner(dom, x):
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]


def test_oracle_properties(oracle):
    dom, grid, gv, cs, a = synthetic.mle_inputs(36, 28, 14, land_blocks=3)
    c, o = _run_oracle(oracle, dom, grid, gv, cs, a)
    dt = a["dt"]
    duh, dvh = (o["uhtr"] - a["uhtr"]) / dt, (o["vhtr"] - a["vhtr"]) / dt         # = uhml, vhml
    assert np.abs(duh).max() > 0 and np.abs(dvh).max() > 0
    # an overturning circulation: no net transport through any face (the sum of a(k) over the column vanishes)
    scale = np.abs(duh).sum(axis=0).max()
    assert np.abs(duh.sum(axis=0)).max() < 1e-12 * scale and np.abs(dvh.sum(axis=0)).max() < 1e-12 * scale
    # nothing through land faces; the column thickness is unchanged; h = h - dt*div (:623-627)
    assert (duh[:, grid["mask2dCu"] == 0] == 0).all() and (dvh[:, grid["mask2dCv"] == 0] == 0).all()
    hi, ho = _inner(dom, a["h"]), _inner(dom, o["h"])
    m = _inner(dom, grid["mask2dT"]) > 0
    assert np.allclose(ho.sum(axis=0)[m], hi.sum(axis=0)[m], rtol=1e-13)
    assert np.abs(ho - hi).max() > 1e-6 and ho.min() >= 0.5 * gv["Angstrom_H"]
    j0, i0 = dom.jsc - dom.jsd, dom.isc - dom.isd
    nj, ni = dom.jec - dom.jsc + 1, dom.iec - dom.isc + 1
    iu0 = i0 + (1 if a["uhtr"].shape[-1] > a["h"].shape[-1] else 0)
    jv0 = j0 + (1 if a["vhtr"].shape[-2] > a["h"].shape[-2] else 0)
    div = ((duh[:, j0:j0 + nj, iu0:iu0 + ni] - duh[:, j0:j0 + nj, iu0 - 1:iu0 + ni - 1]) +
           (dvh[:, jv0:jv0 + nj, i0:i0 + ni] - dvh[:, jv0 - 1:jv0 + nj - 1, i0:i0 + ni]))
    want = np.maximum(hi - dt * _inner(dom, grid["IareaT"])[None] * div, 0.5 * gv["Angstrom_H"])
    assert np.allclose(ho, want, rtol=1e-12, atol=1e-12)
    # the transport is confined to the mixed layer: zero below the deeper of the two neighbouring filtered depths
    # the MLD filter only deepens instantly (:323)
    ext = (slice(dom.jsc - dom.jsd - 1, dom.jec - dom.jsd + 2), slice(dom.isc - dom.isd - 1, dom.iec - dom.isd + 2))
    assert (c["MLD_filtered"][ext] >= (cs["MLE_MLD_stretch"] * a["h_MLD"])[ext]).all()
    assert not np.array_equal(c["MLD_filtered"], cs["MLD_filtered"])


def test_oracle_restratifies(oracle):
    """The overturning flattens isopycnals: with a pure x front, light water moves over dense water, so the upper-layer transport
    is directed from the light (warm) side to the dense side."""
    dom, grid, gv, cs, a = synthetic.mle_inputs(32, 16, 10, front=0.0, eos="LINEAR")
    ii = np.arange(a["T"].shape[-1])
    a["T"][...] = 10.0 + 2.0 * np.sin(2 * np.pi * (ii - (dom.isc - dom.isd)) / 32.0)[None, None, :]
    a["S"][...] = 35.0
    a["h_MLD"][...] = 400.0; a["ustar"][...] = 0.01
    cs["MLE_MLD_decay_time"] = 0.0
    c, o = _run_oracle(oracle, dom, grid, gv, cs, a)
    duh = _inner(dom, o["uhtr"] - a["uhtr"])
    Tx = _inner(dom, np.roll(a["T"], -1, axis=-1) - a["T"])[0]
    sel = (np.abs(Tx) > 0.05) & (np.abs(duh[0][..., :Tx.shape[-1]]) > 0)
    # u-face I sits between cells i and i+1: warm to the west (Tx < 0) -> surface flow eastward (> 0)
    off = 1 if a["uhtr"].shape[-1] > a["h"].shape[-1] else 0
    top = (o["uhtr"] - a["uhtr"])[0][dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd + off:dom.iec - dom.isd + 1 + off]
    sel = (np.abs(Tx) > 0.05) & (top != 0)
    assert sel.sum() > 50 and (np.sign(top[sel]) == -np.sign(Tx[sel])).all()


def test_oracle_rejects_options_outside_the_frozen_set(oracle):
    dom, grid, gv, cs, a = synthetic.mle_inputs(12, 10, 4)
    for bad in (dict(use_Bodner=1), dict(MLE_use_PBL_MLD=0), dict(EOS_form=0), dict(use_Stanley_ML=1)):
        with pytest.raises(RuntimeError):
            _run_oracle(oracle, dom, grid, gv, dict(cs, **bad), a)



@pytest.mark.parametrize("kw", CASES + EXT_CASES)
def test_device_column_code_equals_oracle_on_the_host(oracle, mle_host, kw):
    """The column / face / update functions the kernels of csrc/mle.cu call (early exit at the base of the mixed layer, mu reused
    between a layer's bottom and the next layer's top, h_avail evaluated in place) compiled for the host: bit for bit the oracle."""
    for (ni, nj, nk) in ((44, 40, 20), (31, 9, 3), (20, 22, 75)):
        dom, grid, gv, cs, a = synthetic.mle_inputs(ni, nj, nk, **kw)
        a["uhtr"][:, 5:9, 5:30] = -0.0                                   # signed zeros must survive below the mixed layer
        a["vhtr"][:, 5:9, 5:30] = 0.0
        c, o = _run_oracle(oracle, dom, grid, gv, cs, a)
        got = _run_device_code_on_host(mle_host, dom, grid, gv, cs, a)
        assert np.array_equal(_inner(dom, o["h"]).view(np.int64), _inner(dom, got["h"]).view(np.int64)), kw
        for k in ("uhtr", "vhtr"):
            assert np.array_equal(o[k].view(np.int64), got[k].view(np.int64)), (k, kw)
        ext = (slice(dom.jsc - dom.jsd - 1, dom.jec - dom.jsd + 2), slice(dom.isc - dom.isd - 1, dom.iec - dom.isd + 2))
        for k in ("MLD_filtered", "MLD_filtered_slow"):
            assert np.array_equal(c[k][ext].view(np.int64), got[k][ext].view(np.int64)), (k, kw)
        assert nk < 20 or cs["MLE_density_diff"] > 0 or np.abs(o["uhtr"] - a["uhtr"]).max() > 0


@pytest.mark.gpu
def test_mle_mu_device(oracle, ctx_factory):
    dom, *_ = synthetic.mle_inputs(12, 10, 4)
    ctx = ctx_factory(dom)
    sig = np.array([v[0] for v in MU_VECTORS]); dh = np.array([v[1] for v in MU_VECTORS])
    got = ctx.mle_mu(sig, dh)
    for g, (_, _, want, tol) in zip(got, MU_VECTORS):
        assert abs(g - want) <= max(tol, 0.0 if want in (0.0, 1.0) else 2 * np.finfo(np.float64).eps)
    s = synthetic.rng(5).uniform(-1.6, 0.1, size=5000)
    assert np.array_equal(ctx.mle_mu(s, 0.0), np.array([oracle.mle_mu(x, 0.0) for x in s]))


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_mixedlayer_restrat_bitwise(oracle, ctx_factory, kw):
    for (ni, nj, nk) in ((44, 40, 20), (131, 9, 3), (30, 22, 75)):
        dom, grid, gv, cs, a = synthetic.mle_inputs(ni, nj, nk, **kw)
        c, o = _run_oracle(oracle, dom, grid, gv, cs, a)
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv)
        n0 = ctx.launches
        ctx.mixedlayer_restrat(cs, a["h"], a["uhtr"], a["vhtr"], a["T"], a["S"], a["ustar"], a["dt"], a["h_MLD"], a["Rd_dx_h"])
        assert ctx.launches - n0 >= 4
        assert np.array_equal(_inner(dom, o["h"]).view(np.int64), _inner(dom, a["h"]).view(np.int64)), kw
        for k in ("uhtr", "vhtr"):
            assert np.array_equal(o[k].view(np.int64), a[k].view(np.int64)), (k, kw)
        ext = (slice(dom.jsc - dom.jsd - 1, dom.jec - dom.jsd + 2), slice(dom.isc - dom.isd - 1, dom.iec - dom.isd + 2))
        for k in ("MLD_filtered", "MLD_filtered_slow"):
            assert np.array_equal(c[k][ext].view(np.int64), cs[k][ext].view(np.int64)), (k, kw)
        assert np.abs(o["uhtr"]).max() > 0


@pytest.mark.gpu
def test_mixedlayer_restrat_errors(ctx_factory):
    from mom6_b200.api import Mom6cuError
    dom, grid, gv, cs, a = synthetic.mle_inputs(16, 12, 5)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    for bad in (dict(use_Bodner=1), dict(MLE_use_PBL_MLD=0), dict(EOS_form=0), dict(use_Stanley_ML=1), dict(MLE_tail_dh=0.1)):
        with pytest.raises(Mom6cuError):
            ctx.mixedlayer_restrat(dict(cs, **bad), a["h"], a["uhtr"], a["vhtr"], a["T"], a["S"], a["ustar"], a["dt"], a["h_MLD"], a["Rd_dx_h"])
