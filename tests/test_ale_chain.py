"""ALE_regridding_and_remapping (src/core/MOM.F90:1751-1926) as one device entry, with interpolate_column
(src/ALE/MOM_remapping.F90:1247), ALE_remap_interface_vals / ALE_remap_vertex_vals (src/ALE/MOM_ALE.F90:1303 / :1342).
CPU: the reference's interpolate_column unit-test vectors (MOM_remapping.F90:2648-2682) against the oracle and against the
column code the GPU threads run (csrc/interp_column.cuh through a test-only harness); conservation properties of the oracle's
chain.  GPU: the same vectors through the C ABI; every output of the chain == oracle, bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mom6_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))

# test_interp(test, msg, nsrc, h_src, u_src, ndest, h_dest, u_true), MOM_remapping.F90:2648-2682 (mask_edges = .true. :1826)
INTERP_VECTORS = [
    ("Identity: 3 layer", [1., 2., 3.], [1., 2., 3., 4.], [1., 2., 3.], [1., 2., 3., 4.]),
    ("A: 3 layer to 2", [1., 1., 1.], [1., 2., 3., 4.], [1.5, 1.5], [1., 2.5, 4.]),
    ("B: 2 layer to 3", [1.5, 1.5], [1., 4., 7.], [1., 1., 1.], [1., 3., 5., 7.]),
    ("C: 3 layer (vanished middle) to 2", [1., 0., 2.], [1., 2., 2., 3.], [1., 2.], [1., 2., 3.]),
    ("D: 3 layer (deep) to 3", [1., 2., 3.], [1., 2., 4., 7.], [2., 2.], [1., 3., 5.]),
    ("E: 3 layer to 3 (deep)", [1., 2., 4.], [1., 2., 4., 8.], [2., 3., 4.], [1., 3., 6., 8.]),
    ("F: 3 layer to 4 with vanished top/botton", [1., 2., 4.], [1., 2., 4., 8.], [0., 2., 5., 0.], [0., 1., 3., 8., 0.]),
    ("Fs: 3 layer to 4 with vanished top/botton (shallow)", [1., 2., 4.], [1., 2., 4., 8.], [0., 2., 4., 0.], [0., 1., 3., 7., 0.]),
    ("Fd: 3 layer to 4 with vanished top/botton (deep)", [1., 2., 4.], [1., 2., 4., 8.], [0., 2., 6., 0.], [0., 1., 3., 8., 0.]),
]


@pytest.fixture(scope="module")
def interp_harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("interp_host") / "libinterp_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(HERE, "harness", "interp_host.cpp"), "-o", out])
    lib = C.CDLL(out)
    lib.interp_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.interp_host.restype = None

    def run(h_src, u_src, h_dest, mask_edges):
        h_src, u_src, h_dest = (np.ascontiguousarray(x, dtype=np.float64) for x in (h_src, u_src, h_dest))
        u = np.zeros(len(h_dest) + 1)
        lib.interp_host(len(h_src), h_src.ctypes.data, u_src.ctypes.data, len(h_dest), h_dest.ctypes.data, u.ctypes.data, int(mask_edges))
        return u
    return run


def _random_columns(rng, n):
    for _ in range(n):
        ns, nd = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        hs, hd = rng.uniform(0, 3, ns), rng.uniform(0, 3, nd)
        hs[rng.random(ns) < 0.25] = 0.0
        hd[rng.random(nd) < 0.25] = 0.0
        if rng.random() < 0.5 and hd.sum() > 0:
            hd *= hs.sum() / hd.sum()
        yield hs, rng.standard_normal(ns + 1), hd


def test_interpolate_column_reference_vectors(oracle, interp_harness):
    for msg, hs, us, hd, want in INTERP_VECTORS:
        assert np.array_equal(oracle.interpolate_column(hs, us, hd, True), np.array(want)), msg   # the reference compares with /=
        assert np.array_equal(interp_harness(hs, us, hd, True), np.array(want)), msg


def test_device_column_code_matches_oracle(oracle, interp_harness):
    rng = np.random.default_rng(17)
    for hs, us, hd in _random_columns(rng, 400):
        for mask in (False, True):
            assert np.array_equal(interp_harness(hs, us, hd, mask), oracle.interpolate_column(hs, us, hd, mask))


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_copy(v) for v in x]
    return x


def _inner(dom, x, di=0, dj=0):
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1 + dj, dom.isc - dom.isd:dom.iec - dom.isd + 1 + di]


def test_oracle_chain_properties(oracle):
    dom, grid, gv, ale, cs, a = synthetic.ale_chain_inputs(24, 20, 8, land_blocks=2)
    ref = _copy(a)
    ref["tr"].append(np.full_like(a["h"], 3.25))   # a uniform tracer must stay uniform
    ref["conc_underflow"] = np.array([0.0, 0.0, 1.0e-25, 0.0])
    rcs = _copy(cs)
    oracle.ale_regridding_and_remapping(dom, grid, gv, ale, ref, dyn_cs=rcs)
    assert ale["regridCS"]["old_grid_weight"] == 3600.0 / (3600.0 + 7200.0)
    wet = _inner(dom, grid["mask2dT"]) > 0
    h0, h1 = _inner(dom, a["h"]), _inner(dom, ref["h"])
    assert np.abs(h1 - h0).max() > 1.0   # the grid really moved
    assert np.all(h1 >= 0.0)
    tot0, tot1 = h0.sum(axis=0), h1.sum(axis=0)
    assert np.abs(tot1 - tot0)[wet].max() <= 1e-9 * tot0.max()
    for m in (0, 1):   # tracer inventories of every column
        c0, c1 = (h0 * _inner(dom, a["tr"][m])).sum(axis=0), (h1 * _inner(dom, ref["tr"][m])).sum(axis=0)
        assert np.abs(c1 - c0)[wet].max() <= 1e-9 * np.abs(c0).max()
        t0, t1 = _inner(dom, a["tr"][m]), _inner(dom, ref["tr"][m])
        assert t1[:, wet].max() <= t0[:, wet].max() + 1e-12 and t1[:, wet].min() >= t0[:, wet].min() - 1e-12
    u1 = _inner(dom, ref["tr"][3])
    assert np.abs(u1[:, wet] - 3.25).max() <= 1e-14
    # interface values are interpolated: they stay inside the bounds of their column
    k0, k1 = _inner(dom, a["Kv_shear"]), _inner(dom, ref["Kv_shear"])
    assert np.all(k1[:, wet].max(axis=0) <= k0[:, wet].max(axis=0)) and np.all(k1[:, wet].min(axis=0) >= k0[:, wet].min(axis=0))
    assert not np.array_equal(k0, k1)
    # land columns are untouched
    assert np.array_equal(_inner(dom, a["tr"][0])[:, ~wet], _inner(dom, ref["tr"][0])[:, ~wet])
    assert not np.array_equal(cs["diffu"], rcs["diffu"]) and not np.array_equal(cs["u_av"], rcs["u_av"])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_interpolate_column(oracle, ctx_factory):
    from mom6_b200.api import make_domain
    ctx = ctx_factory(make_domain(8, 8, nk=2))
    for msg, hs, us, hd, want in INTERP_VECTORS:
        assert np.array_equal(ctx.interpolate_column(hs, us, hd, True)[0], np.array(want)), msg
    rng = np.random.default_rng(18)
    for ns, nd in ((5, 7), (12, 3), (1, 1), (75, 75)):
        hs = rng.uniform(0, 3, (64, ns)); hd = rng.uniform(0, 3, (64, nd)); us = rng.standard_normal((64, ns + 1))
        hs[rng.random(hs.shape) < 0.2] = 0.0; hd[rng.random(hd.shape) < 0.2] = 0.0
        for mask in (False, True):
            got = ctx.interpolate_column(hs, us, hd, mask)
            for c in range(64):
                assert np.array_equal(got[c], oracle.interpolate_column(hs[c], us[c], hd[c], mask))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(24, 20, 8), (70, 40, 20)])
def test_gpu_remap_interface_and_vertex_vals(oracle, ctx_factory, shape):
    dom, grid, gv, ale, cs, a = synthetic.ale_chain_inputs(*shape, land_blocks=2)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    rng = np.random.default_rng(19)
    h_old = a["h"]
    h_new = np.ascontiguousarray(h_old * (1.0 + 0.3 * rng.uniform(-1, 1, h_old.shape)))
    h_new[rng.random(h_new.shape) < 0.1] = 0.0
    for vertex, key in ((False, "Kv_shear"), (True, "Kv_shear_Bu")):
        ref, got = a[key].copy(), a[key].copy()
        oracle.ale_remap_vals(dom, grid, h_old, h_new, ref, vertex=vertex)
        (ctx.ale_remap_vertex_vals if vertex else ctx.ale_remap_interface_vals)(h_old, h_new, got)
        assert np.array_equal(ref.view(np.int64), got.view(np.int64)), key
        assert not np.array_equal(ref, a[key])


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(remap_aux_vars=0), dict(store_CAu=0, with_Bu=False, regrid_time_scale=0.0), dict(remapping_scheme=2)])
@pytest.mark.parametrize("shape", [(24, 20, 8), (60, 44, 30)])
def test_gpu_ale_regridding_and_remapping_matches_oracle(oracle, ctx_factory, kw, shape):
    dom, grid, gv, ale, cs, a = synthetic.ale_chain_inputs(*shape, land_blocks=2, **kw)
    ref, rcs, rale = _copy(a), _copy(cs), _copy(ale)
    oracle.ale_regridding_and_remapping(dom, grid, gv, rale, ref, dyn_cs=rcs)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    ctx.ale_regridding_and_remapping(ale, a, dyn_cs=cs)
    assert ale["regridCS"]["old_grid_weight"] == rale["regridCS"]["old_grid_weight"]
    eq = lambda x, y: np.array_equal(np.ascontiguousarray(x).view(np.int64), np.ascontiguousarray(y).view(np.int64))   # noqa: E731
    assert eq(_inner(dom, ref["h"]), _inner(dom, a["h"])), "h"
    assert eq(ref["h"][:, dom.jsc - 2:dom.jec + 1, dom.isc - 2:dom.iec + 1], a["h"][:, dom.jsc - 2:dom.jec + 1, dom.isc - 2:dom.iec + 1]), "h halo 1"
    assert eq(_inner(dom, ref["u"], di=1), _inner(dom, a["u"], di=1)) and eq(_inner(dom, ref["v"], dj=1), _inner(dom, a["v"], dj=1)), "u, v"
    for m in range(3):
        assert eq(_inner(dom, ref["tr"][m]), _inner(dom, a["tr"][m])), f"tracer {m}"
    for k, (di, dj) in (("Kd_shear", (0, 0)), ("Kv_shear", (0, 0)), ("Kv_shear_Bu", (1, 1))):
        if a[k] is not None:
            assert eq(_inner(dom, ref[k], di, dj), _inner(dom, a[k], di, dj)), k
    for k in ("diffu", "u_av", "CAu_pred"):
        assert eq(_inner(dom, rcs[k], di=1), _inner(dom, cs[k], di=1)), k
    for k in ("diffv", "v_av", "CAv_pred"):
        assert eq(_inner(dom, rcs[k], dj=1), _inner(dom, cs[k], dj=1)), k
    assert not eq(_inner(dom, ref["h"]), _inner(dom, _copy(synthetic.ale_chain_inputs(*shape, land_blocks=2, **kw)[5]["h"])))
