"""ALE_regrid, Z* coordinate (src/ALE/MOM_ALE.F90:518 -> MOM_regridding.F90:846-1857, coord_zlike.F90:63).
CPU: properties of the oracle restatement (parity unpinned).  GPU: C ABI == oracle, bit for bit."""
import numpy as np
import pytest

from mom6_b200 import synthetic


def _ext(dom, x):  # (isc-1:iec+1, jsc-1:jec+1), the range the reference writes
    return x[..., dom.jsc - dom.jsd - 1:dom.jec - dom.jsd + 2, dom.isc - dom.isd - 1:dom.iec - dom.isd + 2]


def test_zstar_grid_properties(oracle):
    dom, grid, gv, cs, a = synthetic.regrid_inputs(30, 22, 12, land_blocks=3)
    assert oracle.ale_regrid(dom, grid, gv, cs, a["h"], a["h_new"], a["dzRegrid"]) == 0
    m = _ext(dom, grid["mask2dT"]) > 0
    h, hn, dz = _ext(dom, a["h"]), _ext(dom, a["h_new"]), _ext(dom, a["dzRegrid"])
    assert np.allclose(hn.sum(axis=0)[m], h.sum(axis=0)[m], rtol=1e-13)          # the column thickness is kept
    assert (dz[0] == 0).all() and np.abs(dz[-1][m]).max() < 1e-9                  # top and bottom interfaces stay
    assert hn[:, m].min() >= cs["min_thickness"] * (1 - 1e-12) or hn[:, m].min() >= 0
    assert np.array_equal(hn[:, ~m], h[:, ~m]) and (dz[:, ~m] == 0).all()         # land keeps h (:1035)
    # z* definition: where no layer is inflated, layer k is the stretched target resolution
    tot = h.sum(axis=0); depth = _ext(dom, grid["bathyT"])
    deep = m & (depth > 3999.0)
    k = 2
    with np.errstate(divide="ignore", invalid="ignore"):
        assert np.allclose(hn[k][deep], (cs["coordinateResolution"][k] * tot / depth)[deep], rtol=1e-10)


def test_time_filter_slows_deep_interfaces(oracle):
    dom, grid, gv, cs, a = synthetic.regrid_inputs(24, 18, 10, eta_amp=2.0)
    full = {k: v.copy() for k, v in a.items()}
    assert oracle.ale_regrid(dom, grid, gv, cs, full["h"], full["h_new"], full["dzRegrid"]) == 0
    csf = dict(cs, old_grid_weight=0.75, depth_of_time_filter_shallow=200.0, depth_of_time_filter_deep=1000.0)
    filt = {k: v.copy() for k, v in a.items()}
    assert oracle.ale_regrid(dom, grid, gv, csf, filt["h"], filt["h_new"], filt["dzRegrid"]) == 0
    d0, d1 = np.abs(_ext(dom, full["dzRegrid"])), np.abs(_ext(dom, filt["dzRegrid"]))
    assert (d1 <= d0 + 1e-12).all() and d1.sum() < 0.9 * d0.sum()
    zold = np.cumsum(_ext(dom, a["h"]), axis=0)
    deep = zold[:-1] > 1100.0                                                       # interfaces below the deep filter depth
    sel = deep & (d0[1:-1] > 1e-9)
    ratio = d1[1:-1][sel] / d0[1:-1][sel]                                            # wtd = 1 - old_grid_weight (:1183), unless inflated (:1840)
    assert sel.any() and abs(np.median(ratio) - 0.25) < 1e-9 and (np.abs(ratio - 0.25) < 1e-6).mean() > 0.9


def test_negative_thickness_is_fatal(oracle):
    dom, grid, gv, cs, a = synthetic.regrid_inputs(16, 12, 5)
    a["h"][2, dom.jsc - dom.jsd + 3, dom.isc - dom.isd + 3] = -1.0
    assert oracle.ale_regrid(dom, grid, gv, cs, a["h"], a["h_new"], a["dzRegrid"]) != 0


CASES = [dict(), dict(land_blocks=4, old_grid_weight=0.75, depth_of_time_filter_shallow=200.0, depth_of_time_filter_deep=1000.0),
         dict(land_blocks=2, cyclic_y=True, min_thickness=0.0, eta_amp=3.0), dict(min_thickness=5.0, eta_amp=0.0),
         dict(old_grid_weight=0.5, depth_of_time_filter_shallow=0.0, depth_of_time_filter_deep=0.0, land_blocks=3),
         dict(old_grid_weight=0.9, depth_of_time_filter_shallow=50.0, depth_of_time_filter_deep=3000.0, Z_ref=2.0)]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_ale_regrid_bitwise(oracle, ctx_factory, kw):
    for (ni, nj, nk) in ((44, 40, 20), (131, 9, 3), (30, 22, 75)):
        dom, grid, gv, cs, a = synthetic.regrid_inputs(ni, nj, nk, **kw)
        ref = {k: v.copy() for k, v in a.items()}
        assert oracle.ale_regrid(dom, grid, gv, cs, ref["h"], ref["h_new"], ref["dzRegrid"]) == 0
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv)
        n0 = ctx.launches
        ctx.ale_regrid(cs, a["h"], a["h_new"], a["dzRegrid"])
        assert ctx.launches > n0
        for k in ("h_new", "dzRegrid"):
            assert np.array_equal(_ext(dom, ref[k]).view(np.int64), _ext(dom, a[k]).view(np.int64)), (k, kw)
        assert np.abs(_ext(dom, a["dzRegrid"])).max() > 0


@pytest.mark.gpu
def test_ale_regrid_errors(ctx_factory):
    from mom6_b200.api import Mom6cuError
    dom, grid, gv, cs, a = synthetic.regrid_inputs(16, 12, 5)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    with pytest.raises(Mom6cuError):
        ctx.ale_regrid(dict(cs, regridding_scheme=5), a["h"], a["h_new"], a["dzRegrid"])
    bad = a["h"].copy(); bad[2, dom.jsc - dom.jsd + 3, dom.isc - dom.isd + 3] = -1.0
    with pytest.raises(Mom6cuError):
        ctx.ale_regrid(cs, bad, a["h_new"], a["dzRegrid"])
