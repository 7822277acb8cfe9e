"""Tile-layout invariance (the reference's `layout` regression test, .testing/Makefile: 1 PE vs LAYOUT = 2,1 / 1,2 / 2,2 must agree bit for bit)
of the callers of step_MOM_dynamics / step_MOM_tracer_dyn on the oracle -- mixedlayer_restrat and thickness_diffuse read one halo point of h, T,
S and the 2-D inputs and exchange nothing inside the call; tracer_hordiff is run for the single sweep whose only exchange opens it --: each tile of the layout, cut out of the single-tile inputs with its halos, must reproduce the single-tile answer on
its own computational domain.  (The dycore step, the barotropic solver and the ocean.stats line have their layout tests on the device:
tests/test_step_multigpu.py, test_bt_multigpu.py, test_diag.py.)"""
import numpy as np
import pytest

from mom6_b200 import synthetic
from mom6_b200.api import make_domain


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    return x


def _tile(dom_g, npi, npj, pi, pj):
    NI, NJ = dom_g.iec - dom_g.isc + 1, dom_g.jec - dom_g.jsc + 1
    ni, nj = NI // npi, NJ // npj
    dom = make_domain(ni, nj, nk=dom_g.nk, halo=dom_g.isc - dom_g.isd, cyclic_x=bool(dom_g.cyclic_x), cyclic_y=bool(dom_g.cyclic_y),
                      first_direction=dom_g.first_direction, npi=npi, npj=npj, pi=pi, pj=pj)
    return dom, pi * ni, pj * nj


def _cut_all(dom_g, dom, oi, oj, d, staggers):
    # .copy(): a tile that spans the full width is a contiguous slice, i.e. a view of the single-tile array
    return {k: (synthetic._cut(dom_g, dom, v, staggers[k], False, oi, oj).copy() if isinstance(v, np.ndarray) and v.ndim >= 2 else v) for k, v in d.items()}


def _inner_of_tile(dom_g, dom, oi, oj, x_g, st):
    """The part of a single-tile array that is this tile's computational domain (incl. the symmetric west / south face of staggered fields)."""
    su, sv = int(st in "uq"), int(st in "vq")
    j0, i0 = dom_g.jsc - dom_g.jsd + oj, dom_g.isc - dom_g.isd + oi
    return x_g[..., j0:j0 + (dom.jec - dom.jsc + 1) + sv, i0:i0 + (dom.iec - dom.isc + 1) + su]


def _inner(dom, x, st):
    su, sv = int(st in "uq"), int(st in "vq")
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1 + sv, dom.isc - dom.isd:dom.iec - dom.isd + 1 + su]


LAYOUTS = [(2, 1), (1, 2), (2, 2)]
MLE_ST = dict(h="h", uhtr="u", vhtr="v", T="h", S="h", ustar="h", h_MLD="h", Rd_dx_h="h", MLD_filtered="h", MLD_filtered_slow="h")
TD_ST = dict(h="h", uhtr="u", vhtr="v", T="h", S="h", p_surf="h", Res_fn_u="u", Res_fn_v="v", uhGM="u", vhGM="v", slope_x="u", slope_y="v", cg1="h", MEKE_Kh="h")


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("kw", [dict(), dict(MLE_density_diff=0.1, MLE_MLD_decay_time2=7.776e6, ml_restrat_coef2=0.5)])
def test_mixedlayer_restrat_layout(oracle, layout, kw):
    dom_g, grid_g, gv, cs_g, a_g = synthetic.mle_inputs(24, 20, 24, land_blocks=2, MLE_MLD_stretch=3.0, **kw)
    rc, ra = _copy(cs_g), _copy(a_g)
    oracle.mixedlayer_restrat(dom_g, grid_g, gv, rc, ra["h"], ra["uhtr"], ra["vhtr"], ra["T"], ra["S"], ra["ustar"], ra["dt"], ra["h_MLD"], ra["Rd_dx_h"])
    assert not np.array_equal(ra["h"], a_g["h"])
    npi, npj = layout
    for pj in range(npj):
        for pi in range(npi):
            dom, oi, oj = _tile(dom_g, npi, npj, pi, pj)
            grid = _cut_all(dom_g, dom, oi, oj, grid_g, synthetic.GRID_STAGGER)
            a = _cut_all(dom_g, dom, oi, oj, a_g, MLE_ST)
            cs = _cut_all(dom_g, dom, oi, oj, cs_g, MLE_ST)
            oracle.mixedlayer_restrat(dom, grid, gv, cs, a["h"], a["uhtr"], a["vhtr"], a["T"], a["S"], a["ustar"], a["dt"], a["h_MLD"], a["Rd_dx_h"])
            for k in ("h", "uhtr", "vhtr"):
                assert np.array_equal(_inner_of_tile(dom_g, dom, oi, oj, ra[k], MLE_ST[k]), _inner(dom, a[k], MLE_ST[k])), (k, layout, pi, pj)
            assert np.array_equal(_inner_of_tile(dom_g, dom, oi, oj, rc["MLD_filtered"], "h"), _inner(dom, cs["MLD_filtered"], "h"))


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("kw", [dict(with_GM=True), dict(use_FGNV_streamfn=1, use_stored_slopes=1, use_MEKE_Kh=1, Khth=0.0, use_variable_mixing=1, Resoln_scaled_KhTh=1)])
def test_thickness_diffuse_layout(oracle, layout, kw):
    dom_g, grid_g, gv, cs, a_g = synthetic.thickness_diffuse_inputs(24, 20, 10, land_blocks=2, **kw)
    ra = _copy(a_g)
    oracle.thickness_diffuse(dom_g, grid_g, gv, cs, ra)
    assert not np.array_equal(ra["h"], a_g["h"])
    npi, npj = layout
    for pj in range(npj):
        for pi in range(npi):
            dom, oi, oj = _tile(dom_g, npi, npj, pi, pj)
            grid = _cut_all(dom_g, dom, oi, oj, grid_g, synthetic.GRID_STAGGER)
            a = _cut_all(dom_g, dom, oi, oj, a_g, TD_ST)
            oracle.thickness_diffuse(dom, grid, gv, cs, a)
            for k in ("h", "uhtr", "vhtr", "uhGM", "vhGM"):
                if a.get(k) is not None:
                    assert np.array_equal(_inner_of_tile(dom_g, dom, oi, oj, ra[k], TD_ST[k]), _inner(dom, a[k], TD_ST[k])), (k, layout, pi, pj)


HD_ST = dict(h="h", Res_fn_h="h", Rd_dx_h="h", L2u="u", SN_u="u", L2v="v", SN_v="v", MEKE_Kh="h")


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("kw", [dict(with_df=True), dict(use_variable_mixing=1, KhTr_Slope_Cff=0.05, use_MEKE_Kh=1, Resoln_scaled_KhTr=1, KhTr_max=1500.0)])
def test_tracer_hordiff_layout(oracle, layout, kw):
    """One sweep (num_itts = 1): the only exchange is the halo update that opens it (do_group_pass :541), which the tiles receive here as
    the halos cut from the halo-filled single-tile tracers; the tiles themselves are run with closed edges so that nothing wraps onto a tile."""
    from mom6_b200 import fidx
    dom_g, grid_g, gv, cs, a_g = synthetic.hordiff_inputs(24, 20, 6, land_blocks=2, **kw)
    for m, t in enumerate(a_g["tr"]):                                         # what do_group_pass leaves on one tile
        f = fidx.FA(dom_g.isd, dom_g.ied, dom_g.jsd, dom_g.jed, nk=dom_g.nk); f.a[...] = t; fidx.fill_halo(dom_g, f, "h"); a_g["tr"][m] = np.ascontiguousarray(f.a)
    ra = {k: ([x if x is None else x.copy() for x in v] if isinstance(v, list) else _copy(v)) for k, v in a_g.items()}
    assert oracle.tracer_hordiff(dom_g, grid_g, gv, cs, ra) == 1
    npi, npj = layout
    for pj in range(npj):
        for pi in range(npi):
            dom, oi, oj = _tile(dom_g, npi, npj, pi, pj)
            dom.cyclic_x, dom.cyclic_y = 0, 0
            grid = _cut_all(dom_g, dom, oi, oj, grid_g, synthetic.GRID_STAGGER)
            a = _cut_all(dom_g, dom, oi, oj, {k: v for k, v in a_g.items() if k in HD_ST}, HD_ST)
            a.update(dt=a_g["dt"], conc_underflow=a_g["conc_underflow"], tr=[synthetic._cut(dom_g, dom, t, "h", False, oi, oj).copy() for t in a_g["tr"]])
            for key, st in (("df_x", "u"), ("df_y", "v")):
                if a_g.get(key) is not None:
                    a[key] = [None if f is None else synthetic._cut(dom_g, dom, f, st, False, oi, oj).copy() for f in a_g[key]]
            assert oracle.tracer_hordiff(dom, grid, gv, cs, a) == 1
            for m in range(len(a["tr"])):
                assert np.array_equal(_inner_of_tile(dom_g, dom, oi, oj, ra["tr"][m], "h"), _inner(dom, a["tr"][m], "h")), (m, layout, pi, pj)
            for key, st in (("df_x", "u"), ("df_y", "v")):
                for m, f in enumerate(a.get(key) or []):
                    if f is not None:
                        assert np.array_equal(_inner_of_tile(dom_g, dom, oi, oj, ra[key][m], st), _inner(dom, f, st)), (key, m, layout, pi, pj)
    assert not np.array_equal(ra["tr"][0], a_g["tr"][0])
