"""advect_tracer (src/tracer/MOM_tracer_advect.F90:53-1152).  CPU: properties of the oracle restatement (there are no
known-answer vectors in the reference: parity unpinned).  GPU: C ABI == oracle, bit for bit."""
import numpy as np
import pytest

from mom6_b200 import synthetic


def _copy(a):
    b = dict(a)
    b["tr"] = [t.copy() for t in a["tr"]]
    for k in ("vol_prev", "uhr_out", "vhr_out"):
        if b.get(k) is not None:
            b[k] = b[k].copy()
    return b


def _inner(dom, x):
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]


def _vol0(dom, grid, a):
    """hprev of :188-195."""
    div = np.zeros_like(a["h_end"])
    div[:, 1:-1, 1:-1] = (a["uhtr"][:, 1:-1, 2:-1] - a["uhtr"][:, 1:-1, 1:-2]) + (a["vhtr"][:, 2:-1, 1:-1] - a["vhtr"][:, 1:-2, 1:-1])
    v = np.maximum(0.0, grid["areaT"][None] * a["h_end"] + div)
    return v + np.maximum(0.0, 1.0e-13 * v - grid["areaT"][None] * a["h_end"])


@pytest.mark.parametrize("scheme", [0, 1, 2])
@pytest.mark.parametrize("x_first", [True, False])
def test_conserves_bounds_and_preserves_uniform(oracle, scheme, x_first):
    dom, grid, gv, cs, a = synthetic.advect_inputs(36, 28, 6, land_blocks=3, ntr=3, scheme=scheme, x_first_in=x_first, cfl=0.9)
    a["tr"][2][...] = 1.0
    v0 = _vol0(dom, grid, a)
    b = _copy(a)
    b["vol_prev"] = v0.copy(); b["update_vol_prev"] = True
    b["uhr_out"] = np.zeros_like(a["uhtr"]); b["vhr_out"] = np.zeros_like(a["vhtr"])
    nit = oracle.advect_tracer(dom, grid, gv, cs, b)
    assert 1 <= nit <= 2 * 4 + 1
    m = (_inner(dom, grid["mask2dT"]) > 0)[None]
    for t0, t1 in zip(a["tr"], b["tr"]):
        before = (_inner(dom, t0) * _inner(dom, v0) * m).sum(); after = (_inner(dom, t1) * _inner(dom, b["vol_prev"]) * m).sum()
        assert abs(after - before) <= 1e-11 * abs(before)
        lo, hi = _inner(dom, t0)[np.broadcast_to(m, _inner(dom, t0).shape)].min(), _inner(dom, t0)[np.broadcast_to(m, _inner(dom, t0).shape)].max()
        wet = _inner(dom, t1)[np.broadcast_to(m, _inner(dom, t1).shape)]
        assert wet.min() >= lo - 1e-12 * max(1, abs(lo)) and wet.max() <= hi + 1e-12 * max(1, abs(hi))
    assert np.abs(_inner(dom, b["tr"][2]) - 1.0).max() < 1e-13
    assert np.abs(_inner(dom, b["uhr_out"])).max() == 0.0 and np.abs(_inner(dom, b["vhr_out"])).max() == 0.0   # all transport was used
    assert not np.array_equal(a["tr"][0], b["tr"][0])


def test_limited_fluxes_need_more_passes_and_max_iter_stops_early(oracle):
    dom, grid, gv, cs, a = synthetic.advect_inputs(36, 28, 6, land_blocks=3, cfl=3.5)
    b = _copy(a)
    nit = oracle.advect_tracer(dom, grid, gv, cs, b)
    assert nit > 2
    c = _copy(a); c["max_iter_in"] = 1
    c["uhr_out"] = np.zeros_like(a["uhtr"]); c["vhr_out"] = np.zeros_like(a["vhtr"])
    assert oracle.advect_tracer(dom, grid, gv, cs, c) == 1
    assert np.abs(_inner(dom, c["uhr_out"])).max() > 0.0          # transport left over
    assert not np.array_equal(b["tr"][0], c["tr"][0])


def test_x_first_follows_first_direction(oracle):
    dom, grid, gv, cs, a = synthetic.advect_inputs(24, 20, 4)
    b = _copy(a); oracle.advect_tracer(dom, grid, gv, cs, b)
    c = _copy(a); c["x_first_in"] = True; oracle.advect_tracer(dom, grid, gv, cs, c)
    d = _copy(a); d["x_first_in"] = False; oracle.advect_tracer(dom, grid, gv, cs, d)
    assert np.array_equal(b["tr"][0], c["tr"][0]) and not np.array_equal(c["tr"][0], d["tr"][0])   # first_direction = 0 -> x first (:144)


CASES = [dict(), dict(scheme=2, land_blocks=4), dict(scheme=1, land_blocks=2, cyclic_y=True), dict(cfl=3.5, land_blocks=3, ntr=5),
         dict(x_first_in=False, cfl=3.0, land_blocks=3), dict(max_iter_in=1, cfl=3.5), dict(halo=6, cfl=3.5, land_blocks=2),
         dict(scheme=2, halo=3, cfl=2.5), dict(ntr=1, cyclic_x=False, land_blocks=1)]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_advect_tracer_bitwise(oracle, ctx_factory, kw):
    dom, grid, gv, cs, a = synthetic.advect_inputs(44, 40, 8, **kw)
    ntr = len(a["tr"])
    a["conc_underflow"] = np.array([0.0, 0.0] + [1.0e-3] * (ntr - 2))[:ntr]
    a["advect_scheme"] = [-1, 0, 2, 1, -1][:ntr]
    a["uhr_out"] = np.zeros_like(a["uhtr"]); a["vhr_out"] = np.zeros_like(a["vhtr"])
    ref = _copy(a); oracle.advect_tracer(dom, grid, gv, cs, ref)
    got = _copy(a)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    n0 = ctx.launches
    ctx.advect_tracer(cs, got)
    assert ctx.launches > n0
    for m in range(ntr):
        assert np.array_equal(_inner(dom, ref["tr"][m]).view(np.int64), _inner(dom, got["tr"][m]).view(np.int64)), (m, kw)
    for k in ("uhr_out", "vhr_out"):
        assert np.array_equal(_inner(dom, ref[k]).view(np.int64), _inner(dom, got[k]).view(np.int64)), k


@pytest.mark.gpu
def test_advect_tracer_vol_prev_and_errors(oracle, ctx_factory):
    from mom6_b200.api import Mom6cuError
    dom, grid, gv, cs, a = synthetic.advect_inputs(33, 19, 3, land_blocks=2, cfl=3.2)
    a["vol_prev"] = _vol0(dom, grid, a) * 1.01; a["update_vol_prev"] = True
    ref = _copy(a); oracle.advect_tracer(dom, grid, gv, cs, ref)
    got = _copy(a)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    ctx.advect_tracer(cs, got)
    assert np.array_equal(_inner(dom, ref["vol_prev"]), _inner(dom, got["vol_prev"]))
    assert np.array_equal(_inner(dom, ref["tr"][0]), _inner(dom, got["tr"][0]))
    with pytest.raises(Mom6cuError):
        ctx.advect_tracer(dict(cs, default_advect_scheme=7), _copy(a))
    dom2, grid2, gv2, cs2, a2 = synthetic.advect_inputs(20, 20, 2, halo=2, scheme=2)
    ctx2 = ctx_factory(dom2); ctx2.set_grid(grid2); ctx2.set_vgrid(gv2)
    with pytest.raises(Mom6cuError):
        ctx2.advect_tracer(cs2, a2)        # stencil is wider than the halo (:172)
