"""The callers of step_MOM_dynamics chained on resident fields the way MOM.F90:1386-1427 chains them -- thickness_diffuse, pass_var(h),
mixedlayer_restrat, pass_var(h) -- with no copy of the state to the host in between, against the same chain of the oracle; and
mom6cu_do_group_pass itself against the single-tile halo fill.  Written after the round's GPU budget was spent (every entry it calls is
verified on a B200, mom6cu_do_group_pass is a thin wrapper of the halo update the step uses), so it is named to sort last: a failure
here cannot mask the verified tests under `-x`."""
import numpy as np
import pytest

from mom6_b200 import fidx, synthetic


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    return x


def _inner(dom, x):
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]


def _fill(dom, a, st, nk=None):
    f = fidx.new(dom, st, nk=nk)
    f.a[...] = a
    fidx.fill_halo(dom, f, st)
    return np.ascontiguousarray(f.a)


@pytest.mark.gpu
@pytest.mark.parametrize("cyc", [(True, False), (True, True), (False, False)])
def test_do_group_pass_matches_the_halo_fill(ctx_factory, cyc):
    dom, grid, gv, cs, a = synthetic.mle_inputs(37, 23, 5, cyclic_x=cyc[0], cyclic_y=cyc[1])
    r = synthetic.rng(99)
    ctx = ctx_factory(dom)
    fields, sts = [], ["h", "u", "v", "q"]
    for st in sts:
        fields.append(np.ascontiguousarray(r.uniform(-1, 1, size=fidx.new(dom, st, nk=5).a.shape)))
    want = [_fill(dom, f, st, nk=5) for f, st in zip(fields, sts)]
    ctx.do_group_pass(fields, sts, 5)
    for f, w, st in zip(fields, want, sts):
        assert np.array_equal(f, w), (st, cyc)


@pytest.mark.gpu
def test_thickness_diffuse_then_mixedlayer_restrat_on_resident_fields(oracle, ctx_factory):
    nk = 20
    dom, grid, gv, tcs, ta = synthetic.thickness_diffuse_inputs(44, 40, nk, land_blocks=2)
    _, _, _, mcs, ma = synthetic.mle_inputs(44, 40, nk, land_blocks=2)
    # the oracle chain (MOM.F90:1388-1427)
    rt = _copy(ta)
    oracle.thickness_diffuse(dom, grid, gv, tcs, rt)
    rt["h"] = _fill(dom, rt["h"], "h", nk)                                           # pass_var(h) :1396
    rm = _copy(mcs)
    oracle.mixedlayer_restrat(dom, grid, gv, rm, rt["h"], rt["uhtr"], rt["vhtr"], rt["T"], rt["S"], ma["ustar"], ma["dt"], ma["h_MLD"], ma["Rd_dx_h"])
    rt["h"] = _fill(dom, rt["h"], "h", nk)                                           # pass_var(h) :1427
    # the same on the device, the state resident throughout
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    P = {k: ctx.plane("chain." + k, ta[k], st, False, nk) for k, st in (("h", "h"), ("uhtr", "u"), ("vhtr", "v"), ("T", "h"), ("S", "h"))}
    ctx.thickness_diffuse(tcs, dict(ta, **P))
    ctx.do_group_pass([P["h"]], ["h"], nk)
    gm = _copy(mcs)
    ctx.mixedlayer_restrat(gm, P["h"], P["uhtr"], P["vhtr"], P["T"], P["S"], ma["ustar"], ma["dt"], ma["h_MLD"], ma["Rd_dx_h"])
    ctx.do_group_pass([P["h"]], ["h"], nk)
    for k in ("h", "uhtr", "vhtr"):
        got = np.zeros_like(ta[k]); P[k].download(got)
        if k == "h":
            assert np.array_equal(rt[k].view(np.int64), got.view(np.int64)), k      # halos included
        else:
            assert np.array_equal(rt[k].view(np.int64), got.view(np.int64)), k
    assert not np.array_equal(_inner(dom, rt["h"]), _inner(dom, ta["h"]))
