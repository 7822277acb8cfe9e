"""btstep, btcalc, bt_mass_source (src/core/MOM_barotropic.F90:455-2172, :4360-4605, :5243-5296).
CPU: properties of the oracle restatement.  GPU: through the C ABI == oracle, bit for bit."""
import numpy as np
import pytest

from mom6_b200 import synthetic, fidx
from test_oracle_continuity import _comp


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    return x


OUT_A = ("accel_layer_u", "accel_layer_v", "eta_out", "uhbtav", "vhbtav", "etaav")
OUT_CS = ("eta_cor", "ubtav", "vbtav")


def test_btstep_rest_state_stays_at_rest(oracle):
    dom, grid, gv, cs, a = synthetic.btstep_inputs(20, 16, 3, with_uh0=False)
    a, cs = _copy(a), _copy(cs)
    for k in ("U_in", "V_in", "eta_in", "bc_accel_u", "bc_accel_v", "taux", "tauy", "eta_PF_in", "U_Cor", "V_Cor"):
        a[k][...] = 0.0
    cs["eta_cor"][...] = 0.0
    oracle.btstep(dom, grid, gv, cs, a)
    assert np.abs(a["accel_layer_u"]).max() == 0.0 and np.abs(a["eta_out"]).max() == 0.0 and np.abs(a["uhbtav"]).max() == 0.0


def test_btstep_time_mean_transport_closes_eta(oracle):
    """Volume conservation of the barotropic solver: eta_out - eta_in is finite and uhbtav is nonzero for forced flow;
    the time-averaged transports give back the mean eta tendency to within the filter weighting."""
    dom, grid, gv, cs, a = synthetic.btstep_inputs(24, 18, 4, land_blocks=1)
    a, cs = _copy(a), _copy(cs)
    oracle.btstep(dom, grid, gv, cs, a)
    for k in OUT_A:
        assert np.isfinite(a[k]).all(), k
    assert np.abs(_comp(dom, a["uhbtav"], "u")).max() > 0 and np.abs(a["accel_layer_u"]).max() > 0
    # uhbtav vanishes on land faces
    assert np.abs(a["uhbtav"] * (1 - grid["mask2dCu"])).max() == 0.0


def test_btcalc_fractions_sum_to_one(oracle):
    dom, grid, gv, cs, a = synthetic.btstep_inputs(24, 18, 5, land_blocks=2)
    from mom6_b200 import synthetic as syn
    st = syn.dyn_state(dom, grid)
    for scheme in (1, 2, 3):
        args = dict(h=st["h"], h_u=None, h_v=None, frhatu=fidx.new(dom, "u", nk=5).a, frhatv=fidx.new(dom, "v", nk=5).a,
                    bathyT=grid["bathyT"], hvel_scheme=scheme, may_use_default=0)
        oracle.btcalc(dom, grid, gv, args)
        su = _comp(dom, args["frhatu"].sum(axis=0), "u"); m = _comp(dom, grid["mask2dCu"], "u")
        assert np.allclose(su[m > 0], 1.0, atol=1e-12) and np.abs(su[m == 0]).max() == 0.0


CASES = [
    dict(),
    dict(cs=dict(strong_drag=1)),
    dict(cs=dict(Sadourny=0, strong_drag=1), whalo=8),
    dict(with_uh0=False, with_etaav=False, cs=dict(strong_drag=1)),
    dict(cs=dict(bound_BT_corr=1, BT_cont_bounds=1, strong_drag=1), land_blocks=3),
    dict(cs=dict(bound_BT_corr=1, BT_cont_bounds=0, strong_drag=1)),
    dict(with_bot=True, cs=dict(wt_uv_bug=1, visc_rem_u_uh0=1, strong_drag=1)),
    dict(cs=dict(BT_project_velocity=1, strong_drag=1, bebt=0.2), whalo=4, first_direction=1),
    dict(land_blocks=4, cyclic_y=True, cs=dict(strong_drag=1)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_btstep_bitwise(oracle, ctx_factory, kw):
    kw = dict(kw)
    cs_over = kw.pop("cs", {})
    dom, grid, gv, cs, a = synthetic.btstep_inputs(44, 40, 6, **kw, **cs_over)
    ra, rcs = _copy(a), _copy(cs)
    oracle.btstep(dom, grid, gv, rcs, ra)
    ga, gcs = _copy(a), _copy(cs)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    n0 = ctx.launches
    ctx.btstep(gcs, ga)
    assert ctx.launches > n0 + 10
    bad = []
    for k in OUT_A:
        if ra.get(k) is None:
            continue
        if not np.array_equal(ra[k].view(np.int64), ga[k].view(np.int64)):
            bad.append(f"{k}: {np.count_nonzero(ra[k] != ga[k])} of {ra[k].size} differ, max |d|={np.nanmax(np.abs(ra[k] - ga[k]))}")
    for k in OUT_CS:
        if not np.array_equal(rcs[k].view(np.int64), gcs[k].view(np.int64)):
            bad.append(f"CS%{k}: {np.count_nonzero(rcs[k] != gcs[k])} differ, max |d|={np.nanmax(np.abs(rcs[k] - gcs[k]))}")
    assert not bad, "; ".join(bad)
    assert np.abs(ra["accel_layer_u"]).max() > 0


@pytest.mark.gpu
def test_btcalc_and_mass_source_bitwise(oracle, ctx_factory):
    dom, grid, gv, cs, a = synthetic.btstep_inputs(44, 40, 7, land_blocks=3)
    st = synthetic.dyn_state(dom, grid)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    nk = 7
    hu = np.abs(st["u"]) + 1.0; hv = np.abs(st["v"]) + 1.0
    for scheme, huv in ((1, False), (2, False), (3, False), (4, True)):
        def mk():
            return dict(h=st["h"], h_u=hu if huv else None, h_v=hv if huv else None, frhatu=fidx.new(dom, "u", nk=nk).a,
                        frhatv=fidx.new(dom, "v", nk=nk).a, bathyT=grid["bathyT"], hvel_scheme=scheme, may_use_default=0)
        r, g = mk(), mk()
        oracle.btcalc(dom, grid, gv, r)
        ctx.btcalc(g)
        for k in ("frhatu", "frhatv"):
            assert np.array_equal(r[k].view(np.int64), g[k].view(np.int64)), (scheme, k)
    eta = a["eta_in"]
    for set_cor in (1, 0):
        r = cs["eta_cor"].copy(); g = cs["eta_cor"].copy()
        oracle.bt_mass_source(dom, grid, gv, st["h"], eta, set_cor, r)
        ctx.bt_mass_source(st["h"], eta, set_cor, g)
        assert np.array_equal(r.view(np.int64), g.view(np.int64)), set_cor
