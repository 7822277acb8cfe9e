"""CPU tests of the oracle's btstep_timeloop against the reference's own invariance properties
(the reference pins this path only differentially, .testing/Makefile:448-470)."""
import numpy as np

from mom6_b200 import synthetic, fidx


def _run(oracle, ni, nj, whalo, **kw):
    dom, args = synthetic.bt_timeloop_inputs(ni, nj, whalo=whalo, **kw)
    oracle.btstep_timeloop(dom, args)
    return dom, args


def _comp(dom, a, key, st, wide):
    ilo, ihi, jlo, jhi = fidx.extent(dom, st, wide)
    su = 1 if st == "u" else 0
    sv = 1 if st == "v" else 0
    return a[key][dom.jsc - sv - jlo: dom.jec - jlo + 1, dom.isc - su - ilo: dom.iec - ilo + 1]


def test_wide_halo_width_invariance(oracle):
    """BT_USE_WIDE_HALOS only changes how often halos are exchanged (MOM_barotropic.F90:2418-2422,
    :2510-2518): answers on the computational domain must not depend on the halo width."""
    res = []
    for wh in (2, 5, 10):
        dom, a = _run(oracle, 40, 32, wh, nstep=14, nfilter=3, land_blocks=2)
        res.append({k: _comp(dom, a, k, st, w).copy() for k, st, w in (
            ("eta", "h", True), ("ubt", "u", True), ("vbt", "v", True), ("eta_wtd", "h", True),
            ("u_accel_bt", "u", True), ("v_accel_bt", "v", True), ("uhbtav", "u", False), ("vhbtav", "v", False),
            ("ubt_wtd", "u", False), ("eta_sum", "h", True))})
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]), k
        assert np.array_equal(res[0][k], res[2][k]), k


def test_periodic_edge_consistency(oracle):
    """With a reentrant x direction the duplicated u column I=isc-1 equals I=iec bit for bit."""
    dom, a = _run(oracle, 36, 24, 6, nstep=10, nfilter=2)
    ilo, ihi, jlo, jhi = fidx.extent(dom, "u", True)
    u = a["ubt"]
    assert np.array_equal(u[dom.jsc - jlo: dom.jec - jlo + 1, dom.isc - 1 - ilo],
                          u[dom.jsc - jlo: dom.jec - jlo + 1, dom.iec - ilo])


def test_mass_conservation_and_activity(oracle):
    """sum(eta/IareaT) changes only through eta_src in a closed/reentrant basin (the corrector is in
    flux form, :2724-2725) -- checked to round-off -- and the fields actually evolve."""
    dom, a0 = synthetic.bt_timeloop_inputs(48, 36, whalo=8, nstep=16, nfilter=0)
    a = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a0.items()}
    oracle.btstep_timeloop(dom, a)
    area = 1.0e8
    m = _comp(dom, a, "IareaT_OBCmask", "h", True) > 0
    e0 = _comp(dom, a0, "eta", "h", True)[m].sum() * area
    e1 = _comp(dom, a, "eta", "h", True)[m].sum() * area
    src = 16 * _comp(dom, a, "eta_src", "h", True)[m].sum() * area
    assert abs((e1 - e0) - src) <= 1e-9 * abs(e0) + 1e-3
    assert np.abs(a["ubt"] - a0["ubt"]).max() > 1e-6
    assert np.isfinite(a["eta"]).all()


def test_linear_vs_btcont_limit(oracle):
    """With FA_EE=FA_E0=FA_W0=FA_WW the BT_cont transport (find_uhbt :4610) is u*FA: identical to the
    linear Datu path (:2639) wherever the velocity stays inside (uBT_EE,uBT_WW)."""
    dom, a = synthetic.bt_timeloop_inputs(30, 22, whalo=4, nstep=6, nfilter=1)
    for key, dat in (("BTCL_u", "Datu"), ("BTCL_v", "Datv")):
        b = a[key]
        fa = a[dat]
        for m in range(4):
            b[..., m] = fa
        b[..., 4] = np.where(fa > 0, 1.0e3, 0.0); b[..., 5] = -b[..., 4]
        b[..., 6] = 0.0; b[..., 7] = 0.0
        b[..., 8] = b[..., 4] * ((1.0 / 3.0) * (2.0 * fa + fa)); b[..., 9] = b[..., 5] * ((1.0 / 3.0) * (2.0 * fa + fa))
    a1 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    a2 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    a2["use_BT_cont"] = 0
    oracle.btstep_timeloop(dom, a1)
    oracle.btstep_timeloop(dom, a2)
    for k in ("eta", "ubt", "vbt"):
        assert np.allclose(a1[k], a2[k], rtol=1e-12, atol=1e-14), k
