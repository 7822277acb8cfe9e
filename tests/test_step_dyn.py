"""step_MOM_dyn_split_RK2 (src/core/MOM_dynamics_split_RK2.F90:294-1205): the whole split-explicit baroclinic step.
CPU: the oracle's step is the reference's sequence of the oracle's stages (sanity + consistency properties).
GPU: the device-resident step through the C ABI == the oracle's step, bit for bit, on every field it touches."""
import numpy as np
import pytest

from mom6_b200 import synthetic

STATE = ("u_inst", "v_inst", "h", "uh", "vh", "uhtr", "vhtr", "eta_av")
CSARR = ("CAu", "CAv", "CAu_pred", "CAv_pred", "PFu", "PFv", "diffu", "diffv", "visc_rem_u", "visc_rem_v", "u_accel_bt", "v_accel_bt", "u_av",
         "v_av", "h_av", "pbce", "eta", "eta_PF", "uhbt", "vhbt", "taux_bot", "tauy_bot")


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    return x


def _inner(dom, x):
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]


def test_one_step_moves_the_state_and_conserves_volume(oracle):
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(32, 24, 6, land_blocks=3)
    cs0, a0 = _copy(cs), _copy(a)
    oracle.step_dyn_split_rk2(dom, grid, gv, css, cs, a)
    for k in STATE + ("PFu", "CAu", "diffu", "u_accel_bt", "visc_rem_u", "eta", "h_av"):
        x = a[k] if k in a else cs[k]
        assert np.isfinite(x).all(), k
        assert not np.array_equal(x, (a0[k] if k in a0 else cs0[k])), k
    area = _inner(dom, grid["areaT"] * grid["mask2dT"])
    v0, v1 = (_inner(dom, a0["h"]) * area).sum(), (_inner(dom, a["h"]) * area).sum()
    assert abs(v1 - v0) <= 1e-6 * v0                                     # volume changes only through the barotropic mass-source correction
    assert np.abs(_inner(dom, a["u_inst"])).max() < 1.0 and np.abs(_inner(dom, a["u_inst"]) - _inner(dom, a0["u_inst"])).max() > 1e-6
    # the time-mean transports were accumulated (:1067-1072)
    assert np.array_equal(_inner(dom, a["uhtr"]), _inner(dom, a["uh"]) * a["dt"])
    # eta follows the corrected thickness to roundoff of the barotropic solver's tolerance
    m = _inner(dom, grid["mask2dT"]) > 0
    col = (_inner(dom, a["h"]).sum(axis=0) - _inner(dom, grid["bathyT"]))[m]
    assert np.abs(col - _inner(dom, cs["eta"])[m]).max() < 1e-3


def test_store_cau_skips_the_first_coradcalc(oracle):
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(24, 18, 4, store_CAu=1)
    oracle.step_dyn_split_rk2(dom, grid, gv, css, cs, a)
    assert cs["CAu_pred_stored"] == 1
    kept = cs["CAu_pred"].copy()
    cs2, a2 = _copy(cs), _copy(a)
    oracle.step_dyn_split_rk2(dom, grid, gv, css, cs2, a2)                # the second step starts from the stored accelerations
    assert cs2["CAu_pred_stored"] == 1 and not np.array_equal(cs2["CAu_pred"], kept)


CASES = [dict(), dict(land_blocks=4, store_CAu=1, begw=0.5), dict(land_blocks=2, split_bottom_stress=1, BT_project_velocity=1),
         dict(land_blocks=3, calc_dtbt=1, store_CAu=1),
         dict(size=(36, 28, 75), land_blocks=2, store_CAu=1)]      # OM4 layer count: the nk-dependent kernel variants of the bench


def _dtbt_args(dom, grid, cs, mode):
    a = dict(pbce=cs["pbce"], gtot_est=0.0, have_gtot_est=0, BT_cont=None, eta=None, SSH_add=0.0, frhatu=cs["barotropic"]["frhatu"],
             frhatv=cs["barotropic"]["frhatv"], bathyT=grid["bathyT"], bebt=0.1, G_extra=0.0, dtbt_fraction=0.98, BT_Coriolis_scale=1.0, Z_ref=0.0,
             Nonlinear_continuity=0)
    if mode == "BT_cont":
        a["BT_cont"] = cs["BT_cont"]
    elif mode == "eta":
        a["eta"] = cs["eta"]; a["Nonlinear_continuity"] = 1
    elif mode == "gtot":
        a["pbce"] = None; a["gtot_est"] = 9.8; a["have_gtot_est"] = 1; a["SSH_add"] = 2.0
    return a


def test_set_dtbt_is_the_gravity_wave_limit(oracle):
    """dtbt_max ~ 1/sqrt(g H (1/dx^2 + 1/dy^2) 2 (1+2 bebt)/2 ...): within a factor of 2 of dx / sqrt(2 g H) for the deepest column."""
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(32, 24, 6)
    cs["pbce"][...] = 9.8
    dtbt, dmax = oracle.set_dtbt(dom, grid, gv, _dtbt_args(dom, grid, cs, "bathy"))
    dx = float(grid["dxT"].min()); H = float(grid["bathyT"].max())
    est = dx / np.sqrt(2.0 * 9.8 * H * 1.2)
    assert 0.5 * est < dmax < 2.0 * est and dtbt == 0.98 * dmax


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["BT_cont", "eta", "bathy", "gtot"])
def test_set_dtbt_bitwise(oracle, ctx_factory, mode):
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(44, 40, 8, land_blocks=3)
    oracle.step_dyn_split_rk2(dom, grid, gv, css, cs, a)          # gives pbce, BT_cont, eta realistic values
    args = _dtbt_args(dom, grid, cs, mode)
    ref = oracle.set_dtbt(dom, grid, gv, args)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    got = ctx.set_dtbt(args)
    assert ref == got and got[0] > 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_step_bitwise(oracle, ctx_factory, kw):
    kw = dict(kw)
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(*kw.pop("size", (44, 40, 8)), **kw)
    rcs, ra = _copy(cs), _copy(a)
    gcs, ga = _copy(cs), _copy(a)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    ctx.set_cs_continuity(css["continuity"]); ctx.set_cs_coriolisadv(css["coriolisadv"]); ctx.set_cs_hor_visc(css["hor_visc"])
    ctx.set_cs_pressureforce(css["pressureforce"]); ctx.set_cs_vertvisc(css["vertvisc"])
    for step in range(2):                                                  # two steps: the second starts from the first's CS
        oracle.step_dyn_split_rk2(dom, grid, gv, css, rcs, ra)
        n0 = ctx.launches
        ctx.step_dyn_split_rk2(gcs, ga)
        assert ctx.launches - n0 > 100
        bad = []
        for k in STATE:
            if not np.array_equal(_inner(dom, ra[k]).view(np.int64), _inner(dom, ga[k]).view(np.int64)):
                bad.append((k, int(np.count_nonzero(_inner(dom, ra[k]) != _inner(dom, ga[k])))))
        for k in CSARR:
            if not np.array_equal(_inner(dom, rcs[k]).view(np.int64), _inner(dom, gcs[k]).view(np.int64)):
                bad.append(("CS%" + k, int(np.count_nonzero(_inner(dom, rcs[k]) != _inner(dom, gcs[k])))))
        for k in ("eta_cor", "ubtav", "vbtav", "frhatu", "frhatv"):
            if not np.array_equal(_inner(dom, rcs["barotropic"][k]).view(np.int64), _inner(dom, gcs["barotropic"][k]).view(np.int64)):
                bad.append(("BT%" + k, 0))
        for k, x in rcs["BT_cont"].items():
            if x is not None and not np.array_equal(_inner(dom, x).view(np.int64), _inner(dom, gcs["BT_cont"][k]).view(np.int64)):
                bad.append(("BT_cont%" + k, 0))
        assert not bad, (step, kw, bad)
        assert rcs["CAu_pred_stored"] == gcs["CAu_pred_stored"]
        assert rcs["barotropic"]["dtbt"] == gcs["barotropic"]["dtbt"] and rcs["dtbt_max"] == gcs["dtbt_max"]


@pytest.mark.gpu
def test_step_bitwise_with_page_locked_host_arrays(oracle, ctx_factory):
    """Page-locked caller arrays take the overlapped staging path (u, v, visc%, forces% uploaded under PressureForce, u and v
    copied back under the last continuity call): same bits as the oracle, over repeated calls."""
    import torch

    def pin(x):
        if isinstance(x, np.ndarray) and x.dtype == np.float64:
            t = torch.empty(x.shape, dtype=torch.float64, pin_memory=True)
            y = t.numpy(); y[...] = x
            keep.append(t)
            return y
        return x

    keep = []
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(44, 40, 8, land_blocks=3, store_CAu=1)
    rcs, ra = _copy(cs), _copy(a)
    gcs = _copy(cs)
    ga = {k: pin(v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    ctx.set_cs_continuity(css["continuity"]); ctx.set_cs_coriolisadv(css["coriolisadv"]); ctx.set_cs_hor_visc(css["hor_visc"])
    ctx.set_cs_pressureforce(css["pressureforce"]); ctx.set_cs_vertvisc(css["vertvisc"])
    for step in range(12):
        oracle.step_dyn_split_rk2(dom, grid, gv, css, rcs, ra)
        ctx.step_dyn_split_rk2(gcs, ga)
        for k in STATE:
            assert np.array_equal(_inner(dom, ra[k]).view(np.int64), _inner(dom, ga[k]).view(np.int64)), (step, k)
        for k in CSARR:
            assert np.array_equal(_inner(dom, rcs[k]).view(np.int64), _inner(dom, gcs[k]).view(np.int64)), (step, "CS%" + k)


@pytest.mark.gpu
def test_step_with_resident_control_structure(oracle, ctx_factory):
    """The intended use: the MOM_dyn_split_RK2_CS arrays stay on the device between steps."""
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(36, 28, 5, land_blocks=2)
    rcs, ra = _copy(cs), _copy(a)
    oracle.step_dyn_split_rk2(dom, grid, gv, css, rcs, ra)
    oracle.step_dyn_split_rk2(dom, grid, gv, css, rcs, ra)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    ctx.set_cs_continuity(css["continuity"]); ctx.set_cs_coriolisadv(css["coriolisadv"]); ctx.set_cs_hor_visc(css["hor_visc"])
    ctx.set_cs_pressureforce(css["pressureforce"]); ctx.set_cs_vertvisc(css["vertvisc"])
    ST = dict(CAu="u", CAv="v", CAu_pred="u", CAv_pred="v", PFu="u", PFv="v", diffu="u", diffv="v", visc_rem_u="u", visc_rem_v="v", u_accel_bt="u",
              v_accel_bt="v", u_av="u", v_av="v", h_av="h", pbce="h", eta="h", eta_PF="h", uhbt="u", vhbt="v", taux_bot="u", tauy_bot="v")
    gcs = dict(cs)
    planes = {}
    for k, st in ST.items():
        planes[k] = ctx.plane("cs." + k, cs[k], st, False, dom.nk if cs[k].ndim == 3 else 1)
        gcs[k] = planes[k]
    ga = _copy(a)
    ctx.step_dyn_split_rk2(gcs, ga)
    ctx.step_dyn_split_rk2(gcs, ga)
    for k in STATE:
        assert np.array_equal(_inner(dom, ra[k]).view(np.int64), _inner(dom, ga[k]).view(np.int64)), k
    out = np.zeros_like(cs["u_av"]); planes["u_av"].download(out)
    assert np.array_equal(_inner(dom, rcs["u_av"]), _inner(dom, out))
