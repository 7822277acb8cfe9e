"""The three callers of SURVEY 8f row 2 at BASELINE.json's full size (1440 x 1080 x 75), checked through what their algorithms guarantee by
construction (the oracle is too slow there), on one shared synthetic state:

* thickness_diffuse (MOM_thickness_diffuse.F90:600-616, :1533): the new thickness IS max(h - dt*IareaT*div(uhGM, vhGM), Angstrom_H) of the
  returned transports, recomputed with numpy in the reference's operation order and compared bit for bit; the transports of a face sum to
  zero over the column; nothing crosses a land face;
* mixedlayer_restrat (MOM_mixed_layer_restrat.F90:464-627): an overturning -- no net transport through a face, column thickness kept, the
  thickness floor respected;
* tracer_hordiff (MOM_tracer_hor_diff.F90:537-604): the inventory of every layer conserved, a uniform tracer stays uniform.

Written after the round's GPU budget was spent: not yet run on a B200, and named to sort last so that a failure cannot mask verified tests."""
import numpy as np
import pytest

from mom6_b200 import synthetic

NI, NJ, NK = 1440, 1080, 75


@pytest.mark.gpu
def test_callers_full_size_properties(ctx_factory):
    dom, grid, gv, tcs, a = synthetic.thickness_diffuse_inputs(NI, NJ, NK, land_blocks=40, with_GM=True)
    js, is_ = slice(dom.jsc - dom.jsd, dom.jec - dom.jsd + 1), slice(dom.isc - dom.isd, dom.iec - dom.isd + 1)
    j0, j1, i0, i1 = js.start, js.stop, is_.start, is_.stop
    Ia = grid["IareaT"][js, is_]
    wet = grid["mask2dT"][js, is_] > 0
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    dt = a["dt"]

    # ---- thickness_diffuse
    h0 = a["h"][:, js, is_].copy()
    ctx.thickness_diffuse(tcs, a)
    gu, gv_ = np.where(a["uhGM"] != 7.0, a["uhGM"], 0.0), np.where(a["vhGM"] != 7.0, a["vhGM"], 0.0)
    div = (gu[:, js, i0 + 1:i1 + 1] - gu[:, js, i0:i1]) + (gv_[:, j0 + 1:j1 + 1, is_] - gv_[:, j0:j1, is_])
    want = np.maximum(h0 - dt * Ia[None] * div, gv["Angstrom_H"])
    got = a["h"][:, js, is_]
    assert np.array_equal(want.view(np.int64), got.view(np.int64)), f"{np.count_nonzero(want != got)} cells differ"
    scale = np.abs(gu).sum(axis=0).max()
    assert scale > 0 and np.abs(gu.sum(axis=0)).max() < 1e-11 * scale and np.abs(gv_.sum(axis=0)).max() < 1e-11 * scale
    assert (gu[:, grid["mask2dCu"] == 0] == 0).all() and (gv_[:, grid["mask2dCv"] == 0] == 0).all()
    assert np.allclose(got.sum(axis=0)[wet], h0.sum(axis=0)[wet], rtol=1e-12)
    del gu, gv_, div, want, h0
    ctx.do_group_pass([a["h"]], ["h"], NK)                                   # pass_var(h), MOM.F90:1396

    # ---- mixedlayer_restrat on the state thickness_diffuse left
    mcs, f2 = synthetic.mle_cs_and_forcing(a["h"].shape[1:], MLE_MLD_stretch=4.0)
    h1, u1, v1 = a["h"][:, js, is_].copy(), a["uhtr"].copy(), a["vhtr"].copy()
    ctx.mixedlayer_restrat(mcs, a["h"], a["uhtr"], a["vhtr"], a["T"], a["S"], f2["ustar"], dt, f2["h_MLD"], f2["Rd_dx_h"])
    du, dv = (a["uhtr"] - u1) / dt, (a["vhtr"] - v1) / dt
    scale = np.abs(du).sum(axis=0).max()
    assert scale > 0 and np.abs(du.sum(axis=0)).max() < 1e-10 * scale and np.abs(dv.sum(axis=0)).max() < 1e-10 * scale
    got = a["h"][:, js, is_]
    assert np.allclose(got.sum(axis=0)[wet], h1.sum(axis=0)[wet], rtol=1e-12) and got.min() >= 0.5 * gv["Angstrom_H"]
    assert np.abs(got - h1).max() > 0
    del du, dv, u1, v1, h1
    ctx.do_group_pass([a["h"]], ["h"], NK)                                   # pass_var(h), MOM.F90:1427

    # ---- tracer_hordiff of T, S and a uniform tracer on that state
    hcs = synthetic.hordiff_cs(KhTr=2000.0, check_diffusive_CFL=1)
    uni = np.full_like(a["T"], 3.25)
    vol = (a["h"] * grid["areaT"][None])[:, js, is_]
    inv0 = [(vol * t[:, js, is_]).sum(axis=(1, 2)) for t in (a["T"], a["S"])]
    T0 = a["T"][:, js, is_].copy()
    n = ctx.tracer_hordiff(hcs, dict(h=a["h"], dt=8 * dt, tr=[a["T"], a["S"], uni], conc_underflow=None, Res_fn_h=None, Rd_dx_h=None))
    assert n >= 1
    for t, i0_ in zip((a["T"], a["S"]), inv0):
        assert np.allclose((vol * t[:, js, is_]).sum(axis=(1, 2)), i0_, rtol=1e-11)
    assert (uni[:, js, is_] == 3.25).all() and np.abs(a["T"][:, js, is_] - T0).max() > 0
