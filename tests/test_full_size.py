"""Size-independent properties at BASELINE.json's full size (OM4_025-shaped 1440 x 1080 x 75, configs[1]), where the
oracle is too slow to be the checker.  Every check is something the reference's algorithm guarantees by construction:

* continuity_PPM (MOM_continuity_PPM.F90:348-421): the new thickness IS max(hin - dt*IareaT*div(uh), h_min) of the returned
  transports -- recomputed here with numpy in the same operation order, compared bit for bit; and the Newton solve of
  zonal/meridional_flux_adjust (:1093-1242) leaves sum_k uh within its own tolerance of uhbt;
* ALE_regrid (MOM_regridding.F90:846-972): column thickness kept, top/bottom interfaces fixed, land untouched;
* remapping_core_h (MOM_remapping.F90:234-335): column integrals conserved, no new extrema;
* advect_tracer (MOM_tracer_advect.F90:53-350): tracer inventory conserved, bounds kept, a uniform tracer stays uniform.
"""
import numpy as np
import pytest

from mom6_b200 import synthetic

NI, NJ, NK = 1440, 1080, 75


def _c(dom):
    return slice(dom.jsc - dom.jsd, dom.jec - dom.jsd + 1), slice(dom.isc - dom.isd, dom.iec - dom.isd + 1)


@pytest.mark.gpu
def test_continuity_full_size_properties(ctx_factory):
    dom, grid, gv, cs, a = synthetic.continuity_inputs(NI, NJ, NK, land_blocks=40)
    hin = a["hin"].copy()
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_continuity(cs)
    n0 = ctx.launches
    ctx.continuity(a)
    assert ctx.launches > n0
    js, is_ = _c(dom)
    j0, j1, i0, i1 = js.start, js.stop, is_.start, is_.stop
    dt, h_min = a["dt"], gv["Angstrom_H"]
    Ia = grid["IareaT"][js, is_]
    # continuity_zonal_convergence then continuity_merdional_convergence (x first), same operation order as :371-378, :409-416
    uh, vh = a["uh"], a["vh"]
    # (the first direction is floored at 0, the second at h_min = Angstrom_H: :168 / :175)
    h_mid = np.maximum(hin[:, js, is_] - (dt * Ia)[None] * (uh[:, js, i0 + 1:i1 + 1] - uh[:, js, i0:i1]), 0.0)
    h_new = np.maximum(h_mid - (dt * Ia)[None] * (vh[:, j0 + 1:j1 + 1, is_] - vh[:, j0:j1, is_]), h_min)
    got = a["h"][:, js, is_]
    assert np.array_equal(h_new.view(np.int64), got.view(np.int64)), f"{np.count_nonzero(h_new != got)} cells differ"
    assert np.isfinite(uh).all() and np.isfinite(vh).all()
    # the barotropic constraint: sum_k uh == uhbt to the solver's own tolerance wherever the CFL bounds did not stop it
    IaF = grid["IareaT"]
    for (flux, bt, mask, Imin) in (
            (uh[:, js, i0:i1 + 1], a["uhbt"][js, i0:i1 + 1], grid["mask2dCu"][js, i0:i1 + 1],
             np.minimum(IaF[js, i0 - 1:i1], IaF[js, i0:i1 + 1])),
            (vh[:, j0:j1 + 1, is_], a["vhbt"][j0:j1 + 1, is_], grid["mask2dCv"][j0:j1 + 1, is_],
             np.minimum(IaF[j0 - 1:j1, is_], IaF[j0:j1 + 1, is_]))):
        err = dt * Imin * np.abs(flux.sum(axis=0) - bt)
        wet = mask > 0
        frac = np.count_nonzero(err[wet] <= cs["tol_eta"] * (1 + 1e-9)) / max(1, np.count_nonzero(wet))
        assert frac > 0.99, f"only {frac:.4f} of the wet faces met tol_eta"
    # BT_cont: face areas are non-negative and finite (set_zonal_BT_cont :1391-1407)
    for k, v in a["BT_cont"].items():
        if k.startswith("FA_"):
            assert np.isfinite(v).all() and v.min() >= 0.0, k
    del a, hin, h_mid, h_new


@pytest.mark.gpu
def test_ale_regrid_remap_full_size_properties(ctx_factory):
    dom, grid, gv, cs, a = synthetic.regrid_inputs(NI, NJ, NK, land_blocks=40)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    ctx.ale_regrid(cs, a["h"], a["h_new"], a["dzRegrid"])
    js, is_ = _c(dom)
    m = grid["mask2dT"][js, is_] > 0
    h, hn, dz = a["h"][:, js, is_], a["h_new"][:, js, is_], a["dzRegrid"][:, js, is_]
    assert np.allclose(hn.sum(axis=0)[m], h.sum(axis=0)[m], rtol=1e-13)           # the column thickness is kept
    assert (dz[0] == 0).all() and np.abs(dz[-1][m]).max() < 1e-9                   # top and bottom interfaces stay
    assert hn[:, m].min() >= 0.0
    assert np.array_equal(hn[:, ~m], h[:, ~m]) and (dz[:, ~m] == 0).all()          # land keeps h
    assert not np.array_equal(hn[:, m], h[:, m])
    # remap a smooth profile from h to h_new with each reconstruction the path supports
    T0 = np.ascontiguousarray(20.0 * np.exp(-np.cumsum(a["h"], axis=0) / 800.0) + 0.01 * np.sin(np.arange(NK))[:, None, None])
    for scheme in (2, 4, 5):   # PLM, PPM_H4, PPM_IH4 (MOM_remapping.F90:89-94)
        rm = dict(remapping_scheme=scheme, boundary_extrapolation=0, force_bounds_in_subcell=0, force_bounds_in_target=1,
                  om4_remap_via_sub_cells=1, answer_date=20190101, h_neglect=1.0e-30, h_neglect_edge=1.0e-30)
        T = T0.copy()
        ctx.ale_remap_tracers(rm, a["h"], a["h_new"], [T])
        t0, t1 = T0[:, js, is_], T[:, js, is_]
        before, after = (t0 * h).sum(axis=0), (t1 * hn).sum(axis=0)
        assert np.allclose(before[m], after[m], rtol=1e-12), scheme                # column integrals conserved
        lo, hi = t0.min(axis=0), t0.max(axis=0)
        assert (t1.min(axis=0)[m] >= (lo - 1e-10)[m]).all() and (t1.max(axis=0)[m] <= (hi + 1e-10)[m]).all(), scheme
        assert np.array_equal(t1[:, ~m], t0[:, ~m])                                # land columns untouched
        assert not np.array_equal(t1[:, m], t0[:, m])


@pytest.mark.gpu
def test_advect_tracer_full_size_properties(ctx_factory):
    dom, grid, gv, cs, a = synthetic.advect_inputs(NI, NJ, NK, land_blocks=40, ntr=3, cfl=0.9)
    a["tr"][2][...] = 1.0
    js, is_ = _c(dom)
    areaT = grid["areaT"]
    div = np.zeros_like(a["h_end"])
    div[:, 1:-1, 1:-1] = (a["uhtr"][:, 1:-1, 2:-1] - a["uhtr"][:, 1:-1, 1:-2]) + (a["vhtr"][:, 2:-1, 1:-1] - a["vhtr"][:, 1:-2, 1:-1])
    v0 = np.maximum(0.0, areaT[None] * a["h_end"] + div)
    v0 = v0 + np.maximum(0.0, 1.0e-13 * v0 - areaT[None] * a["h_end"])            # hprev of :188-195
    del div
    tr0 = [t[:, js, is_].copy() for t in a["tr"]]
    a["vol_prev"] = v0.copy(); a["update_vol_prev"] = True
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    nit = ctx.advect_tracer(cs, a)
    assert 1 <= nit <= 2 * 4 + 1
    m = np.broadcast_to((grid["mask2dT"][js, is_] > 0)[None], tr0[0].shape)
    v0c, v1c = v0[:, js, is_], a["vol_prev"][:, js, is_]
    for t0, t1f in zip(tr0, a["tr"]):
        t1 = t1f[:, js, is_]
        before, after = (t0 * v0c)[m].sum(), (t1 * v1c)[m].sum()
        assert abs(after - before) <= 1e-10 * abs(before)                          # inventory conserved
        lo, hi = t0[m].min(), t0[m].max()
        assert t1[m].min() >= lo - 1e-12 * max(1, abs(lo)) and t1[m].max() <= hi + 1e-12 * max(1, abs(hi))
    assert np.abs(a["tr"][2][:, js, is_] - 1.0).max() < 1e-13                       # a uniform tracer stays uniform
    assert not np.array_equal(tr0[0], a["tr"][0][:, js, is_])


@pytest.mark.gpu
def test_reproducing_sums_and_energy_full_size(ctx_factory):
    """The parity metric at 1440 x 1080 x 75 without the oracle: the reference's own known answers (test_reproducing_sum.F90
    :114-135: sum of 1..N == N(N+1)/2 exactly, in any order), order / layout invariance of the extended-fixed-point integers,
    popcount checksums against numpy, and write_energy's totals against float sums of the same fields."""
    from mom6_b200.api import make_domain
    rng = np.random.default_rng(1)
    dom = make_domain(NI, NJ, nk=4, halo=4)
    grid = synthetic.make_grid(dom, 40)
    gv = synthetic.make_vgrid()
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    js, is_ = _c(dom)
    w = dict(isr=dom.isc - (dom.isd - 1), ier=dom.iec - (dom.isd - 1), jsr=dom.jsc - (dom.jsd - 1), jer=dom.jec - (dom.jsd - 1))
    N = NI * NJ
    a = np.zeros((dom.jed, dom.ied))
    a[js, is_] = 1.0 + np.arange(N, dtype=np.float64).reshape(NJ, NI)
    exact = 0.5 * float(N) * float(N + 1)
    r0 = ctx.reproducing_sum(a, want_efp=True, **w)
    assert r0["sum"] == exact
    a[js, is_] = rng.permutation(a[js, is_].ravel()).reshape(NJ, NI)
    r1 = ctx.reproducing_sum(a, want_efp=True, **w)
    assert r1["sum"] == exact and np.array_equal(r0["EFP_sum"], r1["EFP_sum"])
    # a 3-D field: layer sums, and the same integers whether the rows are summed at once or in two halves
    f = np.ascontiguousarray(rng.standard_normal((4, dom.jed, dom.ied)) * 10.0 ** rng.uniform(-8, 8, (4, dom.jed, dom.ied)))
    full = ctx.reproducing_sum(f, want_sums=True, want_lay_efp=True, **w)
    half = (w["jsr"] + w["jer"]) // 2
    lo = ctx.reproducing_sum(f, want_lay_efp=True, isr=w["isr"], ier=w["ier"], jsr=w["jsr"], jer=half)
    hi = ctx.reproducing_sum(f, want_lay_efp=True, isr=w["isr"], ier=w["ier"], jsr=half + 1, jer=w["jer"])
    from mom6_b200 import api, _lib
    lib = _lib.load()
    for k in range(4):
        both = api.efp_op(lib, "plus", lo["EFP_lay_sums"][k], hi["EFP_lay_sums"][k])
        assert api.efp_op(lib, "to_real", both) == full["sums"][k]
        assert abs(full["sums"][k] - f[k][js, is_].sum()) <= 1e-9 * np.abs(f[k][js, is_]).sum()
    # bit-count checksum of the computational domain == numpy's popcount
    bc, kind, st = ctx.chksum(f, 0, haloshift=0, scale=1.0, stats=True)
    want = int(np.unpackbits(np.ascontiguousarray(np.abs(f[:, js, is_])).view(np.uint8)).sum()) % 1000000000
    assert kind == 1 and bc[0] == want
    assert st[1] == f[:, js, is_].min() and st[2] == f[:, js, is_].max()
    # write_energy: mass and kinetic energy against float sums
    h = np.ascontiguousarray(rng.uniform(0.5, 100.0, (4, dom.jed, dom.ied)))
    u = np.ascontiguousarray(0.1 * rng.standard_normal((4, dom.jed, dom.ied + 1)) * grid["mask2dCu"])
    v = np.ascontiguousarray(0.1 * rng.standard_normal((4, dom.jed + 1, dom.ied)) * grid["mask2dCv"])
    cs = dict(do_APE_calc=0, use_temperature=0, dt_in_T=900.0)
    e = ctx.write_energy(cs, u, v, h)
    areaTm = (grid["mask2dT"] * grid["areaT"])[js, is_]
    mass = (h[:, js, is_] * (gv["H_to_RZ"] * areaTm)).sum()
    assert abs(e["mass_tot"] - mass) <= 1e-11 * mass
    j0, i0 = js.start, is_.start
    ke = ((0.25 * gv["H_to_RZ"] * (areaTm * h[:, js, is_])) *
          ((u[:, js, i0:i0 + NI] ** 2 + u[:, js, i0 + 1:i0 + NI + 1] ** 2) + (v[:, j0:j0 + NJ, is_] ** 2 + v[:, j0 + 1:j0 + NJ + 1, is_] ** 2))).sum()
    assert abs(e["KE_tot"] - ke) <= 1e-11 * ke
    cfl = max((np.abs(u[:, js, i0:i0 + NI + 1] * 900.0) * grid["IdxCu"][js, i0:i0 + NI + 1]).max(),
              (np.abs(v[:, j0:j0 + NJ + 1, is_] * 900.0) * grid["IdyCv"][j0:j0 + NJ + 1, is_]).max())
    assert e["max_CFL"][1] == cfl   # a maximum of identically computed products: exact
    e2 = ctx.write_energy(cs, u, v, h)
    assert e2["mass_chg"] == 0.0 and e2["mass_tot"] == e["mass_tot"] and cs["previous_calls"] == 2
