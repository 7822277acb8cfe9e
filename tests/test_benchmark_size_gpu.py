"""Bitwise GPU-vs-oracle parity on BASELINE.json's own grids (configs[2] benchmark 360x180x75; configs[1] double_gyre 44x40x20):
the whole step_MOM_dyn_split_RK2 and every stage called on its own through the C ABI, at the sizes where index arithmetic,
grid-dimension limits and multi-wave scheduling of the 75-layer kernel variants are those of the bench (round-1 review: the
75-layer variants had only been compared on 36x28 / 52x36 horizontal grids).  The oracle runs its OpenMP path (bit-identical
to its serial path: k- and j-parallel loops only, no parallel floating sums).

MOM6CU_TEST_FULL_SIZE=1 adds one whole-step comparison at 1440x1080x75 (the headline grid; ~60 GB of host memory and a few
minutes of oracle time: run by hand, log kept in profiles/)."""
import os

import numpy as np
import pytest

from mom6_b200 import synthetic

NT = max(1, os.cpu_count() or 1)

STATE = ("u_inst", "v_inst", "h", "uh", "vh", "uhtr", "vhtr", "eta_av")
CSARR = ("CAu", "CAv", "CAu_pred", "CAv_pred", "PFu", "PFv", "diffu", "diffv", "visc_rem_u", "visc_rem_v", "u_accel_bt", "v_accel_bt", "u_av",
         "v_av", "h_av", "pbce", "eta", "eta_PF", "uhbt", "vhbt", "taux_bot", "tauy_bot")


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_copy(v) for v in x]
    return x


def _inner(dom, x):
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]


def _same(dom, r, g):
    return np.array_equal(_inner(dom, r).view(np.int64), _inner(dom, g).view(np.int64))


def _diff(name, r, g):
    return f"{name}: {np.count_nonzero(r != g)} of {r.size} differ, max |d|={np.nanmax(np.abs(r - g))}"


def _setup(ctx_factory, dom, grid, gv, css=None):
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    if css is not None:
        ctx.set_cs_continuity(css["continuity"]); ctx.set_cs_coriolisadv(css["coriolisadv"]); ctx.set_cs_hor_visc(css["hor_visc"])
        ctx.set_cs_pressureforce(css["pressureforce"]); ctx.set_cs_vertvisc(css["vertvisc"])
    return ctx


def _step_case(oracle, ctx_factory, shape, nsteps, pgf=None, **kw):
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(*shape, **kw)
    if pgf:
        css["pressureforce"].update(pgf)
    rcs, ra = _copy(cs), _copy(a)
    gcs, ga = cs, a
    ctx = _setup(ctx_factory, dom, grid, gv, css)
    for step in range(nsteps):
        oracle.step_dyn_split_rk2(dom, grid, gv, css, rcs, ra, nthreads=NT)
        ctx.step_dyn_split_rk2(gcs, ga)
        bad = [k for k in STATE if not _same(dom, ra[k], ga[k])]
        bad += ["CS%" + k for k in CSARR if not _same(dom, rcs[k], gcs[k])]
        bad += ["BT%" + k for k in ("eta_cor", "ubtav", "vbtav", "frhatu", "frhatv") if not _same(dom, rcs["barotropic"][k], gcs["barotropic"][k])]
        bad += ["BT_cont%" + k for k, x in rcs["BT_cont"].items() if x is not None and not _same(dom, x, gcs["BT_cont"][k])]
        assert not bad, (shape, step, bad)
        assert rcs["barotropic"]["dtbt"] == gcs["barotropic"]["dtbt"] and rcs["dtbt_max"] == gcs["dtbt_max"]
    ctx.close()


@pytest.mark.gpu
def test_step_bitwise_double_gyre_size(oracle, ctx_factory):
    """configs[1]: 44x40x20, set_dtbt inside the step (DTBT < 0), 3 consecutive steps."""
    _step_case(oracle, ctx_factory, (44, 40, 20), 3, land_blocks=2, store_CAu=1, calc_dtbt=1)


@pytest.mark.gpu
def test_step_bitwise_benchmark_size(oracle, ctx_factory):
    """configs[2]: 360x180x75, 2 consecutive steps, every state field / CS array / BT_cont member bit for bit."""
    _step_case(oracle, ctx_factory, (360, 180, 75), 2, land_blocks=12, store_CAu=1, calc_dtbt=1)


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", [1, 2])
def test_step_bitwise_benchmark_size_with_pressure_reconstruction(oracle, ctx_factory, scheme):
    """The same with RECONSTRUCT_FOR_PRESSURE (PLM, PPM): the reference's default under ALE and what bench.py runs."""
    _step_case(oracle, ctx_factory, (360, 180, 75), 1, pgf=dict(reconstruct=1, Recon_Scheme=scheme), land_blocks=12, store_CAu=1)


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("MOM6CU_TEST_FULL_SIZE", "0") != "1", reason="headline-grid oracle run: set MOM6CU_TEST_FULL_SIZE=1")
def test_step_bitwise_headline_size(oracle, ctx_factory):
    """configs[3]: 1440x1080x75, the grid and land fraction the bench runs, one whole step."""
    _step_case(oracle, ctx_factory, (1440, 1080, 75), 1, pgf=dict(reconstruct=1, Recon_Scheme=1), land_blocks=40, store_CAu=1)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(with_uhbt=False), dict(first_direction=1, with_BT_cont=False, alias_h=True)])
def test_continuity_benchmark_size(oracle, ctx_factory, kw):
    dom, grid, gv, cs, a = synthetic.continuity_inputs(360, 180, 75, land_blocks=12, **kw)
    ref = _copy(a)
    oracle.continuity(dom, grid, gv, cs, ref, nthreads=NT)
    ctx = _setup(ctx_factory, dom, grid, gv)
    ctx.set_cs_continuity(cs)
    ctx.continuity(a)
    bad = []
    for k, r in ref.items():
        if k in ("u", "v", "hin", "visc_rem_u", "visc_rem_v", "uhbt", "vhbt"):
            continue
        if isinstance(r, np.ndarray) and not np.array_equal(r.view(np.int64), a[k].view(np.int64)):
            bad.append(_diff(k, r, a[k]))
        elif isinstance(r, dict):
            bad += [_diff("BT_cont." + kk, rr, a[k][kk]) for kk, rr in r.items() if rr is not None and not np.array_equal(rr.view(np.int64), a[k][kk].view(np.int64))]
    assert not bad, "; ".join(bad)
    ctx.close()


@pytest.mark.gpu
def test_btstep_benchmark_size(oracle, ctx_factory):
    dom, grid, gv, cs, a = synthetic.btstep_inputs(360, 180, 75, land_blocks=12, whalo=10)
    ra, rcs = _copy(a), _copy(cs)
    oracle.btstep(dom, grid, gv, rcs, ra, nthreads=NT)
    ctx = _setup(ctx_factory, dom, grid, gv)
    ctx.btstep(cs, a)
    bad = [_diff(k, r, a[k]) for k, r in ra.items() if isinstance(r, np.ndarray) and r.dtype == np.float64 and not np.array_equal(r.view(np.int64), a[k].view(np.int64))]
    bad += [_diff("CS%" + k, r, cs[k]) for k, r in rcs.items() if isinstance(r, np.ndarray) and r.dtype == np.float64 and not np.array_equal(r.view(np.int64), cs[k].view(np.int64))]
    assert not bad, "; ".join(bad)
    assert np.abs(ra["accel_layer_u"]).max() > 0
    ctx.close()


@pytest.mark.gpu
def test_coradcalc_hor_visc_pressure_force_benchmark_size(oracle, ctx_factory):
    dom, grid, gv, cs, a = synthetic.coradcalc_inputs(360, 180, 75, land_blocks=12)
    ref = _copy(a)
    oracle.coradcalc(dom, grid, gv, cs, ref, nthreads=NT)
    ctx = _setup(ctx_factory, dom, grid, gv)
    ctx.set_cs_coriolisadv(cs)
    ctx.coradcalc(a)
    bad = [_diff(k, ref[k], a[k]) for k in ("CAu", "CAv") if not np.array_equal(ref[k].view(np.int64), a[k].view(np.int64))]
    ctx.close()

    dom, grid, gv, cs, a = synthetic.hor_visc_inputs(360, 180, 75, land_blocks=12)
    ref = _copy(a)
    oracle.horizontal_viscosity(dom, grid, gv, cs, ref, nthreads=NT)
    ctx = _setup(ctx_factory, dom, grid, gv)
    ctx.set_cs_hor_visc(cs)
    ctx.horizontal_viscosity(a)
    bad += [_diff(k, ref[k], a[k]) for k in ("diffu", "diffv") if not np.array_equal(ref[k].view(np.int64), a[k].view(np.int64))]
    ctx.close()

    dom, grid, gv, cs, a = synthetic.pressureforce_inputs(360, 180, 75, land_blocks=12, MassWghtInterp=1)
    ref = _copy(a)
    oracle.pressure_force(dom, grid, gv, cs, ref, nthreads=NT)
    ctx = _setup(ctx_factory, dom, grid, gv)
    ctx.set_cs_pressureforce(cs)
    ctx.pressure_force(a)
    bad += [_diff(k, ref[k], a[k]) for k in ("PFu", "PFv", "pbce", "eta") if ref.get(k) is not None and not np.array_equal(ref[k].view(np.int64), a[k].view(np.int64))]
    ctx.close()
    assert not bad, "; ".join(bad)


@pytest.mark.gpu
def test_tracer_and_ale_pass_benchmark_size(oracle, ctx_factory):
    """advect_tracer (2 tracers, PPM:H3) and ALE_regridding_and_remapping at 360x180x75."""
    dom, grid, gv, cs, a = synthetic.advect_inputs(360, 180, 75, land_blocks=12, cfl=2.5, scheme=1)
    ref = _copy(a)
    oracle.advect_tracer(dom, grid, gv, cs, ref)
    ctx = _setup(ctx_factory, dom, grid, gv)
    ctx.advect_tracer(cs, a)
    for m in range(len(a["tr"])):
        assert _same(dom, ref["tr"][m], a["tr"][m]), ("advect_tracer", m)
    ctx.close()

    dom, grid, gv, ale, dcs, a = synthetic.ale_chain_inputs(360, 180, 75, land_blocks=12)
    ra, rdcs, rale = _copy(a), _copy(dcs), _copy(ale)
    oracle.ale_regridding_and_remapping(dom, grid, gv, rale, ra, dyn_cs=rdcs)
    ctx = _setup(ctx_factory, dom, grid, gv)
    ctx.ale_regridding_and_remapping(ale, a, dyn_cs=dcs)
    sl = (Ellipsis, slice(dom.jsc - 1, dom.jec), slice(dom.isc - 1, dom.iec))
    for k in ("u", "v", "h"):
        assert np.array_equal(ra[k][sl].view(np.int64), a[k][sl].view(np.int64)), ("ALE", k)
    for m in range(len(a["tr"])):
        assert _same(dom, ra["tr"][m], a["tr"][m]), ("ALE tracer", m)
    ctx.close()
