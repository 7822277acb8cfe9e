"""The oracle still reproduces the committed seeded digests (tests/golden/oracle_digests.json, made by
tests/golden/make_digests.py).  Guards the checker itself: the GPU parity tests compare against an oracle that this test
shows to be the one the fixtures were generated with."""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_matches_committed_digests(oracle):
    spec = importlib.util.spec_from_file_location("make_digests", os.path.join(HERE, "golden", "make_digests.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(HERE, "golden", "oracle_digests.json")))
    got = mod.cases()
    assert sorted(got) == sorted(want)
    bad = [k for k in want if want[k] != got[k]]
    assert not bad, bad


import numpy as np
import pytest


@pytest.mark.gpu
def test_device_path_matches_committed_digests(ctx_factory):
    """The same seeded cases through the C ABI on the device, against the committed digests only (no oracle involved)."""
    from mom6_b200 import synthetic, fidx
    spec = importlib.util.spec_from_file_location("make_digests", os.path.join(HERE, "golden", "make_digests.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    dig = mod.dig
    want = json.load(open(os.path.join(HERE, "golden", "oracle_digests.json")))

    def inner(dom, x):
        return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]

    got = {}
    dom, a = synthetic.bt_timeloop_inputs(44, 40, whalo=6, nstep=10, nfilter=3, land_blocks=2)
    ctx = ctx_factory(dom); ctx.btstep_timeloop(a)
    got["btstep_timeloop"] = dig(a["eta"], a["ubt"], a["vbt"], a["uhbtav"], a["vhbtav"])
    dom, grid, gv, cs, a = synthetic.continuity_inputs(44, 40, 8, land_blocks=2)
    ctx = ctx_factory(dom); ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_continuity(cs); ctx.continuity(a)
    got["continuity"] = dig(inner(dom, a["h"]), inner(dom, a["uh"]), inner(dom, a["vh"]))
    dom, grid, gv, cs, a = synthetic.coradcalc_inputs(44, 40, 8, land_blocks=2)
    ctx = ctx_factory(dom); ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_coriolisadv(cs); ctx.coradcalc(a)
    got["coradcalc"] = dig(inner(dom, a["CAu"]), inner(dom, a["CAv"]))
    dom, grid, gv, cs, a = synthetic.hor_visc_inputs(44, 40, 8, land_blocks=2)
    ctx = ctx_factory(dom); ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_hor_visc(cs); ctx.horizontal_viscosity(a)
    got["horizontal_viscosity"] = dig(inner(dom, a["diffu"]), inner(dom, a["diffv"]))
    dom, grid, gv, cs, a = synthetic.pressureforce_inputs(44, 40, 8, land_blocks=2)
    ctx = ctx_factory(dom); ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_pressureforce(cs); ctx.pressure_force(a)
    got["pressure_force"] = dig(inner(dom, a["PFu"]), inner(dom, a["PFv"]), inner(dom, a["pbce"]), inner(dom, a["eta"]))
    dom, grid, gv, cs, coef, sol = synthetic.vertvisc_inputs(44, 40, 8, land_blocks=2)
    z = lambda st, n: np.zeros((n,) + fidx.new(dom, st).a.shape)   # noqa: E731
    a_u, a_v, h_u, h_v = z("u", 9), z("v", 9), z("u", 8), z("v", 8)
    ctx = ctx_factory(dom); ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_vertvisc(cs)
    ctx.vertvisc_coef(coef); ctx.vertvisc_get_coef(a_u, a_v, h_u, h_v); ctx.vertvisc(sol)
    got["vertvisc"] = dig(inner(dom, a_u), inner(dom, h_v), inner(dom, sol["u"]), inner(dom, sol["v"]))
    dom, grid, gv, cs, a = synthetic.advect_inputs(44, 40, 8, land_blocks=2, cfl=3.0)
    ctx = ctx_factory(dom); ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.advect_tracer(cs, a)
    got["advect_tracer"] = dig(*[inner(dom, t) for t in a["tr"]])
    dom, grid, gv, cs, a = synthetic.regrid_inputs(44, 40, 8, land_blocks=2)
    ctx = ctx_factory(dom); ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.ale_regrid(cs, a["h"], a["h_new"], a["dzRegrid"])
    got["ale_regrid"] = dig(inner(dom, a["h_new"]), inner(dom, a["dzRegrid"]))
    dom, grid, cs, a = synthetic.remap_inputs(44, 40, 8, land_blocks=2)
    t = a["tr"][0].copy()
    ctx = ctx_factory(dom); ctx.set_grid(grid); ctx.ale_remap_tracers(cs, a["h_old"], a["h_new"], [t])
    got["ale_remap"] = dig(inner(dom, t))
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(44, 40, 8, whalo=6, land_blocks=2, store_CAu=1)
    ctx = ctx_factory(dom); ctx.set_grid(grid); ctx.set_vgrid(gv)
    ctx.set_cs_continuity(css["continuity"]); ctx.set_cs_coriolisadv(css["coriolisadv"]); ctx.set_cs_hor_visc(css["hor_visc"])
    ctx.set_cs_pressureforce(css["pressureforce"]); ctx.set_cs_vertvisc(css["vertvisc"])
    ctx.step_dyn_split_rk2(cs, a)
    got["step_dyn_split_rk2"] = dig(*[inner(dom, a[k]) for k in ("u_inst", "v_inst", "h", "uh", "vh", "eta_av")], inner(dom, cs["eta"]))
    dom, grid, gv, cs, a = synthetic.mle_inputs(44, 40, 20, land_blocks=2, MLE_MLD_decay_time2=7.776e6, ml_restrat_coef2=0.5)
    ctx = ctx_factory(dom); ctx.set_grid(grid); ctx.set_vgrid(gv)
    ctx.mixedlayer_restrat(cs, a["h"], a["uhtr"], a["vhtr"], a["T"], a["S"], a["ustar"], a["dt"], a["h_MLD"], a["Rd_dx_h"])
    got["mixedlayer_restrat"] = dig(inner(dom, a["h"]), inner(dom, a["uhtr"]), inner(dom, a["vhtr"]), inner(dom, cs["MLD_filtered"]))
    dom, grid, gv, cs, a = synthetic.hordiff_inputs(44, 40, 8, land_blocks=2, KhTr=5.0e4, check_diffusive_CFL=1)
    ctx = ctx_factory(dom); ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.tracer_hordiff(cs, a)
    got["tracer_hordiff"] = dig(*[inner(dom, t) for t in a["tr"]])
    bad = [k for k in want if want[k] != got.get(k)]
    assert not bad, bad
