"""Regenerates tests/golden/oracle_digests.json: SHA-256 digests of the oracle's outputs on fixed seeded inputs, one entry
per stage of the path.  These are *oracle-generated* pins (the reference holds no static vectors for these stages, SURVEY
8c): they freeze the restatement so that an accidental change to oracle/ or to the synthetic generators is caught, and
they give the GPU parity tests a fixture that travels to the GPU box.  Run from the repo root:
    python tests/golden/make_digests.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from mom6_b200 import synthetic, fidx  # noqa: E402


def dig(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def cases():
    """name -> digest of the oracle outputs (computational domain only, so halo conventions do not matter)."""
    out = {}

    def inner(dom, x):
        return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]

    dom, a = synthetic.bt_timeloop_inputs(44, 40, whalo=6, nstep=10, nfilter=3, land_blocks=2)
    oracle.btstep_timeloop(dom, a)
    out["btstep_timeloop"] = dig(a["eta"], a["ubt"], a["vbt"], a["uhbtav"], a["vhbtav"])
    dom, grid, gv, cs, a = synthetic.continuity_inputs(44, 40, 8, land_blocks=2)
    oracle.continuity(dom, grid, gv, cs, a)
    out["continuity"] = dig(inner(dom, a["h"]), inner(dom, a["uh"]), inner(dom, a["vh"]))
    dom, grid, gv, cs, a = synthetic.coradcalc_inputs(44, 40, 8, land_blocks=2)
    oracle.coradcalc(dom, grid, gv, cs, a)
    out["coradcalc"] = dig(inner(dom, a["CAu"]), inner(dom, a["CAv"]))
    dom, grid, gv, cs, a = synthetic.hor_visc_inputs(44, 40, 8, land_blocks=2)
    oracle.horizontal_viscosity(dom, grid, gv, cs, a)
    out["horizontal_viscosity"] = dig(inner(dom, a["diffu"]), inner(dom, a["diffv"]))
    dom, grid, gv, cs, a = synthetic.pressureforce_inputs(44, 40, 8, land_blocks=2)
    oracle.pressure_force(dom, grid, gv, cs, a)
    out["pressure_force"] = dig(inner(dom, a["PFu"]), inner(dom, a["PFv"]), inner(dom, a["pbce"]), inner(dom, a["eta"]))
    dom, grid, gv, cs, coef, sol = synthetic.vertvisc_inputs(44, 40, 8, land_blocks=2)
    z = lambda st, n: np.zeros((n,) + fidx.new(dom, st).a.shape)   # noqa: E731
    a_u, a_v, h_u, h_v = z("u", 9), z("v", 9), z("u", 8), z("v", 8)
    oracle.vertvisc_coef(dom, grid, gv, cs, coef, a_u, a_v, h_u, h_v)
    oracle.vertvisc(dom, grid, gv, cs, sol, a_u, a_v, h_u, h_v)
    out["vertvisc"] = dig(inner(dom, a_u), inner(dom, h_v), inner(dom, sol["u"]), inner(dom, sol["v"]))
    dom, grid, gv, cs, a = synthetic.advect_inputs(44, 40, 8, land_blocks=2, cfl=3.0)
    oracle.advect_tracer(dom, grid, gv, cs, a)
    out["advect_tracer"] = dig(*[inner(dom, t) for t in a["tr"]])
    dom, grid, gv, cs, a = synthetic.regrid_inputs(44, 40, 8, land_blocks=2)
    oracle.ale_regrid(dom, grid, gv, cs, a["h"], a["h_new"], a["dzRegrid"])
    out["ale_regrid"] = dig(inner(dom, a["h_new"]), inner(dom, a["dzRegrid"]))
    dom, grid, cs, a = synthetic.remap_inputs(44, 40, 8, land_blocks=2)
    t = a["tr"][0].copy()
    oracle.ale_remap_scalar(dom, grid, cs, a["h_old"], a["h_new"], t)
    out["ale_remap"] = dig(inner(dom, t))
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(44, 40, 8, whalo=6, land_blocks=2, store_CAu=1)
    oracle.step_dyn_split_rk2(dom, grid, gv, css, cs, a)
    out["step_dyn_split_rk2"] = dig(*[inner(dom, a[k]) for k in ("u_inst", "v_inst", "h", "uh", "vh", "eta_av")], inner(dom, cs["eta"]))
    dom, grid, gv, cs, a = synthetic.mle_inputs(44, 40, 20, land_blocks=2, MLE_MLD_decay_time2=7.776e6, ml_restrat_coef2=0.5)
    oracle.mixedlayer_restrat(dom, grid, gv, cs, a["h"], a["uhtr"], a["vhtr"], a["T"], a["S"], a["ustar"], a["dt"], a["h_MLD"], a["Rd_dx_h"])
    out["mixedlayer_restrat"] = dig(inner(dom, a["h"]), inner(dom, a["uhtr"]), inner(dom, a["vhtr"]), inner(dom, cs["MLD_filtered"]))
    dom, grid, gv, cs, a = synthetic.hordiff_inputs(44, 40, 8, land_blocks=2, KhTr=5.0e4, check_diffusive_CFL=1)
    oracle.tracer_hordiff(dom, grid, gv, cs, a)
    out["tracer_hordiff"] = dig(*[inner(dom, t) for t in a["tr"]])
    return out


if __name__ == "__main__":
    oracle.build()
    d = cases()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_digests.json"), "w") as f:
        json.dump(d, f, indent=1, sort_keys=True)
    print(json.dumps(d, indent=1, sort_keys=True))
