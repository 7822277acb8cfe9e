"""Times mixedlayer_restrat (csrc/mle.cu) on resident fields at a given size (default 1440 x 1080 x 75) and prints one JSON line.
Device time = mom6cu_last_kernel_ms (CUDA events on the launching stream).  Inputs: an OM4-like vertical grid (2 m layers at the
surface growing by 6 % per layer), T with lateral noise so that every face carries an overturning, random transports (timing
only: parity is tests/test_mle.py)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mom6_b200 import synthetic  # noqa: E402
from mom6_b200.api import Context, make_domain  # noqa: E402


def main():
    ni, nj, nk = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1440,1080,75").split(","))
    rng = np.random.default_rng(0)
    dom = make_domain(ni, nj, nk=nk, halo=4)
    grid = synthetic.make_grid(dom, 40)
    gv = synthetic.make_vgrid()
    ctx = Context(dom, 0)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    shp = (nk, dom.jed, dom.ied)
    dz = 2.0 * 1.06 ** np.arange(nk)
    h = np.ascontiguousarray(dz[:, None, None] * (1.0 + 0.1 * rng.random(shp)))
    zmid = -(np.cumsum(dz) - 0.5 * dz)
    P = {"h": ctx.plane("h", h, "h", False, nk)}
    del h
    P["T"] = ctx.plane("T", np.ascontiguousarray(20.0 * np.exp(zmid / 1000.0)[:, None, None] + 0.5 * rng.random(shp)), "h", False, nk)
    P["S"] = ctx.plane("S", np.ascontiguousarray(35.0 + 0.05 * rng.random(shp)), "h", False, nk)
    P["uhtr"] = ctx.plane("uhtr", np.ascontiguousarray(1.0e6 * (rng.random((nk, dom.jed, dom.ied + 1)) - 0.5)), "u", False, nk)
    P["vhtr"] = ctx.plane("vhtr", np.ascontiguousarray(1.0e6 * (rng.random((nk, dom.jed + 1, dom.ied)) - 0.5)), "v", False, nk)
    cs, f2 = synthetic.mle_cs_and_forcing(shp[1:])
    f2 = {k: ctx.plane(k, x, "h", False, 1) for k, x in f2.items()}
    for k in ("MLD_filtered", "MLD_filtered_slow"):
        cs[k] = ctx.plane(k, cs[k], "h", False, 1)
    out = {"size": [ni, nj, nk], "cells": ni * nj * nk}
    n0 = ctx.launches
    ms = []
    for rep in range(4):
        ctx.mixedlayer_restrat(cs, P["h"], P["uhtr"], P["vhtr"], P["T"], P["S"], f2["ustar"], 900.0, f2["h_MLD"], f2["Rd_dx_h"])
        ms.append(ctx.last_kernel_ms)
    out["mixedlayer_restrat_ms"] = ms
    out["launches_per_call"] = (ctx.launches - n0) // 4
    # algorithmic bytes per cell: column pass h, h_avail (+ T, S in the mixed layer) ~ 4; per direction h, h_avail, hml, htr r+w = 5; update
    # uhml, vhml, h r+w = 4  => 18 doubles
    out["algorithmic_B_per_cell"] = 144
    out["achieved_GBps"] = ni * nj * nk * 144 / (min(ms[1:]) * 1e-3) / 1e9
    hh = np.zeros(shp); P["h"].download(hh)
    out["h_min_after"] = float(hh[:, dom.jsc - 1:dom.jec, dom.isc - 1:dom.iec].min())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
