"""Times thickness_diffuse (csrc/thickdiff.cu) and tracer_hordiff (csrc/hordiff.cu) on resident fields at a given size (default
1440 x 1080 x 75) and prints one JSON line.  Device time = mom6cu_last_kernel_ms (CUDA events on the launching stream).  Inputs: an
OM4-like vertical grid (2 m layers at the surface growing by 6 % per layer), T and S with lateral noise, random transports (timing
only: parity is tests/test_thickness_diffuse.py / tests/test_tracer_hordiff.py).  bench.py runs this in a process of its own."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mom6_b200 import synthetic  # noqa: E402
from mom6_b200.api import Context, make_domain  # noqa: E402


def main():
    ni, nj, nk = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1440,1080,75").split(","))
    device = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(0)
    dom = make_domain(ni, nj, nk=nk, halo=4)
    grid = synthetic.make_grid(dom, 40)
    gv = synthetic.make_vgrid()
    ctx = Context(dom, device)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    shp = (nk, dom.jed, dom.ied)
    dz = 2.0 * 1.06 ** np.arange(nk)
    dz *= 3800.0 / dz.sum()                                   # a 3800 m water column (timing only: it rides above the synthetic seamount)
    zmid = -(np.cumsum(dz) - 0.5 * dz)
    h = ctx.plane("h", np.ascontiguousarray(dz[:, None, None] * (1.0 + 0.05 * rng.random(shp))), "h", False, nk)
    T = ctx.plane("T", np.ascontiguousarray(20.0 * np.exp(zmid / 1000.0)[:, None, None] + 0.5 * rng.random(shp)), "h", False, nk)
    S = ctx.plane("S", np.ascontiguousarray(35.0 + 0.05 * rng.random(shp)), "h", False, nk)
    uhtr = ctx.plane("uhtr", np.ascontiguousarray(1.0e6 * (rng.random((nk, dom.jed, dom.ied + 1)) - 0.5)), "u", False, nk)
    vhtr = ctx.plane("vhtr", np.ascontiguousarray(1.0e6 * (rng.random((nk, dom.jed + 1, dom.ied)) - 0.5)), "v", False, nk)
    out = {"size": [ni, nj, nk], "cells": ni * nj * nk}
    tcs = synthetic.thickness_diffuse_cs()
    targs = dict(h=h, uhtr=uhtr, vhtr=vhtr, T=T, S=S, p_surf=None, dt=900.0, Res_fn_u=None, Res_fn_v=None, uhGM=None, vhGM=None)
    ms = []
    for rep in range(3):
        ctx.thickness_diffuse(tcs, targs)
        ms.append(ctx.last_kernel_ms)
    out["thickness_diffuse_ms"] = ms
    # algorithmic bytes per cell: column pass h, T, S in, e, pres, rsum, h_frac, Tf, Sf out (+ c1 round trip) ~ 11; per direction h, e, pres,
    # rsum, h_frac, Tf, Sf in, hD out, htr r+w ~ 10; update 4  => ~35 doubles
    out["thickness_diffuse_GBps"] = ni * nj * nk * 35 * 8 / (min(ms[1:]) * 1e-3) / 1e9
    hcs = synthetic.hordiff_cs(KhTr=2000.0, check_diffusive_CFL=1)
    hargs = dict(h=h, dt=7200.0, tr=[T, S], conc_underflow=None, Res_fn_h=None, Rd_dx_h=None)
    ms = []
    for rep in range(3):
        n_it = ctx.tracer_hordiff(hcs, hargs)
        ms.append(ctx.last_kernel_ms)
    out["tracer_hordiff_ms"] = ms; out["tracer_hordiff_iterations"] = n_it; out["tracers"] = 2
    out["tracer_hordiff_GBps"] = ni * nj * nk * 2 * n_it * 5 * 8 / (min(ms[1:]) * 1e-3) / 1e9   # per tracer and sweep: h, T in, T out, copy back r+w
    out["launches"] = ctx.launches
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
