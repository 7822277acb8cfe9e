"""Times the parity-metric entries (write_energy, chksum, reproducing_sum) and the ALE interface interpolation on resident
fields at a given size (default 1440 x 1080 x 75) and prints one JSON line.  Device times are mom6cu_last_kernel_ms (CUDA
events on the launching stream); inputs are random fields (timing only: parity is tests/test_diag.py)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from mom6_b200 import synthetic  # noqa: E402
from mom6_b200.api import Context, make_domain  # noqa: E402


def main():
    ni, nj, nk = (int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1440,1080,75".split(",")))
    rng = np.random.default_rng(0)
    dom = make_domain(ni, nj, nk=nk, halo=4)
    grid = synthetic.make_grid(dom, 40)
    gv = synthetic.make_vgrid()
    ctx = Context(dom, 0)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    P = {}
    for name, st, shp in (("u", "u", (nk, dom.jed, dom.ied + 1)), ("v", "v", (nk, dom.jed + 1, dom.ied)), ("h", "h", (nk, dom.jed, dom.ied)),
                          ("T", "h", (nk, dom.jed, dom.ied)), ("S", "h", (nk, dom.jed, dom.ied))):
        a = rng.random(shp)
        if name in ("u", "v"):
            a = 0.2 * (a - 0.5)
        elif name == "h":
            a = 1.0 + 50.0 * a
        P[name] = ctx.plane(name, np.ascontiguousarray(a), stagger=st, nk=nk)
        del a
    # a monotone depth list (timing only; the reference builds it once in depth_list_setup)
    depth = np.sort(np.unique(np.round(grid["bathyT"][grid["mask2dT"] > 0], 0)))[::-1].copy()
    area = np.linspace(1.0e9, 3.0e14, len(depth)); vol = np.concatenate(([0.0], np.cumsum(area[:-1] * -np.diff(depth))))
    depth = np.concatenate((depth, depth[-1:])); area = np.concatenate((area, area[-1:])); vol = np.concatenate((vol, vol[-1:] * 1000.0))
    cs = synthetic.sum_output_cs(dom, (depth, area, vol))
    out = {"size": [ni, nj, nk], "cells": ni * nj * nk}
    for rep in range(3):
        t0 = time.perf_counter()
        e = ctx.write_energy(cs, P["u"], P["v"], P["h"], P["T"], P["S"])
        wall = time.perf_counter() - t0
        out["write_energy_ms"] = ctx.last_kernel_ms
        out["write_energy_wall_ms"] = 1e3 * wall
    out["ocean_stats_line"] = ctx.ocean_stats_line(cs, e, 1, 0.0104)
    # algorithmic bytes of write_energy: h 3x (mass, KE, heat/salt) + u, v 2x (KE, CFL) + T, S + h again and PE_pt w+r for the APE pass
    out["write_energy_GBps"] = ni * nj * nk * 8 * (3 + 4 + 2 + 3) / (out["write_energy_ms"] * 1e-3) / 1e9
    for rep in range(2):
        ctx.chksum(P["h"], 0, haloshift=1, stats=True)
        out["chksum_h_haloshift1_stats_ms"] = ctx.last_kernel_ms
        ctx.chksum(P["u"], 1, haloshift=0)
        out["chksum_u_ms"] = ctx.last_kernel_ms
        ctx.reproducing_sum(P["h"], want_sums=True)
        out["reproducing_sum_3d_ms"] = ctx.last_kernel_ms
    out["reproducing_sum_3d_GBps"] = ni * nj * nk * 8 / (out["reproducing_sum_3d_ms"] * 1e-3) / 1e9
    kv = ctx.plane("Kv", np.ascontiguousarray(rng.random((nk + 1, dom.jed, dom.ied))), stagger="h", nk=nk + 1)
    for rep in range(2):
        ctx.ale_remap_interface_vals(P["h"], P["T"], kv)
        out["ale_remap_interface_vals_ms"] = ctx.last_kernel_ms
    out["launches"] = ctx.launches
    print(json.dumps(out))


if __name__ == "__main__":
    main()
