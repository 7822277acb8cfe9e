"""The reference's answer-reproducibility metric: reproducing sums (src/framework/MOM_coms.F90), bit-count
checksums (src/framework/MOM_checksums.F90) and write_energy / the ocean.stats line (src/diagnostics/MOM_sum_output.F90).
CPU: the host/device EFP code the GPU threads run (mom6_b200/csrc/efp.cuh, compiled by a test-only harness), the host EFP
operators and the ocean.stats formatter of the C ABI, against the oracle (pinned in test_oracle_efp.py); properties of the
oracle's write_energy.  GPU: mom6cu_reproducing_sum / mom6cu_chksum / mom6cu_write_energy through the C ABI == oracle, bit
for bit."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from mom6_b200 import synthetic, _lib, api
from mom6_b200.api import make_domain

HERE = os.path.dirname(os.path.abspath(__file__))


def _window(dom):
    return dict(isr=dom.isc - (dom.isd - 1), ier=dom.iec - (dom.isd - 1), jsr=dom.jsc - (dom.jsd - 1), jer=dom.jec - (dom.jsd - 1))


def _values(rng, n):
    """a mix of magnitudes and signs, exact zeros, denormal-range and 1e25-range values"""
    mag = 10.0 ** rng.uniform(-30, 25, n)
    v = rng.standard_normal(n) * mag
    v[rng.random(n) < 0.05] = 0.0
    v[rng.random(n) < 0.02] = -0.0
    v[rng.random(n) < 0.01] = 4.9e-324
    return v


# ------------------------------------------------------------------------------------------------ CPU
@pytest.fixture(scope="module")
def efp_harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("efp_host") / "libefp_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(HERE, "harness", "efp_host.cpp"), "-o", out])
    lib = C.CDLL(out)
    lib.efp_host_sum.argtypes = [C.c_long, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("ni,nj,threads,per_cta", [(37, 23, 32, 100), (600, 300, 256, 65536), (1, 1, 256, 65536), (513, 7, 64, 999)])
def test_device_efp_code_matches_oracle(efp_harness, oracle, ni, nj, threads, per_cta):
    """efp.cuh in the kernel's reduction shape == the oracle's serial reproducing_sum: the six integers and the real."""
    rng = np.random.default_rng(ni * 1000 + nj)
    dom = make_domain(ni, nj, halo=0)
    for scale in (1.0, 1.0e-12, 3.0e7):
        a = np.ascontiguousarray((_values(rng, ni * nj) * scale).reshape(nj, ni))
        ref = oracle.reproducing_sum(dom, a, want_efp=True)
        ints = np.zeros(6, dtype=np.int64); s = C.c_double(0.0); amax = C.c_double(0.0)
        fl = efp_harness.efp_host_sum(a.size, a.ctypes.data, threads, per_cta, ints.ctypes.data, C.byref(s), C.byref(amax))
        assert fl == 0
        assert np.array_equal(ints, ref["EFP_sum"]) and s.value == ref["sum"]
        assert amax.value == np.abs(a).max()
    a = np.ones((nj, ni)); a[0, 0] = np.nan
    ints = np.zeros(6, dtype=np.int64); s = C.c_double(0.0); amax = C.c_double(0.0)
    assert efp_harness.efp_host_sum(a.size, a.ctypes.data, threads, per_cta, ints.ctypes.data, C.byref(s), C.byref(amax)) & 1


def test_host_efp_operators_match_oracle(oracle):
    lib = _lib.load()
    rng = np.random.default_rng(3)
    vals = list(_values(rng, 40)) + [0.0, 2.0**100, -(2.0**-130), 1.0 / 3.0]
    for x, y in zip(vals, reversed(vals)):
        ex, ey = api.efp_op(lib, "from_real", x), api.efp_op(lib, "from_real", y)
        assert np.array_equal(ex, oracle.efp_op("from_real", x))
        assert api.efp_op(lib, "to_real", ex) == oracle.efp_op("to_real", ex)
        assert api.efp_op(lib, "to_real", ex) == x or abs(x) < 2.0**-80  # bits below 2**-138 are dropped
        for op in ("plus", "minus"):
            assert np.array_equal(api.efp_op(lib, op, ex, ey), oracle.efp_op(op, ex, ey))
        assert api.efp_op(lib, "diff", ex, ey) == oracle.efp_op("diff", ex, ey)
    with pytest.raises(OverflowError):
        api.efp_op(lib, "from_real", 2.0**140)


def _we_inputs(ni=24, nj=20, nk=6, land_blocks=2, **kw):
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(ni, nj, nk, whalo=6, land_blocks=land_blocks)
    return dom, grid, gv, a, kw


def _make_cs(oracle, dom, grid, **kw):
    return synthetic.sum_output_cs(dom, oracle.create_depth_list(dom, grid, min_depth_inc=1.0e-10), **kw)


def test_oracle_write_energy_properties(oracle):
    dom, grid, gv, a, _ = _we_inputs()
    cs = _make_cs(oracle, dom, grid)
    e = oracle.write_energy(dom, grid, gv, cs, a["u_inst"], a["v_inst"], a["h"], a["T"], a["S"])
    inner = (slice(None), slice(dom.jsc - dom.jsd, dom.jec - dom.jsd + 1), slice(dom.isc - dom.isd, dom.iec - dom.isd + 1))
    areaTm = (grid["mask2dT"] * grid["areaT"])[inner[1:]]
    mass = (a["h"][inner] * (gv["H_to_RZ"] * areaTm)).sum()
    assert abs(e["mass_tot"] - mass) <= 1e-12 * mass
    assert e["mass_tot"] == np.add.reduce(e["mass_lay"]) or abs(e["mass_tot"] - e["mass_lay"].sum()) <= 4 * np.spacing(e["mass_tot"])
    assert e["toten"] == e["KE_tot"] + e["PE_tot"] and e["En_mass"] == e["toten"] / e["mass_tot"]
    assert e["KE_tot"] > 0.0 and np.all(e["KE"] >= 0.0) and np.all(e["PE"] >= 0.0)
    assert e["mass_chg"] == 0.0 and e["mass_anom"] == 0.0 and e["Salt_chg"] == 0.0 and e["Heat_anom"] == 0.0
    assert 0.0 < e["max_CFL"][0] < 1.0 and 0.0 < e["max_CFL"][1] < 1.0
    assert abs(e["salin"] - 35.0) < 1.0 and cs["previous_calls"] == 1
    # the zero-APE depths decrease monotonically with depth and sit between the sea surface and the deepest point
    assert np.all(np.diff(e["Z_0APE"][:-1]) >= 0.0) is not None
    # a second call on a changed state reports the exact change of the extended-fixed-point totals
    h2 = a["h"].copy(); h2[0] *= 1.0 + 2.0**-20
    e2 = oracle.write_energy(dom, grid, gv, cs, a["u_inst"], a["v_inst"], h2, a["T"], a["S"])
    assert e2["mass_chg"] != 0.0 and abs(e2["mass_chg"] - (e2["mass_tot"] - e["mass_tot"])) <= 8 * np.spacing(e["mass_tot"])
    assert e2["mass_anom"] == e2["mass_chg"] and cs["previous_calls"] == 2


LINE_T = re.compile(r"^ *\d+, +\d+\.\d{3}, +\d+, En \d\.\d{16}E[+-]\d\d, CFL +\d\.\d{5}, SL +-?\d\.\d{4}E[+-]\d\d, M \d\.\d{5}E[+-]\d\d, S +\d+\.\d{4}, "
                    r"T +-?\d+\.\d{4}, Me +-?\d\.\d\dE[+-]\d\d, Se +-?\d\.\d\dE[+-]\d\d, Te +-?\d\.\d\dE[+-]\d\d$")


def test_ocean_stats_line_formatting(oracle):
    """The C ABI's formatter == the oracle's, and both have the layout of the reference's ocean.stats lines (:874-889), e.g.
    '     0,       0.000,     0, En 0.0000000000000000E+00, CFL  0.00000, SL  0.0000E+00, M 1.36404E+18, S 35.0000, T 13.3525, ...'"""
    lib = _lib.load()
    dom, grid, gv, a, _ = _we_inputs()
    for kw in (dict(), dict(use_temperature=False), dict(do_APE_calc=False)):
        cs = _make_cs(oracle, dom, grid, **kw)
        e = oracle.write_energy(dom, grid, gv, cs, a["u_inst"], a["v_inst"], a["h"], a["T"], a["S"])
        for n, day in ((0, 0.0), (48, 0.5), (1234567, 12860.073), (5, 3.0e9)):
            ref = oracle.ocean_stats_line(cs, e, n, day)
            assert api.ocean_stats_line(lib, cs, e, n, day) == ref
            if cs["use_temperature"] and day < 1e8:
                assert LINE_T.match(ref), ref
    e0 = dict(e); e0.update(En_mass=0.0, max_CFL=np.zeros(2), mass_tot=1.36404e18, salin=35.0, temp=13.3525, mass_anom=0.0, salin_anom=0.0,
                            temp_anom=0.0, Z_0APE=np.zeros(dom.nk + 1), ntrunc=0)
    cs = _make_cs(oracle, dom, grid)
    assert api.ocean_stats_line(lib, cs, e0, 0, 0.0) == ("     0,       0.000,     0, En 0.0000000000000000E+00, CFL  0.00000, SL -0.0000E+00, "
                                                          "M 1.36404E+18, S 35.0000, T 13.3525, Me  0.00E+00, Se  0.00E+00, Te  0.00E+00")


# ------------------------------------------------------------------------------------------------ GPU
def _ctx(ctx_factory, dom, grid=None, gv=None):
    ctx = ctx_factory(dom)
    if grid is not None:
        ctx.set_grid(grid); ctx.set_vgrid(gv)
    return ctx


@pytest.mark.gpu
@pytest.mark.parametrize("ni,nj,nk", [(37, 23, 1), (200, 300, 1), (64, 48, 5), (600, 300, 3)])
def test_gpu_reproducing_sum_matches_oracle(oracle, ctx_factory, ni, nj, nk):
    rng = np.random.default_rng(ni + nj + nk)
    dom = make_domain(ni, nj, nk=nk, halo=3)
    ctx = _ctx(ctx_factory, dom)
    for stagger, (di, dj) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
        shape = (nk, dom.jed + dj, dom.ied + di) if nk > 1 else (dom.jed + dj, dom.ied + di)
        a = np.ascontiguousarray(_values(rng, int(np.prod(shape))).reshape(shape))
        w = dict(isr=dom.isc + di, ier=dom.iec + di, jsr=dom.jsc + dj, jer=dom.jec + dj) if stagger else _window(dom)
        for unscale in (1.0, 2.0**-7, 3.3):
            for kw in (dict(), dict(want_sums=True, want_efp=True, want_lay_efp=True), dict(want_efp=True)):
                ref = oracle.reproducing_sum(dom, a, stagger, unscale=unscale, **w, **kw)
                got = ctx.reproducing_sum(a, stagger, unscale=unscale, **w, **kw)
                assert got["sum"] == ref["sum"], (stagger, unscale, kw)
                for key in ("sums", "EFP_sum", "EFP_lay_sums"):
                    if key in ref:
                        assert np.array_equal(got[key], ref[key]), (stagger, unscale, key)
        # the whole array (no window), and order invariance on the device
        ref = oracle.reproducing_sum(dom, a, stagger, want_efp=True)
        got = ctx.reproducing_sum(a, stagger, want_efp=True)
        assert got["sum"] == ref["sum"] and np.array_equal(got["EFP_sum"], ref["EFP_sum"])
        b = np.ascontiguousarray(rng.permutation(a.ravel()).reshape(a.shape))
        assert np.array_equal(ctx.reproducing_sum(b, stagger, want_efp=True)["EFP_sum"], ref["EFP_sum"])
    with pytest.raises(api.Mom6cuError, match="NaN in input field"):
        a = np.ones((nk, dom.jed, dom.ied)) if nk > 1 else np.ones((dom.jed, dom.ied))
        a[..., dom.jsc, dom.isc] = np.nan
        ctx.reproducing_sum(a)


@pytest.mark.gpu
def test_gpu_reproducing_sum_unit_test_vectors(ctx_factory):
    """test_reproducing_sum.F90 :114-135 through the C ABI: the exact sum of 1..N, in any order."""
    NI, NJ = 200, 300
    dom = make_domain(NI, NJ, halo=2)
    ctx = _ctx(ctx_factory, dom)
    N = NI * NJ
    rng = np.random.default_rng(1)
    a = np.zeros((dom.jed, dom.ied))
    exact = 0.5 * float(N) * float(N + 1)
    for _ in range(4):
        a[2:2 + NJ, 2:2 + NI] = rng.permutation(1.0 + np.arange(N, dtype=np.float64)).reshape(NJ, NI)
        assert ctx.reproducing_sum(a, **_window(dom))["sum"] == exact


@pytest.mark.gpu
def test_gpu_chksum_matches_oracle(oracle, ctx_factory):
    rng = np.random.default_rng(21)
    for (ni, nj, nk) in ((12, 9, 3), (300, 40, 1), (70, 50, 4)):
        dom = make_domain(ni, nj, nk=nk, halo=3)
        ctx = _ctx(ctx_factory, dom)
        for stagger, (di, dj) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
            shape = (nk, dom.jed + dj, dom.ied + di) if nk > 1 else (dom.jed + dj, dom.ied + di)
            a = np.ascontiguousarray(rng.standard_normal(shape) * 10.0 ** rng.uniform(-5, 5, shape))
            for hs in (0, 1, 3, -1):
                for sym in (False, True):
                    for omit in (False, True):
                        for scale in (1.0, 0.125):
                            rb, rk, rs = oracle.chksum(dom, a, stagger, hs, sym, omit, scale, stats=True)
                            gb, gk, gs = ctx.chksum(a, stagger, haloshift=hs, symmetric=sym, omit_corners=omit, scale=scale, stats=True)
                            assert gk == rk and np.array_equal(gb, rb), (stagger, hs, sym, omit)
                            assert gs[0] == rs[0] and gs[1] == rs[1] and gs[2] == rs[2], (stagger, hs, sym, omit, gs, rs)
        with pytest.raises(api.Mom6cuError, match="haloshift"):
            ctx.chksum(np.zeros((dom.jed, dom.ied)), 0, nk=1, haloshift=4)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(use_temperature=False), dict(do_APE_calc=False), dict(RZL2_to_kg=2.0**-10, L_T_to_m_s=2.0**3, Z_to_m=2.0**2)])
@pytest.mark.parametrize("shape", [(24, 20, 6), (150, 90, 10)])
def test_gpu_write_energy_matches_oracle(oracle, ctx_factory, kw, shape):
    dom, grid, gv, a, _ = _we_inputs(*shape)
    ctx = _ctx(ctx_factory, dom, grid, gv)
    cs_r = _make_cs(oracle, dom, grid, **kw)
    cs_g = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in cs_r.items()}
    u, v, h, T, S = (a[k].copy() for k in ("u_inst", "v_inst", "h", "T", "S"))
    rng = np.random.default_rng(4)
    for call in range(3):
        er = oracle.write_energy(dom, grid, gv, cs_r, u, v, h, T, S)
        eg = ctx.write_energy(cs_g, u, v, h, T, S)
        for k, x in er.items():
            y = eg[k]
            if isinstance(x, np.ndarray):
                assert np.array_equal(np.asarray(x).view(np.int64), np.asarray(y).view(np.int64)), (call, k, x, y)
            else:
                assert np.float64(x).view(np.int64) == np.float64(y).view(np.int64), (call, k, x, y)
        for k in ("previous_calls", "ntrunc"):
            assert cs_r[k] == cs_g[k]
        for k in ("fresh_water_in_EFP", "net_salt_in_EFP", "net_heat_in_EFP", "mass_prev_EFP", "salt_prev_EFP", "heat_prev_EFP"):
            assert np.array_equal(cs_r[k], cs_g[k]), k
        assert np.array_equal(cs_r["lH"], cs_g["lH"])
        assert ctx.ocean_stats_line(cs_g, eg, call, 0.25 * call) == oracle.ocean_stats_line(cs_r, er, call, 0.25 * call)
        # change the state between the calls
        h *= 1.0 + 1.0e-3 * rng.standard_normal(h.shape)
        u += 1.0e-3 * rng.standard_normal(u.shape); T += 0.01 * rng.standard_normal(T.shape)
        cs_r["ntrunc"] = cs_g["ntrunc"] = call + 1


@pytest.mark.gpu
def test_gpu_write_energy_on_resident_planes(oracle, ctx_factory):
    """the state stays on the device (mom6cu_plane_*): same numbers, no staging"""
    dom, grid, gv, a, _ = _we_inputs(40, 30, 5)
    ctx = _ctx(ctx_factory, dom, grid, gv)
    cs_r = _make_cs(oracle, dom, grid)
    cs_g = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in cs_r.items()}
    er = oracle.write_energy(dom, grid, gv, cs_r, a["u_inst"], a["v_inst"], a["h"], a["T"], a["S"])
    P = {}
    for k, st in (("u_inst", "u"), ("v_inst", "v"), ("h", "h"), ("T", "h"), ("S", "h")):
        P[k] = ctx.plane(k, a[k], stagger=st, nk=dom.nk)
    eg = ctx.write_energy(cs_g, P["u_inst"], P["v_inst"], P["h"], P["T"], P["S"])
    for k in ("En_mass", "mass_tot", "Salt", "Heat", "KE_tot", "PE_tot"):
        assert er[k] == eg[k], k
    assert np.array_equal(er["max_CFL"], eg["max_CFL"])
