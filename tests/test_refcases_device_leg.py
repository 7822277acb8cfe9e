"""The device leg of the reference-digest cases (refcases.run_device), exercised WITHOUT a device: a stand-in for
mom6_b200.api.Context whose every method calls the oracle.  Since the oracle reproduces every digest, run_device through the
stand-in must reproduce them too -- which checks, for all cases including those nobody has yet run on a B200 (refcases.LATE), what is
Python on the device leg: the order of the calls, which dictionaries are copied and which are updated in place, the names the
outputs are collected under.  What the GPU adds is the kernels behind the same calls (tests/test_reference_golden.py -m gpu)."""
import json
import os

import numpy as np
import pytest

import refcases

HERE = os.path.dirname(os.path.abspath(__file__))
WANT = json.load(open(os.path.join(HERE, "golden", "reference_f90_digests.json")))


class OracleContext:
    """the methods of mom6_b200.api.Context that refcases.run_device calls, answered by oracle.pyoracle"""

    def __init__(self, oracle, dom):
        self.o, self.dom, self.css, self.coefs = oracle, dom, {}, None

    def set_grid(self, grid): self.grid = grid
    def set_vgrid(self, gv): self.gv = gv
    def close(self): pass
    def set_cs_continuity(self, cs): self.css["continuity"] = cs
    def set_cs_coriolisadv(self, cs): self.css["coriolisadv"] = cs
    def set_cs_hor_visc(self, cs): self.css["hor_visc"] = cs
    def set_cs_pressureforce(self, cs): self.css["pressureforce"] = cs
    def set_cs_vertvisc(self, cs): self.css["vertvisc"] = cs
    def _g(self): return self.dom, self.grid, self.gv
    def continuity(self, a): self.o.continuity(*self._g(), self.css["continuity"], a)
    def coradcalc(self, a): self.o.coradcalc(*self._g(), self.css["coriolisadv"], a)
    def horizontal_viscosity(self, a): self.o.horizontal_viscosity(*self._g(), self.css["hor_visc"], a)
    def pressure_force(self, a): self.o.pressure_force(*self._g(), self.css["pressureforce"], a)
    def btstep(self, cs, a): self.o.btstep(*self._g(), cs, a)
    def advect_tracer(self, cs, a): self.o.advect_tracer(*self._g(), cs, a)
    def thickness_diffuse(self, cs, a): self.o.thickness_diffuse(*self._g(), cs, a)
    def tracer_hordiff(self, cs, a): self.o.tracer_hordiff(*self._g(), cs, a)
    def mixedlayer_restrat(self, cs, *a): self.o.mixedlayer_restrat(*self._g(), cs, *a)
    def step_dyn_split_rk2(self, cs, a): self.o.step_dyn_split_rk2(*self._g(), self.css, cs, a)
    def ale_regridding_and_remapping(self, ale, a, dyn_cs=None): self.o.ale_regridding_and_remapping(*self._g(), ale, a, dyn_cs=dyn_cs)
    def btcalc(self, a): return self.o.btcalc(*self._g(), a)
    def bt_mass_source(self, h, eta, sc, e): return self.o.bt_mass_source(*self._g(), h, eta, sc, e)
    def set_dtbt(self, a): return self.o.set_dtbt(*self._g(), a)
    def write_energy(self, cs, u, v, h, T=None, S=None): return self.o.write_energy(*self._g(), cs, u, v, h, T, S)
    def ocean_stats_line(self, cs, e, n, reday): return self.o.ocean_stats_line(cs, e, n, reday)

    def chksum(self, array, stagger=0, nk=None, haloshift=0, symmetric=False, omit_corners=False, scale=1.0, stats=False):
        return self.o.chksum(self.dom, array, stagger, haloshift, symmetric, omit_corners, scale, stats=stats)

    # the vertvisc family keeps its coefficients on the device between the calls: here they are held by the stand-in
    def vertvisc_coef(self, coef):
        self.coefs = refcases._coefs(self.dom, int(self.dom.nk))
        self.o.vertvisc_coef(*self._g(), self.css["vertvisc"], coef, *self.coefs)

    def vertvisc_get_coef(self, a_u, a_v, h_u, h_v):
        for dst, src in zip((a_u, a_v, h_u, h_v), self.coefs):
            dst[...] = src

    def vertvisc_remnant(self, vru, vrv, dt, Ray_u, Ray_v):
        self.o.vertvisc_remnant(self.dom, self.grid, self.css["vertvisc"], vru, vrv, dt, *self.coefs, Ray_u, Ray_v)

    def vertvisc(self, s):
        self.ntrunc = self.o.vertvisc(*self._g(), self.css["vertvisc"], s, *self.coefs)

    def vertvisc_ntrunc(self):
        return self.ntrunc


@pytest.mark.parametrize("name", sorted(n for n in refcases.CASES if n not in refcases.DEVICE_REFUSES))
def test_device_leg_reproduces_the_digest_through_an_oracle_backed_context(oracle, name):
    got = refcases.run_device(lambda dom, device=0: OracleContext(oracle, dom), name, refcases.build(name))
    assert sorted(got) == WANT[name]["outputs"], name
    assert refcases.digest(got) == WANT[name]["digest"], name


def test_the_stand_in_has_no_method_the_real_context_lacks():
    from mom6_b200.api import Context
    missing = [m for m in vars(OracleContext) if not m.startswith("_") and not hasattr(Context, m)]
    assert not missing, missing
    assert np.all([callable(getattr(Context, m)) for m in vars(OracleContext) if not m.startswith("_")])
