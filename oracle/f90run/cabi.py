"""A stand-in for libmom6cu.so behind the Fortran bindings, for executing fortran/bodies/*.inc without a GPU.  TEST INFRASTRUCTURE ONLY.

tests/test_fortran_shims_executed.py installs the shims into a shadow copy of the reference's modules (fortran/install_shims.py),
runs the hooked reference procedure through oracle/f90run and lets the `mom6cu_*` C entry points the bindings call land HERE: each
stub rebuilds the dictionaries of oracle/pyoracle.py from the bind(C) structures the Fortran body filled (field by field, following
the ctypes mirrors of mom6_b200/_lib.py, i.e. include/mom6cu.h) and calls the C++ oracle.  What this checks is the Fortran side: that
every member the reference's control structures hand over arrives in the right field, with the right logical -> int conversion,
array and optional-argument handling -- the part of the boundary that cannot be compiled here."""
import ctypes as C

import numpy as np

from mom6_b200 import _lib
from .rt import FArray, NS, mangle


class Abi:
    def __init__(self, oracle, dom, grid, gv):
        self.oracle, self.dom, self.grid, self.gv = oracle, dom, grid, gv
        self.css = {}
        self.calls = []
        self.error = ""

    # ---- bind(C) structure (an NS filled by the Fortran body) -> dict of numpy arrays / scalars
    def to_dict(self, ns, struct, pairs):
        out = {}
        for name, ctype in struct._fields_:
            v = getattr(ns, mangle(name))
            if ctype is C.c_void_p or isinstance(v, FArray):   # a data pointer (also double* const* and int* members)
                if v is None:
                    out[name] = None
                elif isinstance(v, FArray) and v.kind == "o":     # an array of c_ptr (tracer fields): a list of arrays / None
                    lst = []
                    for e in v.tolist():
                        if isinstance(e, FArray):
                            a = np.ascontiguousarray(e.to_numpy(), dtype=np.float64)
                            pairs.append((e, a))
                            lst.append(a)
                        else:
                            lst.append(None)
                    out[name] = lst
                elif isinstance(v, FArray) and v.kind == "i":
                    out[name] = [int(x) for x in v.tolist()]
                elif isinstance(v, FArray):
                    a = np.ascontiguousarray(v.to_numpy(), dtype=np.float64)
                    pairs.append((v, a))
                    out[name] = a
                else:
                    raise TypeError(f"{struct.__name__}%{name}: c_loc of a {type(v).__name__}")
            elif ctype in (C.c_int, C.c_double):
                if v is None:
                    raise KeyError(f"{struct.__name__}%{name} was not set by the Fortran binding")
                out[name] = int(v) if ctype is C.c_int else float(v)
            else:
                out[name] = v   # pointer to a nested structure: handled by the caller
        return out

    @staticmethod
    def back(pairs):
        for fa, a in pairs:
            fa.assign(FArray.from_numpy(a))

    def stubs(self):
        o, dom, grid, gv = self.oracle, self.dom, self.grid, self.gv
        L = _lib

        def set_cs(key, struct):
            def f(ctx, c):
                pairs = []
                self.css[key] = self.to_dict(c, struct, pairs)
                self.calls.append("set_cs_" + key)
                return 0
            return f

        def run(name, fn):
            def f(*a):
                self.calls.append(name)
                try:
                    return fn(*a) or 0
                except RuntimeError as e:   # the oracle's refusal (rc 3) or failure: what mom6cu_last_error would report
                    self.error = str(e)
                    return 3
            return f

        def continuity(ctx, a):
            pairs = []
            d = self.to_dict(a, L.ContinuityArgs, pairs)
            if d["BT_cont"] is not None:
                d["BT_cont"] = self.to_dict(d["BT_cont"], L.BTCont, pairs)
            o.continuity(dom, grid, gv, self.css["continuity"], d)
            self.back(pairs)

        def coradcalc(ctx, a):
            pairs = []
            o.coradcalc(dom, grid, gv, self.css["coriolisadv"], self.to_dict(a, L.CorAdCalcArgs, pairs))
            self.back(pairs)

        def hor_visc(ctx, a):
            pairs = []
            o.horizontal_viscosity(dom, grid, gv, self.css["hor_visc"], self.to_dict(a, L.HorViscArgs, pairs))
            self.back(pairs)

        def pressure_force(ctx, a):
            pairs = []
            o.pressure_force(dom, grid, gv, self._pgf_cs(), self.to_dict(a, L.PressureForceArgs, pairs))
            self.back(pairs)

        def btstep(ctx, c, a):
            pairs = []
            cs = self.to_dict(c, L.BarotropicCS, pairs)
            d = self.to_dict(a, L.BtstepArgs, pairs)
            d["BT_cont"] = self.to_dict(d["BT_cont"], L.BTCont, pairs)
            o.btstep(dom, grid, gv, cs, d)
            self.back(pairs)

        def step(ctx, c, a):
            pairs = []
            cs = self.to_dict(c, L.DynSplitRK2CS, pairs)
            cs["BT_cont"] = self.to_dict(cs["BT_cont"], L.BTCont, pairs)
            bt = cs["barotropic"]
            cs["barotropic"] = self.to_dict(bt, L.BarotropicCS, pairs)
            d = self.to_dict(a, L.StepDynArgs, pairs)
            css = dict(continuity=self.css["continuity"], coriolisadv=self.css["coriolisadv"], hor_visc=self.css["hor_visc"],
                       pressureforce=self._pgf_cs(), vertvisc=self.css["vertvisc"])
            o.step_dyn_split_rk2(dom, grid, gv, css, cs, d)
            self.back(pairs)
            c.cau_pred_stored, c.dtbt_max = int(cs["CAu_pred_stored"]), float(cs["dtbt_max"])
            bt.dtbt = float(cs["barotropic"]["dtbt"])

        def advect_tracer(ctx, c, a):
            pairs = []
            cs = self.to_dict(c, L.TracerAdvectCS, pairs)
            d = self.to_dict(a, L.AdvectTracerArgs, pairs)
            n = d["ntr"]
            d["tr"], d["advect_scheme"] = d["tr"][:n], d["advect_scheme"][:n]
            d["conc_underflow"] = np.ascontiguousarray(d["conc_underflow"][:n])
            d["x_first_in"] = None if d["x_first_in"] < 0 else d["x_first_in"]
            d["max_iter_in"] = None if d["max_iter_in"] < 0 else d["max_iter_in"]
            o.advect_tracer(dom, grid, gv, cs, d)
            self.back(pairs)

        def mixedlayer_restrat(ctx, c, h, uhtr, vhtr, T, S, ustar, dt, h_MLD, Rd):
            pairs = []
            cs = self.to_dict(c, L.MleCS, pairs)

            def arr(x):
                if x is None:
                    return None
                a = np.ascontiguousarray(x.to_numpy(), dtype=np.float64)
                pairs.append((x, a))
                return a
            o.mixedlayer_restrat(dom, grid, gv, cs, arr(h), arr(uhtr), arr(vhtr), arr(T), arr(S), arr(ustar), float(dt), arr(h_MLD), arr(Rd))
            self.back(pairs)

        def thickness_diffuse(ctx, c, a):
            pairs = []
            o.thickness_diffuse(dom, grid, gv, self.to_dict(c, L.ThicknessDiffuseCS, pairs), self.to_dict(a, L.ThicknessDiffuseArgs, pairs))
            self.back(pairs)

        def tracer_hordiff(ctx, c, a):
            pairs = []
            cs = self.to_dict(c, L.TracerHorDiffCS, pairs)
            d = self.to_dict(a, L.TracerHordiffArgs, pairs)
            n = d["ntr"]
            d["tr"] = d["tr"][:n]
            d["conc_underflow"] = np.ascontiguousarray(d["conc_underflow"][:n])
            for k in ("df_x", "df_y"):
                d[k] = d[k][:n] if d[k] is not None else None
            o.tracer_hordiff(dom, grid, gv, cs, d)
            self.back(pairs)

        def ale(ctx, c, dptr, a):
            pairs = []
            cs = dict(regridCS=self.to_dict(c.regridcs, L.RegriddingCS, pairs), remapCS=self.to_dict(c.remapcs, L.RemappingCS, pairs),
                      vel_remapCS=self.to_dict(c.vel_remapcs, L.RemappingCS, pairs), regrid_time_scale=float(c.regrid_time_scale),
                      remap_uv_using_old_alg=int(c.remap_uv_using_old_alg), do_conv_adj=int(c.do_conv_adj),
                      use_hybgen_unmix=int(c.use_hybgen_unmix), remap_aux_vars=int(c.remap_aux_vars))
            cs["regridCS"]["coordinateResolution"] = np.ascontiguousarray(cs["regridCS"]["coordinateResolution"]).ravel()
            d = self.to_dict(a, L.AleArgs, pairs)
            n = d["ntr"]
            d["tr"] = d["tr"][:n]
            d["conc_underflow"] = np.ascontiguousarray(d["conc_underflow"][:n])
            dyn = None
            if dptr is not None:
                dyn = {k: v for k, v in self.to_dict(dptr, L.DynSplitRK2CS, pairs).items() if k not in ("BT_cont", "barotropic")}
                dyn["BT_cont"] = {k: None for k, _ in L.BTCont._fields_}
                dyn["barotropic"] = self._null_bt()
            o.ale_regridding_and_remapping(dom, grid, gv, cs, d, dyn_cs=dyn)
            self.back(pairs)
            c.regridcs.old_grid_weight = float(cs["regridCS"]["old_grid_weight"])

        def vertvisc_coef(ctx, a):
            pairs = []
            d = self.to_dict(a, L.VertviscCoefArgs, pairs)
            nk = d["h"].shape[0]
            zu, zv = np.zeros((nk + 1,) + d["u"].shape[1:]), np.zeros((nk + 1,) + d["v"].shape[1:])
            self.coef = [zu, zv, np.zeros_like(d["u"]), np.zeros_like(d["v"])]     # the library's resident CS%a_u, a_v, h_u, h_v
            o.vertvisc_coef(dom, grid, gv, self.css["vertvisc"], d, *self.coef)

        def vertvisc_get_coef(ctx, a_u, a_v, h_u, h_v):
            for fa, x in zip((a_u, a_v, h_u, h_v), self.coef):
                if fa is not None:
                    fa.assign(FArray.from_numpy(x))

        def vertvisc(ctx, a):
            pairs = []
            d = self.to_dict(a, L.VertviscArgs, pairs)
            self.ntrunc = getattr(self, "ntrunc", 0) + o.vertvisc(dom, grid, gv, self.css["vertvisc"], d, *self.coef)
            self.back(pairs)

        def vertvisc_remnant(ctx, Ray_u, Ray_v, vru, vrv, dt):
            pairs = []

            def arr(x):
                if x is None:
                    return None
                y = np.ascontiguousarray(x.to_numpy(), dtype=np.float64)
                pairs.append((x, y))
                return y
            o.vertvisc_remnant(dom, grid, self.css["vertvisc"], arr(vru), arr(vrv), float(dt), *self.coef, arr(Ray_u), arr(Ray_v))
            self.back(pairs)

        return {
            "mom6cu_vertvisc_coef": run("vertvisc_coef", vertvisc_coef), "mom6cu_vertvisc_get_coef": run("vertvisc_get_coef", vertvisc_get_coef),
            "mom6cu_vertvisc": run("vertvisc", vertvisc), "mom6cu_vertvisc_remnant": run("vertvisc_remnant", vertvisc_remnant),
            "mom6cu_vertvisc_ntrunc": lambda ctx: getattr(self, "ntrunc", 0),
            "mom6cu_ale_regridding_and_remapping": run("ale_regridding_and_remapping", ale),
            "mom6cu_advect_tracer": run("advect_tracer", advect_tracer),
            "mom6cu_mixedlayer_restrat": run("mixedlayer_restrat", mixedlayer_restrat),
            "mom6cu_thickness_diffuse": run("thickness_diffuse", thickness_diffuse),
            "mom6cu_tracer_hordiff": run("tracer_hordiff", tracer_hordiff),
            "c_loc": lambda x: x, "c_null_ptr": None, "c_associated": lambda p, q=None: p is not None, "c_null_char": "\0",
            "c_int": 4, "c_double": 8, "c_size_t": 8, "c_long_long": 8, "c_char": 1, "c_ptr": None,
            "mom6cu_last_error": lambda ctx, buf, n: 0,
            "mom6cu_set_cs_continuity": set_cs("continuity", L.ContinuityCS),
            "mom6cu_set_cs_coriolisadv": set_cs("coriolisadv", L.CoriolisAdvCS),
            "mom6cu_set_cs_hor_visc": set_cs("hor_visc", L.HorViscCS),
            "mom6cu_set_cs_pressureforce": set_cs("pressureforce", L.PressureForceCS),
            "mom6cu_set_cs_vertvisc": set_cs("vertvisc", L.VertviscCS),
            "mom6cu_continuity": run("continuity", continuity), "mom6cu_coradcalc": run("coradcalc", coradcalc),
            "mom6cu_horizontal_viscosity": run("horizontal_viscosity", hor_visc),
            "mom6cu_pressure_force": run("pressure_force", pressure_force), "mom6cu_btstep": run("btstep", btstep),
            "mom6cu_step_dyn_split_rk2": run("step_dyn_split_rk2", step),
        }

    @staticmethod
    def _null_bt():
        d = {}
        for k, ct in _lib.BarotropicCS._fields_:
            d[k] = None if ct is C.c_void_p else (0 if ct is C.c_int else 0.0)
        return d

    def _pgf_cs(self):
        d = dict(self.css["pressureforce"])
        for k in ("Rlay", "g_prime"):   # GV%Rlay(1:nk), GV%g_prime(1:nk+1): 1-D host arrays
            if d.get(k) is not None:
                d[k] = np.ascontiguousarray(d[k]).ravel()
        return d
