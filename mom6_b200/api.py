"""Host-side mirror of the reference interface for the hot path.

Python is only the test/bench harness here (the reference host language is Fortran: see
fortran/ and INTEGRATION.md); every method is a thin marshalling layer over one C-ABI entry
point of include/mom6cu.h and carries the name of the reference subroutine it stands for.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Domain, BtTimeloopArgs, fill_struct


class Mom6cuError(RuntimeError):
    """A FATAL from the library (the Fortran shim maps this to MOM_error(FATAL, msg))."""


class Context:
    """One rank's device context (mom6cu_ctx)."""

    def __init__(self, dom, device=0):
        self.lib = _lib.load()
        self.dom = dom if isinstance(dom, Domain) else make_domain(**dom)
        self._h = C.c_void_p()
        rc = self.lib.mom6cu_create(C.byref(self._h), C.byref(self.dom), device)
        if rc != 0:
            raise Mom6cuError({1: "no CUDA device visible (there is no CPU fallback)",
                               2: "bad domain / device argument"}.get(rc, f"mom6cu_create failed rc={rc}"))

    def close(self):
        if self._h:
            self.lib.mom6cu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc > 0:
            buf = C.create_string_buffer(1024)
            self.lib.mom6cu_last_error(self._h, buf, 1024)
            raise Mom6cuError(f"rc={rc}: {buf.value.decode()}")
        return rc

    @property
    def launches(self):
        return int(self.lib.mom6cu_launch_count(self._h))

    @property
    def last_kernel_ms(self):
        return float(self.lib.mom6cu_last_kernel_ms(self._h))

    STAGES = ("pressure_force", "coradcalc", "vertvisc", "continuity", "btcalc", "btstep", "horizontal_viscosity")

    def last_step_stage_ms(self):
        """Device time each stage took inside the most recent step_dyn_split_rk2 call (summed over its calls in the step)."""
        buf = (C.c_double * len(self.STAGES))()
        n = self.lib.mom6cu_last_step_stage_ms(self._h, buf, len(self.STAGES))
        return {self.STAGES[i]: float(buf[i]) for i in range(n)}

    @property
    def total_kernel_ms(self):
        return float(self.lib.mom6cu_total_kernel_ms(self._h))

    def attach_comm(self, dist):
        """Create the NCCL communicator for halo exchanges; torch.distributed only ferries the
        ncclUniqueId (plumbing)."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            self._check(self.lib.mom6cu_comm_unique_id(idbuf, 128))
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().numpy().tobytes())
        self._check(self.lib.mom6cu_comm_init(self._h, raw, len(raw), rank, world))

    # ---- field residency (mom6cu_plane_*)
    def plane(self, name, host=None, stagger="h", wide=False, nk=1):
        """Allocate a resident plane (and upload `host` into it)."""
        p = self.lib.mom6cu_plane_alloc(self._h, name.encode(), nk)
        if not p:
            raise Mom6cuError("mom6cu_plane_alloc failed")
        pl = Plane(self, p, stagger, wide, nk)
        if host is not None:
            pl.upload(host)
        return pl

    def sync(self):
        self._check(self.lib.mom6cu_sync(self._h))

    # ---- MOM_barotropic.F90:2175 btstep_timeloop
    def btstep_timeloop(self, args, reps=1, download=True):
        keep = []
        st = fill_struct(BtTimeloopArgs(), args, keep)
        return self._check(self.lib.mom6cu_btstep_timeloop_resident(self._h, C.byref(st), reps, 1 if download else 0))


class Plane:
    """A device-resident field (mom6cu_plane_alloc); pass it wherever an array argument is expected."""
    ST = {"h": 0, "u": 1, "v": 2, "q": 3}

    def __init__(self, ctx, ptr, stagger, wide, nk):
        self.ctx, self.ptr, self.stagger, self.wide, self.nk = ctx, int(ptr), stagger, bool(wide), nk

    def upload(self, host):
        self.ctx._check(self.ctx.lib.mom6cu_plane_upload(self.ctx._h, self.ptr, host.ctypes.data, self.ST[self.stagger],
                                                         int(self.wide), self.nk))

    def zero(self):
        self.ctx._check(self.ctx.lib.mom6cu_plane_zero(self.ctx._h, self.ptr, self.nk))

    def download(self, host):
        self.ctx._check(self.ctx.lib.mom6cu_plane_download(self.ctx._h, self.ptr, host.ctypes.data, self.ST[self.stagger],
                                                           int(self.wide), self.nk))
        return host


def _ctx_methods():
    from . import marshal

    def set_grid(self, g):
        """Upload the ocean_grid_type metrics once (src/core/MOM_grid.F90:75-175)."""
        keep = []
        return self._check(self.lib.mom6cu_set_grid(self._h, C.byref(marshal.grid(g, keep))))

    def set_vgrid(self, gv):
        return self._check(self.lib.mom6cu_set_vgrid(self._h, C.byref(marshal.vgrid(gv))))

    def set_cs_continuity(self, cs):
        """continuity_PPM_init's resolved parameters (MOM_continuity_PPM.F90:2674-2754)."""
        return self._check(self.lib.mom6cu_set_cs_continuity(self._h, C.byref(marshal.continuity_cs(cs))))

    def continuity(self, args):
        """continuity_PPM, MOM_continuity_PPM.F90:86."""
        keep = []
        st = marshal.continuity_args(args, keep)
        return self._check(self.lib.mom6cu_continuity(self._h, C.byref(st)))

    def set_unit_scale(self, us=None):
        return self._check(self.lib.mom6cu_set_unit_scale(self._h, C.byref(marshal.unit_scale(us))))

    def set_cs_coriolisadv(self, cs):
        """CoriolisAdv_init's resolved parameters (MOM_CoriolisAdv.F90:1054-1200)."""
        return self._check(self.lib.mom6cu_set_cs_coriolisadv(self._h, C.byref(marshal.coriolisadv_cs(cs))))

    def coradcalc(self, args):
        """CorAdCalc, MOM_CoriolisAdv.F90:125."""
        keep = []
        st = marshal.coradcalc_args(args, keep)
        return self._check(self.lib.mom6cu_coradcalc(self._h, C.byref(st)))

    def set_cs_hor_visc(self, cs):
        """hor_visc_init's resolved parameters and static arrays (MOM_hor_visc.F90:2322-3302)."""
        keep = []
        return self._check(self.lib.mom6cu_set_cs_hor_visc(self._h, C.byref(marshal.hor_visc_cs(cs, keep))))

    def horizontal_viscosity(self, args):
        """horizontal_viscosity, MOM_hor_visc.F90:266."""
        keep = []
        st = marshal.hor_visc_args(args, keep)
        return self._check(self.lib.mom6cu_horizontal_viscosity(self._h, C.byref(st)))

    def btstep(self, cs, args):
        """btstep, MOM_barotropic.F90:455 (cs = the barotropic_CS members it reads / updates)."""
        keep = []
        return self._check(self.lib.mom6cu_btstep(self._h, C.byref(marshal.barotropic_cs(cs, keep)),
                                                  C.byref(marshal.btstep_args(args, keep))))

    def btcalc(self, args):
        """btcalc, MOM_barotropic.F90:4360."""
        keep = []
        return self._check(self.lib.mom6cu_btcalc(self._h, C.byref(marshal.btcalc_args(args, keep))))

    def bt_mass_source(self, h, eta, set_cor, eta_cor):
        """bt_mass_source, MOM_barotropic.F90:5243."""
        ptr = lambda x: x.ptr if isinstance(x, Plane) else x.ctypes.data  # noqa: E731
        return self._check(self.lib.mom6cu_bt_mass_source(self._h, ptr(h), ptr(eta), int(set_cor), ptr(eta_cor)))

    def set_cs_pressureforce(self, cs):
        """PressureForce_FV_init's resolved parameters (MOM_PressureForce_FV.F90:2020+) and the EOS parameters."""
        keep = []
        return self._check(self.lib.mom6cu_set_cs_pressureforce(self._h, C.byref(marshal.pressureforce_cs(cs, keep))))

    def pressure_force(self, args):
        """PressureForce, MOM_PressureForce.F90:40 (-> PressureForce_FV_Bouss)."""
        keep = []
        return self._check(self.lib.mom6cu_pressure_force(self._h, C.byref(marshal.pressureforce_args(args, keep))))

    def _p(x):
        """Address of a numpy array / Plane / None for a plain-pointer argument."""
        if x is None:
            return None
        if isinstance(x, Plane):
            return x.ptr
        if isinstance(x, np.ndarray):
            if not x.flags["C_CONTIGUOUS"] or x.dtype != np.float64:
                raise Mom6cuError("array arguments must be C-contiguous float64")
            return x.ctypes.data
        return int(x)

    def ale_remap_tracers(self, cs, h_old, h_new, tracers, conc_underflow=None):
        """ALE_remap_tracers, src/ALE/MOM_ALE.F90:760 (the column loop :806-826) for a list of h-point fields."""
        n = len(tracers)
        ptrs = (C.c_void_p * max(n, 1))(*[_p(t) for t in tracers])
        cu = None if conc_underflow is None else np.ascontiguousarray(conc_underflow, dtype=np.float64)
        return self._check(self.lib.mom6cu_ale_remap_tracers(self._h, C.byref(marshal.remapping_cs(cs)), _p(h_old), _p(h_new), n, ptrs,
                                                             None if cu is None else cu.ctypes.data))

    def ale_remap_set_h_vel(self, h_new, h_u, h_v):
        """ALE_remap_set_h_vel, MOM_ALE.F90:882."""
        return self._check(self.lib.mom6cu_ale_remap_set_h_vel(self._h, _p(h_new), _p(h_u), _p(h_v)))

    def ale_remap_velocities(self, cs, h_old_u, h_old_v, h_new_u, h_new_v, u, v):
        """ALE_remap_velocities, MOM_ALE.F90:1089."""
        return self._check(self.lib.mom6cu_ale_remap_velocities(self._h, C.byref(marshal.remapping_cs(cs)), _p(h_old_u), _p(h_old_v),
                                                                _p(h_new_u), _p(h_new_v), _p(u), _p(v)))

    def remapping_core_h(self, cs, h0, u0, h1):
        """remapping_core_h, MOM_remapping.F90:234, on a batch of columns: h0, u0 (ncol, n0); h1 (ncol, n1) -> u1 (ncol, n1)."""
        h0, u0, h1 = (np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64) for x in (h0, u0, h1))
        u1 = np.zeros_like(h1)
        self._check(self.lib.mom6cu_remapping_core_h(self._h, C.byref(marshal.remapping_cs(cs)), h0.shape[0], h0.shape[1], _p(h0), _p(u0),
                                                     h1.shape[1], _p(h1), _p(u1)))
        return u1

    def advect_tracer(self, cs, args):
        """advect_tracer, src/tracer/MOM_tracer_advect.F90:53; returns the number of passes made."""
        keep = []
        self._check(self.lib.mom6cu_advect_tracer(self._h, C.byref(marshal.tracer_advect_cs(cs)), C.byref(marshal.advect_tracer_args(args, keep))))
        return int(self.lib.mom6cu_last_iterations(self._h))

    def ale_regrid(self, cs, h, h_new, dzRegrid):
        """ALE_regrid, src/ALE/MOM_ALE.F90:518 (Z* coordinate)."""
        keep = []
        return self._check(self.lib.mom6cu_ale_regrid(self._h, C.byref(marshal.regridding_cs(cs, keep)), _p(h), _p(h_new), _p(dzRegrid)))

    def set_cs_vertvisc(self, cs):
        """vertvisc_init, MOM_vert_friction.F90:2929 (resolved values)."""
        return self._check(self.lib.mom6cu_set_cs_vertvisc(self._h, C.byref(marshal.vertvisc_cs(cs))))

    def vertvisc_coef(self, args):
        """vertvisc_coef, MOM_vert_friction.F90:1357: sets the resident CS%a_u, a_v, h_u, h_v."""
        keep = []
        return self._check(self.lib.mom6cu_vertvisc_coef(self._h, C.byref(marshal.vertvisc_coef_args(args, keep))))

    def vertvisc_get_coef(self, a_u=None, a_v=None, h_u=None, h_v=None):
        return self._check(self.lib.mom6cu_vertvisc_get_coef(self._h, _p(a_u), _p(a_v), _p(h_u), _p(h_v)))

    def vertvisc_ntrunc(self):
        """CS%ntrunc: velocity truncations made by vertvisc_limit_vel (MOM_vert_friction.F90:2926) in this context so far."""
        n = int(self.lib.mom6cu_vertvisc_ntrunc(self._h))
        if n < 0:
            self._check(4)
        return n

    def vertvisc(self, args):
        """vertvisc, MOM_vert_friction.F90:557."""
        keep = []
        return self._check(self.lib.mom6cu_vertvisc(self._h, C.byref(marshal.vertvisc_args(args, keep))))

    def vertvisc_remnant(self, visc_rem_u, visc_rem_v, dt, Ray_u=None, Ray_v=None):
        """vertvisc_remnant, MOM_vert_friction.F90:1229."""
        return self._check(self.lib.mom6cu_vertvisc_remnant(self._h, _p(Ray_u), _p(Ray_v), _p(visc_rem_u), _p(visc_rem_v), float(dt)))

    def step_dyn_split_rk2(self, cs, args):
        """step_MOM_dyn_split_RK2, src/core/MOM_dynamics_split_RK2.F90:294; cs["CAu_pred_stored"] is updated."""
        keep = []
        st = marshal.dyn_split_rk2_cs(cs, keep)
        rc = self._check(self.lib.mom6cu_step_dyn_split_rk2(self._h, C.byref(st), C.byref(marshal.step_dyn_args(args, keep))))
        cs["CAu_pred_stored"] = int(st.CAu_pred_stored)
        cs["dtbt_max"] = float(st.dtbt_max)
        cs["barotropic"]["dtbt"] = float(st.barotropic.contents.dtbt)
        return rc

    def set_dtbt(self, args):
        """set_dtbt, MOM_barotropic.F90:3509; returns (CS%dtbt, CS%dtbt_max)."""
        keep = []
        dtbt, dmax = C.c_double(0.0), C.c_double(0.0)
        self._check(self.lib.mom6cu_set_dtbt(self._h, C.byref(marshal.set_dtbt_args(args, keep)), C.byref(dtbt), C.byref(dmax)))
        return dtbt.value, dmax.value

    def remap_dyn_split_rk2_aux_vars(self, remap_cs, cs, h_old_u, h_old_v, h_new_u, h_new_v):
        """remap_dyn_split_RK2_aux_vars, MOM_dynamics_split_RK2.F90:1302."""
        keep = []
        st = marshal.dyn_split_rk2_cs(cs, keep)
        return self._check(self.lib.mom6cu_remap_dyn_split_rk2_aux_vars(self._h, C.byref(marshal.remapping_cs(remap_cs)), C.byref(st), _p(h_old_u),
                                                                       _p(h_old_v), _p(h_new_u), _p(h_new_v)))

    # ---- the parity metric: reproducing sums, checksums, write_energy (csrc/diag.cu)
    def reproducing_sum(self, array, stagger=0, nk=None, isr=0, ier=0, jsr=0, jer=0, unscale=1.0, only_on_PE=False,
                        want_sums=False, want_efp=False, want_lay_efp=False):
        """reproducing_sum, src/framework/MOM_coms.F90:227 (2-D) / :337 (3-D).  array: (nk, nj, ni) / (nj, ni) host array or a Plane."""
        from ._lib import Efp
        if nk is None:
            nk = array.nk if isinstance(array, Plane) else (1 if array.ndim == 2 else array.shape[0])
        total = C.c_double(0.0)
        sums = np.zeros(nk) if want_sums else None
        e = Efp() if want_efp else None
        le = (Efp * nk)() if want_lay_efp else None
        self._check(self.lib.mom6cu_reproducing_sum(self._h, _p(array), stagger, nk, isr, ier, jsr, jer, float(unscale), int(only_on_PE),
                                                    C.byref(total), _p(sums), C.byref(e) if e is not None else None, le))
        r = {"sum": total.value}
        if want_sums:
            r["sums"] = sums
        if want_efp:
            r["EFP_sum"] = marshal.efp_back(e)
        if want_lay_efp:
            r["EFP_lay_sums"] = np.array([marshal.efp_back(le[k]) for k in range(nk)])
        return r

    def chksum(self, array, stagger=0, nk=None, haloshift=0, symmetric=False, omit_corners=False, scale=1.0, stats=False):
        """hchksum / uvchksum / Bchksum, src/framework/MOM_checksums.F90; returns (bc[5], kind, (mean, min, max) or None)."""
        if nk is None:
            nk = array.nk if isinstance(array, Plane) else (1 if array.ndim == 2 else array.shape[0])
        bc = (C.c_int * 5)()
        kind = C.c_int(0)
        st = (C.c_double * 3)()
        self._check(self.lib.mom6cu_chksum(self._h, _p(array), stagger, nk, haloshift, int(symmetric), int(omit_corners), float(scale), bc,
                                           C.byref(kind), st if stats else None))
        return np.array(list(bc), dtype=np.int64), kind.value, (np.array(list(st)) if stats else None)

    def write_energy(self, cs, u, v, h, T=None, S=None):
        """write_energy, src/diagnostics/MOM_sum_output.F90:321; cs (dict) is updated like the reference's Sum_output_CS."""
        keep = []
        st = marshal.sum_output_cs(cs, keep)
        out, arrs = marshal.energy_out(self.dom.nk, keep)
        self._check(self.lib.mom6cu_write_energy(self._h, C.byref(st), _p(u), _p(v), _p(h), _p(T), _p(S), C.byref(out)))
        marshal.sum_output_cs_back(st, cs)
        return marshal.energy_out_back(out, arrs)

    def ocean_stats_line(self, cs, e, n, reday):
        """The line write_energy appends to ocean.stats (MOM_sum_output.F90:874-902)."""
        return format_ocean_stats_line(self.lib, cs, e, n, reday)

    for f in (reproducing_sum, chksum, write_energy, ocean_stats_line):
        setattr(Context, f.__name__, f)
    # ---- the thermodynamic-cadence ALE pass (csrc/ale.cu)
    def interpolate_column(self, h_src, u_src, h_dest, mask_edges=False):
        """interpolate_column, src/ALE/MOM_remapping.F90:1247, for (ncol, n) arrays of columns."""
        h_src, u_src, h_dest = (np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64) for x in (h_src, u_src, h_dest))
        u_dest = np.zeros((h_dest.shape[0], h_dest.shape[1] + 1))
        self._check(self.lib.mom6cu_interpolate_column(self._h, h_src.shape[0], h_src.shape[1], _p(h_src), _p(u_src), h_dest.shape[1], _p(h_dest),
                                                       _p(u_dest), int(mask_edges)))
        return u_dest

    def ale_remap_interface_vals(self, h_old, h_new, int_val):
        """ALE_remap_interface_vals, src/ALE/MOM_ALE.F90:1303."""
        return self._check(self.lib.mom6cu_ale_remap_interface_vals(self._h, _p(h_old), _p(h_new), _p(int_val)))

    def ale_remap_vertex_vals(self, h_old, h_new, vert_val):
        """ALE_remap_vertex_vals, src/ALE/MOM_ALE.F90:1342."""
        return self._check(self.lib.mom6cu_ale_remap_vertex_vals(self._h, _p(h_old), _p(h_new), _p(vert_val)))

    def ale_regridding_and_remapping(self, cs, args, dyn_cs=None):
        """ALE_regridding_and_remapping, src/core/MOM.F90:1751; cs["regridCS"]["old_grid_weight"] is updated."""
        keep = []
        st = marshal.ale_cs(cs, keep)
        dyn = marshal.dyn_split_rk2_cs(dyn_cs, keep) if dyn_cs is not None else None
        rc = self._check(self.lib.mom6cu_ale_regridding_and_remapping(self._h, C.byref(st), C.byref(dyn) if dyn is not None else None,
                                                                     C.byref(marshal.ale_args(args, keep))))
        cs["regridCS"]["old_grid_weight"] = float(st.regridCS.old_grid_weight)
        return rc

    # ---- mixedlayer_restrat (csrc/mle.cu)
    def mixedlayer_restrat(self, cs, h, uhtr, vhtr, T, S, ustar, dt, h_MLD, Rd_dx_h=None):
        """mixedlayer_restrat -> mixedlayer_restrat_OM4, src/parameterizations/lateral/MOM_mixed_layer_restrat.F90:149 / :189;
        h, uhtr, vhtr and cs["MLD_filtered"], cs["MLD_filtered_slow"] are updated in place."""
        keep = []
        return self._check(self.lib.mom6cu_mixedlayer_restrat(self._h, C.byref(marshal.mle_cs(cs, keep)), _p(h), _p(uhtr), _p(vhtr), _p(T), _p(S),
                                                              _p(ustar), float(dt), _p(h_MLD), _p(Rd_dx_h)))

    def mle_mu(self, sigma, dh):
        """mu(sigma, dh), MOM_mixed_layer_restrat.F90:717, elementwise on the device."""
        sigma, dh = np.broadcast_arrays(np.asarray(sigma, dtype=np.float64), np.asarray(dh, dtype=np.float64))
        sigma, dh = np.ascontiguousarray(sigma).ravel(), np.ascontiguousarray(dh).ravel()
        out = np.zeros_like(sigma)
        self._check(self.lib.mom6cu_mle_mu(self._h, sigma.size, _p(sigma), _p(dh), _p(out)))
        return out

    def tracer_hordiff(self, cs, args):
        """tracer_hordiff, src/tracer/MOM_tracer_hor_diff.F90:119 (the along-surface path); returns the number of iterations made."""
        keep = []
        self._check(self.lib.mom6cu_tracer_hordiff(self._h, C.byref(marshal.tracer_hor_diff_cs(cs)), C.byref(marshal.tracer_hordiff_args(args, keep))))
        return int(self.lib.mom6cu_last_iterations(self._h))

    def thickness_diffuse(self, cs, args):
        """thickness_diffuse, src/parameterizations/lateral/MOM_thickness_diffuse.F90:134; h, uhtr, vhtr (and uhGM, vhGM) updated in place."""
        keep = []
        return self._check(self.lib.mom6cu_thickness_diffuse(self._h, C.byref(marshal.thickness_diffuse_cs(cs)),
                                                             C.byref(marshal.thickness_diffuse_args(args, keep))))

    def do_group_pass(self, fields, staggers, nk):
        """pass_var / pass_vector / do_group_pass (MOM_domains.F90) on host arrays or resident planes; staggers: 'h', 'u', 'v', 'q' per field."""
        n = len(fields)
        ptrs = (C.c_void_p * max(n, 1))(*[_p(f) for f in fields])
        st = (C.c_int * max(n, 1))(*[{"h": 0, "u": 1, "v": 2, "q": 3}[s] for s in staggers])
        return self._check(self.lib.mom6cu_do_group_pass(self._h, n, ptrs, st, int(nk)))

    setattr(Context, "do_group_pass", do_group_pass)
    setattr(Context, "thickness_diffuse", thickness_diffuse)
    setattr(Context, "tracer_hordiff", tracer_hordiff)
    for f in (mixedlayer_restrat, mle_mu):
        setattr(Context, f.__name__, f)
    for f in (interpolate_column, ale_remap_interface_vals, ale_remap_vertex_vals, ale_regridding_and_remapping):
        setattr(Context, f.__name__, f)
    setattr(Context, "remap_dyn_split_rk2_aux_vars", remap_dyn_split_rk2_aux_vars)
    setattr(Context, "set_dtbt", set_dtbt)
    setattr(Context, "step_dyn_split_rk2", step_dyn_split_rk2)
    for f in (set_cs_vertvisc, vertvisc_coef, vertvisc_get_coef, vertvisc, vertvisc_remnant, vertvisc_ntrunc):
        setattr(Context, f.__name__, f)
    setattr(Context, "ale_regrid", ale_regrid)
    setattr(Context, "advect_tracer", advect_tracer)
    for f in (ale_remap_tracers, ale_remap_set_h_vel, ale_remap_velocities, remapping_core_h):
        setattr(Context, f.__name__, f)
    for f in (set_grid, set_vgrid, set_cs_continuity, continuity, set_unit_scale, set_cs_coriolisadv, coradcalc,
              set_cs_hor_visc, horizontal_viscosity, btstep, btcalc, bt_mass_source, set_cs_pressureforce, pressure_force):
        setattr(Context, f.__name__, f)


_ctx_methods()


def make_domain(ni, nj, nk=1, halo=4, whalo=None, cyclic_x=True, cyclic_y=False, first_direction=0,
                npi=1, npj=1, pi=0, pj=0):
    """Index bounds the way MOM_domains / hor_index_init set them: isc = halo+1 (1-based)."""
    d = Domain()
    d.isc, d.iec, d.jsc, d.jec = halo + 1, halo + ni, halo + 1, halo + nj
    d.isd, d.ied, d.jsd, d.jed = 1, ni + 2 * halo, 1, nj + 2 * halo
    wh = halo if whalo is None else whalo
    d.isdw, d.iedw, d.jsdw, d.jedw = d.isc - wh, d.iec + wh, d.jsc - wh, d.jec + wh
    d.nk = nk
    d.cyclic_x, d.cyclic_y, d.first_direction = int(cyclic_x), int(cyclic_y), first_direction
    d.npi, d.npj, d.pi, d.pj = npi, npj, pi, pj
    return d


def format_ocean_stats_line(lib, cs, e, n, reday):
    """mom6cu_ocean_stats_line on a write_energy result (host formatting only: needs no device)."""
    from . import marshal
    from ._lib import EnergyOut, _EO_SCALARS, _EO_SCALARS2
    keep = []
    st = marshal.sum_output_cs(cs, keep)
    out = EnergyOut()
    for k in _EO_SCALARS + _EO_SCALARS2:
        setattr(out, k, float(e[k]))
    out.max_CFL[0], out.max_CFL[1] = float(e["max_CFL"][0]), float(e["max_CFL"][1])
    out.ntrunc = int(e["ntrunc"])
    z = np.ascontiguousarray(e["Z_0APE"], dtype=np.float64)
    out.Z_0APE = z.ctypes.data
    buf = C.create_string_buffer(512)
    rc = lib.mom6cu_ocean_stats_line(C.byref(st), C.byref(out), int(n), float(reday), buf, 512)
    if rc != 0:
        raise Mom6cuError(f"mom6cu_ocean_stats_line rc={rc}")
    return buf.value.decode()


def efp_op(lib, op, a, b=None):
    """EFP_plus / EFP_minus / EFP_to_real / real_to_EFP / EFP_real_diff (MOM_coms.F90:737-815) through the C ABI; host-only."""
    from . import marshal
    from ._lib import Efp
    if op == "from_real":
        out = Efp()
        rc = lib.mom6cu_real_to_efp(float(a), C.byref(out))
        if rc:
            raise OverflowError("Overflow in real_to_EFP conversion")
        return marshal.efp_back(out)
    ea = marshal.efp(a)
    if op == "to_real":
        return float(lib.mom6cu_efp_to_real(C.byref(ea)))
    eb = marshal.efp(b)
    if op == "diff":
        return float(lib.mom6cu_efp_real_diff(C.byref(ea), C.byref(eb)))
    out, ov = Efp(), C.c_int(0)
    getattr(lib, "mom6cu_efp_" + op)(C.byref(ea), C.byref(eb), C.byref(out), C.byref(ov))
    return marshal.efp_back(out)


ocean_stats_line = format_ocean_stats_line
