#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MOM6 dycore hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one whole baroclinic dynamics step (step_MOM_dyn_split_RK2, MOM_dynamics_split_RK2.F90:294-1205: pressure
force, CorAdCalc, vertvisc_coef/vertvisc/vertvisc_remnant, 3 x continuity_PPM, 2 x btstep with its barotropic subcycle,
horizontal_viscosity and the seven group passes) of the OM4_025-shaped synthetic configuration (1440 x 1080 x 75,
BASELINE.json configs[1]) as ONE call of the C ABI (mom6cu_step_dyn_split_rk2).
value = cell-updates/s = ni*nj*nk*K / t with every field resident in HBM (mom6cu_plane_*), t = device time of the K
steps (CUDA events on the launching stream, max over ranks); ms_per_step_wall is the host wall clock of the same loop.
e2e = the same call with the model state, T, S, visc% and forces% in pinned HOST arrays (staging copies inside the timed
region, results copied back).  in_step = device time of each stage INSIDE the timed steps (CUDA events around its kernels,
mom6cu_last_step_stage_ms); roofline = the dominant of those (continuity_PPM) against the measured HBM peak;
per_stage_isolated = each stage called alone on synthetic stage inputs; thermo_pass = tracer advection + ALE regrid/remap, which run once per DT_THERM/DT dynamics steps.
cpu_baseline / --impl reference = the oracle restatement of the reference CPU path, one single-threaded rank per host
core on the tiles of an MPI-style decomposition (the Fortran reference cannot be built here: no Fortran compiler, MPI,
netCDF or FMS in the image).  N > 1: the same global grid on a 2x1 / 2x2 / 4x2 tile layout, one rank per GPU, NCCL halo
exchanges (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NI, NJ, NK = 1440, 1080, 75
SAMPLE = (90, 135)              # CPU-baseline tile: 1440x1080 over a 16x8 rank layout, all 75 layers
BT_BYTES_PER_PT_SUBSTEP = 552   # 69 fp64 operands on the BT_cont path
# stage -> (calls per baroclinic step [MOM_dynamics_split_RK2.F90 line], algorithmic bytes per cell per call [SURVEY 8d])
STEP = [("pressure_force", 1, 48), ("continuity", 3, 96), ("btcalc", 1, 32), ("bt_mass_source", 2, 8), ("btstep", 2, 136), ("coradcalc", 2, 56),
        ("horizontal_viscosity", 1, 40)]
# DT, the prescribed DTBT > 0 (synthetic.btstep_inputs: 0.45 dx / sqrt(2 g H) with dx = 25 km, H = 4000 m) and DT_BT_FILTER of the workload
DT, DTBT, DT_BT_FILTER = 900.0, 0.9 * 0.5 * 2.5e4 / (9.8 * 4000.0 * 2) ** 0.5, -0.25
LAND_BLOCKS = 40                               # synthetic land: ~24 % of the 1440x1080 points (masks exercised as in a global ocean)
PGF_RECON = dict(reconstruct=1, Recon_Scheme=1, boundary_extrap=0)   # RECONSTRUCT_FOR_PRESSURE=True, PRESSURE_RECONSTRUCTION_SCHEME=1 (PLM):
#                                                                      the reference's defaults under ALE (MOM_PressureForce_FV.F90:2172-2190)


def bt_substeps():
    """nstep + nfilter of btstep for DT, DTBT, DT_BT_FILTER (MOM_barotropic.F90:780-802, :1678-1700)."""
    import math
    nstep = int(math.ceil(DT / DTBT - 0.0001))
    dtbt = DT / nstep
    dt_filt = 0.5 * max(0.0, min(DT_BT_FILTER, 2.0 * DT)) if DT_BT_FILTER >= 0.0 else 0.5 * max(0.0, DT * min(-DT_BT_FILTER, 2.0))
    return nstep, int(math.ceil(dt_filt / dtbt))


def workload_config():
    """The `config` object: identical in both arms (the driver compares them)."""
    nstep, nfilter = bt_substeps()
    return {"workload": f"OM4_025-shaped {NI}x{NJ}x{NK} split-RK2 dynamics step",
            "stages": ["step_MOM_dyn_split_RK2 (MOM_dynamics_split_RK2.F90:294-1205) as ONE call: PressureForce_FV_Bouss (Wright EOS, "
                       "RECONSTRUCT_FOR_PRESSURE with PLM T,S profiles: int_density_dz_generic_plm) x1, CorAdCalc x2(+1 with store_CAu), "
                       "vertvisc_coef x3 + vertvisc x2 + vertvisc_remnant x3, continuity_PPM x3, btcalc x2, bt_mass_source x2, "
                       f"btstep x2 ({nstep} + {nfilter} = {nstep + nfilter} barotropic substeps each: DT = {DT:g} s, DTBT = {DTBT:.2f} s prescribed, "
                       f"DT_BT_FILTER = {DT_BT_FILTER:g}), horizontal_viscosity x1, the elementwise glue and the 7 group passes"],
            "stages_missing": MISSING, "land_blocks": LAND_BLOCKS, "seed": "synthetic.SEED (one global field set; tiles are cut from it)",
            "l2": "inputs larger than L2 (about 28 GB of resident fields per 1440x1080x75 tile set are swept every step)"}


MISSING = ["set_viscous_ML (a no-op with DYNAMIC_VISCOUS_ML=False)", "set_dtbt (calc_dtbt=False: CS%dtbt as stored)", "diagnostics / checksums"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def tile_of(rank, n):
    """LAYOUT for n ranks (MOM_domains.F90:154-222): minimise halo perimeter."""
    layouts = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}
    npi, npj = layouts.get(n, (n, 1))
    return npi, npj, rank % npi, rank // npi


STAGGER = dict(u="u", v="v", hin="h", h="h", uh="u", vh="v", visc_rem_u="u", visc_rem_v="v", uhbt="u", vhbt="v", u_cor="u", v_cor="v",
               CAu="u", CAv="v", diffu="u", diffv="v", U_in="u", V_in="v", eta_in="h", bc_accel_u="u", bc_accel_v="v", taux="u",
               tauy="v", pbce="h", eta_PF_in="h", U_Cor="u", V_Cor="v", accel_layer_u="u", accel_layer_v="v", eta_out="h",
               uhbtav="u", vhbtav="v", uh0="u", vh0="v", u_uh0="u", v_vh0="v", etaav="h", h_u="u", h_v="v", frhatu="u",
               frhatv="v", bathyT="h", eta="h", eta_cor="h", FA_u_EE="u", FA_u_E0="u", FA_u_W0="u", FA_u_WW="u", uBT_WW="u",
               uBT_EE="u", FA_v_NN="v", FA_v_N0="v", FA_v_S0="v", FA_v_SS="v", vBT_SS="v", vBT_NN="v", IDatu="u", IDatv="v",
               T="h", S="h", PFu="u", PFv="v", eta_cor_bound="h", ubtav="u", vbtav="v", IareaT="h", IareaT_OBCmask="h", IdxCu="u", IdyCv="v", q_D="q",
               D_u_Cor="u", D_v_Cor="v", ua_polarity="h", va_polarity="h", OBCmask_u="u", OBCmask_v="v")
WIDE = {"IareaT", "IareaT_OBCmask", "IdxCu", "IdyCv", "q_D", "D_u_Cor", "D_v_Cor", "ua_polarity", "va_polarity", "OBCmask_u", "OBCmask_v"}


def make_resident(ctx, stages, nk):
    """Upload every array argument of every stage once; the same host array maps to the same plane."""
    cache = {}

    def res(name, arr, wide=False):
        if not isinstance(arr, np.ndarray) or name.startswith("_"):
            return arr
        key = id(arr)
        if key not in cache:
            n3 = nk if arr.ndim == 3 else 1
            cache[key] = ctx.plane(f"{name}.{len(cache)}", arr, STAGGER[name], wide, n3)
        return cache[key]

    out = {}
    for st, (cs, a) in stages.items():
        ra = {k: ({kk: res(kk, vv) for kk, vv in v.items()} if isinstance(v, dict) else res(k, v)) for k, v in a.items()}
        rcs = cs
        if st == "btstep":
            rcs = {k: res(k, v, k in WIDE or k == "bathyT") for k, v in cs.items()}
        out[st] = (rcs, ra)
    return out, sum(v.nk for v in cache.values())


CS_ST = dict(CAu="u", CAv="v", CAu_pred="u", CAv_pred="v", PFu="u", PFv="v", diffu="u", diffv="v", visc_rem_u="u", visc_rem_v="v", u_accel_bt="u",
             v_accel_bt="v", u_av="u", v_av="v", h_av="h", pbce="h", eta="h", eta_PF="h", uhbt="u", vhbt="v", taux_bot="u", tauy_bot="v")
ARG_ST = dict(u_inst="u", v_inst="v", h="h", T="h", S="h", Kv_bbl_u="u", Kv_bbl_v="v", bbl_thick_u="u", bbl_thick_v="v", Kv_shear="h", taux="u",
              tauy="v", ustar="h", uh="u", vh="v", uhtr="u", vhtr="v", eta_av="h")


def step_resident(ctx, dom, cs, sa):
    """Upload everything a step touches once: MOM_dyn_split_RK2_CS arrays, BT_cont, barotropic_CS and the arguments."""
    n = [0]

    def pl(name, arr, st, wide=False):
        nk = arr.shape[0] if arr.ndim == 3 else 1
        n[0] += nk
        return ctx.plane(name, arr, st, wide, nk)

    rcs = dict(cs)
    for k, st in CS_ST.items():
        rcs[k] = pl("cs." + k, cs[k], st)
    rcs["BT_cont"] = {k: (pl("btc." + k, v, "u" if ("_u" in k or k.startswith("uBT")) else "v") if isinstance(v, np.ndarray) else v)
                      for k, v in cs["BT_cont"].items()}
    rcs["barotropic"] = {k: (pl("bt." + k, v, STAGGER[k], k in WIDE or k == "bathyT") if isinstance(v, np.ndarray) else v)
                         for k, v in cs["barotropic"].items()}
    rsa = dict(sa)
    for k, st in ARG_ST.items():
        if isinstance(sa.get(k), np.ndarray):
            rsa[k] = pl("arg." + k, sa[k], st)
    return rcs, rsa, n[0]


THERMO_EVERY = 8   # DT_THERM / DT of OM4_025 (7200 s / 900 s)
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the four kernels of one continuity_PPM call at 1440x1080x75 on one B200, from the
# committed `ncu --set full` capture profiles/r02_cont_all_ncu.md (the predictor call shape: uhbt, u_cor and BT_cont all present)
CONT_TRAFFIC = {"cont_flux_tiled<zonal>": 2.968966e9 + 2.912770e9, "cont_convergence_kernel<zonal>": 2.165341e9 + 0.912571e9,
                "cont_flux_tiled<meridional>": 2.907649e9 + 2.869879e9, "cont_convergence_kernel<meridional>": 2.020983e9 + 0.899442e9}


def thermo_pass(Context, synthetic, ni, nj, device):
    """advect_tracer (T, S; PLM) + ALE_regrid (Z*) + ALE_remap_tracers (T, S) + ALE_remap_set_h_vel x2 + ALE_remap_velocities on
    resident fields; returns the device times [ms]."""
    out = {}
    dom, grid, gv, cs, a = synthetic.advect_inputs(ni, nj, NK, land_blocks=40, cfl=1.5, dt=900.0 * THERMO_EVERY, dt_dyn=900.0)
    ctx = Context(dom, device)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    ra = dict(a)
    for k, st in (("h_end", "h"), ("uhtr", "u"), ("vhtr", "v")):
        ra[k] = ctx.plane("adv." + k, a[k], st, False, NK)
    ra["tr"] = [ctx.plane(f"adv.tr{m}", t, "h", False, NK) for m, t in enumerate(a["tr"])]
    ctx.advect_tracer(cs, ra)
    it = ctx.advect_tracer(cs, ra)
    out["advect_tracer_ms"] = ctx.last_kernel_ms; out["advect_tracer_iterations"] = it
    ctx.close()
    dom, grid, gv, rcs, ga = synthetic.regrid_inputs(ni, nj, NK, land_blocks=40)
    _, _, mcs, ma = synthetic.remap_inputs(ni, nj, NK, land_blocks=40)
    ctx = Context(dom, device)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    h = ctx.plane("ale.h", ga["h"], "h", False, NK); hn = ctx.plane("ale.h_new", ga["h_new"], "h", False, NK)
    dz = ctx.plane("ale.dz", ga["dzRegrid"], "h", False, NK + 1)
    ctx.ale_regrid(rcs, h, hn, dz); ctx.ale_regrid(rcs, h, hn, dz)
    out["ale_regrid_ms"] = ctx.last_kernel_ms
    T = ctx.plane("ale.T", ma["tr"][0], "h", False, NK); S = ctx.plane("ale.S", ma["tr"][1], "h", False, NK)
    ctx.ale_remap_tracers(mcs, h, hn, [T, S]); ctx.ale_remap_tracers(mcs, h, hn, [T, S])
    out["ale_remap_tracers_ms"] = ctx.last_kernel_ms
    u = ctx.plane("ale.u", ma["u"], "u", False, NK); v = ctx.plane("ale.v", ma["v"], "v", False, NK)
    hu0 = ctx.plane("ale.hu0", ma["u"], "u", False, NK); hv0 = ctx.plane("ale.hv0", ma["v"], "v", False, NK)
    hu1 = ctx.plane("ale.hu1", ma["u"], "u", False, NK); hv1 = ctx.plane("ale.hv1", ma["v"], "v", False, NK)
    ctx.ale_remap_set_h_vel(h, hu0, hv0); ctx.ale_remap_set_h_vel(hn, hu1, hv1)
    out["ale_remap_set_h_vel_ms"] = 2.0 * ctx.last_kernel_ms
    ctx.ale_remap_velocities(mcs, hu0, hv0, hu1, hv1, u, v); ctx.ale_remap_velocities(mcs, hu0, hv0, hu1, hv1, u, v)
    out["ale_remap_velocities_ms"] = ctx.last_kernel_ms
    # the parity metric on the same resident state (not part of the pass: ocean.stats is written every ENERGYSAVEDAYS)
    diag = {}
    try:
        so = dict(do_APE_calc=0, use_temperature=1, dt_in_T=900.0)
        ctx.write_energy(so, u, v, h, T, S); e = ctx.write_energy(so, u, v, h, T, S)
        diag = {"write_energy_ms": ctx.last_kernel_ms, "ocean_stats_line": ctx.ocean_stats_line(so, e, 1, 0.0)}
        ctx.chksum(h, 0, haloshift=1); ctx.chksum(h, 0, haloshift=1)
        diag["hchksum_haloshift1_ms"] = ctx.last_kernel_ms
    except Exception as ex:   # an auxiliary leg must never cost the bench line
        diag["error"] = repr(ex)
    # mixedlayer_restrat (MOM.F90:1422: every dynamics step of an OM4-like run, between the dycore and the tracer advection) on the
    # same resident h, T, S; reported on its own, not part of the headline step
    mle = {}
    try:
        mcs_, f2 = synthetic.mle_cs_and_forcing(ga["h"].shape[1:])
        uhtr = ctx.plane("mlearg.uhtr", ma["u"], "u", False, NK); vhtr = ctx.plane("mlearg.vhtr", ma["v"], "v", False, NK)
        f2 = {k: ctx.plane("mlearg." + k, x, "h", False, 1) for k, x in f2.items()}
        for k in ("MLD_filtered", "MLD_filtered_slow"):
            mcs_[k] = ctx.plane("mlearg." + k, mcs_[k], "h", False, 1)
        for _ in range(2):
            ctx.mixedlayer_restrat(mcs_, h, uhtr, vhtr, T, S, f2["ustar"], 900.0, f2["h_MLD"], f2["Rd_dx_h"])
        mle = {"mixedlayer_restrat_ms": ctx.last_kernel_ms, "calls_per_dynamics_step": 1}
    except Exception as ex:
        mle["error"] = repr(ex)
    ctx.close()
    out["total_ms"] = sum(v for k, v in out.items() if k.endswith("_ms"))
    out["diagnostics"] = diag
    out["mixedlayer_restrat"] = mle
    # the other two callers of SURVEY 8f row 2 (thickness_diffuse, MOM.F90:1388; tracer_hordiff, MOM.F90:1526), timed by tools/time_callers.py
    # in a process of its own after this context is closed: an auxiliary leg must never be able to cost the bench line
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_callers.py"), f"{ni},{nj},{NK}", str(device)], capture_output=True,
                           text=True, timeout=240)
        out["thickness_diffuse_and_tracer_hordiff"] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else \
            {"error": (r.stderr or r.stdout)[-300:]}
    except Exception as ex:
        out["thickness_diffuse_and_tracer_hordiff"] = {"error": repr(ex)}
    return out


def _flatten(tree, path=()):
    if isinstance(tree, dict):
        for k in sorted(tree, key=str):
            yield from _flatten(tree[k], path + (k,))
    else:
        yield path, tree


def _unflatten(items):
    out = {}
    for path, v in items:
        d = out
        for k in path[:-1]:
            d = d.setdefault(k, {})
        d[path[-1]] = v
    return out


def tile_inputs(synthetic, torch, dist, world, rank, gni, gnj, npi, npj, nk=None, whalo=10):
    """This rank's tile of ONE global synthetic field set (seed synthetic.SEED), so that every N runs the same problem and the state
    checksums after the K steps can be compared across N (the reference's layout test, .testing/Makefile test.layout).  Rank 0 builds
    the global inputs, cuts every rank's tile (synthetic.split_step_tile) and sends it over NCCL; nothing of this is timed."""
    nk = NK if nk is None else nk
    if world == 1:
        return synthetic.step_dyn_inputs(gni, gnj, nk, whalo=whalo, land_blocks=LAND_BLOCKS, store_CAu=1)
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")   # cpu: the gloo test
    if rank == 0:
        dom_g, grid_g, gv, css, cs_g, a_g = synthetic.step_dyn_inputs(gni, gnj, nk, whalo=whalo, land_blocks=LAND_BLOCKS, store_CAu=1)
        css_rest = {k: v for k, v in css.items() if k != "hor_visc"}
        mine = None
        for r in list(range(1, world)) + [0]:
            dom, g, c, t, hv = synthetic.split_step_tile(dom_g, grid_g, cs_g, a_g, npi, npj, r % npi, r // npi, css["hor_visc"])
            tree = {"grid": g, "cs": c, "a": t, "hv": hv, "gv": gv, "css": css_rest}
            if r == 0:
                mine = tree
                break
            items = list(_flatten(tree))
            meta = [(p, ("nd", v.shape, str(v.dtype)) if isinstance(v, np.ndarray) else ("py", v)) for p, v in items]
            dist.send_object_list([meta], dst=r)
            for p, v in items:
                if isinstance(v, np.ndarray):
                    dist.send(torch.from_numpy(np.ascontiguousarray(v)).to(dev), dst=r)
            del tree, items
        del dom_g, grid_g, cs_g, a_g
        tree = mine
    else:
        box = [None]
        dist.recv_object_list(box, src=0)
        items = []
        for p, m in box[0]:
            if m[0] == "nd":
                buf = torch.empty(tuple(m[1]), dtype=getattr(torch, m[2]), device=dev)
                dist.recv(buf, src=0)
                items.append((p, buf.cpu().numpy()))
            else:
                items.append((p, m[1]))
        tree = _unflatten(items)
    pi, pj = rank % npi, rank // npi
    dom = synthetic.make_domain(gni // npi, gnj // npj, nk=nk, halo=4, whalo=whalo, cyclic_x=True, cyclic_y=False, first_direction=0,
                                npi=npi, npj=npj, pi=pi, pj=pj)
    css = dict(tree["css"]); css["hor_visc"] = tree["hv"]
    return dom, tree["grid"], tree["gv"], css, tree["cs"], tree["a"]


def resident_cycle_e2e(ctx, torch, synthetic, rcs, rsa, sa, world, cycles, barrier):
    """e2e through the public API the way an OM4-like run would drive it (step_MOM, src/core/MOM.F90:1234-1658), the model state resident:
    per dynamics step    forces% (taux, tauy, ustar) and visc% (Kv_bbl_u/v, bbl_thick_u/v) come from pinned HOST arrays, eta_av goes back;
                         step_MOM_dyn_split_RK2 -> thickness_diffuse -> pass_var(h) -> mixedlayer_restrat -> pass_var(h)   (:1312-1427)
    per DT_THERM/DT = 8  visc%Kv_shear (3-D) and visc%h_ML from the host; advect_tracer(T, S) -> uhtr = vhtr = 0 -> tracer_hordiff ->
                         ALE_regridding_and_remapping (:1523-1541, :1658), then write_energy (the ocean.stats numbers) read back to the host.
    u, v, h, T, S, uhtr, vhtr and every control-structure array stay in HBM throughout.  At N > 1 the callers outside the dycore are left out
    (their multi-tile device paths are not validated): the cycle is the 8 dynamics steps with the same host traffic.
    Returns (seconds for `cycles` cycles, h2d bytes per step, d2h bytes per step, description)."""
    def pin(x):
        t = torch.empty(x.shape, dtype=torch.float64, pin_memory=True)
        y = t.numpy(); y[...] = x
        keep.append(t)
        return y

    keep = []
    nk = NK
    HOST2D = ("taux", "tauy", "ustar", "Kv_bbl_u", "Kv_bbl_v", "bbl_thick_u", "bbl_thick_v")
    host = {k: pin(sa[k]) for k in HOST2D}
    host_eta, host_kvs = pin(sa["eta_av"]), pin(sa["Kv_shear"])
    esa = dict(rsa)
    esa.update(host); esa["eta_av"] = host_eta
    shp2 = sa["eta_av"].shape
    chain = world == 1
    u, v, h, T, S, uhtr, vhtr = (rsa[k] for k in ("u_inst", "v_inst", "h", "T", "S", "uhtr", "vhtr"))
    if chain:
        tcs = synthetic.thickness_diffuse_cs()
        td_args = dict(h=h, uhtr=uhtr, vhtr=vhtr, T=T, S=S, dt=DT)
        mcs, f2 = synthetic.mle_cs_and_forcing(shp2)
        for k in ("MLD_filtered", "MLD_filtered_slow"):
            mcs[k] = ctx.plane("cyc." + k, mcs[k], "h", False, 1)
        h_MLD = pin(f2["h_MLD"]); Rd = ctx.plane("cyc.Rd_dx_h", f2["Rd_dx_h"], "h", False, 1)
        acs = dict(dt=DT, default_advect_scheme=0, useHuynhStencilBug=0)
        adv_args = dict(h_end=h, uhtr=uhtr, vhtr=vhtr, dt=DT * THERMO_EVERY, tr=[T, S])
        hcs = synthetic.hordiff_cs()
        hd_args = dict(h=h, dt=DT * THERMO_EVERY, tr=[T, S], conc_underflow=np.zeros(2))
        w = np.linspace(1.0, 3.0, nk); w /= w.sum()
        remapCS = dict(remapping_scheme=4, boundary_extrapolation=0, force_bounds_in_subcell=0, force_bounds_in_target=1, om4_remap_via_sub_cells=1,
                       answer_date=20190101, h_neglect=1.0e-30, h_neglect_edge=1.0e-30)
        ale = dict(regridCS=dict(regridding_scheme=2, nk=nk, min_thickness=1.0e-3, old_grid_weight=0.0, depth_of_time_filter_shallow=0.0,
                                 depth_of_time_filter_deep=0.0, Z_ref=0.0, coordinateResolution=np.ascontiguousarray(4000.0 * w)),
                   remapCS=remapCS, vel_remapCS=dict(remapCS), regrid_time_scale=3600.0, remap_aux_vars=1)
        ale_args = dict(u=u, v=v, h=h, tr=[T, S], conc_underflow=np.zeros(2), iT=0, iS=1, dtdia=DT * THERMO_EVERY, Kd_shear=None,
                        Kv_shear=rsa["Kv_shear"], Kv_shear_Bu=None)
    so = dict(do_APE_calc=0, use_temperature=1, dt_in_T=DT)

    def cycle(chain):
        rsa["Kv_shear"].upload(host_kvs)
        for _ in range(THERMO_EVERY):
            ctx.step_dyn_split_rk2(rcs, esa)
            if chain:
                ctx.thickness_diffuse(tcs, td_args)
                ctx.do_group_pass([h], ["h"], nk)
                ctx.mixedlayer_restrat(mcs, h, uhtr, vhtr, T, S, host["ustar"], DT, h_MLD, Rd)
                ctx.do_group_pass([h], ["h"], nk)
        if chain:
            ctx.advect_tracer(acs, adv_args)
            uhtr.zero(); vhtr.zero()
            ctx.tracer_hordiff(hcs, hd_args)
            ctx.ale_regridding_and_remapping(ale, ale_args, dyn_cs=rcs)
        return ctx.write_energy(so, u, v, h, T, S)

    def timed(chain_on):
        e = cycle(chain_on)          # first cycle: allocates the staging planes, untimed
        if not np.isfinite(e["En_mass"]) or not np.isfinite(e["mass_tot"]):
            raise RuntimeError("the resident cycle produced a non-finite state")
        barrier()
        t0 = time.perf_counter()
        for _ in range(cycles):
            e = cycle(chain_on)
        barrier()
        t = time.perf_counter() - t0
        if not np.isfinite(e["En_mass"]):
            raise RuntimeError("the resident cycle produced a non-finite state")
        return t, e

    # (1) the dynamics steps alone: the work of `value` and of the reference arm, plus the host traffic of the forcing
    dyn_s, e = timed(False)
    # (2) the whole cycle with the callers of the dycore inside the timed region (N = 1)
    full = None
    if chain:
        try:
            full_s, e = timed(True)
            full = dict(seconds=full_s, En_mass=float(e["En_mass"]))
        except Exception as ex:
            full = dict(error=repr(ex))
    dt_s = dyn_s
    per_step_2d = sum(x.nbytes for x in host.values()) + host_eta.nbytes
    h2d = per_step_2d + host_kvs.nbytes // THERMO_EVERY
    d2h = host_eta.nbytes + 4096 // THERMO_EVERY
    desc = (f"{THERMO_EVERY} consecutive calls of step_MOM_dyn_split_RK2 with the model state resident: forces% and visc% 2-D fields from pinned host "
            "arrays every step, eta_av back every step, visc%Kv_shear (3-D) from the host and the write_energy numbers (ocean.stats) back once per "
            f"{THERMO_EVERY} steps; u, v, h, T, S, uhtr, vhtr and every control-structure array stay in HBM between steps, which is what running the "
            "callers of the dycore on the device buys (full_cycle times them too)")
    if full is not None and "seconds" in full:
        full["what"] = (f"one thermodynamic cycle of step_MOM (MOM.F90:1234-1658) per {THERMO_EVERY} dynamics steps, everything inside the timed region: each "
                        "step is step_MOM_dyn_split_RK2 -> thickness_diffuse -> pass_var(h) -> mixedlayer_restrat -> pass_var(h); the cycle ends with "
                        "advect_tracer(T,S) -> uhtr = vhtr = 0 -> tracer_hordiff -> ALE_regridding_and_remapping -> write_energy; same host traffic "
                        "plus forces%ustar / visc%h_ML for mixedlayer_restrat")
        full["h2d_bytes_per_step"] = h2d + host["ustar"].nbytes + h_MLD.nbytes
    e = dict(e); e["full_cycle"] = full
    return dt_s, h2d, d2h, desc, e


def full_cycle_line(cells, cyc):
    f = cyc.get("full")
    if not f:
        return None
    if "seconds" not in f:
        return f
    return {"value": cells * THERMO_EVERY * cyc["cycles"] / f["seconds"], "unit": "cell-updates/s (dynamics steps per second x cells, the callers' time included)",
            "h2d_bytes_per_step": f["h2d_bytes_per_step"], "d2h_bytes_per_step": cyc["d2h"], "what": f["what"], "En_mass_after": f["En_mass"]}


def e2e_line(cells, e2e_s, e2e_steps, h2d, d2h, cyc):
    """The e2e object.  value: the dynamics step through the C ABI with the state resident between steps and the forcing crossing PCIe every step
    (like for like with `value` and with the reference arm: the same work).  full_cycle: the same with the callers of the dycore that make the
    residency possible timed too.  host_state_every_step: the whole model state crossing PCIe every step (a host that adopts only the dycore)."""
    host_state = None if e2e_s is None else {"value": cells * e2e_steps / e2e_s, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                             "what": "state u,v,h + T,S + visc% + forces% from pinned host arrays each step, u,v,h,eta_av back"}
    if cyc and "seconds" in cyc:
        return {"value": cells * THERMO_EVERY * cyc["cycles"] / cyc["seconds"], "unit": "cell-updates/s", "h2d_bytes_per_step": cyc["h2d"],
                "d2h_bytes_per_step": cyc["d2h"], "what": cyc["desc"], "dynamics_steps_timed": THERMO_EVERY * cyc["cycles"],
                "En_mass_after": cyc["En_mass"], "full_cycle": full_cycle_line(cells, cyc), "host_state_every_step": host_state}
    if host_state is None:
        return None
    if cyc:
        host_state["resident_cycle_error"] = cyc.get("error")
    return host_state


def run_step(ctx, stages, times=None):
    """One baroclinic step: the implemented stages in the reference's call counts."""
    for name, calls, _ in STEP:
        cs, a = stages[name]
        for _ in range(calls):
            if name == "btstep":
                ctx.btstep(cs, a)
            elif name == "bt_mass_source":
                ctx.bt_mass_source(a["h"], a["eta"], 1, a["eta_cor"])
            else:
                getattr(ctx, name)(a)
            if times is not None:
                times[name] = times.get(name, 0.0) + ctx.last_kernel_ms


def oracle_step(orc, dom, grid, gv, stages, cores):
    for name, calls, _ in STEP:
        cs, a = stages[name]
        for _ in range(calls):
            if name == "continuity":
                orc.continuity(dom, grid, gv, cs, a, nthreads=cores)
            elif name == "coradcalc":
                orc.coradcalc(dom, grid, gv, cs, a, nthreads=cores)
            elif name == "horizontal_viscosity":
                orc.horizontal_viscosity(dom, grid, gv, cs, a, nthreads=cores)
            elif name == "btstep":
                orc.btstep(dom, grid, gv, cs, a, nthreads=cores)
            elif name == "pressure_force":
                orc.pressure_force(dom, grid, gv, cs, a, nthreads=cores)
            elif name == "btcalc":
                orc.btcalc(dom, grid, gv, a, nthreads=cores)
            elif name == "bt_mass_source":
                orc.bt_mass_source(dom, grid, gv, a["h"], a["eta"], 1, a["eta_cor"])


def cpu_reference(steps, warmup):
    """The oracle restatement of the reference CPU path, run the way the reference runs on a node: one single-threaded
    worker per host core, each stepping its own tile of the MPI-style decomposition of the workload (SAMPLE = the tile of
    a 16x8 layout of 1440x1080) through the whole step_MOM_dyn_split_RK2, no halo exchange between tiles (which only
    flatters the CPU arm)."""
    import oracle
    import threading
    from mom6_b200 import synthetic
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ni, nj = SAMPLE
    def clone(x):
        if isinstance(x, np.ndarray):
            return x.copy()
        if isinstance(x, dict):
            return {k: clone(v) for k, v in x.items()}
        return x

    work = []
    for w in range(cores):   # every rank its own tile: distinct seeds, the land-block count and PressureForce options of the GPU arm
        t = synthetic.step_dyn_inputs(ni, nj, NK, whalo=10, land_blocks=LAND_BLOCKS, seed=synthetic.SEED + w, store_CAu=1)
        t[3]["pressureforce"].update(PGF_RECON)
        work.append(t)

    def run(w, n):
        dom, grid, gv, css, cs, a = work[w]
        for _ in range(n):
            oracle.step_dyn_split_rk2(dom, grid, gv, css, cs, a, nthreads=1)

    def all_workers(n):
        th = [threading.Thread(target=run, args=(w, n)) for w in range(cores)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    if warmup:
        all_workers(warmup)
    t0 = time.perf_counter()
    all_workers(steps)
    t = time.perf_counter() - t0
    land = float(np.mean([1.0 - w[1]["mask2dT"][4:-4, 4:-4].mean() for w in work]))
    return (cores * ni * nj * NK * steps / t, t, cores,
            f"{warmup} warm-up + {steps} timed step(s) of step_MOM_dyn_split_RK2 on {cores} concurrent {ni}x{nj}x{NK} tiles (the tile size of a "
            f"16x8 MPI-style decomposition of the workload, one single-threaded rank per core, each tile its own seed, land_blocks={LAND_BLOCKS} "
            f"(land fraction {land:.2f}), no halo exchange)")


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; see module docstring)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    val, t, cores, sample = cpu_reference(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "cell-updates/sec", "value": val, "unit": "cell-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": val, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-array leg")
    ap.add_argument("--no-stages", action="store_true", help="skip the per-stage breakdown")
    ap.add_argument("--no-thermo", action="store_true", help="skip the tracer-advection / ALE pass that runs every DT_THERM/DT steps")
    ap.add_argument("--size", default=None, help="ni,nj override (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) out of it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mom6_b200 import synthetic
    from mom6_b200.api import Context

    gni, gnj = (NI, NJ) if not args.size else tuple(int(x) for x in args.size.split(","))
    # strong scaling: the global domain is split into npi x npj tiles (one per GPU)
    npi, npj, pi, pj = tile_of(rank, world)
    ni, nj = gni // npi, gnj // npj
    dom, grid, gv, css, cs, sa = tile_inputs(synthetic, torch, dist, world, rank, gni, gnj, npi, npj)
    css["pressureforce"].update(PGF_RECON)
    ctx = Context(dom, local)
    if world > 1:
        ctx.attach_comm(dist)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    ctx.set_cs_continuity(css["continuity"]); ctx.set_cs_coriolisadv(css["coriolisadv"]); ctx.set_cs_hor_visc(css["hor_visc"])
    ctx.set_cs_pressureforce(css["pressureforce"]); ctx.set_cs_vertvisc(css["vertvisc"])
    rcs, rsa, nplanes = step_resident(ctx, dom, cs, sa)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up steps, then exactly K steps of the whole baroclinic step
    for _ in range(args.warmup):
        ctx.step_dyn_split_rk2(rcs, rsa)
    barrier()
    n1 = ctx.launches
    dev_ms = 0.0
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        in_step = {}
        for _ in range(args.steps):
            ctx.step_dyn_split_rk2(rcs, rsa)
            dev_ms += ctx.last_kernel_ms
            for k, v in ctx.last_step_stage_ms().items():   # CUDA events around each stage's kernels inside this step
                in_step[k] = in_step.get(k, 0.0) + v
        barrier()
        wall = time.perf_counter() - t0
    launches = ctx.launches - n1
    # the reference's layout-invariance metric (bit-count checksums, MOM_checksums.F90:387-557) of the state after the W + K steps: every N
    # steps the same global problem, so these must be identical in every line of a scaling run
    state_chk = {}
    try:
        for k, stg in (("u_inst", 1), ("v_inst", 2), ("h", 0)):
            bc, _, st = ctx.chksum(rsa[k], stg, haloshift=0, stats=True)
            state_chk[k] = {"bitcount": int(bc[0]), "mean": float(st[0]), "min": float(st[1]), "max": float(st[2])}
    except Exception as ex:
        state_chk["error"] = repr(ex)
    # ---- e2e: the model state, tracers and forcing come from pinned HOST arrays every step and the new state goes back to
    # the host; the control structure (MOM_dyn_split_RK2_CS, BT_cont, barotropic_CS) and the transports stay on the device,
    # as they do between the reference's own steps.
    e2e_s, e2e_steps, h2d, d2h = None, 0, 0, 0
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 3))
        HOST_IN = ("u_inst", "v_inst", "h", "T", "S", "Kv_bbl_u", "Kv_bbl_v", "bbl_thick_u", "bbl_thick_v", "Kv_shear", "taux", "tauy", "ustar")
        HOST_OUT = ("u_inst", "v_inst", "h", "eta_av")
        esa = dict(rsa)
        for k in set(HOST_IN + HOST_OUT):
            if isinstance(sa.get(k), np.ndarray):
                t = torch.empty(sa[k].shape, dtype=torch.float64, pin_memory=True)
                esa[k] = t.numpy()
                esa[k][...] = sa[k]
        h2d = sum(esa[k].nbytes for k in HOST_IN if isinstance(esa.get(k), np.ndarray)) + esa["eta_av"].nbytes
        d2h = sum(esa[k].nbytes for k in HOST_OUT)
        ctx.step_dyn_split_rk2(rcs, esa)      # first call allocates the staging planes
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.step_dyn_split_rk2(rcs, esa)
        barrier()
        e2e_s = time.perf_counter() - t0
    # ---- e2e with the state resident between steps (what adopting the callers of the dycore buys): see resident_cycle_e2e
    cyc = None
    if not args.no_e2e:
        try:
            ncyc = 1 if args.steps < 2 * THERMO_EVERY else args.steps // THERMO_EVERY
            cyc_s, cyc_h2d, cyc_d2h, cyc_desc, cyc_e = resident_cycle_e2e(ctx, torch, synthetic, rcs, rsa, sa, world, ncyc, barrier)
            cyc = dict(seconds=cyc_s, cycles=ncyc, h2d=cyc_h2d, d2h=cyc_d2h, desc=cyc_desc, En_mass=float(cyc_e["En_mass"]), full=cyc_e.get("full_cycle"))
        except Exception as ex:      # the host-state leg above remains the e2e number
            cyc = dict(error=repr(ex))
    # ---- per-stage breakdown (separate stage calls on resident fields; not part of the headline)
    times, stage_passes = {}, 0
    if world == 1 and not args.no_stages:
        ctx.close()
        domS, gridS, gvS, stages = synthetic.step_inputs(ni, nj, NK, whalo=10, land_blocks=LAND_BLOCKS, seed=synthetic.SEED + rank)
        stages["pressure_force"][0].update(PGF_RECON)
        ctx = Context(domS, local)
        ctx.set_grid(gridS); ctx.set_vgrid(gvS)
        ctx.set_cs_continuity(stages["continuity"][0]); ctx.set_cs_coriolisadv(stages["coradcalc"][0])
        ctx.set_cs_hor_visc(stages["horizontal_viscosity"][0]); ctx.set_cs_pressureforce(stages["pressure_force"][0])
        resident, _ = make_resident(ctx, stages, NK)
        run_step(ctx, resident)
        stage_passes = 2
        for _ in range(stage_passes):
            run_step(ctx, resident, times)
        barrier()

    # ---- the thermodynamic-cadence pass (every DT_THERM/DT = 8 steps in OM4_025): advect_tracer of T and S over the 8 steps'
    # transports, Z* regrid, remap of T, S, u, v -- timed on its own, reported next to the dynamics step
    thermo = None
    if world == 1 and not args.no_thermo:
        thermo = thermo_pass(Context, synthetic, ni, nj, local)
        barrier()

    cyc_ok = 1.0 if (cyc and "seconds" in cyc) else 0.0
    tmax = torch.tensor([dev_ms, e2e_s or 0.0, wall, cyc["seconds"] if cyc_ok else 0.0, -cyc_ok], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms, e2e_max, wall, cyc_max, cyc_bad = [float(x) for x in tmax.cpu()]
    e2e_s = e2e_max if e2e_s is not None else None
    if cyc_ok and cyc_bad < 0.0:         # every rank completed the cycle
        cyc["seconds"] = cyc_max
    elif cyc and "seconds" in cyc:
        cyc = dict(error="a rank failed in the resident cycle")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cells = gni * gnj * NK
    tile_cells = ni * nj * NK
    value = cells * args.steps / (dev_ms * 1e-3)
    peak, peak_src = peaks()
    per_stage = {}
    for name, calls, bpc in (STEP if times else []):
        ms = times[name] / (stage_passes * calls)
        per_stage[name] = {"calls_per_step": calls, "ms_per_call": ms, "algorithmic_B_per_cell": bpc,
                           "achieved_GBps": tile_cells * bpc / (ms * 1e-3) / 1e9, "frac_of_peak": tile_cells * bpc / (ms * 1e-3) / 1e9 / peak}
    # the same stages timed INSIDE the K timed steps (this rank's tile): calls per step as the step makes them
    IN_STEP = {"pressure_force": (1, 48), "coradcalc": (2, 56), "vertvisc": (1, 512), "continuity": (3, 96), "btcalc": (2, 32), "btstep": (2, 136),
               "horizontal_viscosity": (1, 40)}
    in_step_tab = {}
    for name, tot in in_step.items():
        calls, bpc = IN_STEP[name]
        ms_step = tot / args.steps
        gbs = tile_cells * bpc * calls / (ms_step * 1e-3) / 1e9 if ms_step > 0 else 0.0
        in_step_tab[name] = {"calls_per_step": calls, "ms_per_step": ms_step, "share_of_step": ms_step / (dev_ms / args.steps),
                             "algorithmic_B_per_cell_per_call": bpc, "achieved_GBps": gbs, "frac_of_peak": gbs / peak}
    if "pressure_force" in in_step_tab:   # not an HBM-bound stage with RECONSTRUCT_FOR_PRESSURE: say so next to the HBM fraction
        in_step_tab["pressure_force"]["bound"] = ("fp64 pipe, not HBM: 35 equation-of-state evaluations per cell (~2.9 k fp64 instructions); ncu: fp64 pipe "
                                                  "65 % active, issue slots 58 % (profiles/r02_pgf_recon_ncu.md)")
    if in_step_tab:
        covered = sum(v["ms_per_step"] for v in in_step_tab.values() if "ms_per_step" in v)
        in_step_tab["glue_and_halo_updates"] = {"ms_per_step": dev_ms / args.steps - covered, "share_of_step": 1.0 - covered / (dev_ms / args.steps)}
    if in_step_tab and world == 1:
        dom_stage = max(IN_STEP, key=lambda k: in_step_tab[k]["ms_per_step"])
        t = in_step_tab[dom_stage]
        ds = {"achieved_GBps": t["achieved_GBps"], "frac_of_peak": t["frac_of_peak"], "algorithmic_B_per_cell": IN_STEP[dom_stage][1],
              "launch_ms": t["ms_per_step"] / t["calls_per_step"], "share": t["share_of_step"]}
    elif per_stage:
        dom_stage = max(per_stage, key=lambda k: per_stage[k]["ms_per_call"] * per_stage[k]["calls_per_step"])
        ds = per_stage[dom_stage]
    else:   # multi-GPU runs: the whole step against the sum of the stages' algorithmic bytes
        dom_stage = "step"
        bpc = sum(calls * b for _, calls, b in STEP)
        ms = dev_ms / args.steps
        ds = {"achieved_GBps": tile_cells * bpc / (ms * 1e-3) / 1e9, "frac_of_peak": tile_cells * bpc / (ms * 1e-3) / 1e9 / peak,
              "algorithmic_B_per_cell": bpc}
    kernel_of = {"continuity": "cont_flux_tiled<zonal|meridional> + cont_convergence_kernel", "btstep": "bt_substep_kernel x26 + bt_col_kernel + bt_layer_accel_kernel", "vertvisc": "vv_coef_kernel + vv_solve_kernel",
                 "coradcalc": "corad_kernel", "horizontal_viscosity": "hor_visc_kernel", "btcalc": "btcalc_kernel", "bt_mass_source": "bt_mass_source_kernel",
                 "pressure_force": "pgf_ts_edges_kernel + pgf_recon_kernel (fp64-pipe bound: 35 EOS evaluations per cell)", "step": "all stage kernels of one step"}
    line = {"metric": "cell-updates/sec", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config() if not args.size else dict(workload_config(), workload=f"OM4_025-shaped {gni}x{gnj}x{NK} split-RK2 dynamics step"),
            "run": {"tiles": f"{npi}x{npj}", "resident_planes": nplanes, "resident_GB_per_rank": nplanes * ni * nj * 8 / 1e9,
                    "e2e": "state u,v,h + T,S + visc% + forces% from pinned host arrays each step, u,v,h,eta_av back; CS arrays and transports resident"},
            "state_checksum_after_steps": {"steps_run": args.warmup + args.steps, "fields": state_chk,
                                           "note": "bit-count checksum + mean/min/max over the global domain; identical for every N"},
            "e2e": e2e_line(cells, e2e_s, e2e_steps, h2d, d2h, cyc),
            "gpu_launches": launches, "ms_per_step_wall": 1e3 * wall / args.steps,
            "roofline": {"bound": "hbm", "kernel": kernel_of[dom_stage], "stage": dom_stage, "achieved": ds["achieved_GBps"], "peak": peak,
                         "unit": "GB/s", "frac": ds["frac_of_peak"],
                         "traffic": (sum(CONT_TRAFFIC.values()) if (dom_stage == "continuity" and world == 1 and (gni, gnj) == (NI, NJ)) else None),
                         "traffic_source": "profiles/r02_cont_all_ncu.md: DRAM bytes of the 2 flux + 2 convergence launches of one call (1.58 x the "
                                           "96 B/cell algorithmic figure, which credits uh and the intermediate h as on-chip; each flux kernel alone "
                                           "moves 1.05 x its own compulsory 48 B/cell)", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": tile_cells * ds["algorithmic_B_per_cell"],
                         "avg_launch_ms": ds.get("launch_ms"), "share_of_step": ds.get("share"),
                         "note": "algorithmic bytes of one stage call / its average CUDA-event time INSIDE the timed steps (in_step); "
                                 "per_stage_isolated = the same stages called alone on synthetic stage inputs"},
            "in_step": in_step_tab, "per_stage_isolated": per_stage, "clocks": clk.summary()}
    if thermo:
        per8 = THERMO_EVERY * dev_ms / args.steps + thermo["total_ms"]
        thermo["every_n_steps"] = THERMO_EVERY
        thermo["cell_updates_per_s_with_thermo"] = cells * THERMO_EVERY / (per8 * 1e-3)
        line["thermo_pass"] = thermo
    # BASELINE.json configs[4]: btstep microbench, 4320x3240 subcycle sweep at 1 GPU
    if world == 1 and not args.size:
        ctx.close()
        domb, ab = synthetic.bt_timeloop_inputs(4320, 3240, whalo=10, halo=4, nstep=60, nfilter=8, land_blocks=40)
        ctxb = Context(domb, local)
        ctxb.btstep_timeloop(ab, reps=3, download=False)
        msb = ctxb.last_kernel_ms
        gbs = 4320 * 3240 * 68 * BT_BYTES_PER_PT_SUBSTEP / (msb * 1e-3) / 1e9
        line["btstep_microbench"] = {"grid": "4320x3240", "substeps": 68, "ms_per_call": msb, "achieved_GBps": gbs, "frac_of_peak": gbs / peak,
                                     "kernel": "bt_substep_kernel", "algorithmic_B_per_pt_substep": BT_BYTES_PER_PT_SUBSTEP}
        ctxb.close()
    if not args.no_cpu:
        val, t, cores, sample = cpu_reference(1, 0)
        line["cpu_baseline"] = {"value": val, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
