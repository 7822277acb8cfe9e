#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MOM6 dycore hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one baroclinic dynamics step of the OM4_025-shaped synthetic configuration
(1440 x 1080 x 75, BASELINE.json configs[3]) through the stages implemented so far (listed in
config.stages); value = cell-updates/s = ni*nj*nk*K / t, t = device time (CUDA events on the
launching stream) with every input already resident in HBM.  e2e = the same metric through the
reference-facing C ABI with HOST arrays (host->device and device->host copies inside the timed
region).  roofline = the dominant kernel (fused barotropic substep) against the measured HBM peak.
cpu_baseline / --impl reference = the oracle restatement of the reference CPU path (the Fortran
reference cannot be built: no Fortran/MPI/netCDF/FMS in the image) on the host cores.
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
NSTEP, NFILTER = 60, 8          # barotropic substeps per btstep call
BT_CALLS_PER_STEP = 2           # predictor + corrector (MOM_dynamics_split_RK2.F90:673,:939)
BT_BYTES_PER_PT_SUBSTEP = 552   # 69 fp64 operands on the BT_cont path
STAGES = ["btstep_timeloop x2 (predictor+corrector barotropic subcycling, 68 substeps each)"]


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


def make_inputs(ni, nj):
    from mom6_b200 import synthetic
    return synthetic.bt_timeloop_inputs(ni, nj, whalo=10, halo=4, nstep=NSTEP, nfilter=NFILTER, land_blocks=40)


def synthetic_bt(ni, nj):
    return make_inputs(ni, nj)


def run_reference(args):
    """--impl reference: the oracle restatement of the reference CPU path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    cores = os.cpu_count() or 1
    dom, a = make_inputs(NI, NJ)
    times = []
    for s in range(args.warmup + args.steps):
        b = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()}
        t0 = time.perf_counter()
        for _ in range(BT_CALLS_PER_STEP):
            oracle.btstep_timeloop(dom, b, nthreads=cores)
        t1 = time.perf_counter()
        if s >= args.warmup:
            times.append(t1 - t0)
    t = sum(times)
    val = NI * NJ * NK * args.steps / t
    line = {"impl": "reference", "metric": "cell-updates/sec", "value": val, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"OM4_025-shaped {NI}x{NJ}x{NK} split-RK2 dynamics step", "stages": STAGES},
            "cpu_baseline": {"value": val, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} full steps of the same workload (oracle C++ restatement, OpenMP)"},
            "e2e": {"value": val, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

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
    from mom6_b200.api import Context

    # strong scaling: the global 1440x1080 domain is split into npi x npj tiles (one per GPU)
    npi, npj, pi, pj = tile_of(rank, world)
    ni, nj = NI // npi, NJ // npj
    dom, a = make_inputs(ni, nj)
    if world > 1:
        dom.npi, dom.npj, dom.pi, dom.pj = npi, npj, pi, pj
    ctx = Context(dom, local)
    if world > 1:
        ctx.attach_comm(dist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up then exactly K steps
    n0 = ctx.launches
    ctx.btstep_timeloop(a, reps=max(1, args.warmup * BT_CALLS_PER_STEP), download=False)
    barrier()
    n1 = ctx.launches
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        ctx.btstep_timeloop(a, reps=args.steps * BT_CALLS_PER_STEP, download=False)
        barrier()
        wall = time.perf_counter() - t0
    dev_ms = ctx.total_kernel_ms
    launches = ctx.launches - n1
    # ---- e2e: host arrays through the C ABI, copies inside the timed region
    e2e_steps = max(1, min(args.steps, 3))
    h2d = sum(v.nbytes for k, v in a.items() if isinstance(v, np.ndarray)) * BT_CALLS_PER_STEP
    outs = ["eta", "ubt", "vbt", "u_accel_bt", "v_accel_bt", "eta_wtd", "eta_sum", "ubtav", "vbtav", "uhbtav", "vhbtav",
            "ubt_wtd", "vbt_wtd"]
    d2h = sum(a[k].nbytes for k in outs) * BT_CALLS_PER_STEP
    ctx.btstep_timeloop({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a.items()})
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps * BT_CALLS_PER_STEP):
        ctx.btstep_timeloop(a)
    barrier()
    e2e_s = time.perf_counter() - t0

    tmax = torch.tensor([dev_ms, e2e_s, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, wall = [float(x) for x in tmax.cpu()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cells = NI * NJ * NK
    value = cells * args.steps / (dev_ms * 1e-3)
    e2e_val = cells * e2e_steps / e2e_s
    peak, peak_src = peaks()
    n_sub = (NSTEP + NFILTER) * BT_CALLS_PER_STEP * args.steps
    k_ms = dev_ms / n_sub                     # average substep-kernel duration (events over the timed region)
    alg_bytes = ni * nj * BT_BYTES_PER_PT_SUBSTEP
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    line = {"metric": "cell-updates/sec", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"OM4_025-shaped {NI}x{NJ}x{NK} split-RK2 dynamics step", "stages": STAGES,
                       "tiles": f"{npi}x{npj}", "l2": "inputs larger than L2 (>= 0.7 GB of 2-D planes per substep sweep)"},
            "e2e": {"value": e2e_val, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "wall_ms_per_step": 1e3 * wall / args.steps,
            "roofline": {"bound": "hbm", "kernel": "bt_substep_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes},
            "clocks": clk.summary()}
    # BASELINE.json configs[4]: btstep microbench, 4320x3240 subcycle sweep at 1 GPU
    if world == 1:
        domb, ab = synthetic_bt(4320, 3240)
        ctxb = Context(domb, local)
        ctxb.btstep_timeloop(ab, reps=3, download=False)
        msb = ctxb.last_kernel_ms
        gbs = 4320 * 3240 * (NSTEP + NFILTER) * BT_BYTES_PER_PT_SUBSTEP / (msb * 1e-3) / 1e9
        line["btstep_microbench"] = {"grid": "4320x3240", "substeps": NSTEP + NFILTER, "ms_per_call": msb,
                                     "achieved_GBps": gbs, "frac_of_peak": gbs / peak}
        ctxb.close()
    if not args.no_cpu:
        import oracle
        cores = os.cpu_count() or 1
        dom1, a1 = make_inputs(NI, NJ)
        oracle.btstep_timeloop(dom1, {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in a1.items()}, nthreads=cores)
        t0 = time.perf_counter()
        for _ in range(BT_CALLS_PER_STEP):
            oracle.btstep_timeloop(dom1, a1, nthreads=cores)
        ts = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": cells / ts, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                                "sample": "1 full step of the same workload (oracle C++ restatement of the reference, OpenMP over j)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
