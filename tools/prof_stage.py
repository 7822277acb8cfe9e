"""Per-stage timing: python tools/prof_stage.py <stage> [ni nj nk reps]   (run under ncu for captures)
Device time = CUDA events around the stage kernels (mom6cu_last_kernel_ms).  Algorithmic bytes per cell
are SURVEY 8d's figures."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mom6_b200 import synthetic
from mom6_b200.api import Context
stage = sys.argv[1]
ni = int(sys.argv[2]) if len(sys.argv) > 2 else 1440
nj = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
nk = int(sys.argv[4]) if len(sys.argv) > 4 else 75
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
BYTES = {"continuity": 96, "corad": 56, "hor_visc": 56, "pgf": 48, "remap": 32, "btstep": 136, "advect": 128, "vertvisc": 144}
t0 = time.time()
if stage == "corad":
    dom, grid, gv, cs, a = synthetic.coradcalc_inputs(ni, nj, nk, land_blocks=40)
elif stage == "continuity":
    dom, grid, gv, cs, a = synthetic.continuity_inputs(ni, nj, nk, land_blocks=40)
elif stage == "hor_visc":
    dom, grid, gv, cs, a = synthetic.hor_visc_inputs(ni, nj, nk, land_blocks=40)
elif stage == "pgf":   # MOM6CU_PGF_RECON=1|2: RECONSTRUCT_FOR_PRESSURE with PLM | PPM profiles (the reference's ALE default is 1)
    rs = int(os.environ.get("MOM6CU_PGF_RECON", "0"))
    dom, grid, gv, cs, a = synthetic.pressureforce_inputs(ni, nj, nk, land_blocks=40, **(dict(reconstruct=1, Recon_Scheme=rs, MassWghtInterp=1) if rs else {}))
elif stage == "remap":
    dom, grid, cs, a = synthetic.remap_inputs(ni, nj, nk, land_blocks=40)
    gv = synthetic.make_vgrid()
elif stage == "advect":   # 2 tracers, 1 iteration = x pass + y pass: 2 x (hprev, uhr|vhr, 2 T) x (read + write) = 128 B/cell
    dom, grid, gv, cs, a = synthetic.advect_inputs(ni, nj, nk, land_blocks=40, cfl=0.9)
elif stage == "btstep":
    dom, grid, gv, cs, a = synthetic.btstep_inputs(ni, nj, nk, whalo=10, land_blocks=40)
elif stage == "vertvisc":   # vertvisc_coef (u+v: 2 x 40 B) + vertvisc (2 x 32 B) per cell
    dom, grid, gv, cs, a, sol = synthetic.vertvisc_inputs(ni, nj, nk, land_blocks=40)
else:
    raise SystemExit("unknown stage " + stage)
print(f"inputs built in {time.time()-t0:.1f}s", flush=True)
ctx = Context(dom, 0)
ctx.set_grid(grid); ctx.set_vgrid(gv)
if stage == "corad":
    ctx.set_cs_coriolisadv(cs); run = ctx.coradcalc
elif stage == "continuity":
    ctx.set_cs_continuity(cs); run = ctx.continuity
elif stage == "hor_visc":
    ctx.set_cs_hor_visc(cs); run = ctx.horizontal_viscosity
elif stage == "pgf":
    ctx.set_cs_pressureforce(cs); run = ctx.pressure_force
elif stage == "remap":   # one tracer per call: 32 B/cell = h_old + h_new + tracer in + tracer out
    run = lambda a: ctx.ale_remap_tracers(cs, a["h_old"], a["h_new"], [a["tr"][0].copy()])
elif stage == "advect":
    def run(a):
        b = dict(a); b["tr"] = [t.copy() for t in a["tr"]]
        print("iterations", ctx.advect_tracer(cs, b))
elif stage == "btstep":
    run = lambda a: ctx.btstep(cs, a)
elif stage == "vertvisc":
    ctx.set_cs_vertvisc(cs)
    def run(a):
        ctx.vertvisc_coef(a); t1 = ctx.last_kernel_ms
        ctx.vertvisc(sol); t2 = ctx.last_kernel_ms
        print(f"vertvisc_coef {t1:.3f} ms, vertvisc {t2:.3f} ms")
for r in range(reps):
    run(a)
    ms = ctx.last_kernel_ms
    print(f"{stage} {ni}x{nj}x{nk}: {ms:.3f} ms, {ni*nj*nk/ms/1e6:.3f} Gcell/s, "
          f"{ni*nj*nk*BYTES[stage]/ms/1e6:.1f} GB/s algorithmic ({BYTES[stage]} B/cell)", flush=True)
ctx.close()
