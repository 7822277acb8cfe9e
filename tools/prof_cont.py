"""continuity_PPM timing: python tools/prof_cont.py [ni nj nk reps] (run under ncu for captures).
Device time = CUDA events around the stage kernels (mom6cu_last_kernel_ms); 96 B/cell algorithmic
(5 in + 5 out + BT_cont%h_u,h_v)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mom6_b200 import synthetic
from mom6_b200.api import Context
ni = int(sys.argv[1]) if len(sys.argv) > 1 else 1440
nj = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
nk = int(sys.argv[3]) if len(sys.argv) > 3 else 75
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
t0 = time.time()
dom, grid, gv, cs, a = synthetic.continuity_inputs(ni, nj, nk, land_blocks=40)
print(f"inputs built in {time.time()-t0:.1f}s", flush=True)
ctx = Context(dom, 0)
ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_continuity(cs)
for r in range(reps):
    ctx.continuity(a)
    ms = ctx.last_kernel_ms
    print(f"continuity {ni}x{nj}x{nk}: {ms:.3f} ms, {ni*nj*nk/ms/1e6:.3f} Gcell/s, {ni*nj*nk*96/ms/1e6:.1f} GB/s algorithmic (96 B/cell)", flush=True)
ctx.close()
