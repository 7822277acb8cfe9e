"""btstep microbench (BASELINE.json configs[4]): 4320x3240 eta/ubt/vbt subcycle sweep, 60+8 substeps.
Usage: python tools/prof_bt.py [ni nj reps]   (run under ncu for the roofline capture)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mom6_b200 import synthetic
from mom6_b200.api import Context
ni = int(sys.argv[1]) if len(sys.argv) > 1 else 4320
nj = int(sys.argv[2]) if len(sys.argv) > 2 else 3240
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dom, args = synthetic.bt_timeloop_inputs(ni, nj, whalo=10, nstep=60, nfilter=8, land_blocks=40)
ctx = Context(dom, 0)
ctx.btstep_timeloop(args, reps=reps, download=False)
ms = ctx.last_kernel_ms
print(f"btstep microbench {ni}x{nj}: {ms:.3f} ms per 68-substep call, {ni*nj*68*552/ms/1e6:.1f} GB/s algorithmic "
      f"(552 B/pt/substep), launches={ctx.launches}", flush=True)
ctx.close()
