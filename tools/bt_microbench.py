"""The barotropic microbench of BASELINE.json configs[4] alone: 4320x3240, 60 + 8 substeps, BT_cont transports; prints ms and achieved GB/s on the
algorithmic 552 B/pt/substep.  MOM6CU_BT_OPT selects the kernel variant (bt_timeloop.cu).  usage: python tools/bt_microbench.py [ni nj]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mom6_b200 import synthetic
from mom6_b200.api import Context
ni = int(sys.argv[1]) if len(sys.argv) > 1 else 4320
nj = int(sys.argv[2]) if len(sys.argv) > 2 else 3240
dom, args = synthetic.bt_timeloop_inputs(ni, nj, whalo=10, halo=4, nstep=60, nfilter=8, land_blocks=40)
ctx = Context(dom, 0)
ctx.btstep_timeloop(args, reps=3, download=False)
ms = ctx.last_kernel_ms
print(f"MOM6CU_BT_OPT={os.environ.get('MOM6CU_BT_OPT', 'default')} {ni}x{nj}: {ms:.2f} ms per 68 substeps, {ni * nj * 68 * 552 / ms / 1e6:.1f} GB/s algorithmic", flush=True)
ctx.close()
