import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from mom6_b200 import synthetic
from mom6_b200.api import Context
for (ni,nj) in ((1440,1080),(4320,3240)):
    dom,args = synthetic.bt_timeloop_inputs(ni,nj,whalo=10,nstep=60,nfilter=8)
    ctx = Context(dom,0)
    ctx.btstep_timeloop(args, reps=3, download=False)
    ms = ctx.last_kernel_ms
    pts = ni*nj; n=68
    print(ni,nj,'ms',ms,'GB/s alg', pts*n*552/ms/1e6, 'launches',ctx.launches, flush=True)
    ctx.close()
