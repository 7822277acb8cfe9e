"""Multi-GPU layout invariance (the reference's `layout` test, .testing/Makefile:607): btstep_timeloop
on npi x npj tiles with NCCL halo exchanges == the single-tile oracle, bit for bit.
Needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_bt_multigpu.py -m gpu`."""
import os
import socket

import numpy as np
import pytest

from mom6_b200 import synthetic, fidx

KEYS = [("eta", "h", True), ("ubt", "u", True), ("vbt", "v", True), ("eta_wtd", "h", True), ("u_accel_bt", "u", True),
        ("v_accel_bt", "v", True), ("uhbtav", "u", False), ("vhbtav", "v", False), ("ubt_wtd", "u", False),
        ("vbt_wtd", "v", False), ("eta_sum", "h", True)]


def _comp(dom, arr, st, wide):
    ilo, ihi, jlo, jhi = fidx.extent(dom, st, wide)
    su = 1 if st == "u" else 0
    sv = 1 if st == "v" else 0
    return arr[dom.jsc - sv - jlo: dom.jec - jlo + 1, dom.isc - su - ilo: dom.iec - ilo + 1]


def _worker(rank, world, port, npi, npj, NI, NJ, whalo, nstep, nfilter, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mom6_b200.api import Context
    dom_g, a_g = synthetic.bt_timeloop_inputs(NI, NJ, whalo=whalo, nstep=nstep, nfilter=nfilter, land_blocks=4)
    dom, a = synthetic.split_tile(dom_g, a_g, npi, npj, rank % npi, rank // npi)
    ctx = Context(dom, rank)
    ctx.attach_comm(dist)
    ctx.btstep_timeloop(a)
    out = {k: _comp(dom, a[k], st, w).copy() for k, st, w in KEYS}
    q.put((rank, out))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("npi,npj", [(2, 1), (1, 2), (2, 2), (4, 2)])
def test_bt_timeloop_two_tiles_bitwise(oracle, npi, npj):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < npi * npj:
        pytest.skip("needs %d GPUs" % (npi * npj))
    NI, NJ, whalo, nstep, nfilter = 96, 64, 6, 17, 4
    dom_g, a_g = synthetic.bt_timeloop_inputs(NI, NJ, whalo=whalo, nstep=nstep, nfilter=nfilter, land_blocks=4)
    oracle.btstep_timeloop(dom_g, a_g)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    world = npi * npj
    procs = [ctxm.Process(target=_worker, args=(r, world, port, npi, npj, NI, NJ, whalo, nstep, nfilter, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    ni, nj = NI // npi, NJ // npj
    for rank, out in res.items():
        pi, pj = rank % npi, rank // npi
        for k, st, w in KEYS:
            g = _comp(dom_g, a_g[k], st, w)
            su = 1 if st == "u" else 0
            sv = 1 if st == "v" else 0
            ref = g[pj * nj: pj * nj + nj + sv, pi * ni: pi * ni + ni + su]
            assert np.array_equal(ref, out[k]), f"rank {rank} {k}"
