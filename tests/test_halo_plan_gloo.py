"""N>1 host logic on CPU: the neighbour/box planning used by the NCCL halo exchange
(mom6cu_halo_plan, mom6_b200/csrc/halo_nccl.cu) driven over gloo with world_size 2 and 4, checked
against a single-tile halo fill of the global field (the reference's `layout` test idea,
.testing/Makefile:607: 1 PE vs LAYOUT=2,1 must agree)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mom6_b200 import _lib, fidx
from mom6_b200.api import make_domain

OPP = [1, 0, 3, 2, 5, 4, 7, 6]
ST = {"h": 0, "u": 1, "v": 2, "q": 3}


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _global_field(NI, NJ, st, seed):
    """a globally defined, periodic-consistent function of the global index"""
    r = np.random.default_rng(seed)
    su = 1 if st in ("u", "q") else 0
    sv = 1 if st in ("v", "q") else 0
    g = r.uniform(-1, 1, size=(NJ, NI))
    def val(gi, gj):  # global 1-based point index -> value, periodic
        return g[(gj - 1) % NJ, (gi - 1) % NI]
    return val


def _worker(rank, world, port, npi, npj, NI, NJ, halo, cyc_x, cyc_y, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _lib.load()
    ni, nj = NI // npi, NJ // npj
    pi, pj = rank % npi, rank // npi
    dom = make_domain(ni, nj, halo=halo, cyclic_x=cyc_x, cyclic_y=cyc_y, npi=npi, npj=npj, pi=pi, pj=pj)
    ok = True
    for st in ("h", "u", "v", "q"):
        val = _global_field(NI, NJ, st, 11 + ST[st])
        f = fidx.new(dom, st)
        su = 1 if st in ("u", "q") else 0
        sv = 1 if st in ("v", "q") else 0
        # computational (symmetric) domain from the global function; halos start as NaN
        f.a[...] = np.nan
        for j in range(dom.jsc - sv, dom.jec + 1):
            for i in range(dom.isc - su, dom.iec + 1):
                f.s(i, i, j, j)[...] = val(pi * ni + i - dom.isc + 1, pj * nj + j - dom.jsc + 1)
        sb, rb = (C.c_int * 4)(), (C.c_int * 4)()
        sends, recvs = [], []
        for d in range(8):
            p = lib.mom6cu_halo_plan(C.byref(dom), ST[st], 0, -1, d, sb, rb)
            if p >= 0:
                sends.append((d, p, f.s(*sb).copy()))
        for d in range(8):
            rd = OPP[d]
            p = lib.mom6cu_halo_plan(C.byref(dom), ST[st], 0, -1, rd, sb, rb)
            if p >= 0:
                recvs.append((rd, p, tuple(rb)))
        reqs = []
        for d, p, buf in sends:
            if p != rank:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(buf)), p, tag=0))
        self_msgs = {OPP[d]: buf for d, p, buf in sends if p == rank}
        for rd, p, box in recvs:
            tgt = f.s(*box)
            if p == rank:
                tgt[...] = self_msgs[rd]
            else:
                t = torch.empty(tgt.shape, dtype=torch.float64)
                dist.recv(t, p, tag=0)
                tgt[...] = t.numpy()
        for r_ in reqs:
            r_.wait()
        # expected: every halo point reachable through a periodic / interior neighbour holds the global value
        for j in range(f.jlo, f.jhi + 1):
            for i in range(f.ilo, f.ihi + 1):
                gi, gj = pi * ni + i - dom.isc + 1, pj * nj + j - dom.jsc + 1
                in_x = (1 - su <= gi <= NI) or cyc_x
                in_y = (1 - sv <= gj <= NJ) or cyc_y
                v = f.s(i, i, j, j)[0, 0]
                if in_x and in_y:
                    if not v == val(gi, gj):
                        ok = False
                else:
                    if not np.isnan(v):
                        ok = False
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("npi,npj,cyc_x,cyc_y", [(2, 1, True, False), (1, 2, True, True), (2, 2, True, False),
                                                 (2, 1, False, False), (4, 2, True, False)])   # 4x2: the 8-GPU layout of bench.py
def test_halo_plan_matches_global_fill(npi, npj, cyc_x, cyc_y):
    world = npi * npj
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, npi, npj, 24, 16, 3, cyc_x, cyc_y, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
