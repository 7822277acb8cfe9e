"""Tile-layout invariance of the whole baroclinic step (the reference's `layout` test, .testing/Makefile:607):
step_MOM_dyn_split_RK2 on 2 (2x1, 1x2) or 4 (2x2, corner exchanges) tiles with NCCL halo exchanges == the single-tile oracle,
bit for bit.  Needs 2 / 4 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_step_multigpu.py -m gpu`."""
import os
import socket

import numpy as np
import pytest

from mom6_b200 import synthetic

STATE = ("u_inst", "v_inst", "h", "uh", "vh", "uhtr", "vhtr", "eta_av")
CSARR = ("CAu", "PFu", "PFv", "diffu", "diffv", "visc_rem_u", "u_accel_bt", "v_accel_bt", "u_av", "v_av", "h_av", "pbce", "eta", "uhbt", "vhbt")
SIZE = (64, 48, 6)


def _inner(dom, x):
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]


def _cp(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _cp(v) for k, v in x.items()}
    return x


def _worker(rank, world, port, npi, npj, q, depth_list):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mom6_b200.api import Context
    dom_g, grid_g, gv, css, cs_g, a_g = synthetic.step_dyn_inputs(*SIZE, whalo=6, land_blocks=3, store_CAu=1, calc_dtbt=1)
    dom, grid, cs, a, hv = synthetic.split_step_tile(dom_g, grid_g, cs_g, a_g, npi, npj, rank % npi, rank // npi, css["hor_visc"])
    ctx = Context(dom, rank)
    ctx.attach_comm(dist)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    ctx.set_cs_continuity(css["continuity"]); ctx.set_cs_coriolisadv(css["coriolisadv"]); ctx.set_cs_hor_visc(hv)
    ctx.set_cs_pressureforce(css["pressureforce"]); ctx.set_cs_vertvisc(css["vertvisc"])
    for _ in range(2):
        ctx.step_dyn_split_rk2(cs, a)
    out = {k: _inner(dom, a[k]).copy() for k in STATE}
    out.update({"CS%" + k: _inner(dom, cs[k]).copy() for k in CSARR})
    out["dtbt"] = cs["barotropic"]["dtbt"]
    # the reference's own layout-invariance metric: the ocean.stats line and the bit-count checksums of the final state
    so = synthetic.sum_output_cs(dom, depth_list)
    lines = []
    for n in range(2):   # the second call reports the (zero) change since the first
        e = ctx.write_energy(so, a["u_inst"], a["v_inst"], a["h"], a["T"], a["S"])
        lines.append(ctx.ocean_stats_line(so, e, n, 0.0))
    out["stats_lines"] = lines
    out["energy"] = {k: e[k] for k in ("En_mass", "mass_tot", "KE_tot", "PE_tot", "Salt", "Heat", "max_CFL", "KE", "PE", "mass_lay", "Z_0APE")}
    out["chk"] = [ctx.chksum(a["h"], 0, haloshift=0, stats=True), ctx.chksum(a["u_inst"], 1, haloshift=0, stats=True),
                  ctx.chksum(a["v_inst"], 2, haloshift=0, stats=True)]
    q.put((rank, out))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("npi,npj", [(2, 1), (1, 2), (2, 2), (4, 2), (2, 4)])
def test_step_two_tiles_bitwise(oracle, npi, npj):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < npi * npj:
        pytest.skip("needs %d GPUs" % (npi * npj))
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(*SIZE, whalo=6, land_blocks=3, store_CAu=1, calc_dtbt=1)
    for _ in range(2):
        oracle.step_dyn_split_rk2(dom, grid, gv, css, cs, a, nthreads=4)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    world = npi * npj
    depth_list = oracle.create_depth_list(dom, grid)
    procs = [ctxm.Process(target=_worker, args=(r, world, port, npi, npj, q, depth_list)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    ni, nj = SIZE[0] // npi, SIZE[1] // npj
    # ocean.stats and checksums of the tiled run == the single-tile oracle (every rank reports the global numbers)
    so = synthetic.sum_output_cs(dom, depth_list)
    ref_lines = []
    for n in range(2):
        e_ref = oracle.write_energy(dom, grid, gv, so, a["u_inst"], a["v_inst"], a["h"], a["T"], a["S"])
        ref_lines.append(oracle.ocean_stats_line(so, e_ref, n, 0.0))
    ref_chk = [oracle.chksum(dom, a["h"], 0, 0, stats=True), oracle.chksum(dom, a["u_inst"], 1, 0, stats=True),
               oracle.chksum(dom, a["v_inst"], 2, 0, stats=True)]
    for rank, out in res.items():
        assert out["stats_lines"] == ref_lines, (rank, out["stats_lines"], ref_lines)
        for k, v in out["energy"].items():
            assert np.array_equal(np.asarray(v, dtype=np.float64).view(np.int64), np.asarray(e_ref[k], dtype=np.float64).view(np.int64)), (rank, k)
        for got, want in zip(out["chk"], ref_chk):
            assert np.array_equal(got[0], want[0]) and got[1] == want[1] and np.array_equal(got[2], want[2]), (rank, got, want)
    bad = []
    for rank, out in res.items():
        pi, pj = rank % npi, rank // npi
        assert out["dtbt"] == cs["barotropic"]["dtbt"]
        for k in STATE + tuple("CS%" + c for c in CSARR):
            g = _inner(dom, a[k] if k in a else cs[k[3:]])
            ref = g[..., pj * nj:(pj + 1) * nj, pi * ni:(pi + 1) * ni]
            if not np.array_equal(ref.view(np.int64), out[k].view(np.int64)):
                bad.append((rank, k, int(np.count_nonzero(ref != out[k]))))
    assert not bad, bad


def test_split_step_tile_reassembles(oracle):
    """CPU check of the tile splitter itself: the tiles' interiors tile the global interior exactly."""
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(32, 24, 3, whalo=6)
    for npi, npj in ((2, 1), (1, 2), (2, 2)):
        ni, nj = 32 // npi, 24 // npj
        for r in range(npi * npj):
            pi, pj = r % npi, r // npi
            d, g, c, t, hv = synthetic.split_step_tile(dom, grid, cs, a, npi, npj, pi, pj, css["hor_visc"])
            assert np.array_equal(_inner(d, t["h"]), _inner(dom, a["h"])[:, pj * nj:(pj + 1) * nj, pi * ni:(pi + 1) * ni])
            assert np.array_equal(_inner(d, g["bathyT"]), _inner(dom, grid["bathyT"])[pj * nj:(pj + 1) * nj, pi * ni:(pi + 1) * ni])
            assert t["u_inst"].shape[-1] == ni + 2 * 4 + 1 and c["barotropic"]["IdxCu"].shape[-1] == ni + 2 * 6 + 1
            # the western halo of an interior tile is the eastern interior of its neighbour
            if pi > 0:
                assert np.array_equal(t["h"][:, 4:-4, :4], _inner(dom, a["h"])[:, pj * nj:(pj + 1) * nj, pi * ni - 4:pi * ni])
