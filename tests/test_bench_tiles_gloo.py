"""bench.py at N > 1: every rank's inputs are its tile of ONE global synthetic field set, built on rank 0 and sent to the others
(bench.tile_inputs), so that the state checksums bench.py prints are comparable across N (the reference's layout test).  World-size-2
gloo run on the CPU: what the ranks receive == synthetic.split_step_tile of the global inputs, array for array."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from mom6_b200 import synthetic
    npi, npj = 2, 1
    dom, grid, gv, css, cs, a = bench.tile_inputs(synthetic, torch, dist, world, rank, 48, 32, npi, npj, nk=4, whalo=6)
    dom_g, grid_g, gv_g, css_g, cs_g, a_g = synthetic.step_dyn_inputs(48, 32, 4, whalo=6, land_blocks=bench.LAND_BLOCKS, store_CAu=1)
    d2, g2, c2, t2, hv2 = synthetic.split_step_tile(dom_g, grid_g, cs_g, a_g, npi, npj, rank % npi, rank // npi, css_g["hor_visc"])

    def same(x, y, path):
        if isinstance(x, dict):
            assert set(x) == set(y), (path, set(x) ^ set(y))
            return all(same(x[k], y[k], path + (k,)) for k in x)
        if isinstance(x, np.ndarray):
            assert isinstance(y, np.ndarray) and x.shape == y.shape and np.array_equal(x, y, equal_nan=True), path
            return True
        assert x == y or (x is None and y is None), (path, x, y)
        return True

    ok = same(g2, grid, ("grid",)) and same(c2, cs, ("cs",)) and same(t2, a, ("a",)) and same(hv2, css["hor_visc"], ("hv",)) and same(gv_g, gv, ("gv",))
    ok = ok and all(same(css_g[k], css[k], (k,)) for k in css_g if k != "hor_visc")
    ok = ok and all(getattr(dom, f) == getattr(d2, f) for f in ("isc", "iec", "jsc", "jec", "isd", "ied", "isdw", "iedw", "nk", "npi", "npj", "pi", "pj",
                                                                "cyclic_x", "cyclic_y", "first_direction"))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_every_rank_gets_its_tile_of_the_global_inputs():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_both_arms_print_the_same_config():
    import bench
    c = bench.workload_config()
    assert "26 barotropic substeps" in c["stages"][0] and c["land_blocks"] == bench.LAND_BLOCKS
    assert bench.bt_substeps() == (23, 3)
