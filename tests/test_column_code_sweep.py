"""Seeded sweeps of shapes and options for the host builds of the device column / cell code (csrc/mle_column.cuh, hordiff_cell.cuh,
thickdiff_column.cuh) against the oracle: ragged and tiny tiles (a single row or column of cells, two layers), closed and reentrant
edges, land fractions from none to most of the tile, and random draws of every option of the frozen sets.  Bit for bit, CPU only."""
import numpy as np
import pytest

from mom6_b200 import synthetic
from test_mle import _run_device_code_on_host as mle_host_run, _run_oracle as mle_oracle, mle_host  # noqa: F401
from test_thickness_diffuse import _run_device_code_on_host as td_host_run, _assert_same as td_same, td_host  # noqa: F401
from test_tracer_hordiff import _run_device_code_on_host as hd_host_run, _assert_same as hd_same, hd_host  # noqa: F401

SHAPES = [(1, 1), (1, 7), (9, 1), (2, 2), (5, 3), (33, 4), (17, 19), (64, 5), (40, 36)]


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_copy(v) for v in x]
    return x


def _inner(dom, x):
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]


def _draws(seed, n):
    r = np.random.default_rng(seed)
    for c in range(n):
        ni, nj = SHAPES[int(r.integers(len(SHAPES)))]
        yield c, r, dict(ni=ni, nj=nj, halo=int(r.integers(2, 6)), cyclic_x=bool(r.integers(2)), cyclic_y=bool(r.integers(2)),
                         land_blocks=int(r.integers(0, 4)), seed=int(r.integers(1, 10 ** 6)))


def test_mixedlayer_restrat_sweep(oracle, mle_host):   # noqa: F811
    moved = 0
    for c, r, g in _draws(101, 60):
        nk = int(r.choice([2, 3, 9, 24, 75]))
        kw = dict(eos=str(r.choice(["LINEAR", "WRIGHT"])), dt=float(r.choice([300.0, 900.0, 7200.0])), front=float(r.uniform(0, 6)),
                  ml_restrat_coef=float(r.choice([1.0, 5.0, 60.0])), ml_restrat_coef2=float(r.choice([0.0, 0.5, 5.0])),
                  front_length=float(r.choice([0.0, 200.0, 500.0])), MLE_MLD_decay_time=float(r.choice([0.0, 86400.0, 2.592e6])),
                  MLE_MLD_decay_time2=float(r.choice([0.0, 7.776e6])), MLE_MLD_stretch=float(r.choice([1.0, 1.5, 4.0])),
                  MLE_density_diff=float(r.choice([-9.0e9, -9.0e9, 0.03, 0.3])))
        dom, grid, gv, cs, a = synthetic.mle_inputs(g["ni"], g["nj"], nk, halo=g["halo"], seed=g["seed"], land_blocks=g["land_blocks"],
                                                    cyclic_x=g["cyclic_x"], cyclic_y=g["cyclic_y"], **kw)
        a["uhtr"][::2] = -0.0
        cref, ref = mle_oracle(oracle, dom, grid, gv, cs, a)
        got = mle_host_run(mle_host, dom, grid, gv, cs, a)
        assert np.array_equal(_inner(dom, ref["h"]).view(np.int64), _inner(dom, got["h"]).view(np.int64)), (c, g, kw)
        for k in ("uhtr", "vhtr"):
            assert np.array_equal(ref[k].view(np.int64), got[k].view(np.int64)), (c, k, g, kw)
        moved += int(not np.array_equal(ref["h"], a["h"]))
    assert moved > 10


def test_tracer_hordiff_sweep(oracle, hd_host):   # noqa: F811
    many = 0
    for c, r, g in _draws(202, 60):
        nk = int(r.choice([1, 2, 5, 30]))
        vm = int(r.integers(2))
        kw = dict(ntr=int(r.integers(1, 5)), dt=float(r.choice([900.0, 7200.0])), with_df=bool(r.integers(2)), KhTr=float(r.choice([1.0, 2000.0, 5.0e4, 3.0e5])),
                  check_diffusive_CFL=int(r.integers(2)), max_diff_CFL=float(r.choice([-1.0, 0.3, 2.5])), use_variable_mixing=vm,
                  Resoln_scaled_KhTr=int(r.integers(2)) * vm, KhTr_max=float(r.choice([0.0, 1500.0])), KhTr_min=float(r.choice([0.0, 100.0])),
                  KhTr_passivity_coeff=float(r.choice([0.0, 2.0])), KhTr_Slope_Cff=float(r.choice([0.0, 0.1])), use_MEKE_Kh=int(r.integers(2)),
                  MEKE_KhTr_fac=float(r.choice([0.5, 1.0])))
        dom, grid, gv, cs, a = synthetic.hordiff_inputs(g["ni"], g["nj"], nk, halo=g["halo"], seed=g["seed"], land_blocks=g["land_blocks"],
                                                        cyclic_x=g["cyclic_x"], cyclic_y=g["cyclic_y"], **kw)
        ref = _copy(a)
        n_ref = oracle.tracer_hordiff(dom, grid, gv, cs, ref)
        n, got = hd_host_run(hd_host, dom, grid, gv, cs, a)
        assert n == n_ref, (c, g, kw)
        hd_same(dom, ref, got, (c, g, kw))
        many += int(n_ref > 1)
    assert many > 5


def test_thickness_diffuse_sweep(oracle, td_host):   # noqa: F811
    moved = 0
    for c, r, g in _draws(303, 80):
        nk = int(r.choice([2, 3, 8, 40]))
        vm = int(r.integers(2))
        kw = dict(dt=float(r.choice([300.0, 900.0, 7200.0])), front=float(r.uniform(0, 6)), with_p_surf=bool(r.integers(2)), with_GM=bool(r.integers(2)),
                  EOS_form=int(r.choice([1, 3])), Khth=float(r.choice([10.0, 600.0, 5000.0])), Khth_Max=float(r.choice([0.0, 400.0])),
                  Khth_Min=float(r.choice([0.0, 50.0])), max_Khth_CFL=float(r.choice([0.1, 0.8])), slope_max=float(r.choice([1.0e-3, 1.0e-2])),
                  kappa_smooth=float(r.choice([0.0, 1.0e-6, 1.0e-3])), use_variable_mixing=vm, Resoln_scaled_KhTh=int(r.integers(2)) * vm,
                  use_stored_slopes=int(r.integers(2)), use_FGNV_streamfn=int(r.integers(2)), use_MEKE_Kh=int(r.integers(2)),
                  FGNV_scale=float(r.choice([0.1, 1.0])))
        dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(g["ni"], g["nj"], nk, halo=g["halo"], seed=g["seed"], land_blocks=g["land_blocks"],
                                                                  cyclic_x=g["cyclic_x"], cyclic_y=g["cyclic_y"], **kw)
        ref = _copy(a)
        oracle.thickness_diffuse(dom, grid, gv, cs, ref)
        got = td_host_run(td_host, dom, grid, gv, cs, a)
        td_same(dom, ref, got, (c, g, kw))
        moved += int(not np.array_equal(ref["h"], a["h"]))
    assert moved > 10
