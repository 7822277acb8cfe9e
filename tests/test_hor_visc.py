"""horizontal_viscosity (src/parameterizations/lateral/MOM_hor_visc.F90:266-2317).
CPU: properties of the oracle restatement (no acceleration at rest or in solid translation; dissipation of kinetic
energy).  GPU: mom6cu_horizontal_viscosity (through the C ABI) == oracle, bit for bit, per option set."""
import numpy as np
import pytest

from mom6_b200 import synthetic
from test_oracle_continuity import _copy, _comp


def test_uniform_flow_has_no_stress(oracle):
    dom, grid, gv, cs, a = synthetic.hor_visc_inputs(24, 20, 2, cyclic_y=True)
    a = _copy(a)
    a["u"][...] = 0.0; a["v"][...] = 0.0
    oracle.horizontal_viscosity(dom, grid, gv, cs, a)
    assert np.abs(_comp(dom, a["diffu"], "u")).max() == 0.0 and np.abs(_comp(dom, a["diffv"], "v")).max() == 0.0


@pytest.mark.parametrize("kw", [dict(), dict(Laplacian=True, biharmonic=False, Kh=500.0, Smagorinsky_Kh=True)])
def test_viscosity_dissipates_energy(oracle, kw):
    """sum_faces h_u*areaCu*u*diffu + (v) < 0: the stress tensor form is energetically dissipative."""
    dom, grid, gv, cs, a = synthetic.hor_visc_inputs(32, 24, 3, land_blocks=2, **kw)
    a = _copy(a)
    oracle.horizontal_viscosity(dom, grid, gv, cs, a)
    h = a["h"]
    hu = np.zeros_like(a["u"]); hu[:, :, 1:-1] = 0.5 * (h[:, :, :-1] + h[:, :, 1:])
    hv = np.zeros_like(a["v"]); hv[:, 1:-1, :] = 0.5 * (h[:, :-1, :] + h[:, 1:, :])
    wu = _comp(dom, hu * grid["areaCu"][None] * a["u"] * a["diffu"], "u")[:, :, 1:]
    wv = _comp(dom, hv * grid["areaCv"][None] * a["v"] * a["diffv"], "v")[:, 1:, :]
    assert np.isfinite(wu).all() and np.isfinite(wv).all()
    assert wu.sum() + wv.sum() < 0.0


CASES = [
    dict(),                                                                    # biharmonic Smagorinsky, better bounds (benchmark-like)
    dict(land_blocks=4, Ah=1.0e11),
    dict(land_blocks=4, Laplacian=True, Kh=800.0, Smagorinsky_Kh=True),         # Laplacian + biharmonic, both better-bounded
    dict(land_blocks=3, Laplacian=True, biharmonic=False, Kh_vel_scale=0.01),   # tc-like Laplacian only
    dict(land_blocks=3, Laplacian=True, Smagorinsky_Kh=True, better_bound_Kh=False, better_bound_Ah=False),  # legacy bounds
    dict(land_blocks=3, bound_Coriolis=True),
    dict(land_blocks=3, no_slip=True, Laplacian=True, Kh=300.0),
    dict(land_blocks=3, use_land_mask=True, add_LES_viscosity=True, Laplacian=True, Smagorinsky_Kh=True, Kh_bg_min=50.0),
    dict(land_blocks=3, Re_Ah=20.0),
    dict(land_blocks=3, cont_thick=True),
    dict(land_blocks=2, Laplacian=True, better_bound_Ah=False, Kh=100.0),       # better_bound_Kh only
    dict(land_blocks=2, bound_Ah=False, better_bound_Ah=False, Smagorinsky_Ah=False, Ah_vel_scale=0.01, cyclic_y=True),
]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_hor_visc_bitwise(oracle, ctx_factory, kw):
    dom, grid, gv, cs, a = synthetic.hor_visc_inputs(44, 40, 5, **kw)
    ref = _copy(a)
    oracle.horizontal_viscosity(dom, grid, gv, cs, ref)
    got = _copy(a)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_hor_visc(cs)
    n0 = ctx.launches
    ctx.horizontal_viscosity(got)
    assert ctx.launches > n0
    bad = [f"{k}: {np.count_nonzero(ref[k] != got[k])} of {ref[k].size} differ, max |d|={np.nanmax(np.abs(ref[k] - got[k]))}"
           for k in ("diffu", "diffv") if not np.array_equal(ref[k].view(np.int64), got[k].view(np.int64))]
    assert not bad, "; ".join(bad)
    assert np.abs(ref["diffu"]).max() > 0


@pytest.mark.gpu
def test_hor_visc_ragged_sizes(oracle, ctx_factory):
    for ni, nj, nk, kw in ((33, 17, 1, dict(cyclic_y=True)), (97, 71, 3, dict(land_blocks=5)), (360, 180, 2, dict(land_blocks=12))):
        dom, grid, gv, cs, a = synthetic.hor_visc_inputs(ni, nj, nk, **kw)
        ref = _copy(a)
        oracle.horizontal_viscosity(dom, grid, gv, cs, ref)
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_hor_visc(cs)
        ctx.horizontal_viscosity(a)
        for k in ("diffu", "diffv"):
            assert np.array_equal(ref[k].view(np.int64), a[k].view(np.int64)), (ni, nj, k)


@pytest.mark.gpu
def test_hor_visc_rejects_unsupported_and_uninitialised(ctx_factory):
    from mom6_b200.api import Mom6cuError
    dom, grid, gv, cs, a = synthetic.hor_visc_inputs(20, 16, 2)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    with pytest.raises(Mom6cuError):          # "Module must be initialized before it is used" (:504)
        ctx.horizontal_viscosity(a)
    bad = dict(cs); bad["unsupported"] = 1
    with pytest.raises(Mom6cuError):
        ctx.set_cs_hor_visc(bad)
