"""GPU parity: mom6cu_continuity (through the C ABI) == oracle, bit for bit.
Reference: continuity_PPM, src/core/MOM_continuity_PPM.F90:86-194."""
import numpy as np
import pytest

from mom6_b200 import synthetic
from test_oracle_continuity import _copy

CASES = [
    # ni, nj, nk, kwargs
    (44, 40, 20, {}),                                              # double_gyre-sized, all optionals (corrector-type call)
    (44, 40, 20, dict(first_direction=1)),                         # y first
    (36, 28, 6, dict(with_uhbt=False)),                            # the BT_cont-setting call (:646)
    (36, 28, 6, dict(with_BT_cont=False, alias_h=True)),           # the corrector call with hin == h (:1043)
    (30, 22, 5, dict(with_uhbt=False, with_visc_rem=False, with_BT_cont=False)),  # plain advective call
    (40, 30, 8, dict(land_blocks=4, cyclic_y=True)),
    (40, 30, 8, dict(land_blocks=3, cs_over=dict(monotonic=1))),
    (40, 30, 8, dict(land_blocks=3, cs_over=dict(simple_2nd=1))),
    (40, 30, 4, dict(cs_over=dict(upwind_1st=1))),
    (40, 30, 8, dict(cs_over=dict(aggress_adjust=1, vol_CFL=1))),
    (40, 30, 8, dict(cs_over=dict(better_iter=0, use_visc_rem_max=0, marginal_faces=0))),
    (120, 60, 15, dict(land_blocks=8, first_direction=1)),
    (52, 36, 75, dict(land_blocks=3)),                             # OM4 layer count: 5 layers per thread in the tiled flux kernel
    (40, 30, 40, dict(land_blocks=2, with_uhbt=False)),            # 3 layers per thread
    (33, 21, 130, dict(land_blocks=2)),                            # deeper than the tiled kernel: one thread per column
]


def _flat(a):
    out = {}
    for k, v in a.items():
        if isinstance(v, np.ndarray):
            out[k] = v
        elif isinstance(v, dict):
            for kk, vv in v.items():
                out["BT_cont." + kk] = vv
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("ni,nj,nk,kw", CASES)
def test_continuity_bitwise(oracle, ctx_factory, ni, nj, nk, kw):
    dom, grid, gv, cs, a = synthetic.continuity_inputs(ni, nj, nk, **kw)
    ref = _copy(a)
    oracle.continuity(dom, grid, gv, cs, ref)
    got = _copy(a)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_continuity(cs)
    n0 = ctx.launches
    ctx.continuity(got)
    assert ctx.launches > n0
    fr, fg = _flat(ref), _flat(got)
    for k in fr:
        if k in ("u", "v", "hin", "visc_rem_u", "visc_rem_v", "uhbt", "vhbt"):
            continue
        assert np.array_equal(fr[k].view(np.int64), fg[k].view(np.int64)), (
            f"{k}: {np.count_nonzero(fr[k] != fg[k])} of {fr[k].size} differ, max |d|={np.nanmax(np.abs(fr[k] - fg[k]))}")
    assert np.abs(ref["h"] - a["hin"]).max() > 0.0


@pytest.mark.gpu
def test_continuity_requires_init_and_pairs(ctx_factory):
    from mom6_b200.api import Mom6cuError
    dom, grid, gv, cs, a = synthetic.continuity_inputs(20, 16, 3)
    ctx = ctx_factory(dom)
    with pytest.raises(Mom6cuError):          # "Module must be initialized before it is used" (:154)
        ctx.continuity(_copy(a))
    ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_continuity(cs)
    bad = _copy(a); bad["visc_rem_v"] = None
    with pytest.raises(Mom6cuError):          # :159-161
        ctx.continuity(bad)
