"""thickness_diffuse on the device: C ABI == oracle, bit for bit (the CPU half is tests/test_thickness_diffuse.py, which already runs the
column / face code the kernels call, compiled for the host, against the oracle).  This file was written after the round's GPU budget was
spent, so it has not run on a B200 yet; it is named to sort last so that a failure here cannot mask the verified tests under `-x`."""
import numpy as np
import pytest

from mom6_b200 import synthetic
from test_thickness_diffuse import CASES, _assert_same, _copy


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_thickness_diffuse_bitwise(oracle, ctx_factory, kw):
    for (ni, nj, nk) in ((44, 40, 20), (131, 9, 2), (30, 22, 75)):
        dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(ni, nj, nk, **kw)
        ref = _copy(a)
        oracle.thickness_diffuse(dom, grid, gv, cs, ref)
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv)
        n0 = ctx.launches
        ctx.thickness_diffuse(cs, a)
        assert ctx.launches - n0 >= 4
        _assert_same(dom, ref, a, kw)


@pytest.mark.gpu
def test_thickness_diffuse_errors(ctx_factory):
    from mom6_b200.api import Mom6cuError
    dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(16, 12, 5)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    for bad in (dict(use_FGNV_streamfn=1), dict(use_stored_slopes=1), dict(use_MEKE_Kh=1), dict(EOS_form=0), dict(find_work=1)):
        with pytest.raises(Mom6cuError):
            ctx.thickness_diffuse(dict(cs, **bad), a)
