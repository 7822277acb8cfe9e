"""thickness_diffuse with stored slopes / the FGNV streamfunction / the MEKE diffusivity on the device: C ABI == oracle, bit for bit.  The
column code these kernels call (m6td::face_ext) is already checked against the oracle on the host (tests/test_thickness_diffuse.py,
tests/test_column_code_sweep.py); this GPU half was written after the round's GPU budget was spent, so it has not run on a B200 yet and
is named to sort last: a failure here cannot mask the verified tests under `-x`."""
import pytest

from mom6_b200 import synthetic
from test_thickness_diffuse import EXT_CASES, _assert_same, _copy


@pytest.mark.gpu
@pytest.mark.parametrize("kw", EXT_CASES)
def test_thickness_diffuse_ext_bitwise(oracle, ctx_factory, kw):
    for (ni, nj, nk) in ((44, 40, 20), (131, 9, 2), (30, 22, 75)):
        dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(ni, nj, nk, **kw)
        ref = _copy(a)
        oracle.thickness_diffuse(dom, grid, gv, cs, ref)
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv)
        ctx.thickness_diffuse(cs, a)
        _assert_same(dom, ref, a, kw)
