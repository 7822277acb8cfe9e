"""tracer_hordiff with the Eady (KHTR_SLOPE_CFF) and MEKE diffusivity terms on the device: C ABI == oracle, bit for bit.  The face code is
already checked against the oracle on the host (tests/test_tracer_hordiff.py); this GPU half was written after the round's GPU budget was
spent, so it has not run on a B200 yet and is named to sort last: a failure here cannot mask the verified tests under `-x`."""
import pytest

from mom6_b200 import synthetic
from test_tracer_hordiff import EXT_CASES, _assert_same, _copy


@pytest.mark.gpu
@pytest.mark.parametrize("kw", EXT_CASES)
def test_tracer_hordiff_ext_bitwise(oracle, ctx_factory, kw):
    for (ni, nj, nk) in ((44, 40, 20), (131, 9, 3)):
        dom, grid, gv, cs, a = synthetic.hordiff_inputs(ni, nj, nk, **kw)
        ref = _copy(a)
        n_ref = oracle.tracer_hordiff(dom, grid, gv, cs, ref)
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv)
        assert ctx.tracer_hordiff(cs, a) == n_ref
        _assert_same(dom, ref, a, kw)
