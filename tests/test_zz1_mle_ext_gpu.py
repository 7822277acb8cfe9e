"""mixedlayer_restrat with the mixed-layer depth detected from the density profile (MLE_DENSITY_DIFF > 0, detect_mld) on the device: C ABI
== oracle, bit for bit.  The column code is already checked against the oracle on the host (tests/test_mle.py); this GPU half was written
after the round's GPU budget was spent, so it has not run on a B200 yet and is named to sort last: a failure here cannot mask the verified
tests under `-x`."""
import numpy as np
import pytest

from mom6_b200 import synthetic
from test_mle import EXT_CASES, _inner, _run_oracle


@pytest.mark.gpu
@pytest.mark.parametrize("kw", EXT_CASES)
def test_mixedlayer_restrat_detect_mld_bitwise(oracle, ctx_factory, kw):
    for (ni, nj, nk) in ((44, 40, 20), (30, 22, 75)):
        dom, grid, gv, cs, a = synthetic.mle_inputs(ni, nj, nk, **kw)
        c, o = _run_oracle(oracle, dom, grid, gv, cs, a)
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv)
        ctx.mixedlayer_restrat(cs, a["h"], a["uhtr"], a["vhtr"], a["T"], a["S"], a["ustar"], a["dt"], None, a["Rd_dx_h"])
        assert np.array_equal(_inner(dom, o["h"]).view(np.int64), _inner(dom, a["h"]).view(np.int64)), kw
        for k in ("uhtr", "vhtr"):
            assert np.array_equal(o[k].view(np.int64), a[k].view(np.int64)), (k, kw)
