"""CorAdCalc (src/core/MOM_CoriolisAdv.F90:125-965).
CPU: properties of the oracle restatement that the reference's schemes guarantee.
GPU: mom6cu_coradcalc (through the C ABI) == oracle, bit for bit, for every scheme."""
import numpy as np
import pytest

from mom6_b200 import synthetic, fidx
from mom6_b200 import _lib as L
from test_oracle_continuity import _copy, _comp


def test_rest_state(oracle):
    """u = v = 0 gives CAu = CAv = 0."""
    dom, grid, gv, cs, a = synthetic.coradcalc_inputs(24, 20, 3)
    a = _copy(a)
    for k in ("u", "v", "uh", "vh"):
        a[k][...] = 0.0
    oracle.coradcalc(dom, grid, gv, cs, a)
    assert np.abs(_comp(dom, a["CAu"], "u")).max() == 0.0 and np.abs(_comp(dom, a["CAv"], "v")).max() == 0.0


def test_sadourny_energy_conserves_energy(oracle):
    """Sadourny (1975) energy scheme: the Coriolis part does no work, sum(uh*CAu_cor*dx + vh*CAv_cor*dy) = 0
    (checked with the KE gradient removed through the gradKE diagnostics) on a reentrant channel with islands."""
    dom, grid, gv, cs, a = synthetic.coradcalc_inputs(32, 24, 2, land_blocks=2, diags=True)
    a = _copy(a)
    oracle.coradcalc(dom, grid, gv, cs, a)
    cor_u = a["CAu"] - a["gradKEu"]
    cor_v = a["CAv"] - a["gradKEv"]
    # work = sum_u uh * CAu * dxCu + sum_v vh * CAv * dyCv over the periodic computational domain (each face once)
    iu = (slice(None), slice(dom.jsc - dom.jsd, dom.jec - dom.jsd + 1), slice(dom.isc - dom.isd + 1, dom.iec - dom.isd + 2))
    iv = (slice(None), slice(dom.jsc - dom.jsd + 1, dom.jec - dom.jsd + 2), slice(dom.isc - dom.isd, dom.iec - dom.isd + 1))
    wu = (a["uh"] * cor_u * grid["dxCu"][None])[iu]
    wv = (a["vh"] * cor_v * grid["dyCv"][None])[iv]
    scale = np.abs(wu).sum() + np.abs(wv).sum()
    assert scale > 0
    assert abs(wu.sum() + wv.sum()) <= 1e-12 * scale


SCHEMES = [
    dict(),
    dict(Coriolis_Scheme=L.ARAKAWA_HSU90),
    dict(Coriolis_Scheme=L.SADOURNY75_ENSTRO),
    dict(Coriolis_Scheme=L.ARAKAWA_LAMB81),
    dict(Coriolis_Scheme=L.AL_BLEND),
    dict(Coriolis_Scheme=L.AL_BLEND, F_eff_max_blend=2.5, wt_lin_blend=0.3),
    dict(Coriolis_Scheme=L.ROBUST_ENSTRO),
    dict(Coriolis_Scheme=L.ROBUST_ENSTRO, PV_Adv_Scheme=L.PV_ADV_UPWIND1),
    dict(bound_Coriolis=1),                               # tc1
    dict(Coriolis_En_Dis=1),                              # tc4
    dict(Coriolis_Scheme=L.ARAKAWA_HSU90, bound_Coriolis=1, no_slip=1),
    dict(KE_Scheme=L.KE_SIMPLE_GUDONOV),
    dict(KE_Scheme=L.KE_GUDONOV, Coriolis_Scheme=L.ARAKAWA_LAMB81),
]


def test_schemes_agree_for_uniform_flow(oracle):
    """All Coriolis discretisations are consistent: for uniform h and uniform flow they agree closely."""
    dom, grid, gv, cs0, a0 = synthetic.coradcalc_inputs(24, 20, 2)
    res = []
    for over in SCHEMES[:5]:
        a = _copy(a0)
        a["h"][...] = 100.0
        a["u"][...] = 0.1 * grid["mask2dCu"][None]; a["v"][...] = 0.05 * grid["mask2dCv"][None]
        a["uh"][...] = a["u"] * 100.0 * grid["dy_Cu"][None]; a["vh"][...] = a["v"] * 100.0 * grid["dx_Cv"][None]
        cs = synthetic.coriolisadv_cs(**over)
        oracle.coradcalc(dom, grid, gv, cs, a)
        res.append(_comp(dom, a["CAu"], "u")[:, 4:-4, 4:-4].copy())
    for r in res[1:]:
        assert np.allclose(r, res[0], rtol=0, atol=2e-2 * np.abs(res[0]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("over", SCHEMES)
def test_coradcalc_bitwise(oracle, ctx_factory, over):
    dom, grid, gv, cs, a = synthetic.coradcalc_inputs(44, 40, 6, land_blocks=3, cs_over=over, diags=True,
                                                      por=bool(over.get("Coriolis_En_Dis")))
    ref = _copy(a)
    oracle.coradcalc(dom, grid, gv, cs, ref)
    got = _copy(a)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_coriolisadv(cs)
    n0 = ctx.launches
    ctx.coradcalc(got)
    assert ctx.launches > n0
    bad = [f"{k}: {np.count_nonzero(ref[k] != got[k])} of {ref[k].size} differ, max |d|={np.nanmax(np.abs(ref[k] - got[k]))}"
           for k in ("CAu", "CAv", "RV", "PV", "gradKEu", "gradKEv")
           if not np.array_equal(ref[k].view(np.int64), got[k].view(np.int64))]
    assert not bad, "; ".join(bad)
    assert np.abs(ref["CAu"]).max() > 0


@pytest.mark.gpu
def test_coradcalc_ragged_and_large(oracle, ctx_factory):
    """Tile-edge cases: sizes that are not multiples of the CTA tile, cyclic in y, one layer."""
    for ni, nj, nk, kw in ((33, 9, 1, dict(cyclic_y=True)), (97, 71, 3, dict(land_blocks=5)), (360, 180, 2, dict(land_blocks=12))):
        dom, grid, gv, cs, a = synthetic.coradcalc_inputs(ni, nj, nk, **kw)
        ref = _copy(a)
        oracle.coradcalc(dom, grid, gv, cs, ref)
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_coriolisadv(cs)
        ctx.coradcalc(a)
        for k in ("CAu", "CAv"):
            assert np.array_equal(ref[k].view(np.int64), a[k].view(np.int64)), k


@pytest.mark.gpu
def test_coradcalc_requires_init(ctx_factory):
    from mom6_b200.api import Mom6cuError
    dom, grid, gv, cs, a = synthetic.coradcalc_inputs(20, 16, 2)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    with pytest.raises(Mom6cuError):          # "Module must be initialized before it is used" (:236)
        ctx.coradcalc(a)
