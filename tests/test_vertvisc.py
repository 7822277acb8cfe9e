"""vertvisc_coef / vertvisc / vertvisc_remnant (src/parameterizations/vertical/MOM_vert_friction.F90:1357, :557, :1229).
CPU: properties of the oracle restatement (parity unpinned).  GPU: C ABI == oracle, bit for bit."""
import numpy as np
import pytest

from mom6_b200 import synthetic
from mom6_b200 import fidx


def _coefs(dom, nk):
    return (np.zeros((nk + 1,) + fidx.new(dom, "u").a.shape), np.zeros((nk + 1,) + fidx.new(dom, "v").a.shape),
            fidx.new(dom, "u", nk=nk).a, fidx.new(dom, "v", nk=nk).a)


def _cu(dom, x):   # u-points (IscB:IecB, jsc:jec)
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 2]


def _cv(dom, x):   # v-points (isc:iec, JscB:JecB)
    return x[..., dom.jsc - dom.jsd:dom.jec - dom.jsd + 2, dom.isc - dom.isd:dom.iec - dom.isd + 1]


def test_solver_conserves_momentum_without_stress_and_drag(oracle):
    """With no wind stress and no bottom coupling the tridiagonal solve only redistributes h_u*u within a column."""
    dom, grid, gv, cs, coef, sol = synthetic.vertvisc_inputs(24, 18, 8, land_blocks=2, bottomdraglaw=0)
    cs = dict(cs, Kv=1.0e-2)
    a_u, a_v, h_u, h_v = _coefs(dom, 8)
    oracle.vertvisc_coef(dom, grid, gv, cs, coef, a_u, a_v, h_u, h_v)
    a_u[-1] = 0.0; a_v[-1] = 0.0                          # no bottom drag
    sol["taux"][...] = 0.0; sol["tauy"][...] = 0.0
    u0, v0 = sol["u"].copy(), sol["v"].copy()
    oracle.vertvisc(dom, grid, gv, cs, sol, a_u, a_v, h_u, h_v)
    mu0, mu1 = (_cu(dom, h_u) * _cu(dom, u0)).sum(axis=0), (_cu(dom, h_u) * _cu(dom, sol["u"])).sum(axis=0)
    assert np.allclose(mu0, mu1, rtol=1e-12, atol=1e-12) and not np.array_equal(u0, sol["u"])
    mv0, mv1 = (_cv(dom, h_v) * _cv(dom, v0)).sum(axis=0), (_cv(dom, h_v) * _cv(dom, sol["v"])).sum(axis=0)
    assert np.allclose(mv0, mv1, rtol=1e-12, atol=1e-12)
    assert np.abs(_cu(dom, sol["taux_bot"])).max() == 0.0


def test_remnant_is_the_response_to_a_unit_velocity(oracle):
    """visc_rem is what is left of a unit velocity after the implicit step: in (0, 1], smaller near the bottom drag,
    and equal to vertvisc applied to u = 1 with no stress."""
    dom, grid, gv, cs, coef, sol = synthetic.vertvisc_inputs(24, 18, 8, land_blocks=2, with_Ray=True)
    a_u, a_v, h_u, h_v = _coefs(dom, 8)
    oracle.vertvisc_coef(dom, grid, gv, cs, coef, a_u, a_v, h_u, h_v)
    vru, vrv = np.zeros_like(sol["u"]), np.zeros_like(sol["v"])
    oracle.vertvisc_remnant(dom, grid, cs, vru, vrv, sol["dt"], a_u, a_v, h_u, h_v, sol["Ray_u"], sol["Ray_v"])
    m = (_cu(dom, grid["mask2dCu"]) > 0)[None]
    r = _cu(dom, vru)[np.broadcast_to(m, _cu(dom, vru).shape)]
    assert r.min() > 0.0 and r.max() <= 1.0
    one = dict(sol, u=np.ones_like(sol["u"]), v=np.ones_like(sol["v"]), taux=np.zeros_like(sol["taux"]), tauy=np.zeros_like(sol["tauy"]),
               taux_bot=None, tauy_bot=None)
    oracle.vertvisc(dom, grid, gv, cs, one, a_u, a_v, h_u, h_v)
    assert np.allclose(_cu(dom, one["u"])[np.broadcast_to(m, _cu(dom, vru).shape)], r, rtol=1e-13)


def test_coupling_coefficients_feel_the_bottom_boundary_layer(oracle):
    dom, grid, gv, cs, coef, sol = synthetic.vertvisc_inputs(24, 18, 8, with_shear=False)
    a_u, a_v, h_u, h_v = _coefs(dom, 8)
    oracle.vertvisc_coef(dom, grid, gv, cs, coef, a_u, a_v, h_u, h_v)
    m = _cu(dom, grid["mask2dCu"]) > 0
    au = _cu(dom, a_u)
    assert (au[0] == 0).all() and (au[1:, m] > 0).all()
    assert (au[-1][m] > 5.0 * au[2][m]).mean() > 0.9           # kv_bbl / min(h/2, bbl_thick) >> Kv / h in the interior
    nodrag = dict(cs, bottomdraglaw=0)
    b_u, b_v, g_u, g_v = _coefs(dom, 8)
    oracle.vertvisc_coef(dom, grid, gv, nodrag, coef, b_u, b_v, g_u, g_v)
    assert (_cu(dom, b_u)[-1][m] < au[-1][m]).all()


CASES = [dict(), dict(land_blocks=4, with_Bu=True, with_Ray=True), dict(harmonic_visc=1, land_blocks=3), dict(bottomdraglaw=0, Kv_extra_bbl=5e-3),
         dict(bottomdraglaw=0, land_blocks=2, cyclic_y=True), dict(Kvml_invZ2=1e-3, land_blocks=2), dict(harm_BL_val=0.5, with_Ray=True),
         dict(fixed_LOTW_ML=1, land_blocks=3), dict(apply_LOTW_floor=1), dict(fixed_LOTW_ML=1, apply_LOTW_floor=1, Kvml_invZ2=1e-3, harmonic_visc=1),
         dict(direct_stress=1, land_blocks=2, with_Ray=True),
         # vertvisc_limit_vel (:2926-3120): CFL-based truncation, MAXVEL truncation, VEL_UNDERFLOW
         dict(land_blocks=2, dt=60000.0), dict(land_blocks=1, dt=30000.0, CFL_trunc=0.2, with_Ray=True),
         dict(land_blocks=2, CFL_based_trunc=0, maxvel=0.2), dict(land_blocks=2, vel_underflow=0.05)]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_vertvisc_family_bitwise(oracle, ctx_factory, kw):
    for (ni, nj, nk) in ((44, 40, 12), (130, 7, 3)):
        dom, grid, gv, cs, coef, sol = synthetic.vertvisc_inputs(ni, nj, nk, **kw)
        a_u, a_v, h_u, h_v = _coefs(dom, nk)
        oracle.vertvisc_coef(dom, grid, gv, cs, coef, a_u, a_v, h_u, h_v)
        ref = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in sol.items()}
        ntrunc = oracle.vertvisc(dom, grid, gv, cs, ref, a_u, a_v, h_u, h_v)
        rru, rrv = np.zeros_like(sol["u"]), np.zeros_like(sol["v"])
        oracle.vertvisc_remnant(dom, grid, cs, rru, rrv, sol["dt"], a_u, a_v, h_u, h_v, sol["Ray_u"], sol["Ray_v"])
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_vertvisc(cs)
        n0 = ctx.launches
        ctx.vertvisc_coef(coef)
        assert ctx.launches > n0
        g = _coefs(dom, nk)
        ctx.vertvisc_get_coef(*g)
        for name, x, y, cut in (("a_u", a_u, g[0], _cu), ("a_v", a_v, g[1], _cv), ("h_u", h_u, g[2], _cu), ("h_v", h_v, g[3], _cv)):
            assert np.array_equal(cut(dom, x).view(np.int64), cut(dom, y).view(np.int64)), (name, kw, np.count_nonzero(cut(dom, x) != cut(dom, y)))
        got = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in sol.items()}
        ctx.vertvisc(got)
        assert ctx.vertvisc_ntrunc() == ntrunc, (kw, ntrunc)
        if "dt" in kw or "maxvel" in kw:
            assert ntrunc > 0, kw
        for name, cut in (("u", _cu), ("v", _cv), ("taux_bot", _cu), ("tauy_bot", _cv)):
            assert np.array_equal(cut(dom, ref[name]).view(np.int64), cut(dom, got[name]).view(np.int64)), (name, kw)
        gru, grv = np.zeros_like(sol["u"]), np.zeros_like(sol["v"])
        ctx.vertvisc_remnant(gru, grv, sol["dt"], sol["Ray_u"], sol["Ray_v"])
        assert np.array_equal(_cu(dom, rru).view(np.int64), _cu(dom, gru).view(np.int64)) and np.array_equal(_cv(dom, rrv).view(np.int64), _cv(dom, grv).view(np.int64))
        assert not np.array_equal(ref["u"], sol["u"])


@pytest.mark.gpu
def test_vertvisc_errors(ctx_factory):
    from mom6_b200.api import Mom6cuError
    dom, grid, gv, cs, coef, sol = synthetic.vertvisc_inputs(16, 12, 4)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv)
    with pytest.raises(Mom6cuError):
        ctx.vertvisc_coef(coef)                          # not initialised (:1479)
    for bad in (dict(unsupported=1), dict(dynamic_viscous_ML=1), dict(nkml=2), dict(answer_date=20181231)):
        with pytest.raises(Mom6cuError):
            ctx.set_cs_vertvisc(dict(cs, **bad))
    ctx.set_cs_vertvisc(cs)
    with pytest.raises(Mom6cuError):
        ctx.vertvisc_coef(dict(coef, Kv_bbl_u=None))
