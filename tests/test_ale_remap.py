"""ALE remapping (src/ALE/MOM_remapping.F90 remapping_core_h, src/ALE/MOM_ALE.F90 ALE_remap_tracers /
ALE_remap_set_h_vel / ALE_remap_velocities).  The oracle is pinned to the reference's unit-test vectors in
test_oracle_remap_kat.py; here the C ABI is compared with it bit for bit, and with the vectors directly."""
import numpy as np
import pytest

from mom6_b200 import synthetic
from mom6_b200.api import make_domain
from remap_cases import columns, cs_variants, SHAPES, KINDS
from test_oracle_remap_kat import CS_PPM_H4


def _beq(a, b):
    return np.array_equal(a.view(np.int64), b.view(np.int64))


def test_oracle_3d_remap_conserves_and_skips_land(oracle):
    dom, grid, cs, a = synthetic.remap_inputs(24, 18, 10, land_blocks=3)
    t0 = a["tr"][0].copy()
    t = t0.copy()
    oracle.ale_remap_scalar(dom, grid, cs, a["h_old"], a["h_new"], t)
    js = slice(dom.jsc - dom.jsd, dom.jec - dom.jsd + 1); is_ = slice(dom.isc - dom.isd, dom.iec - dom.isd + 1)
    m = grid["mask2dT"][js, is_] > 0
    before = (t0 * a["h_old"]).sum(axis=0)[js, is_]; after = (t * a["h_new"]).sum(axis=0)[js, is_]
    assert np.allclose(before[m], after[m], rtol=1e-12)
    assert np.array_equal(t[:, js, is_][:, ~m], t0[:, js, is_][:, ~m])            # land columns untouched
    assert not np.array_equal(t[:, js, is_][:, m], t0[:, js, is_][:, m])


@pytest.mark.gpu
def test_core_h_known_answers_through_the_c_abi(ctx_factory):
    """MOM_remapping.F90:2155-2162 vectors, directly against the device path."""
    ctx = ctx_factory(make_domain(8, 8, nk=4))
    h0 = [0.75] * 4; u0 = [9., 3., -3., -9.]
    assert np.array_equal(ctx.remapping_core_h(CS_PPM_H4, h0, u0, [0.5] * 6)[0], [10., 6., 2., -2., -6., -10.])
    assert np.array_equal(ctx.remapping_core_h(CS_PPM_H4, h0, u0, [.125] * 6)[0], [11.5, 10.5, 9.5, 8.5, 7.5, 6.5])
    assert np.array_equal(ctx.remapping_core_h(CS_PPM_H4, h0, u0, [2.25, 1.5, 1.])[0], [3., -10.5, -12.])


@pytest.mark.gpu
def test_core_h_batch_bitwise(oracle, ctx_factory):
    ctx = ctx_factory(make_domain(8, 8, nk=4))
    rng = np.random.default_rng(21)
    bad = []
    variants = list(cs_variants())
    for (n0, n1) in SHAPES:
        for kind in KINDS:
            h0, u0, h1 = columns(rng, 40, n0, n1, kind)
            for cs in variants[::3] if (n0, n1) != (75, 75) else variants:
                got = ctx.remapping_core_h(cs, h0, u0, h1)
                ref = np.stack([oracle.remapping_core_h(cs, h0[c], u0[c], h1[c])[0] for c in range(h0.shape[0])])
                if not np.array_equal(ref, got, equal_nan=True):
                    bad.append((n0, n1, kind, cs["remapping_scheme"], cs["boundary_extrapolation"], cs["om4_remap_via_sub_cells"]))
    assert not bad, bad[:10]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(remapping_scheme=2, boundary_extrapolation=1, om4_remap_via_sub_cells=0, land_blocks=4),
                                dict(remapping_scheme=5, kind="uniform", land_blocks=3, cyclic_y=True),
                                dict(remapping_scheme=0, ntr=1), dict(ntr=19, land_blocks=2, force_bounds_in_subcell=1),
                                dict(size=(36, 28, 75), land_blocks=2),                       # OM4 layer count (the 80-layer kernel variant)
                                dict(size=(36, 28, 75), remapping_scheme=2, land_blocks=2),
                                dict(size=(20, 16, 110), remapping_scheme=5, ntr=2)])         # the 128-layer variant
def test_ale_remap_tracers_and_velocities_bitwise(oracle, ctx_factory, kw):
    kw = dict(kw)
    dom, grid, cs, a = synthetic.remap_inputs(*kw.pop("size", (44, 40, 12)), **kw)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid)
    ntr = len(a["tr"])
    under = np.array([0.0, 1.0e-35] + [1.0e-20] * (ntr - 2))[:ntr]
    ref = [t.copy() for t in a["tr"]]
    for m in range(ntr):
        oracle.ale_remap_scalar(dom, grid, cs, a["h_old"], a["h_new"], ref[m], conc_underflow=under[m])
    got = [t.copy() for t in a["tr"]]
    n0 = ctx.launches
    ctx.ale_remap_tracers(cs, a["h_old"], a["h_new"], got, under)
    assert ctx.launches > n0
    for m in range(ntr):
        assert _beq(ref[m], got[m]), (m, np.count_nonzero(ref[m] != got[m]))
    # velocities: h at the velocity points before and after, then the remap
    hu = {k: np.zeros_like(a["u"]) for k in ("ro", "rn", "go", "gn")}; hv = {k: np.zeros_like(a["v"]) for k in ("ro", "rn", "go", "gn")}
    oracle.ale_remap_set_h_vel(dom, grid, a["h_old"], hu["ro"], hv["ro"]); oracle.ale_remap_set_h_vel(dom, grid, a["h_new"], hu["rn"], hv["rn"])
    ctx.ale_remap_set_h_vel(a["h_old"], hu["go"], hv["go"]); ctx.ale_remap_set_h_vel(a["h_new"], hu["gn"], hv["gn"])
    assert _beq(hu["ro"], hu["go"]) and _beq(hv["rn"], hv["gn"]) and _beq(hu["rn"], hu["gn"]) and _beq(hv["ro"], hv["go"])
    ur, vr, ug, vg = a["u"].copy(), a["v"].copy(), a["u"].copy(), a["v"].copy()
    oracle.ale_remap_velocities(dom, grid, cs, hu["ro"], hv["ro"], hu["rn"], hv["rn"], ur, vr)
    ctx.ale_remap_velocities(cs, hu["go"], hv["go"], hu["gn"], hv["gn"], ug, vg)
    assert _beq(ur, ug) and _beq(vr, vg)
    assert not _beq(ur, a["u"])


@pytest.mark.gpu
def test_remap_rejects_what_is_not_implemented(ctx_factory):
    from mom6_b200.api import Mom6cuError
    ctx = ctx_factory(make_domain(8, 8, nk=4))
    for bad in (dict(remapping_scheme=7), dict(answer_date=20181231)):
        with pytest.raises(Mom6cuError):
            ctx.remapping_core_h(dict(CS_PPM_H4, **bad), [1., 1.], [1., 2.], [2.])
    with pytest.raises(Mom6cuError):
        ctx.remapping_core_h(CS_PPM_H4, np.ones((1, 200)), np.ones((1, 200)), np.ones((1, 3)))


@pytest.mark.gpu
def test_remap_dyn_split_rk2_aux_vars_bitwise(oracle, ctx_factory):
    """MOM_dynamics_split_RK2.F90:1302: the auxiliary restart variables follow the grid with the velocities."""
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(44, 40, 8, whalo=6, land_blocks=2, store_CAu=1)
    oracle.step_dyn_split_rk2(dom, grid, gv, css, cs, a)                  # realistic u_av, CAu_pred, diffu
    _, _, rcs, ra = synthetic.remap_inputs(44, 40, 8, land_blocks=2)
    hu = {k: np.zeros_like(a["u_inst"]) for k in ("o", "n")}; hv = {k: np.zeros_like(a["v_inst"]) for k in ("o", "n")}
    oracle.ale_remap_set_h_vel(dom, grid, ra["h_old"], hu["o"], hv["o"]); oracle.ale_remap_set_h_vel(dom, grid, ra["h_new"], hu["n"], hv["n"])
    ref = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in cs.items()}
    got = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in cs.items()}
    oracle.remap_dyn_split_rk2_aux_vars(dom, grid, rcs, ref, hu["o"], hv["o"], hu["n"], hv["n"])
    ctx = ctx_factory(dom)
    ctx.set_grid(grid)
    ctx.remap_dyn_split_rk2_aux_vars(rcs, got, hu["o"], hv["o"], hu["n"], hv["n"])
    for k in ("u_av", "v_av", "CAu_pred", "CAv_pred", "diffu", "diffv"):
        assert _beq(ref[k], got[k]), k
    assert not _beq(ref["u_av"], cs["u_av"])
