"""Index-rotation invariance -- the reference's `rotate` regression test (ROTATE_INDEX = True must reproduce ocean.stats bit for
bit; .testing/Makefile, src/core/MOM.F90:2805-2920) re-expressed on the inputs of the hot path (tests/rotation.py): the whole
step_MOM_dyn_split_RK2, advect_tracer, mixedlayer_restrat and the ocean.stats line of the result are computed on an index map
turned by a quarter (x-first becomes y-first, reentrant-in-x becomes reentrant-in-y, vectors change sign) and must equal the
un-rotated answers bit for bit once turned back.  Every u-branch of the restatement is checked against its independently
written v-twin this way (the substitute pin for the stages the reference holds no vector for).
CPU: the oracle.  GPU: the device path through the C ABI on the rotated map == the oracle on the original one."""
import numpy as np
import pytest

import rotation as R
from mom6_b200 import synthetic

ST = synthetic.STEP_STAGGER


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_copy(v) for v in x]
    return x


def _inner(dom, x, st="h"):
    """The computational domain including the symmetric south / west edge of staggered fields."""
    su = 1 if st in "uq" else 0
    sv = 1 if st in "vq" else 0
    return x[..., dom.jsc - dom.jsd - sv + sv:dom.jec - dom.jsd + 1 + sv, dom.isc - dom.isd - su + su:dom.iec - dom.isd + 1 + su]


def _step_mismatches(dom, ref_cs, ref_a, cs, a):
    bad = []
    for name, d0, d1 in (("", ref_a, a), ("CS%", ref_cs, cs), ("BT_cont%", ref_cs["BT_cont"], cs["BT_cont"]),
                         ("BT%", ref_cs["barotropic"], cs["barotropic"])):
        for k, v in d0.items():
            if isinstance(v, np.ndarray) and v.ndim >= 2 and k in ST and k not in synthetic.BT_WIDE:
                A, B = _inner(dom, v, ST[k]), _inner(dom, d1[k], ST[k])
                if not np.array_equal(A, B):                      # == on values: -0.0 of a negated land point equals +0.0
                    bad.append((name + k, int(np.count_nonzero(A != B))))
    return bad


STEP_CASES = [dict(land_blocks=2, store_CAu=1), dict(land_blocks=3, split_bottom_stress=1, BT_project_velocity=1, begw=0.5),
              dict(land_blocks=1, calc_dtbt=1), dict(size=(18, 26, 9), land_blocks=2, bound_BT_corr=1)]


@pytest.mark.parametrize("kw", STEP_CASES)
def test_oracle_step_is_rotation_invariant(oracle, kw):
    kw = dict(kw)
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(*kw.pop("size", (22, 16, 6)), whalo=6, **kw)
    u0 = a["u_inst"].copy()
    rcs, ra = _copy(cs), _copy(a)
    domr, gridr, cssr, csr, ar = R.rotate_step(dom, grid, css, cs, a)
    assert (domr.cyclic_x, domr.cyclic_y, domr.first_direction) == (0, 1, 1)
    for step in range(2):
        oracle.step_dyn_split_rk2(dom, grid, gv, css, rcs, ra)
        oracle.step_dyn_split_rk2(domr, gridr, gv, cssr, csr, ar)
        csb, ab = R.unrotate_step(csr, ar)
        assert not _step_mismatches(dom, rcs, ra, csb, ab), (step, kw)
        assert rcs["barotropic"]["dtbt"] == csr["barotropic"]["dtbt"] and rcs["dtbt_max"] == csr["dtbt_max"]
    assert np.abs(_inner(dom, ra["u_inst"] - u0, "u")).max() > 1e-6          # the steps did move the state


@pytest.mark.parametrize("scheme", [0, 1, 2])
def test_oracle_advect_tracer_is_rotation_invariant(oracle, scheme):
    dom, grid, gv, cs, a = synthetic.advect_inputs(20, 14, 5, land_blocks=2, cfl=2.5, scheme=scheme, ntr=3)
    ref = _copy(a)
    oracle.advect_tracer(dom, grid, gv, cs, ref)
    ar = R.rotate_fields(a, R.STEP_VEC, R.STEP_PAIR)
    oracle.advect_tracer(R.rotate_domain(dom), R.rotate_grid(grid), gv, cs, ar)      # first_direction 1: y first = the original x first
    back = R.unrotate_fields(ar, R.STEP_VEC, R.STEP_PAIR)
    for m in range(3):
        assert np.array_equal(_inner(dom, ref["tr"][m]), _inner(dom, back["tr"][m])), (scheme, m)
    assert not np.array_equal(_inner(dom, ref["tr"][0]), _inner(dom, a["tr"][0]))


@pytest.mark.parametrize("kw", [dict(), dict(MLE_density_diff=0.1)])
def test_oracle_mixedlayer_restrat_is_rotation_invariant(oracle, kw):
    dom, grid, gv, cs, a = synthetic.mle_inputs(20, 14, 24, land_blocks=2, MLE_MLD_decay_time2=7.776e6, ml_restrat_coef2=0.5, MLE_MLD_stretch=3.0, **kw)
    cr, ar = R.rotate_fields(cs), R.rotate_fields(a, R.STEP_VEC, R.STEP_PAIR)
    c0, a0 = _copy(cs), _copy(a)
    oracle.mixedlayer_restrat(dom, grid, gv, c0, a0["h"], a0["uhtr"], a0["vhtr"], a0["T"], a0["S"], a0["ustar"], a0["dt"], a0["h_MLD"], a0["Rd_dx_h"])
    oracle.mixedlayer_restrat(R.rotate_domain(dom), R.rotate_grid(grid), gv, cr, ar["h"], ar["uhtr"], ar["vhtr"], ar["T"], ar["S"], ar["ustar"],
                              ar["dt"], ar["h_MLD"], ar["Rd_dx_h"])
    back = R.unrotate_fields(ar, R.STEP_VEC, R.STEP_PAIR)
    for k, st in (("h", "h"), ("uhtr", "u"), ("vhtr", "v")):
        assert np.array_equal(_inner(dom, a0[k], st), _inner(dom, back[k], st)), k
    assert np.array_equal(_inner(dom, c0["MLD_filtered"]), _inner(dom, R.unrot(cr["MLD_filtered"])))
    assert not np.array_equal(a0["h"], a["h"])


def test_ocean_stats_line_is_rotation_invariant(oracle):
    """The reference's criterion itself: the ocean.stats line of the state after a step is the same text on the rotated map."""
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(22, 16, 6, whalo=6, land_blocks=2)
    domr, gridr, cssr, csr, ar = R.rotate_step(dom, grid, css, cs, a)
    oracle.step_dyn_split_rk2(dom, grid, gv, css, cs, a)
    oracle.step_dyn_split_rk2(domr, gridr, gv, cssr, csr, ar)
    so = synthetic.sum_output_cs(dom, oracle.create_depth_list(dom, grid))
    sor = synthetic.sum_output_cs(domr, oracle.create_depth_list(domr, gridr))
    e = oracle.write_energy(dom, grid, gv, so, a["u_inst"], a["v_inst"], a["h"], a["T"], a["S"])
    er = oracle.write_energy(domr, gridr, gv, sor, ar["u_inst"], ar["v_inst"], ar["h"], ar["T"], ar["S"])
    assert oracle.ocean_stats_line(so, e, 1, 0.0104) == oracle.ocean_stats_line(sor, er, 1, 0.0104)
    for k in ("En_mass", "mass_tot", "KE_tot", "PE_tot", "Salt", "Heat"):
        assert e[k] == er[k], k
    assert e["max_CFL"][0] == er["max_CFL"][0]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", STEP_CASES[:2])
def test_gpu_step_on_the_rotated_map(oracle, ctx_factory, kw):
    kw = dict(kw)
    dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(*kw.pop("size", (44, 40, 8)), whalo=6, **kw)
    rcs, ra = _copy(cs), _copy(a)
    domr, gridr, cssr, csr, ar = R.rotate_step(dom, grid, css, cs, a)
    ctx = ctx_factory(domr)
    ctx.set_grid(gridr); ctx.set_vgrid(gv)
    ctx.set_cs_continuity(cssr["continuity"]); ctx.set_cs_coriolisadv(cssr["coriolisadv"]); ctx.set_cs_hor_visc(cssr["hor_visc"])
    ctx.set_cs_pressureforce(cssr["pressureforce"]); ctx.set_cs_vertvisc(cssr["vertvisc"])
    for step in range(2):
        oracle.step_dyn_split_rk2(dom, grid, gv, css, rcs, ra)
        ctx.step_dyn_split_rk2(csr, ar)
        csb, ab = R.unrotate_step(csr, ar)
        assert not _step_mismatches(dom, rcs, ra, csb, ab), (step, kw)


def test_oracle_ale_regridding_and_remapping_is_rotation_invariant(oracle):
    dom, grid, gv, ale, dcs, a = synthetic.ale_chain_inputs(20, 14, 8, land_blocks=2)
    ref_a, ref_cs, ref_ale = _copy(a), _copy(dcs), _copy(ale)
    oracle.ale_regridding_and_remapping(dom, grid, gv, ref_ale, ref_a, dyn_cs=ref_cs)
    ar = R.rotate_fields(a, R.STEP_VEC, R.STEP_PAIR, keep=("conc_underflow",))
    csr = R.rotate_fields({k: v for k, v in dcs.items() if not isinstance(v, dict)}, R.STEP_VEC, R.STEP_PAIR)
    csr["BT_cont"] = R.rotate_bt_cont(dcs["BT_cont"]); csr["barotropic"] = R.rotate_fields(dcs["barotropic"], R.STEP_VEC, R.STEP_PAIR)
    aler = _copy(ale)
    oracle.ale_regridding_and_remapping(R.rotate_domain(dom), R.rotate_grid(grid), gv, aler, ar, dyn_cs=csr)
    back = R.unrotate_fields(ar, R.STEP_VEC, R.STEP_PAIR, keep=("conc_underflow",))
    csb = R.unrotate_fields({k: v for k, v in csr.items() if not isinstance(v, dict)}, R.STEP_VEC, R.STEP_PAIR)
    for k, st in (("u", "u"), ("v", "v"), ("h", "h"), ("Kd_shear", "h"), ("Kv_shear", "h"), ("Kv_shear_Bu", "q")):
        assert np.array_equal(_inner(dom, ref_a[k], st), _inner(dom, back[k], st)), k
    for m in range(len(a["tr"])):
        assert np.array_equal(_inner(dom, ref_a["tr"][m]), _inner(dom, back["tr"][m])), m
    for k in ("diffu", "diffv", "CAu_pred", "CAv_pred", "u_av", "v_av"):
        assert np.array_equal(_inner(dom, ref_cs[k], ST[k]), _inner(dom, csb[k], ST[k])), k
    assert not np.array_equal(ref_a["h"], a["h"]) and aler["regridCS"]["old_grid_weight"] == ref_ale["regridCS"]["old_grid_weight"]


@pytest.mark.parametrize("kw", [dict(KhTr=5.0e4, check_diffusive_CFL=1, with_df=True),
                                dict(use_variable_mixing=1, Resoln_scaled_KhTr=1, KhTr_max=1500.0, KhTr_min=100.0, KhTr_passivity_coeff=2.0),
                                dict(use_variable_mixing=1, KhTr_Slope_Cff=0.05, use_MEKE_Kh=1, KhTr=10.0, check_diffusive_CFL=1)])
def test_oracle_tracer_hordiff_is_rotation_invariant(oracle, kw):
    dom, grid, gv, cs, a = synthetic.hordiff_inputs(20, 14, 5, land_blocks=2, **kw)
    ref = _copy(a)
    n = oracle.tracer_hordiff(dom, grid, gv, cs, ref)
    ar = R.rotate_fields(a, pair=[("L2u", "L2v"), ("SN_u", "SN_v")], keep=("conc_underflow", "df_x", "df_y"))
    if a.get("df_x"):                                              # flux diagnostics: (df_x, df_y) is a vector
        ar["df_x"] = [None if f is None else -R.rot(f) for f in a["df_y"]]
        ar["df_y"] = [None if f is None else R.rot(f) for f in a["df_x"]]
    assert oracle.tracer_hordiff(R.rotate_domain(dom), R.rotate_grid(grid), gv, cs, ar) == n
    for m in range(len(a["tr"])):
        assert np.array_equal(_inner(dom, ref["tr"][m]), _inner(dom, R.unrot(ar["tr"][m]))), m
    if a.get("df_x"):
        for m in range(len(a["tr"])):
            if ref["df_x"][m] is not None:
                assert np.array_equal(_inner(dom, ref["df_x"][m], "u"), _inner(dom, R.unrot(ar["df_y"][m]), "u")), m
            if ref["df_y"][m] is not None:
                assert np.array_equal(_inner(dom, ref["df_y"][m], "v"), _inner(dom, -R.unrot(ar["df_x"][m]), "v")), m
    assert not np.array_equal(ref["tr"][0], a["tr"][0])


@pytest.mark.parametrize("kw", [dict(with_GM=True, land_blocks=2), dict(use_variable_mixing=1, Resoln_scaled_KhTh=1, Khth_Max=400.0, Khth_Min=50.0, with_p_surf=True),
                                dict(EOS_form=1, Khth=3000.0, max_Khth_CFL=0.2, kappa_smooth=1.0e-4), dict(use_stored_slopes=1, land_blocks=2),
                                dict(use_FGNV_streamfn=1, with_GM=True, land_blocks=2),
                                dict(use_FGNV_streamfn=1, use_stored_slopes=1, use_MEKE_Kh=1, Khth=0.0, use_variable_mixing=1, Resoln_scaled_KhTh=1)])
def test_oracle_thickness_diffuse_is_rotation_invariant(oracle, kw):
    dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(20, 14, 10, **kw)
    VEC, PAIR = R.STEP_VEC + [("uhGM", "vhGM"), ("slope_x", "slope_y")], R.STEP_PAIR + [("Res_fn_u", "Res_fn_v")]
    ref = _copy(a)
    oracle.thickness_diffuse(dom, grid, gv, cs, ref)
    ar = R.rotate_fields(a, VEC, PAIR)
    oracle.thickness_diffuse(R.rotate_domain(dom), R.rotate_grid(grid), gv, cs, ar)
    back = R.unrotate_fields(ar, VEC, PAIR)
    for k, st in (("h", "h"), ("uhtr", "u"), ("vhtr", "v"), ("uhGM", "u"), ("vhGM", "v")):
        if ref.get(k) is not None:
            assert np.array_equal(_inner(dom, ref[k], st), _inner(dom, back[k], st)), k
    assert not np.array_equal(ref["h"], a["h"])
