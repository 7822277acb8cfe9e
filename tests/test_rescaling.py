"""Power-of-2 dimensional rescaling -- the reference's `dim` regression tests (.testing test.dim.t / .l / .h / .z: T_RESCALE_POWER,
L_RESCALE_POWER, H_RESCALE_POWER, Z_RESCALE_POWER must reproduce ocean.stats bit for bit; src/framework/MOM_unit_scaling.F90)
re-expressed on the inputs of the hot path (tests/rescale.py): every dimensional input, metric and parameter of a stage is multiplied
by the power of two its dimension [T^a L^b H^c Z^d] implies, the oracle is run on the rescaled problem and its answers, scaled back,
must equal the un-scaled answers bit for bit.  A restatement that dropped a unit-conversion factor, mixed H with Z, or carries a
dimensional constant of the wrong units cannot pass.  (the last of the substitute pins.)  PressureForce_FV is covered through the
EOS conversion factors the reference carries for exactly this purpose (EOS%RL2_T2_to_Pa etc., MOM_EOS.F90:140-150)."""
import numpy as np
import pytest

import rescale as RS
from mom6_b200 import fidx, synthetic

POWERS = [(3, 0, 0, 0), (0, 5, 0, 0), (0, 0, -4, 0), (0, 0, 0, 6), (-2, 3, 7, 1)]


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_copy(v) for v in x]
    return x


def _same(a, b):
    if isinstance(a, dict):
        return all(_same(v, b[k]) for k, v in a.items())
    if isinstance(a, list):
        return all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, np.ndarray):
        return np.array_equal(a, b)
    return True


def _grids(grid, gv, p):
    return RS.scale(grid, RS.DIMS_GRID, p), RS.scale(gv, RS.DIMS_GV, p)


@pytest.mark.parametrize("p", POWERS)
def test_continuity(oracle, p):
    for cs_over in (None, dict(monotonic=1), dict(simple_2nd=1), dict(vol_CFL=1, aggress_adjust=1)):
        dom, grid, gv, cs, a = synthetic.continuity_inputs(20, 14, 5, land_blocks=2, cs_over=cs_over)
        ref = _copy(a); oracle.continuity(dom, grid, gv, cs, ref)
        gs, gvs = _grids(grid, gv, p)
        s = RS.scale(a, RS.CONT, p)
        oracle.continuity(dom, gs, gvs, RS.scale(cs, RS.with_flags(RS.CONT_CS, cs), p), s)
        assert _same(ref, RS.scale(s, RS.CONT, p, inverse=True)), (p, cs_over)
        assert np.abs(ref["uh"]).max() > 0


@pytest.mark.parametrize("p", POWERS)
def test_coradcalc(oracle, p):
    for over in (dict(), dict(Coriolis_Scheme=2), dict(Coriolis_Scheme=3), dict(Coriolis_Scheme=5), dict(Coriolis_Scheme=6), dict(bound_Coriolis=1),
                 dict(KE_Scheme=11), dict(Coriolis_En_Dis=1)):
        dom, grid, gv, cs, a = synthetic.coradcalc_inputs(20, 14, 5, land_blocks=2, cs_over=over, diags=True, por=True)
        ref = _copy(a); oracle.coradcalc(dom, grid, gv, cs, ref)
        gs, gvs = _grids(grid, gv, p)
        s = RS.scale(a, RS.CORAD, p)
        oracle.coradcalc(dom, gs, gvs, cs, s, us=RS.unit_scale(p))
        assert _same(ref, RS.scale(s, RS.CORAD, p, inverse=True)), (p, over)


@pytest.mark.parametrize("p", POWERS)
def test_horizontal_viscosity(oracle, p):
    for kw in (dict(), dict(Laplacian=True, Smagorinsky_Kh=True, Kh=500.0), dict(Smagorinsky_Ah=False, Ah=1e11), dict(cont_thick=True),
               dict(Re_Ah=10.0), dict(better_bound_Ah=False)):
        dom, grid, gv, cs, a = synthetic.hor_visc_inputs(20, 14, 5, land_blocks=2, **kw)
        ref = _copy(a); oracle.horizontal_viscosity(dom, grid, gv, cs, ref)
        gs, gvs = _grids(grid, gv, p)
        s = RS.scale(a, RS.HORVISC, p)
        oracle.horizontal_viscosity(dom, gs, gvs, RS.scale(cs, RS.with_flags(RS.HORVISC_CS, cs), p), s)
        assert _same(ref, RS.scale(s, RS.HORVISC, p, inverse=True)), (p, kw)
        assert np.abs(ref["diffu"]).max() > 0


def _vertvisc_family(oracle, dom, grid, gv, cs, coef, sol, us=None):
    nk = dom.nk
    out = dict(a_u=fidx.new(dom, "u", nk=nk + 1).a, a_v=fidx.new(dom, "v", nk=nk + 1).a, h_u=fidx.new(dom, "u", nk=nk).a, h_v=fidx.new(dom, "v", nk=nk).a,
               visc_rem_u=fidx.new(dom, "u", nk=nk).a, visc_rem_v=fidx.new(dom, "v", nk=nk).a)
    oracle.vertvisc_coef(dom, grid, gv, cs, coef, out["a_u"], out["a_v"], out["h_u"], out["h_v"], us=us)
    oracle.vertvisc(dom, grid, gv, cs, sol, out["a_u"], out["a_v"], out["h_u"], out["h_v"])
    oracle.vertvisc_remnant(dom, grid, cs, out["visc_rem_u"], out["visc_rem_v"], sol["dt"], out["a_u"], out["a_v"], out["h_u"], out["h_v"],
                            sol.get("Ray_u"), sol.get("Ray_v"))
    return out


@pytest.mark.parametrize("p", POWERS)
def test_vertvisc_family(oracle, p):
    for kw in (dict(), dict(harmonic_visc=1), dict(bottomdraglaw=0), dict(Kv_extra_bbl=1e-3), dict(Kvml_invZ2=1e-2), dict(fixed_LOTW_ML=1),
               dict(apply_LOTW_floor=1), dict(direct_stress=1), dict(with_Ray=True, with_Bu=True)):
        dom, grid, gv, cs, coef, sol = synthetic.vertvisc_inputs(20, 14, 6, land_blocks=2, **kw)
        c0, s0 = _copy(coef), _copy(sol)
        r0 = _vertvisc_family(oracle, dom, grid, gv, cs, c0, s0)
        gs, gvs = _grids(grid, gv, p)
        s1 = RS.scale(sol, RS.VERTVISC, p)
        r1 = _vertvisc_family(oracle, dom, gs, gvs, RS.scale(cs, RS.with_flags(RS.VERTVISC_CS, cs), p), RS.scale(coef, RS.VERTVISC_COEF, p), s1,
                              us=RS.unit_scale(p))
        assert _same(r0, RS.scale(r1, RS.VERTVISC_OUT, p, inverse=True)), (p, kw)
        assert _same(s0, RS.scale(s1, RS.VERTVISC, p, inverse=True)), (p, kw)


@pytest.mark.parametrize("p", POWERS)
def test_btstep(oracle, p):
    for kw in (dict(), dict(BT_project_velocity=1), dict(bound_BT_corr=1), dict(Sadourny=0), dict(strong_drag=1), dict(visc_rem_u_uh0=1)):
        dom, grid, gv, cs, a = synthetic.btstep_inputs(20, 14, 5, whalo=6, land_blocks=2, **kw)
        dcs = RS.with_flags(RS.BTSTEP_CS, cs)
        c0, a0 = _copy(cs), _copy(a); oracle.btstep(dom, grid, gv, c0, a0)
        gs, gvs = _grids(grid, gv, p)
        c1, a1 = RS.scale(cs, dcs, p), RS.scale(a, RS.BTSTEP, p)
        oracle.btstep(dom, gs, gvs, c1, a1)
        assert _same(a0, RS.scale(a1, RS.BTSTEP, p, inverse=True)), (p, kw)
        assert _same(c0, RS.scale(c1, dcs, p, inverse=True)), (p, kw)
        assert np.abs(a0["accel_layer_u"]).max() > 0 and np.abs(a0["eta_out"] - a["eta_in"]).max() > 0


@pytest.mark.parametrize("p", POWERS)
def test_tracer_advection_and_diffusion(oracle, p):
    for scheme in (0, 1, 2):
        dom, grid, gv, cs, a = synthetic.advect_inputs(20, 14, 5, land_blocks=2, cfl=2.5, scheme=scheme, ntr=3)
        ref = _copy(a); n = oracle.advect_tracer(dom, grid, gv, cs, ref)
        gs, gvs = _grids(grid, gv, p)
        s = RS.scale(a, RS.ADVECT, p)
        assert oracle.advect_tracer(dom, gs, gvs, RS.scale(cs, RS.ADVECT_CS, p), s) == n
        assert _same(ref["tr"], s["tr"]), (p, scheme)
    for kw in (dict(KhTr=5.0e4, check_diffusive_CFL=1, with_df=True), dict(KhTr=8.0e4, max_diff_CFL=2.5),
               dict(use_variable_mixing=1, Resoln_scaled_KhTr=1, KhTr_max=1500.0, KhTr_min=100.0, KhTr_passivity_coeff=2.0),
               dict(use_variable_mixing=1, KhTr_Slope_Cff=0.05, use_MEKE_Kh=1, KhTr=10.0, check_diffusive_CFL=1)):
        dom, grid, gv, cs, a = synthetic.hordiff_inputs(20, 14, 5, land_blocks=2, **kw)
        ref = _copy(a); n = oracle.tracer_hordiff(dom, grid, gv, cs, ref)
        gs, gvs = _grids(grid, gv, p)
        s = RS.scale(a, RS.HORDIFF, p)
        assert oracle.tracer_hordiff(dom, gs, gvs, RS.scale(cs, RS.with_flags(RS.HORDIFF_CS, cs), p), s) == n
        assert _same(ref, RS.scale(s, RS.HORDIFF, p, inverse=True)), (p, kw)


@pytest.mark.parametrize("p", POWERS)
def test_mixedlayer_restrat(oracle, p):
    for kw in (dict(), dict(MLE_MLD_decay_time2=7.776e6, ml_restrat_coef2=0.5, land_blocks=2), dict(front_length=0.0, ml_restrat_coef=60.0),
               dict(MLE_density_diff=0.1)):
        dom, grid, gv, cs, a = synthetic.mle_inputs(20, 14, 24, MLE_MLD_stretch=3.0, **kw)
        dcs = RS.with_flags(RS.MLE_CS, cs)
        c0, a0 = _copy(cs), _copy(a)
        oracle.mixedlayer_restrat(dom, grid, gv, c0, a0["h"], a0["uhtr"], a0["vhtr"], a0["T"], a0["S"], a0["ustar"], a0["dt"], a0["h_MLD"], a0["Rd_dx_h"])
        gs, gvs = _grids(grid, gv, p)
        c1, a1 = RS.scale(cs, dcs, p), RS.scale(a, RS.MLE, p)
        oracle.mixedlayer_restrat(dom, gs, gvs, c1, a1["h"], a1["uhtr"], a1["vhtr"], a1["T"], a1["S"], a1["ustar"], a1["dt"], a1["h_MLD"], a1["Rd_dx_h"])
        assert _same(a0, RS.scale(a1, RS.MLE, p, inverse=True)), (p, kw)
        assert _same(c0, RS.scale(c1, dcs, p, inverse=True)), (p, kw)
        assert not np.array_equal(a0["h"], a["h"])


@pytest.mark.parametrize("p", [(0, 0, -4, 0), (0, 0, 0, 6), (0, 0, 7, -3)])
def test_thickness_diffuse(oracle, p):
    for kw in (dict(with_GM=True, land_blocks=2), dict(use_variable_mixing=1, Resoln_scaled_KhTh=1, Khth_Max=400.0, Khth_Min=50.0),
               dict(EOS_form=1, Khth=3000.0, max_Khth_CFL=0.2, kappa_smooth=1.0e-4), dict(use_stored_slopes=1), dict(use_FGNV_streamfn=1, N2_floor=1.0e-12),
               dict(use_FGNV_streamfn=1, use_stored_slopes=1, use_MEKE_Kh=1, Khth=0.0, use_variable_mixing=1, Resoln_scaled_KhTh=1)):
        dom, grid, gv, cs, a = synthetic.thickness_diffuse_inputs(20, 14, 10, **kw)
        ref = _copy(a); oracle.thickness_diffuse(dom, grid, gv, cs, ref)
        gs, gvs = _grids(grid, gv, p)
        s = RS.scale(a, RS.THICKDIFF, p)
        oracle.thickness_diffuse(dom, gs, gvs, RS.scale(cs, RS.with_flags(RS.THICKDIFF_CS, cs), p), s, us=RS.unit_scale(p))
        assert _same(ref, RS.scale(s, RS.THICKDIFF, p, inverse=True)), (p, kw)
        assert not np.array_equal(ref["h"], a["h"])


@pytest.mark.parametrize("p", POWERS)
def test_pressure_force(oracle, p):
    """PressureForce_FV_Bouss under T, L, H, Z rescaling: analytic Wright / linear integrals with rho_scale / pres_scale (int_density_dz_wright,
    MOM_EOS_Wright.F90:497-534), the layered form, and the quadrature integrals of RECONSTRUCT_FOR_PRESSURE through calculate_density's own
    unit conversion (MOM_EOS.F90:332-352); Set_pbce_Bouss and the GFS_scale term included."""
    from mom6_b200 import marshal
    for kw in (dict(), dict(eos="LINEAR", dRho_dp=4.0e-6, MassWghtInterp=3, land_blocks=2), dict(eos="NONE"),
               dict(with_p_atm=True, use_SSH_in_Z0p=1, MassWghtInterp=1, GFS_scale=0.5, land_blocks=2, Z_ref=1.5),
               dict(reconstruct=1, Recon_Scheme=1, MassWghtInterp=1, land_blocks=2), dict(reconstruct=1, Recon_Scheme=2, boundary_extrap=1, with_p_atm=True),
               dict(reconstruct=1, Recon_Scheme=1, eos="LINEAR", dRho_dp=4.0e-6, use_inaccurate_pgf_rho_anom=1, MassWghtInterpVanOnly=1, h_nonvanished=1.0e-3)):
        dom, grid, gv, cs, a = synthetic.pressureforce_inputs(20, 14, 8, **kw)
        cs = dict(marshal.PGF_RECON_DEFAULTS, **cs)
        ref = _copy(a); oracle.pressure_force(dom, grid, gv, cs, ref)
        gs, gvs = _grids(grid, gv, p)
        s = RS.scale(a, RS.PGF, p)
        oracle.pressure_force(dom, gs, gvs, RS.scale(cs, RS.with_flags(RS.PGF_CS, cs), p), s)
        back = RS.scale(s, RS.PGF, p, inverse=True)
        for k in ("PFu", "PFv", "pbce", "eta"):
            assert np.array_equal(ref[k], back[k]), (p, kw, k, np.abs(ref[k] - back[k]).max())
        assert np.abs(ref["PFu"]).max() > 0


@pytest.mark.parametrize("p", POWERS)
def test_set_dtbt(oracle, p):
    from test_step_dyn import _dtbt_args
    for mode in ("pbce", "BT_cont", "eta", "gtot"):
        dom, grid, gv, css, cs, a = synthetic.step_dyn_inputs(20, 14, 5, whalo=6, land_blocks=2)
        oracle.step_dyn_split_rk2(dom, grid, gv, css, cs, a)                     # a step, so that pbce and BT_cont are set
        args = _dtbt_args(dom, grid, cs, mode)
        dtbt, dmax = oracle.set_dtbt(dom, grid, gv, args)
        gs, gvs = _grids(grid, gv, p)
        dtbt_s, dmax_s = oracle.set_dtbt(dom, gs, gvs, RS.scale(args, RS.SET_DTBT, p), us=RS.unit_scale(p))
        f = RS.factor(RS.TIME, p)
        assert (dtbt_s / f, dmax_s / f) == (dtbt, dmax) and dtbt > 0, (p, mode)


@pytest.mark.parametrize("p", [(0, 0, -4, 0), (0, 0, 0, 6), (0, 0, 7, -3)])
def test_ale_regrid_and_remap(oracle, p):
    """Z* regridding (target resolution in Z, thicknesses in H) and the conservative remap (homogeneous in the thicknesses)."""
    for kw in (dict(), dict(land_blocks=3, old_grid_weight=0.75, depth_of_time_filter_shallow=200.0, depth_of_time_filter_deep=1000.0), dict(min_thickness=5.0)):
        dom, grid, gv, cs, a = synthetic.regrid_inputs(20, 14, 8, **kw)
        ref = _copy(a)
        assert oracle.ale_regrid(dom, grid, gv, cs, ref["h"], ref["h_new"], ref["dzRegrid"]) == 0
        gs, gvs = _grids(grid, gv, p)
        s = RS.scale(a, RS.REGRID, p)
        css = RS.scale(cs, RS.REGRID_CS, p)
        css["coordinateResolution"] = np.ascontiguousarray(cs["coordinateResolution"] * RS.factor(RS.ZL, p))
        assert oracle.ale_regrid(dom, gs, gvs, css, s["h"], s["h_new"], s["dzRegrid"], us=RS.unit_scale(p)) == 0
        assert _same(ref, RS.scale(s, RS.REGRID, p, inverse=True)), (p, kw)
        assert np.abs(ref["dzRegrid"]).max() > 0
    for scheme in (2, 4, 5):                                                    # PLM, PPM_H4, PPM_IH4
        dom, grid, cs, a = synthetic.remap_inputs(20, 14, 8, land_blocks=2, remapping_scheme=scheme)
        f = RS.factor(RS.THK, p)
        t0, t1 = a["tr"][0].copy(), a["tr"][0].copy()
        oracle.ale_remap_scalar(dom, grid, cs, a["h_old"], a["h_new"], t0)
        css = dict(cs, h_neglect=cs["h_neglect"] * f, h_neglect_edge=cs["h_neglect_edge"] * f)
        oracle.ale_remap_scalar(dom, RS.scale(grid, RS.DIMS_GRID, p), css, np.ascontiguousarray(a["h_old"] * f), np.ascontiguousarray(a["h_new"] * f), t1)
        assert np.array_equal(t0, t1) and not np.array_equal(t0, a["tr"][0]), (p, scheme)
