"""PressureForce_FV_Bouss (src/core/MOM_PressureForce_FV.F90:947-2017) with the analytic LINEAR / WRIGHT layer
integrals and Set_pbce_Bouss.  CPU: properties of the oracle restatement.  GPU: C ABI == oracle, bit for bit."""
import numpy as np
import pytest

from mom6_b200 import synthetic
from test_oracle_continuity import _comp


def _copy(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, dict):
        return {k: _copy(v) for k, v in x.items()}
    return x


def test_resting_uniform_ocean_has_no_pressure_force(oracle):
    """Flat interfaces and horizontally uniform T,S: the finite-volume PGF vanishes identically (Adcroft et al. 2008)."""
    dom, grid, gv, cs, a = synthetic.pressureforce_inputs(24, 18, 6, eos="WRIGHT")
    a = _copy(a)
    grid = dict(grid); grid["bathyT"] = np.full_like(grid["bathyT"], 3000.0)
    nk = 6
    a["h"][...] = 3000.0 / nk
    for k in range(nk):
        a["T"][k] = 20.0 - 2.0 * k; a["S"][k] = 35.0 + 0.1 * k
    oracle.pressure_force(dom, grid, gv, cs, a)
    assert np.abs(_comp(dom, a["PFu"], "u")).max() < 1e-12 and np.abs(_comp(dom, a["PFv"], "v")).max() < 1e-12
    assert np.abs(_comp(dom, a["eta"], "h")).max() < 1e-9
    assert (_comp(dom, a["pbce"], "h") > 0).all()


def test_sea_surface_slope_drives_barotropic_force(oracle):
    """A uniform-density ocean with a surface tilt feels PFu = -g d(eta)/dx in every layer (rho_ref = Rho0 = const)."""
    dom, grid, gv, cs, a = synthetic.pressureforce_inputs(24, 18, 4, eos="LINEAR", dRho_dT=0.0, dRho_dS=0.0, Rho_T0_S0=1035.0)
    a = _copy(a)
    grid = dict(grid); grid["bathyT"] = np.full_like(grid["bathyT"], 1000.0)
    ii = np.arange(a["h"].shape[2])[None, None, :]
    a["h"][...] = 250.0
    a["h"][0] = 250.0 + 0.01 * np.sin(2 * np.pi * ii[0] / 24.0)
    oracle.pressure_force(dom, grid, gv, cs, a)
    eta = a["eta"]
    expect = -gv["g_Earth"] * (eta[:, 1:] - eta[:, :-1]) * grid["IdxCu"][:, 1:-1]
    got = a["PFu"][:, :, 1:-1]
    js = slice(dom.jsc - dom.jsd, dom.jec - dom.jsd + 1); is_ = slice(dom.isc - dom.isd, dom.iec - dom.isd)
    for k in range(4):
        assert np.allclose(got[k][js, is_], expect[js, is_], rtol=1e-6, atol=1e-12)


RECON = [dict(reconstruct=1, Recon_Scheme=1), dict(reconstruct=1, Recon_Scheme=2), dict(reconstruct=1, Recon_Scheme=1, boundary_extrap=1),
         dict(reconstruct=1, Recon_Scheme=2, boundary_extrap=1)]


@pytest.mark.parametrize("eos", ["WRIGHT", "LINEAR"])
@pytest.mark.parametrize("rc", RECON[:2])
def test_reconstructed_profiles_reduce_to_the_analytic_integrals(oracle, eos, rc):
    """RECONSTRUCT_FOR_PRESSURE (int_density_dz_generic_plm / _ppm, MOM_density_integrals.F90:418/874) with vertically uniform T,S in every
    column has zero slopes and curvatures, so its Boole quadrature of the PINNED density (tests/test_oracle_eos_kat.py) must agree with the
    analytic layer integrals (int_density_dz_wright MOM_EOS_Wright.F90:389 / _linear) to the quadrature's truncation error: an independent
    check of the transcription of the analytic integrals' series and constants."""
    dom, grid, gv, cs, a = synthetic.pressureforce_inputs(28, 20, 8, eos=eos, land_blocks=2, MassWghtInterp=1, dRho_dp=4.0e-6)
    r = np.random.default_rng(3)
    if eos == "LINEAR":   # linear in T, S and p: interpolating T,S along a face (generic) or the EOS coefficients (analytic) is the same
        T2 = 10.0 + 8.0 * r.uniform(-1, 1, size=a["T"].shape[1:]); S2 = 35.0 + r.uniform(-1, 1, size=a["S"].shape[1:])
        a["T"][...] = T2[None]; a["S"][...] = S2[None]
    else:                 # WRIGHT: uniform T,S; the force comes from the surface slope and the pressure dependence of density over the
        a["T"][...] = 12.0; a["S"][...] = 34.5   # sloping layers: a test of the analytic series in eps (C1_3, 0.2, C1_7, C1_9) and Boole weights
        a["h"][0] *= 1.0 + 0.02 * r.uniform(-1, 1, size=a["h"].shape[1:])
    ref = _copy(a); oracle.pressure_force(dom, grid, gv, cs, ref)
    got = _copy(a); oracle.pressure_force(dom, grid, gv, dict(cs, **rc), got)
    for k in ("PFu", "PFv"):
        x, y = _comp(dom, ref[k], k[-1]), _comp(dom, got[k], k[-1])
        assert np.abs(x).max() > 1e-6
        assert np.abs(x - y).max() < 2e-9 * np.abs(x).max(), (k, np.abs(x - y).max(), np.abs(x).max())
    assert np.array_equal(ref["pbce"], got["pbce"]) and np.array_equal(ref["eta"], got["eta"])


@pytest.mark.parametrize("rc", RECON)
def test_reconstructed_resting_ocean_has_no_pressure_force(oracle, rc):
    """Flat interfaces, horizontally uniform stratified T,S: every face integral equals the column integrals either side of it."""
    dom, grid, gv, cs, a = synthetic.pressureforce_inputs(24, 18, 8, eos="WRIGHT")
    a = _copy(a)
    grid = dict(grid); grid["bathyT"] = np.full_like(grid["bathyT"], 3000.0)
    a["h"][...] = 3000.0 / 8
    for k in range(8):
        a["T"][k] = 20.0 - 2.0 * k - 0.1 * k * k; a["S"][k] = 35.0 + 0.1 * k
    oracle.pressure_force(dom, grid, gv, dict(cs, **rc), a)
    assert np.abs(_comp(dom, a["PFu"], "u")).max() < 1e-12 and np.abs(_comp(dom, a["PFv"], "v")).max() < 1e-12


@pytest.mark.parametrize("rc", RECON[:2])
def test_reconstruction_changes_the_answer_for_stratified_columns(oracle, rc):
    """With sloping layers over the seamount the sub-layer profiles matter: PLM / PPM differ from the piecewise-constant answer by a
    fraction of a percent of the force, and from each other."""
    dom, grid, gv, cs, a = synthetic.pressureforce_inputs(28, 20, 10, eos="WRIGHT", land_blocks=2)
    ref = _copy(a); oracle.pressure_force(dom, grid, gv, cs, ref)
    got = _copy(a); oracle.pressure_force(dom, grid, gv, dict(cs, **rc), got)
    x, y = _comp(dom, ref["PFu"], "u"), _comp(dom, got["PFu"], "u")
    d = np.abs(x - y).max() / np.abs(x).max()
    assert 1e-7 < d < 0.2, d
    assert np.isfinite(y).all()


CASES = [
    dict(),                                                        # WRIGHT, defaults (benchmark-like)
    dict(eos="LINEAR"),                                            # tc4-like
    dict(eos="NONE"),                                              # layered, no EOS
    dict(land_blocks=4, MassWghtInterp=1),
    dict(land_blocks=4, MassWghtInterp=3, eos="LINEAR", dRho_dp=4.0e-6),
    dict(with_p_atm=True, use_SSH_in_Z0p=1, MassWghtInterp=2),
    dict(GFS_scale=0.5, rho_ref_bug=1, rho_ref=1030.0),
    dict(GFS_scale=0.2, eos="NONE", with_pbce=False, with_eta=False),
    dict(cyclic_y=True, land_blocks=3, Z_ref=1.5, with_p_atm=True),
    # RECONSTRUCT_FOR_PRESSURE (the reference's default with ALE): PLM / PPM sub-layer profiles, quadrature integrals
    dict(reconstruct=1, Recon_Scheme=1),
    dict(reconstruct=1, Recon_Scheme=2),
    dict(reconstruct=1, Recon_Scheme=1, boundary_extrap=1, land_blocks=4, MassWghtInterp=1),
    dict(reconstruct=1, Recon_Scheme=2, boundary_extrap=1, land_blocks=4, MassWghtInterp=3, with_p_atm=True, use_SSH_in_Z0p=1),
    dict(reconstruct=1, Recon_Scheme=1, eos="LINEAR", dRho_dp=4.0e-6, MassWghtInterp=1, MassWghtInterpVanOnly=1, h_nonvanished=1.0e-3),
    dict(reconstruct=1, Recon_Scheme=2, eos="LINEAR", use_inaccurate_pgf_rho_anom=1, land_blocks=2),
    dict(reconstruct=1, Recon_Scheme=1, use_inaccurate_pgf_rho_anom=1, GFS_scale=0.5, cyclic_y=True, land_blocks=3),
    dict(reconstruct=1, Recon_Scheme=1, eos="NONE"),                 # use_ALE needs an equation of state: falls back to the layered form
]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_pressure_force_bitwise(oracle, ctx_factory, kw):
    dom, grid, gv, cs, a = synthetic.pressureforce_inputs(44, 40, 8, **kw)
    ref = _copy(a)
    oracle.pressure_force(dom, grid, gv, cs, ref)
    got = _copy(a)
    ctx = ctx_factory(dom)
    ctx.set_grid(grid); ctx.set_vgrid(gv); ctx.set_cs_pressureforce(cs)
    n0 = ctx.launches
    ctx.pressure_force(got)
    assert ctx.launches > n0
    bad = [f"{k}: {np.count_nonzero(ref[k] != got[k])} of {ref[k].size} differ, max |d|={np.nanmax(np.abs(ref[k] - got[k]))}"
           for k in ("PFu", "PFv", "pbce", "eta") if ref.get(k) is not None and not np.array_equal(ref[k].view(np.int64), got[k].view(np.int64))]
    assert not bad, "; ".join(bad)
    assert np.abs(ref["PFu"]).max() > 0


@pytest.mark.gpu
def test_pressure_force_ragged_and_errors(oracle, ctx_factory):
    from mom6_b200.api import Mom6cuError
    for ni, nj, nk in ((33, 9, 2), (130, 67, 5)):
        dom, grid, gv, cs, a = synthetic.pressureforce_inputs(ni, nj, nk, land_blocks=3)
        ref = _copy(a)
        oracle.pressure_force(dom, grid, gv, cs, ref)
        ctx = ctx_factory(dom)
        ctx.set_grid(grid); ctx.set_vgrid(gv)
        with pytest.raises(Mom6cuError):
            ctx.pressure_force(_copy(a))      # not initialised (:1117)
        bad = dict(cs); bad["unsupported"] = 1
        with pytest.raises(Mom6cuError):
            ctx.set_cs_pressureforce(bad)
        ctx.set_cs_pressureforce(cs)
        ctx.pressure_force(a)
        for k in ("PFu", "PFv", "pbce", "eta"):
            assert np.array_equal(ref[k].view(np.int64), a[k].view(np.int64)), (ni, nj, k)
