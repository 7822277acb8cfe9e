"""PINS the ALE-remapping oracle to the reference's own known-answer vectors: every vector below is copied from
remapping_unit_tests, /root/reference/src/ALE/MOM_remapping.F90:2072-2560 (answer_date = 20190101, h_neglect = 1e-30),
with the reference line cited per test.  test%real_arr with no tolerance is an exact comparison in the reference
(MOM_unit_testing.F90), so exact equality is asserted here too unless the reference passes tol=."""
import numpy as np

# remapping_CS defaults (:47-66): boundary_extrapolation=.true., force_bounds_in_target=.true., om4_remap_via_sub_cells=.false.
CS_PPM_H4 = dict(remapping_scheme=4, boundary_extrapolation=1, force_bounds_in_subcell=0, force_bounds_in_target=1,
                 om4_remap_via_sub_cells=0, answer_date=20190101, h_neglect=1.0e-30, h_neglect_edge=1.0e-30)


def test_remapping_core_h_ppm_h4(oracle):
    """:2155-2162  'remapping_core_h() 2/3/4' (initialize_remapping(CS,'PPM_H4',force_bounds_in_subcell=.false.))"""
    h0 = [0.75, 0.75, 0.75, 0.75]; u0 = [9., 3., -3., -9.]
    u2, _ = oracle.remapping_core_h(CS_PPM_H4, h0, u0, [0.5] * 6)
    assert np.array_equal(u2, [10., 6., 2., -2., -6., -10.])
    u2, _ = oracle.remapping_core_h(CS_PPM_H4, h0, u0, [.125] * 6)
    assert np.array_equal(u2, [11.5, 10.5, 9.5, 8.5, 7.5, 6.5])
    u2, _ = oracle.remapping_core_h(CS_PPM_H4, h0, u0, [2.25, 1.5, 1.])
    assert np.array_equal(u2, [3., -10.5, -12.])


def test_pcm_plm_reconstructions(oracle):
    """:2174-2212  PCM and PLM reconstruction vectors"""
    E, c = oracle.remap_reconstruct("PCM", [1., 1., 1.], [1., 2., 4.])
    assert np.array_equal(E[0], [1., 2., 4.]) and np.array_equal(E[1], [1., 2., 4.]) and np.array_equal(c[0], [1., 2., 4.])
    for u, EL, ER, P0, P1 in (([1., 3., 5.], [1., 2., 5.], [1., 4., 5.], [1., 2., 5.], [0., 2., 0.]),      # Unlim PLM
                              ([1., 2., 7.], [1., 1., 7.], [1., 3., 7.], [1., 1., 7.], [0., 2., 0.]),      # Left lim PLM
                              ([1., 6., 7.], [1., 5., 7.], [1., 7., 7.], [1., 5., 7.], [0., 2., 0.])):     # Right lim PLM
        E, c = oracle.remap_reconstruct("PLM", [1., 1., 1.], u)
        assert np.array_equal(E[0], EL) and np.array_equal(E[1], ER) and np.array_equal(c[0], P0) and np.array_equal(c[1], P1)
    E, c = oracle.remap_reconstruct("PLM", [1., 2., 3.], [1., 4., 9.])                                   # Non-uniform line PLM
    assert np.array_equal(E[0], [1., 2., 9.]) and np.array_equal(E[1], [1., 6., 9.])
    assert np.array_equal(c[0], [1., 2., 9.]) and np.array_equal(c[1], [0., 4., 0.])


def test_edge_values_h4_and_ppm(oracle):
    """:2214-2260  'Line H4', 'Line PPM', 'Parabola H4', 'Parabola PPM', 'Limits PPM'"""
    ones = [1.] * 5
    E, _ = oracle.remap_reconstruct("edge_h4", ones, [1., 3., 5., 7., 9.], h_neglect=1e-10)
    assert np.abs(E[0] - [0., 2., 4., 6., 8.]).max() <= 8.0e-15 and np.abs(E[1] - [2., 4., 6., 8., 10.]).max() <= 1.0e-14
    E, c = oracle.remap_reconstruct("PPM", ones, [1., 3., 5., 7., 9.], E=[[0., 2., 4., 6., 8.], [2., 4., 6., 8., 10.]])
    assert np.array_equal(c[0], [1., 2., 4., 6., 9.]) and np.array_equal(c[1], [0., 2., 2., 2., 0.]) and np.array_equal(c[2], [0.] * 5)
    E, _ = oracle.remap_reconstruct("edge_h4", ones, [1., 1., 7., 19., 37.], h_neglect=1e-10)
    assert np.abs(E[0] - [3., 0., 3., 12., 27.]).max() <= 2.7e-14 and np.abs(E[1] - [0., 3., 12., 27., 48.]).max() <= 4.8e-14
    E, c = oracle.remap_reconstruct("PPM", ones, [0., 1., 7., 19., 37.], E=[[0., 0., 3., 12., 27.], [0., 3., 12., 27., 48.]])
    assert np.array_equal(E[0], [0., 0., 3., 12., 37.]) and np.array_equal(E[1], [0., 3., 12., 27., 37.])
    assert np.array_equal(c[0], [0., 0., 3., 12., 37.]) and np.array_equal(c[1], [0., 0., 6., 12., 0.]) and np.array_equal(c[2], [0., 3., 3., 3., 0.])
    E, c = oracle.remap_reconstruct("PPM", ones, [0., 5., 7., 16., 15.], E=[[0., 0., 6., 10., 15.], [0., 6., 12., 17., 15.]])
    assert np.array_equal(E[0], [0., 3., 6., 16., 15.]) and np.array_equal(E[1], [0., 6., 9., 16., 15.])
    assert np.array_equal(c[0], [0., 3., 6., 16., 15.]) and np.array_equal(c[1], [0., 6., 0., 0., 0.]) and np.array_equal(c[2], [0., -3., 3., 0., 0.])


def _chk(o, **exp):
    for k, v in exp.items():
        assert np.array_equal(o[k], v), (k, o[k], v)


def test_intersect_src_tgt_grids(oracle):
    """:2262-2470  intersect_src_tgt_grids tests 1-5"""
    _chk(oracle.remap_intersect([3., 3.], [2., 2., 2.]), h_sub=[0., 2., 1., 1., 2., 0.], h0_eff=[3., 3.], isrc_start=[1, 4], isrc_end=[3, 5],
         isrc_max=[2, 5], itgt_start=[1, 3, 5], itgt_end=[2, 4, 6], isub_src=[1, 1, 1, 2, 2, 2])
    _chk(oracle.remap_intersect([2., 2., 2.], [3., 3.]), h_sub=[0., 2., 1., 1., 2., 0.], h0_eff=[2., 2., 2.], isrc_start=[1, 3, 5],
         isrc_end=[2, 4, 5], isrc_max=[2, 4, 5], itgt_start=[1, 4], itgt_end=[3, 6], isub_src=[1, 1, 2, 2, 3, 3])
    _chk(oracle.remap_intersect([2., 4.], [2., 2., 2.]), h_sub=[0., 2., 0., 2., 2., 0.], h0_eff=[2., 4.], isrc_start=[1, 3], isrc_end=[2, 5],
         isrc_max=[2, 5], itgt_start=[1, 4, 5], itgt_end=[3, 4, 6], isub_src=[1, 1, 2, 2, 2, 2])
    _chk(oracle.remap_intersect([2., 4.], [2., 2., 1.]), h_sub=[0., 2., 0., 2., 1., 1.], h0_eff=[2., 3.], isrc_start=[1, 3], isrc_end=[2, 6],
         isrc_max=[2, 4], itgt_start=[1, 4, 5], itgt_end=[3, 4, 5], isub_src=[1, 1, 2, 2, 2, 2])
    _chk(oracle.remap_intersect([2., 2., 1.], [2., 4.]), h_sub=[0., 2., 0., 2., 1., 1.], h0_eff=[2., 2., 1.], isrc_start=[1, 3, 5],
         isrc_end=[2, 4, 5], isrc_max=[2, 4, 5], itgt_start=[1, 4], itgt_end=[3, 6], isub_src=[1, 1, 2, 2, 3, 3])


def test_src_to_sub_and_sub_to_tgt(oracle):
    """:2341-2362 (test 3), :2396-2405 (test 4), :2437-2462 (test 5): 'u_sub om4', 'u_sub', 'u1'"""
    for om4 in (1, 0):
        us, u1 = oracle.remap_plm_sub(om4, [2., 4.], [2., 5.], [2., 2., 2.])
        assert np.array_equal(us, [1., 2., 3., 4., 6., 7.]) and np.array_equal(u1, [2., 4., 6.])
        us, u1 = oracle.remap_plm_sub(om4, [2., 2., 1.], [2., 4., 5.5], [2., 4.])
        assert np.array_equal(us, [1., 2., 3., 4., 5.5, 6.]) and np.array_equal(u1, [2., 4.875])
    us, _ = oracle.remap_plm_sub(0, [2., 4.], [2., 5.], [2., 2., 1.])
    assert np.array_equal(us, [1., 2., 3., 4., 5.5, 6.5])


def test_conservation_and_bounds_property(oracle):
    """The reference's brute-force property tests (:2590-2640 'remapping_core_h() conservation/bounds'):
    sum(u1*h1) == sum(u0*h0) to round-off and u1 within the range of u0, for random grids, all schemes, both sub-cell variants."""
    r = np.random.default_rng(7)
    for scheme in (0, 2, 4, 5):
        for extrap in (0, 1):
            for om4 in (0, 1):
                cs = dict(CS_PPM_H4, remapping_scheme=scheme, boundary_extrapolation=extrap, om4_remap_via_sub_cells=om4)
                for _ in range(40):
                    n0, n1 = int(r.integers(1, 12)), int(r.integers(1, 12))
                    h0 = r.uniform(0, 1, n0) * (r.uniform(0, 1, n0) > 0.2); h1 = r.uniform(0, 1, n1) * (r.uniform(0, 1, n1) > 0.2)
                    if h0.sum() == 0 or h1.sum() == 0:
                        continue
                    h1 *= h0.sum() / h1.sum()
                    u0 = r.uniform(-1, 1, n0)
                    u1, err = oracle.remapping_core_h(cs, h0, u0, h1)
                    assert abs((u1 * h1).sum() - (u0 * h0).sum()) <= max(4 * err, 1e-13), (scheme, extrap, om4, n0, n1)
                    if not extrap:
                        assert u1.min() >= u0.min() - 1e-12 and u1.max() <= u0.max() + 1e-12


def test_recon1d_unit_test_vectors_pin_the_edge_values_of_the_pressure_force(oracle):
    """The reference's class-based reconstructions hold their own check values: Recon1d_MPLM_WA.unit_tests (src/ALE/Recon1d_MPLM_WA.F90:248-262:
    the OM4-era monotonized PLM, i.e. PLM_slope_wa + PLM_monotonized_slope) gives, for h = (2,2,2), u = (1,3,5), left edges (1,2,5) and right
    edges (1,4,5); Recon1d_PPM_H4_2019.unit_tests (:522-530) gives, for five cells of thickness 2 and u = (1,3,5,7,9), left edges (1,2,4,6,9)
    and right edges (1,4,6,8,9) (to 2 bits of roundoff).  The column routines RECONSTRUCT_FOR_PRESSURE uses (ALE_PLM_edge_values / TS_PPM_edge_values,
    MOM_ALE.F90:1518-1660 -- the same PLM functions; PPM with the implicit-h4 edge values, exact for a linear profile) must reproduce them, and so
    must the remapping reconstructions they share code with."""
    qt, qb = oracle.ale_edge_values(1, [2., 2., 2.], [1., 3., 5.])
    assert np.array_equal(qt, [1., 2., 5.]) and np.array_equal(qb, [1., 4., 5.])
    qt, qb = oracle.ale_edge_values(2, [2.] * 5, [1., 3., 5., 7., 9.])
    assert np.allclose(qt, [1., 2., 4., 6., 9.], rtol=0, atol=4 * np.finfo(float).eps * 9) and np.allclose(qb, [1., 4., 6., 8., 9.], rtol=0, atol=4e-15)
    # with boundary extrapolation the linear profile is continued into the end cells exactly (PLM_extrapolate_slope :160)
    qt, qb = oracle.ale_edge_values(1, [2., 2., 2.], [1., 3., 5.], bdry_extrap=True)
    assert np.array_equal(qt, [0., 2., 4.]) and np.array_equal(qb, [2., 4., 6.])
    # the remapping reconstructions (PLM = scheme 2, PPM_H4 = scheme 4) on the same vectors
    E, _ = oracle.remap_reconstruct("PLM", [2., 2., 2.], [1., 3., 5.])
    assert np.array_equal(E[0], [1., 2., 5.]) and np.array_equal(E[1], [1., 4., 5.])
    E, _ = oracle.remap_reconstruct("edge_h4", [2.] * 5, [1., 3., 5., 7., 9.])              # explicit h4 edge values, then the PPM limiter
    E, _ = oracle.remap_reconstruct("PPM", [2.] * 5, [1., 3., 5., 7., 9.], E=E)
    assert np.allclose(E[0], [1., 2., 4., 6., 9.], rtol=0, atol=4e-15) and np.allclose(E[1], [1., 4., 6., 8., 9.], rtol=0, atol=4e-15)
