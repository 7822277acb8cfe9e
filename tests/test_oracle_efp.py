"""Pins the oracle's extended-fixed-point sums with the reference's own unit test,
config_src/drivers/unit_tests/test_reproducing_sum.F90: the benchmark-topography-like array (:192-207) whose reproducing sum
must agree with the plain sum to the random-walk round-off bound (:76-98), fast == checked conversion (:100-110), the exact
sum of 1..N (:114-125), and order invariance under random element swaps for integers and for random numbers (:127-150).
Also the EFP operators (MOM_coms.F90:737-815) and the bit-count checksum windows (MOM_checksums.F90)."""
import numpy as np
import pytest

from mom6_b200.api import make_domain

NI, NJ = 200, 300  # the unit test's fallback n_global (:52)


def _dom(ni=NI, nj=NJ, nk=1, halo=2):
    return make_domain(ni, nj, nk=nk, halo=halo)  # the unit test's domain: halo 2, not reentrant (:63)


def generate_array_of_values(dom):
    """generate_array_of_values :180-207 (note: y uses idg_offset, as the reference does)."""
    D = np.zeros((dom.jed - dom.jsd + 1, dom.ied - dom.isd + 1))
    PI = 4.0 * np.arctan(1.0)
    h = dom.isc - dom.isd
    for j in range(1, NJ + 1):
        for i in range(1, NI + 1):
            x = float(i) / float(NI)
            y = float(j) / float(NJ)
            d = -3000.0 * (y * (1.0 + 0.6 * np.cos(4.0 * PI * x)) + 0.75 * np.exp(-6.0 * y) + 0.05 * np.cos(10.0 * PI * x) - 0.7)
            if d > 3000.0:
                d = 3000.0
            if d < 1.0:
                d = 0.0
            D[j - 1 + h, i - 1 + h] = d
    return D


def window(dom):
    o = dom.isd - 1
    return dict(isr=dom.isc - o, ier=dom.iec - o, jsr=dom.jsc - (dom.jsd - 1), jer=dom.jec - (dom.jsd - 1))


def randomly_swap_elements(rng, dom, a):
    """randomly_swap_elements :155-177"""
    h = dom.isc - dom.isd
    ni, nj = dom.iec - dom.isc, dom.jec - dom.jsc
    n_swaps = ni * nj
    r = rng.random((n_swaps, 4))
    i0 = h + (r[:, 0] * ni).astype(int); j0 = h + (r[:, 1] * nj).astype(int)
    i1 = h + (r[:, 2] * ni).astype(int); j1 = h + (r[:, 3] * nj).astype(int)
    for q in range(n_swaps):
        if i0[q] != i1[q] and j0[q] != j1[q]:
            a[j0[q], i0[q]], a[j1[q], i1[q]] = a[j1[q], i1[q]], a[j0[q], i0[q]]


def test_reproducing_sum_unit_test(oracle):
    dom = _dom()
    w = window(dom)
    a = generate_array_of_values(dom)
    inner = a[dom.jsc - dom.jsd:dom.jec - dom.jsd + 1, dom.isc - dom.isd:dom.iec - dom.isd + 1]
    # error estimate :76-88
    eps = np.finfo(np.float64).eps
    error_bound, tot = 0.0, 0.0
    for v in inner.ravel():
        error_bound += max(abs(tot), abs(v)) * eps
        tot += v
    N = NI * NJ
    likely_error = tot * eps * np.sqrt(float(N))
    assert likely_error <= error_bound
    tot_std = oracle.reproducing_sum(dom, a, reproducing=False, **w)["sum"]
    tot_R = oracle.reproducing_sum(dom, a, **w)["sum"]
    tot_fastR = oracle.reproducing_sum(dom, a, overflow_check=False, **w)["sum"]
    assert abs(tot_std - tot_R) <= likely_error
    assert tot_fastR == tot_R
    # the exact sum of 1..N :114-125
    h = dom.isc - dom.isd
    a[:] = 0.0
    a[h:h + NJ, h:h + NI] = 1.0 + np.arange(N, dtype=np.float64).reshape(NJ, NI)
    exact = 0.5 * float(N) * float(N + 1)
    assert oracle.reproducing_sum(dom, a, **w)["sum"] == exact
    rng = np.random.default_rng(5)
    for _ in range(5):
        randomly_swap_elements(rng, dom, a)
        assert oracle.reproducing_sum(dom, a, **w)["sum"] == exact
    # random numbers :137-150
    a = rng.random(a.shape)
    ref = oracle.reproducing_sum(dom, a, want_efp=True, **w)
    for _ in range(5):
        randomly_swap_elements(rng, dom, a)
        got = oracle.reproducing_sum(dom, a, want_efp=True, **w)
        assert got["sum"] == ref["sum"] and np.array_equal(got["EFP_sum"], ref["EFP_sum"])


def test_large_windows_take_the_row_carry_branch(oracle):
    """More than max_count_prec = 131071 elements (:171-183): per-row carries give the same sum as any other grouping."""
    dom = _dom(600, 300, halo=2)
    rng = np.random.default_rng(6)
    a = (rng.random((dom.jed, dom.ied)) - 0.3) * 1.0e9
    w = window(dom)
    full = oracle.reproducing_sum(dom, a, want_efp=True, **w)
    o = dom.isd - 1
    half = (dom.jsc + dom.jec) // 2
    s1 = oracle.reproducing_sum(dom, a, want_efp=True, isr=w["isr"], ier=w["ier"], jsr=w["jsr"], jer=half - o)
    s2 = oracle.reproducing_sum(dom, a, want_efp=True, isr=w["isr"], ier=w["ier"], jsr=half - o + 1, jer=w["jer"])
    both = oracle.efp_op("plus", s1["EFP_sum"], s2["EFP_sum"])
    assert oracle.efp_op("to_real", both) == full["sum"]
    import math
    assert abs(full["sum"] - math.fsum(a[dom.jsc - 1:dom.jec, dom.isc - 1:dom.iec].ravel())) <= 2.0 * np.spacing(full["sum"])


def test_3d_layer_sums_and_unscale(oracle):
    dom = _dom(30, 20, nk=5, halo=3)
    rng = np.random.default_rng(7)
    a = rng.standard_normal((5, dom.jed, dom.ied)) * 1e3
    w = window(dom)
    r = oracle.reproducing_sum(dom, a, want_sums=True, want_efp=True, want_lay_efp=True, **w)
    tot = 0.0
    for k in range(5):
        one = oracle.reproducing_sum(dom, a[k], **w)["sum"]
        assert one == r["sums"][k]
        assert oracle.efp_op("to_real", r["EFP_lay_sums"][k]) == one
        tot = tot + one
    assert tot == r["sum"]
    # the single-accumulator form converts the exact total once (:528-529)
    r1 = oracle.reproducing_sum(dom, a, want_efp=True, **w)
    assert r1["sum"] == oracle.efp_op("to_real", r["EFP_sum"])
    # power-of-two unscaling is exact and is undone in the returned value (:535-543)
    r2 = oracle.reproducing_sum(dom, a * 2.0**-20, unscale=2.0**20, want_sums=True, **w)
    assert np.array_equal(r2["sums"], r["sums"] * 2.0**-20) and r2["sum"] == r["sum"] * 2.0**-20


def test_efp_operators(oracle):
    for x in (0.0, 1.0, -1.0, 3.5e10, -7.25e-12, 2.0**100, -(2.0**-130), 1.0 / 3.0, -1.0e30):
        e = oracle.efp_op("from_real", x)
        assert oracle.efp_op("to_real", e) == x
    a, b = oracle.efp_op("from_real", 1.0e15 + 0.25), oracle.efp_op("from_real", 1.0e15)
    assert oracle.efp_op("diff", a, b) == 0.25
    assert oracle.efp_op("to_real", oracle.efp_op("plus", a, b)) == 2.0e15 + 0.25
    assert oracle.efp_op("to_real", oracle.efp_op("minus", b, a)) == -0.25
    with pytest.raises(OverflowError):
        oracle.efp_op("from_real", 2.0**140)  # real_to_EFP overflows beyond prec * pr(1) = 2**138
    with pytest.raises(RuntimeError):
        dom = _dom(8, 8, halo=1)
        a = np.ones((dom.jed, dom.ied)); a[3, 3] = np.nan
        oracle.reproducing_sum(dom, a)


def test_bitcount_checksums(oracle):
    """subchk: the sum of popcnt(|scale*x|) over the shifted h-point window, mod 10**9 (MOM_checksums.F90:520-529)."""
    dom = _dom(12, 9, nk=3, halo=3)
    rng = np.random.default_rng(8)

    def bc_window(a, i0, i1, j0, j1, ilo, jlo, scale):
        sub = np.abs(scale * a[..., j0 - jlo:j1 - jlo + 1, i0 - ilo:i1 - ilo + 1])
        return int(np.unpackbits(np.ascontiguousarray(sub).view(np.uint8)).sum()) % 1000000000

    for stagger, (di, dj) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
        ilo, jlo = dom.isd - di, dom.jsd - dj
        a = rng.standard_normal((3, dom.jed - jlo + 1, dom.ied - ilo + 1))
        for hs in (0, 1, 3):
            for sym in (False, True):
                if stagger == 0 and sym:
                    continue
                for omit in (False, True):
                    bc, kind, st = oracle.chksum(dom, a, stagger, hs, sym, omit, scale=0.5, stats=True)
                    assert bc[0] == bc_window(a, dom.isc, dom.iec, dom.jsc, dom.jec, ilo, jlo, 0.5)
                    # the rank-3 B-point form widens its corner windows with or without `symmetric` (chksum_B_3d :1698-1706)
                    ex = 1 if ((sym or stagger == 3) and di) else 0
                    ey = 1 if ((sym or stagger == 3) and dj) else 0
                    if kind == 2:
                        assert bc[1] == bc_window(a, dom.isc - hs - ex, dom.iec - hs - ex, dom.jsc - hs - ey, dom.jec - hs - ey, ilo, jlo, 0.5)
                        assert bc[4] == bc_window(a, dom.isc + hs, dom.iec + hs, dom.jsc + hs, dom.jec + hs, ilo, jlo, 0.5)
                    elif kind == 4:
                        assert stagger == 1 and hs == 0 and sym
                        assert bc[1] == bc_window(a, dom.isc - 1, dom.iec - 1, dom.jsc, dom.jec, ilo, jlo, 0.5)
                    elif kind == 5:
                        assert stagger == 2 and hs == 0 and sym
                        assert bc[1] == bc_window(a, dom.isc, dom.iec, dom.jsc - 1, dom.jec - 1, ilo, jlo, 0.5)
                    elif kind == 1:
                        assert hs == 0 and not sym
                    inner = 0.5 * a[:, dom.jsc - jlo:dom.jec - jlo + 1, dom.isc - ilo:dom.iec - ilo + 1]
                    assert abs(st[0] - inner.mean()) < 1e-12
                    sym_stats = sym or hs > 0
                    wide = 0.5 * a[:, dom.jsc - jlo - (1 if (sym_stats and dj) else 0):dom.jec - jlo + 1,
                                   dom.isc - ilo - (1 if (sym_stats and di) else 0):dom.iec - ilo + 1]
                    assert st[1] == wide.min() and st[2] == wide.max()
    with pytest.raises(RuntimeError):
        oracle.chksum(dom, np.zeros((dom.jed, dom.ied)), 0, 4)  # halo shift wider than the halo: the FATAL of :465-471
