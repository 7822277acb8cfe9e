"""Shared column generators for the remapping parity tests (host harness and GPU)."""
import itertools
import numpy as np


def columns(rng, ncol, n0, n1, kind):
    """ncol random (h0, u0, h1) columns with equal total thickness (except kind 'unequal')."""
    h0 = rng.uniform(0.1, 2.0, (ncol, n0)); h1 = rng.uniform(0.1, 2.0, (ncol, n1))
    u0 = rng.uniform(-2.0, 30.0, (ncol, n0))
    if kind == "vanished":          # many zero-thickness layers on both grids
        h0 *= rng.uniform(0, 1, (ncol, n0)) > 0.4; h1 *= rng.uniform(0, 1, (ncol, n1)) > 0.4
        h0[:, 0] += 0.05; h1[:, 0] += 0.05
    elif kind == "zstar":           # the regridding case: the target is a small perturbation of the source
        if n0 == n1:
            h1 = h0 * (1.0 + 0.05 * rng.uniform(-1, 1, (ncol, n0)))
    elif kind == "smooth":          # smooth profile, strongly stretched grid
        z = np.cumsum(h0, axis=1); u0 = 20.0 * np.exp(-z / z[:, -1:]) + 0.01 * rng.uniform(-1, 1, (ncol, n0))
        h0 *= np.linspace(0.01, 3.0, n0)[None, :]
    elif kind == "tiny":            # thicknesses down to 1e-12 mixed with O(1)
        h0 *= 10.0 ** rng.integers(-12, 1, (ncol, n0)); h1 *= 10.0 ** rng.integers(-12, 1, (ncol, n1))
    if kind != "unequal":
        h1 *= (h0.sum(axis=1) / h1.sum(axis=1))[:, None]
    return np.ascontiguousarray(h0), np.ascontiguousarray(u0), np.ascontiguousarray(h1)


def cs_variants():
    for scheme, extrap, om4, fbs, fbt in itertools.product((0, 2, 4, 5), (0, 1), (0, 1), (0, 1), (1, 0)):
        yield dict(remapping_scheme=scheme, boundary_extrapolation=extrap, force_bounds_in_subcell=fbs, force_bounds_in_target=fbt,
                   om4_remap_via_sub_cells=om4, answer_date=20190101, h_neglect=1.0e-30, h_neglect_edge=1.0e-30)


SHAPES = [(1, 1), (1, 5), (2, 2), (3, 4), (4, 3), (5, 5), (7, 12), (12, 7), (35, 35), (75, 75), (75, 50), (100, 128)]
KINDS = ["plain", "vanished", "zstar", "smooth", "tiny", "unequal"]
