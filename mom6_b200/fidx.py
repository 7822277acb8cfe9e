"""Fortran-indexed numpy arrays: A[i0:i1, j0:j1] in the reference's inclusive (i,j[,k]) indices.

Storage is C-contiguous with i fastest -- byte-for-byte the Fortran column-major array the
C ABI expects: 2-D (nj, ni); 3-D (nk, nj, ni); array-of-structs (nj, ni, nm) with m fastest.
"""
import numpy as np


class FA:
    def __init__(self, ilo, ihi, jlo, jhi, nk=None, nm=None, fill=0.0):
        self.ilo, self.ihi, self.jlo, self.jhi, self.nk, self.nm = ilo, ihi, jlo, jhi, nk, nm
        ni, nj = ihi - ilo + 1, jhi - jlo + 1
        shape = (nj, ni)
        if nk is not None:
            shape = (nk, nj, ni)
        if nm is not None:
            shape = shape + (nm,)
        self.a = np.full(shape, fill, dtype=np.float64)

    def s(self, i0, i1, j0, j1):
        """View of the inclusive index box (all k / m)."""
        sl = (slice(j0 - self.jlo, j1 - self.jlo + 1), slice(i0 - self.ilo, i1 - self.ilo + 1))
        if self.nk is not None:
            sl = (slice(None),) + sl
        return self.a[sl]

    def copy(self):
        o = FA.__new__(FA)
        o.__dict__.update(self.__dict__)
        o.a = self.a.copy()
        return o


def extent(dom, stagger, wide=False):
    """(ilo, ihi, jlo, jhi) of a field of the given stagger ('h','u','v','q')."""
    su = 1 if stagger in ("u", "q") else 0
    sv = 1 if stagger in ("v", "q") else 0
    if wide:
        return dom.isdw - su, dom.iedw, dom.jsdw - sv, dom.jedw
    return dom.isd - su, dom.ied, dom.jsd - sv, dom.jed


def new(dom, stagger, wide=False, nk=None, nm=None, fill=0.0):
    return FA(*extent(dom, stagger, wide), nk=nk, nm=nm, fill=fill)


def fill_halo(dom, f, stagger):
    """Single-tile halo update with the symmetric-memory semantics of mpp_update_domains
    (config_src/infra/FMS2/MOM_domain_infra.F90:171-216): reentrant directions wrap, closed
    edges are untouched, the shared edge of staggered fields is never overwritten."""
    su = 1 if stagger in ("u", "q") else 0
    sv = 1 if stagger in ("v", "q") else 0
    ni, nj = dom.iec - dom.isc + 1, dom.jec - dom.jsc + 1
    if dom.cyclic_x:
        w0, w1 = f.ilo, dom.isc - 1 - su
        if w1 >= w0:
            f.s(w0, w1, f.jlo, f.jhi)[...] = f.s(w0 + ni, w1 + ni, f.jlo, f.jhi)
        e0, e1 = dom.iec + 1, f.ihi
        if e1 >= e0:
            f.s(e0, e1, f.jlo, f.jhi)[...] = f.s(e0 - ni, e1 - ni, f.jlo, f.jhi)
    if dom.cyclic_y:
        s0, s1 = f.jlo, dom.jsc - 1 - sv
        if s1 >= s0:
            f.s(f.ilo, f.ihi, s0, s1)[...] = f.s(f.ilo, f.ihi, s0 + nj, s1 + nj)
        n0, n1 = dom.jec + 1, f.jhi
        if n1 >= n0:
            f.s(f.ilo, f.ihi, n0, n1)[...] = f.s(f.ilo, f.ihi, n0 - nj, n1 - nj)
    return f
