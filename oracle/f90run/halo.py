"""Halo updates for the translated reference routines on ONE PE: pass_var / pass_vector / group passes of MOM_domains.
TEST INFRASTRUCTURE ONLY.

FMS (the library behind MOM_domains) is not part of the reference tree, so this is a restatement of what its halo update does
on a single PE with symmetric memory: points outside the computational domain are filled from their periodic image in every
reentrant direction, closed edges are left untouched, and the shared edge of a staggered field (I = isc-1 for u/q points,
J = jsc-1 for v/q points) belongs to the computational domain and is not overwritten.  No tripolar fold."""
from .rt import FArray, NS


def domain_type(dom):
    """MOM_domain_type stand-in: computational-domain bounds and reentrancy (index space shared by G and the wide BT domain)"""
    return NS(isc=int(dom.isc), iec=int(dom.iec), jsc=int(dom.jsc), jec=int(dom.jec), isd=int(dom.isd), ied=int(dom.ied),
              jsd=int(dom.jsd), jed=int(dom.jed), cyclic_x=bool(dom.cyclic_x), cyclic_y=bool(dom.cyclic_y), is_domain=True)


def _src(lo, hi, n, clo, chi, cyclic):
    """for each index lo..hi: the computational index it is filled from (itself inside clo..chi), or None"""
    out = []
    for i in range(lo, hi + 1):
        if clo <= i <= chi:
            out.append(i)
        elif cyclic:
            s = i
            while s < clo:
                s += n
            while s > chi:
                s -= n
            out.append(s if clo <= s <= chi else None)
        else:
            out.append(None)
    return out


def update(a, D, xstag=None, ystag=None):
    """fill the halo of FArray a (rank 2 or 3, horizontal dimensions first) on domain D.  The array may be seen through any
    lower bounds (an array section passed to create_group_pass starts at 1): its position in the domain follows from its
    extents, which are those of a memory domain with equal halos on both sides."""
    ni, nj = D.iec - D.isc + 1, D.jec - D.jsc + 1
    nxh, nyh = a.shape[0], a.shape[1]
    if xstag is None:
        xstag = ((nxh - ni) % 2) == 1
    if ystag is None:
        ystag = ((nyh - nj) % 2) == 1
    wi, wj = (nxh - ni - (1 if xstag else 0)) // 2, (nyh - nj - (1 if ystag else 0)) // 2
    ilo, jlo = D.isc - wi - (1 if xstag else 0), D.jsc - wj - (1 if ystag else 0)   # true index of the first element
    ihi, jhi = ilo + nxh - 1, jlo + nyh - 1
    si = _src(ilo, ihi, ni, D.isc - (1 if xstag else 0), D.iec, D.cyclic_x)
    sj = _src(jlo, jhi, nj, D.jsc - (1 if ystag else 0), D.jec, D.cyclic_y)
    v, s = a.v, a.s
    b = a.b + a.lb[0] * s[0] + a.lb[1] * s[1]    # offset of the first horizontal element
    nk = a.shape[2] if len(a.shape) == 3 else 1
    s2 = s[2] if len(a.shape) == 3 else 0
    b += (a.lb[2] * s2) if len(a.shape) == 3 else 0
    for jj, j in enumerate(range(jlo, jhi + 1)):
        js = sj[jj]
        for ii, i in enumerate(range(ilo, ihi + 1)):
            is_ = si[ii]
            if is_ == i and js == j:
                continue
            if is_ is None or js is None:
                continue
            d = b + ii * s[0] + jj * s[1]
            o = b + (is_ - ilo) * s[0] + (js - jlo) * s[1]
            for k in range(nk):
                v[d + k * s2] = v[o + k * s2]


def pass_var(array=None, mom_dom=None, sideflag=None, complete=None, position=None, halo=None, inner_halo=None, clock=None):
    update(array, mom_dom)


def pass_vector(u_cmpt=None, v_cmpt=None, mom_dom=None, direction=None, stagger=None, complete=None, halo=None, clock=None):
    update(u_cmpt, mom_dom)
    update(v_cmpt, mom_dom)


def create_group_pass(group, a=None, b=None, c=None, *rest, **kw):
    # a group that has already been used is re-pointed at the new arrays (mpp_reset_group_update_field), in the same order
    if group.fields is None or group.used:
        group.fields = []
        group.used = False
    if isinstance(b, FArray):   # (group, u_cmpt, v_cmpt, MOM_dom, ...)
        group.fields += [a, b]
    else:                        # (group, array, MOM_dom, ...)
        group.fields.append(a)


def do_group_pass(group=None, mom_dom=None, clock=None):
    for f in (group.fields or []):
        update(f, mom_dom)
    group.used = True


def start_group_pass(group=None, mom_dom=None, clock=None):
    do_group_pass(group, mom_dom)


def complete_group_pass(group=None, mom_dom=None, clock=None):
    return None


STUBS = {"pass_var": pass_var, "pass_vector": pass_vector, "create_group_pass": create_group_pass, "do_group_pass": do_group_pass,
         "start_group_pass": start_group_pass, "complete_group_pass": complete_group_pass,
         "pass_var_start": lambda *a, **k: (pass_var(*a[:2]) or 0), "pass_var_complete": lambda *a, **k: None,
         "pass_vector_start": lambda *a, **k: (pass_vector(*a[:3]) or 0), "pass_vector_complete": lambda *a, **k: None}
