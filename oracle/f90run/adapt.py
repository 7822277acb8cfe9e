"""Build the reference's derived types (ocean_grid_type, verticalGrid_type, unit_scale_type, control structures) for the
translated reference routines from the dictionaries mom6_b200.synthetic produces.  TEST INFRASTRUCTURE ONLY.

numpy arrays here are C-ordered [k, j, i] over the memory domain, i.e. exactly the Fortran (i, j, k) arrays the reference
declares with SZI_/SZIB_/SZJ_/SZJB_ (symmetric memory): only the lower bounds have to be attached."""
import numpy as np

from .rt import FArray, NS


def stagger_lb(dom, st):
    """lower bounds (i, j) of an array at h / u / v / q points of G's memory domain (symmetric memory)"""
    return {"h": (dom.isd, dom.jsd), "u": (dom.isd - 1, dom.jsd), "v": (dom.isd, dom.jsd - 1), "q": (dom.isd - 1, dom.jsd - 1)}[st]


def stagger_of(dom, a):
    """the stagger of a numpy array over G's memory domain, from its horizontal shape"""
    ni, nj = dom.ied - dom.isd + 1, dom.jed - dom.jsd + 1
    sh = a.shape[-2:]
    for st, (di, dj) in (("h", (0, 0)), ("u", (1, 0)), ("v", (0, 1)), ("q", (1, 1))):
        if sh == (nj + dj, ni + di):
            return st
    raise ValueError(f"array of horizontal shape {sh} is not on G's memory domain")


def farr(dom, a, st=None, klb=1):
    """numpy [k,j,i] or [j,i] array over G's memory domain -> FArray with the reference's bounds"""
    if a is None:
        return None
    a = np.asarray(a)
    if st is None:
        st = stagger_of(dom, a)
    lb = stagger_lb(dom, st)
    if a.ndim == 3:
        lb = lb + (klb,)
    return FArray.from_numpy(a, lb)


def back(fa, out):
    """copy an FArray's elements into the numpy array it was made from"""
    out[...] = fa.to_numpy().reshape(out.shape)


_GRID_ST = {"T": "h", "Cu": "u", "Cv": "v", "Bu": "q"}


def grid_type(dom, grid):
    """ocean_grid_type (src/core/MOM_grid.F90:30-200) with the members the hot path reads"""
    G = NS()
    for k in ("isc", "iec", "jsc", "jec", "isd", "ied", "jsd", "jed"):
        setattr(G, k, int(getattr(dom, k)))
    G.iscb, G.iecb, G.jscb, G.jecb = G.isc - 1, G.iec, G.jsc - 1, G.jec
    G.isdb, G.iedb, G.jsdb, G.jedb = G.isd - 1, G.ied, G.jsd - 1, G.jed
    G.ke = int(dom.nk)
    G.symmetric = True
    G.first_direction = int(dom.first_direction)
    G.nonblocking_updates = False
    G.isd_global, G.jsd_global = G.isd, G.jsd
    G.idg_offset, G.jdg_offset = 0, 0
    G.z_ref = 0.0
    G.domain = NS(cyclic_x=bool(dom.cyclic_x), cyclic_y=bool(dom.cyclic_y))
    G.hi = NS(isc=G.isc, iec=G.iec, jsc=G.jsc, jec=G.jec, isd=G.isd, ied=G.ied, jsd=G.jsd, jed=G.jed,
              iscb=G.iscb, iecb=G.iecb, jscb=G.jscb, jecb=G.jecb, isdb=G.isdb, iedb=G.iedb, jsdb=G.jsdb, jedb=G.jedb)
    for name, a in grid.items():
        if not isinstance(a, np.ndarray) or a.ndim != 2:
            if np.isscalar(a):
                setattr(G, name.lower(), a)
            continue
        setattr(G, name.lower(), farr(dom, a))
    return G


def vgrid_type(dom, gv):
    """verticalGrid_type (src/core/MOM_verticalGrid.F90:22-90)"""
    GV = NS(ke=int(dom.nk))
    for k, v in gv.items():
        if isinstance(v, np.ndarray):
            setattr(GV, k.lower(), FArray.from_numpy(v, (1,)))
        else:
            setattr(GV, k.lower(), (bool(v) if k == "Boussinesq" else float(v)))
    return GV


def unit_scale_type(us=None):
    """unit_scale_type (src/framework/MOM_unit_scaling.F90:12-60): every factor 1 unless given"""
    U = NS()
    names = ("m_to_Z Z_to_m m_to_L L_to_m s_to_T T_to_s R_to_kg_m3 kg_m3_to_R Q_to_J_kg J_kg_to_Q C_to_degC degC_to_C "
             "S_to_ppt ppt_to_S Z_to_L L_to_Z L_T_to_m_s m_s_to_L_T L_T2_to_m_s2 Z2_T_to_m2_s m2_s_to_Z2_T kg_m3_to_R "
             "RZ_to_kg_m2 kg_m2s_to_RZ_T RZ_T_to_kg_m2s RZ3_T3_to_W_m2 W_m2_to_RZ3_T3 L_T_to_m_s Pa_to_RL2_T2 RL2_T2_to_Pa "
             "Pa_to_RLZ_T2 RLZ_T2_to_Pa QRZ_T_to_W_m2 W_m2_to_QRZ_T").split()
    for n in names:
        setattr(U, n.lower(), 1.0)
    for k, v in (us or {}).items():
        setattr(U, k.lower(), float(v))
    return U


def cs_type(cs, logical=()):
    """a control structure from a dict: integers named in `logical` (or holding 0/1 flags by convention) become logicals"""
    C = NS()
    for k, v in cs.items():
        if isinstance(v, np.ndarray):
            continue
        if k in logical:
            v = bool(v)
        setattr(C, k.lower(), v)
    return C
