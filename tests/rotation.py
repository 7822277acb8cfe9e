"""Quarter-turn index rotation of the inputs of the hot path: the reference's ROTATE_INDEX test (.testing `rotate` target;
src/framework/MOM_array_transform.F90:73-94, :176-243; rotate_dyn_horgrid src/framework/MOM_dyn_horgrid.F90:300-380;
rotate_initial_state src/core/MOM.F90:4357) re-expressed on dict inputs.  A field A(i,j) becomes A'(i',j') = A(j', n-i'+1)
("transpose, then row reverse"); a vector (A_u, A_v) becomes (-rot(A_v), rot(A_u)); a u/v pair of scalars swaps; x-first becomes
y-first; reentrant-in-x becomes reentrant-in-y."""
import numpy as np

from mom6_b200.api import make_domain


def rot(a):
    """rotate_array, turns = 1, on [..., j, i] storage."""
    return np.ascontiguousarray(np.swapaxes(a, -1, -2)[..., ::-1])


def unrot(a):
    """rotate_array, turns = -1."""
    return np.ascontiguousarray(np.swapaxes(a[..., ::-1], -1, -2))


# rotated key <- source key (rotate_dyn_horgrid; the inverse metrics are set_derived_dyn_horgrid of the rotated ones)
GRID_ROT = {"dxT": "dyT", "dyT": "dxT", "IdxT": "IdyT", "IdyT": "IdxT", "areaT": "areaT", "IareaT": "IareaT", "mask2dT": "mask2dT",
            "bathyT": "bathyT",
            "dxCu": "dyCv", "dyCu": "dxCv", "IdxCu": "IdyCv", "IdyCu": "IdxCv", "dy_Cu": "dx_Cv", "areaCu": "areaCv", "IareaCu": "IareaCv",
            "mask2dCu": "mask2dCv",
            "dxCv": "dyCu", "dyCv": "dxCu", "IdxCv": "IdyCu", "IdyCv": "IdxCu", "dx_Cv": "dy_Cu", "areaCv": "areaCu", "IareaCv": "IareaCu",
            "mask2dCv": "mask2dCu",
            "dxBu": "dyBu", "dyBu": "dxBu", "IdxBu": "IdyBu", "IdyBu": "IdxBu", "areaBu": "areaBu", "IareaBu": "IareaBu", "mask2dBu": "mask2dBu",
            "CoriolisBu": "CoriolisBu", "Coriolis2Bu": "Coriolis2Bu"}


def rotate_domain(dom):
    ni, nj = dom.iec - dom.isc + 1, dom.jec - dom.jsc + 1
    halo, whalo = dom.isc - dom.isd, dom.isc - dom.isdw
    return make_domain(nj, ni, nk=dom.nk, halo=halo, whalo=whalo, cyclic_x=bool(dom.cyclic_y), cyclic_y=bool(dom.cyclic_x),
                       first_direction=(dom.first_direction + 1) % 2)


def rotate_grid(grid):
    assert set(grid) == set(GRID_ROT), sorted(set(grid) ^ set(GRID_ROT))
    return {k: rot(grid[src]) for k, src in GRID_ROT.items()}


def rotate_fields(d, vec=(), pair=(), keep=()):
    """Rotate every array of d.  vec: (A_u, A_v) vector pairs; pair: (A_u, A_v) scalar pairs; keep: keys passed through (1-D
    arrays are always passed through).  A missing / None partner of a pair stays None."""
    out = dict(d)
    done = set(keep)
    for a, b in vec:
        if a in d or b in d:
            out[a] = None if d.get(b) is None else -rot(d[b])
            out[b] = None if d.get(a) is None else rot(d[a])
            done |= {a, b}
    for a, b in pair:
        if a in d or b in d:
            out[a] = None if d.get(b) is None else rot(d[b])
            out[b] = None if d.get(a) is None else rot(d[a])
            done |= {a, b}
    for k, v in d.items():
        if k in done:
            continue
        if isinstance(v, np.ndarray) and v.ndim >= 2:
            out[k] = rot(v)
        elif isinstance(v, list) and v and isinstance(v[0], np.ndarray):
            out[k] = [rot(x) for x in v]
    return out


def unrotate_fields(d, vec=(), pair=(), keep=()):
    """The inverse of rotate_fields (turns = -1): A_u = unrot(A'_v), A_v = -unrot(A'_u)."""
    out = dict(d)
    done = set(keep)
    for a, b in vec:
        if a in d or b in d:
            out[a] = None if d.get(b) is None else unrot(d[b])
            out[b] = None if d.get(a) is None else -unrot(d[a])
            done |= {a, b}
    for a, b in pair:
        if a in d or b in d:
            out[a] = None if d.get(b) is None else unrot(d[b])
            out[b] = None if d.get(a) is None else unrot(d[a])
            done |= {a, b}
    for k, v in d.items():
        if k in done:
            continue
        if isinstance(v, np.ndarray) and v.ndim >= 2:
            out[k] = unrot(v)
        elif isinstance(v, list) and v and isinstance(v[0], np.ndarray):
            out[k] = [unrot(x) for x in v]
    return out


# ---- the whole step (step_MOM_dyn_split_RK2): which keys are vectors / scalar pairs
STEP_VEC = [("u_inst", "v_inst"), ("uh", "vh"), ("uhtr", "vhtr"), ("taux", "tauy"), ("CAu", "CAv"), ("CAu_pred", "CAv_pred"), ("PFu", "PFv"),
            ("diffu", "diffv"), ("u_accel_bt", "v_accel_bt"), ("u_av", "v_av"), ("uhbt", "vhbt"), ("taux_bot", "tauy_bot"), ("ubtav", "vbtav"),
            ("u", "v"), ("u_cor", "v_cor"), ("U_in", "V_in"), ("bc_accel_u", "bc_accel_v"), ("U_Cor", "V_Cor"),
            ("accel_layer_u", "accel_layer_v"), ("uhbtav", "vhbtav"), ("uh0", "vh0"), ("u_uh0", "v_vh0"), ("du_cor", "dv_cor")]
STEP_PAIR = [("visc_rem_u", "visc_rem_v"), ("Kv_bbl_u", "Kv_bbl_v"), ("bbl_thick_u", "bbl_thick_v"), ("Ray_u", "Ray_v"), ("h_u", "h_v"),
             ("frhatu", "frhatv"), ("IDatu", "IDatv"), ("D_u_Cor", "D_v_Cor"), ("OBCmask_u", "OBCmask_v"), ("IdxCu", "IdyCv"),
             ("ua_polarity", "va_polarity"), ("por_face_areaU", "por_face_areaV"),
             # hor_visc_CS (hor_visc_init, MOM_hor_visc.F90:2322-3302)
             ("dx2q", "dy2q"), ("DX_dyBu", "DY_dxBu"), ("dx2h", "dy2h"), ("DX_dyT", "DY_dxT"), ("Idx2dyCu", "Idxdy2v"), ("Idxdy2u", "Idx2dyCv")]
# BT_cont_type (MOM_variables.F90:315-350): east of the rotated grid is south of the original one, north is east
BT_CONT_ROT = {"FA_u_EE": ("FA_v_SS", 1), "FA_u_E0": ("FA_v_S0", 1), "FA_u_W0": ("FA_v_N0", 1), "FA_u_WW": ("FA_v_NN", 1),
               "uBT_EE": ("vBT_SS", -1), "uBT_WW": ("vBT_NN", -1),
               "FA_v_NN": ("FA_u_EE", 1), "FA_v_N0": ("FA_u_E0", 1), "FA_v_S0": ("FA_u_W0", 1), "FA_v_SS": ("FA_u_WW", 1),
               "vBT_NN": ("uBT_EE", 1), "vBT_SS": ("uBT_WW", 1), "h_u": ("h_v", 1), "h_v": ("h_u", 1)}


def rotate_bt_cont(b):
    out = dict(b)
    for k, (src, sgn) in BT_CONT_ROT.items():
        if isinstance(b.get(src), np.ndarray):
            out[k] = rot(b[src]) if sgn > 0 else -rot(b[src])
    return out


def unrotate_bt_cont(b):
    out = dict(b)
    for k, (src, sgn) in BT_CONT_ROT.items():          # b[k] = sgn * rot(orig[src])  =>  orig[src] = sgn * unrot(b[k])
        if isinstance(b.get(k), np.ndarray):
            out[src] = unrot(b[k]) if sgn > 0 else -unrot(b[k])
    return out


def rotate_step(dom, grid, css, cs, a):
    """The inputs of step_MOM_dyn_split_RK2 on the index map turned by a quarter."""
    domr, gridr = rotate_domain(dom), rotate_grid(grid)
    cssr = {k: rotate_fields(v, STEP_VEC, STEP_PAIR) for k, v in css.items()}
    csr = rotate_fields({k: v for k, v in cs.items() if not isinstance(v, dict)}, STEP_VEC, STEP_PAIR)
    csr["BT_cont"] = rotate_bt_cont(cs["BT_cont"])
    csr["barotropic"] = rotate_fields(cs["barotropic"], STEP_VEC, STEP_PAIR)
    return domr, gridr, cssr, csr, rotate_fields(a, STEP_VEC, STEP_PAIR)


def unrotate_step(cs, a):
    csb = unrotate_fields({k: v for k, v in cs.items() if not isinstance(v, dict)}, STEP_VEC, STEP_PAIR)
    csb["BT_cont"] = unrotate_bt_cont(cs["BT_cont"])
    csb["barotropic"] = unrotate_fields(cs["barotropic"], STEP_VEC, STEP_PAIR)
    return csb, unrotate_fields(a, STEP_VEC, STEP_PAIR)
