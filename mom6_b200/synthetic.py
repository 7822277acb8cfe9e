"""Seeded synthetic inputs for the hot path: beta-plane Cartesian grid, reentrant
in x, closed in y, Gaussian seamount, optional land blocks.  Used by tests/ and bench.py only.
"""
import numpy as np

from . import fidx
from .api import make_domain

SEED = 102030405  # echoes config_src/drivers/timing_tests/time_MOM_remapping.F90:53


def rng(seed=SEED):
    return np.random.Generator(np.random.MT19937(seed))


def _sym_u(dom, f):
    """make a u-field periodic-consistent on the shared edge and fill its halos"""
    if dom.cyclic_x:
        f.s(dom.isc - 1, dom.isc - 1, f.jlo, f.jhi)[...] = f.s(dom.iec, dom.iec, f.jlo, f.jhi)
    return fidx.fill_halo(dom, f, "u")


def _sym_v(dom, f):
    if dom.cyclic_y:
        f.s(f.ilo, f.ihi, dom.jsc - 1, dom.jsc - 1)[...] = f.s(f.ilo, f.ihi, dom.jec, dom.jec)
    return fidx.fill_halo(dom, f, "v")


def masks_and_depth(dom, wide=True, land_blocks=0, seed=SEED, dx=1.0e4):
    """mask2dT/Cu/Cv and bathyT on the (wide) memory domain."""
    r = rng(seed + 7)
    ni, nj = dom.iec - dom.isc + 1, dom.jec - dom.jsc + 1
    mT = fidx.new(dom, "h", wide)
    D = fidx.new(dom, "h", wide)
    ii = np.arange(dom.isc, dom.iec + 1)[None, :]
    jj = np.arange(dom.jsc, dom.jec + 1)[:, None]
    x = (ii - dom.isc + 0.5) / ni - 0.5
    y = (jj - dom.jsc + 0.5) / nj - 0.5
    depth = 4000.0 - 2000.0 * np.exp(-((x / 0.15) ** 2 + (y / 0.15) ** 2))
    m = np.ones((nj, ni))
    if not dom.cyclic_y:
        m[0, :] = 0.0
        m[-1, :] = 0.0
    if not dom.cyclic_x:
        m[:, 0] = 0.0
        m[:, -1] = 0.0
    for _ in range(land_blocks):
        bi, bj = int(r.integers(0, ni)), int(r.integers(1, max(2, nj - 1)))
        wi, wj = int(r.integers(1, max(2, ni // 6))), int(r.integers(1, max(2, nj // 6)))
        m[bj:bj + wj, bi:bi + wi] = 0.0
    mT.s(dom.isc, dom.iec, dom.jsc, dom.jec)[...] = m
    D.s(dom.isc, dom.iec, dom.jsc, dom.jec)[...] = depth * m
    fidx.fill_halo(dom, mT, "h")
    fidx.fill_halo(dom, D, "h")
    mU = fidx.new(dom, "u", wide)
    mV = fidx.new(dom, "v", wide)
    # a face is open when both neighbouring cells are ocean
    mU.s(mU.ilo + 1, mU.ihi - 1, mU.jlo, mU.jhi)[...] = (
        mT.s(mT.ilo, mT.ihi - 1, mT.jlo, mT.jhi) * mT.s(mT.ilo + 1, mT.ihi, mT.jlo, mT.jhi))
    mV.s(mV.ilo, mV.ihi, mV.jlo + 1, mV.jhi - 1)[...] = (
        mT.s(mT.ilo, mT.ihi, mT.jlo, mT.jhi - 1) * mT.s(mT.ilo, mT.ihi, mT.jlo + 1, mT.jhi))
    return mT, mU, mV, D


def bt_timeloop_inputs(ni, nj, whalo=10, halo=4, nstep=60, nfilter=8, seed=SEED, use_BT_cont=True,
                       project=False, find_etaav=True, cyclic_x=True, cyclic_y=False, first_direction=0,
                       land_blocks=0, bracket_bug=False, dx=1.0e4, dtbt=None):
    """Inputs of btstep_timeloop (MOM_barotropic.F90:2175) on the wide-halo domain -- the BASELINE.json
    'btstep microbench' shape: eta~0.1U, ubt,vbt~0.05U, gtot=9.8, bt_rem=1-1e-4, f_4=f/4*D/D, ..."""
    dom = make_domain(ni, nj, nk=1, halo=halo, whalo=whalo, cyclic_x=cyclic_x, cyclic_y=cyclic_y,
                      first_direction=first_direction)
    r = rng(seed)
    mT, mU, mV, D = masks_and_depth(dom, True, land_blocks, seed, dx)

    def U(st):
        # random numbers live on the (symmetric) computational domain only, so the inputs do not
        # depend on the halo width; halos are filled afterwards by the halo update.
        f = fidx.new(dom, st, True)
        f.s(dom.isc - 1, dom.iec, dom.jsc - 1, dom.jec)[...] = r.uniform(-1.0, 1.0, size=(nj + 1, ni + 1))
        return f.a
    g = 9.8
    if dtbt is None:
        dtbt = 0.5 * dx / np.sqrt(2.0 * g * 4000.0)  # gravity-wave CFL ~ 0.5

    def hfield(scale, mask=True):
        f = fidx.new(dom, "h", True)
        f.a[...] = scale * U("h") * (mT.a if mask else 1.0)
        return fidx.fill_halo(dom, f, "h")

    def ufield(scale):
        f = fidx.new(dom, "u", True)
        f.a[...] = scale * U("u") * mU.a
        return _sym_u(dom, f)

    def vfield(scale):
        f = fidx.new(dom, "v", True)
        f.a[...] = scale * U("v") * mV.a
        return _sym_v(dom, f)

    a = {}
    a["eta"] = hfield(0.1)
    a["ubt"] = ufield(0.05)
    a["vbt"] = vfield(0.05)
    a["uhbt0"] = ufield(1.0e2)
    a["vhbt0"] = vfield(1.0e2)
    a["eta_src"] = hfield(1.0e-6)
    a["eta_PF"] = hfield(0.05)
    for k in ("gtot_E", "gtot_W", "gtot_N", "gtot_S"):
        f = fidx.new(dom, "h", True)
        f.a[...] = g * (1.0 + 1.0e-3 * U("h")) * mT.a
        a[k] = fidx.fill_halo(dom, f, "h")
    # depths at faces
    Du = fidx.new(dom, "u", True)
    Du.s(Du.ilo + 1, Du.ihi - 1, Du.jlo, Du.jhi)[...] = 0.5 * (
        D.s(D.ilo, D.ihi - 1, D.jlo, D.jhi) + D.s(D.ilo + 1, D.ihi, D.jlo, D.jhi))
    Du.a *= mU.a
    Dv = fidx.new(dom, "v", True)
    Dv.s(Dv.ilo, Dv.ihi, Dv.jlo + 1, Dv.jhi - 1)[...] = 0.5 * (
        D.s(D.ilo, D.ihi, D.jlo, D.jhi - 1) + D.s(D.ilo, D.ihi, D.jlo + 1, D.jhi))
    Dv.a *= mV.a
    # Sadourny-style Coriolis weights ~ f/4 (btstep_find_Cor :2866-2880), here simply f/4 with noise
    f0, beta = 1.0e-4, 2.0e-11
    jv = np.arange(dom.jsdw - 1, dom.jedw + 1)
    f4u = fidx.new(dom, "u", True, nm=4)
    f4v = fidx.new(dom, "v", True, nm=4)
    yu = (np.arange(dom.jsdw, dom.jedw + 1) - dom.jsc) * dx
    yv = (jv - dom.jsc + 0.5) * dx
    for m in range(4):
        fu = fidx.new(dom, "u", True)
        fu.a[...] = 0.25 * (f0 + beta * yu)[:, None] * (1.0 + 0.01 * U("u")) * mU.a
        f4u.a[..., m] = _sym_u(dom, fu).a
        fv = fidx.new(dom, "v", True)
        fv.a[...] = 0.25 * (f0 + beta * yv)[:, None] * (1.0 + 0.01 * U("v")) * mV.a
        f4v.a[..., m] = _sym_v(dom, fv).a
    a["f_4_u"], a["f_4_v"] = f4u, f4v
    for nm, mk, sym, st in (("bt_rem_u", mU, _sym_u, "u"), ("bt_rem_v", mV, _sym_v, "v")):
        f = fidx.new(dom, st, True)
        f.a[...] = (1.0 - 1.0e-4 * (1.0 + 0.5 * U(st))) * mk.a
        a[nm] = sym(dom, f)
    a["BT_force_u"] = ufield(1.0e-6)
    a["BT_force_v"] = vfield(1.0e-6)
    a["Cor_ref_u"] = ufield(1.0e-7)
    a["Cor_ref_v"] = vfield(1.0e-7)
    IareaT = fidx.new(dom, "h", True)
    IareaT.a[...] = mT.a / (dx * dx)
    a["IareaT_OBCmask"] = IareaT
    Idx = fidx.new(dom, "u", True); Idx.a[...] = 1.0 / dx
    Idy = fidx.new(dom, "v", True); Idy.a[...] = 1.0 / dx
    a["IdxCu"], a["IdyCv"] = Idx, Idy
    # transport closure
    FA0u = fidx.new(dom, "u", True); FA0u.a[...] = dx * Du.a * (1.0 + 0.01 * U("u")); _sym_u(dom, FA0u)
    FA0v = fidx.new(dom, "v", True); FA0v.a[...] = dx * Dv.a * (1.0 + 0.01 * U("v")); _sym_v(dom, FA0v)
    a["Datu"], a["Datv"] = FA0u, FA0v
    C1_3 = 1.0 / 3.0

    def btcl(FA0):
        # set_local_BT_cont_types, MOM_barotropic.F90:4956-4977; field order :367-390
        b = fidx.FA(FA0.ilo, FA0.ihi, FA0.jlo, FA0.jhi, nm=10)
        FA_EE = 1.02 * FA0.a; FA_E0 = FA0.a; FA_W0 = FA0.a * 1.0; FA_WW = 1.02 * FA0.a
        uBT_WW = np.where(FA0.a > 0.0, 0.03, 0.0); uBT_EE = -uBT_WW
        uh_EE = uBT_EE * (C1_3 * (2.0 * FA_E0 + FA_EE))
        uh_WW = uBT_WW * (C1_3 * (2.0 * FA_W0 + FA_WW))
        with np.errstate(divide="ignore", invalid="ignore"):
            crvW = np.where(np.abs(uBT_WW) > 0.0, (C1_3 * (FA_WW - FA_W0)) / uBT_WW ** 2, 0.0)
            crvE = np.where(np.abs(uBT_EE) > 0.0, (C1_3 * (FA_EE - FA_E0)) / uBT_EE ** 2, 0.0)
        for m, v in enumerate((FA_EE, FA_E0, FA_W0, FA_WW, uBT_WW, uBT_EE, crvW, crvE, uh_WW, uh_EE)):
            b.a[..., m] = v
        return b

    a["BTCL_u"], a["BTCL_v"] = btcl(FA0u), btcl(FA0v)
    # accumulators
    a["u_accel_bt"] = fidx.new(dom, "u", True)
    a["v_accel_bt"] = fidx.new(dom, "v", True)
    a["eta_sum"] = fidx.new(dom, "h", True)
    a["eta_wtd"] = fidx.new(dom, "h", True)
    for k, st in (("ubtav", "u"), ("vbtav", "v"), ("uhbtav", "u"), ("vhbtav", "v"), ("ubt_wtd", "u"), ("vbt_wtd", "v")):
        a[k] = fidx.new(dom, st, False)
    # filter weights, MOM_barotropic.F90:1727-1781 (answer_date >= 20190101 branch)
    wts = bt_weights(nstep, nfilter, dtbt)
    arrays = {k: (v.a if isinstance(v, fidx.FA) else v) for k, v in a.items()}
    arrays.update(wts)
    arrays.update(dict(dtbt=float(dtbt), dgeo_de=1.0, bebt=0.1, vel_underflow=1.0e-30, nstep=nstep, nfilter=nfilter,
                       use_BT_cont=int(use_BT_cont), find_etaav=int(find_etaav), BT_project_velocity=int(project),
                       use_old_coriolis_bracket_bug=int(bracket_bug), use_wide_halos=1, min_stencil=0))
    return dom, arrays


def bt_weights(nstep, nfilter, dtbt):
    """wt_vel, wt_eta, wt_accel, wt_trans, wt_accel2 -- MOM_barotropic.F90:1738-1781 with the ramp
    filter width implied by nfilter (dt_filt = nfilter*dtbt)."""
    dt_filt = nfilter * dtbt
    n_tot = nstep + nfilter
    wt_vel = np.zeros(n_tot); wt_eta = np.zeros(n_tot)
    wt_trans = np.zeros(n_tot + 1); wt_accel = np.zeros(n_tot + 1); wt_accel2 = np.zeros(n_tot + 1)
    sum_wt_vel = sum_wt_eta = sum_wt_accel = sum_wt_trans = 0.0
    for n in range(1, n_tot + 1):
        if (n == nstep) or (dt_filt - abs(n - nstep) * dtbt >= 0.0):
            wt_vel[n - 1] = 1.0; wt_eta[n - 1] = 1.0
        elif dtbt + dt_filt - abs(n - nstep) * dtbt > 0.0:
            wt_vel[n - 1] = 1.0 + (dt_filt / dtbt) - abs(n - nstep); wt_eta[n - 1] = wt_vel[n - 1]
        else:
            wt_vel[n - 1] = 0.0; wt_eta[n - 1] = 0.0
        sum_wt_vel += wt_vel[n - 1]; sum_wt_eta += wt_eta[n - 1]
    for n in range(n_tot, 0, -1):
        wt_trans[n - 1] = wt_trans[n] + wt_eta[n - 1]
        wt_accel[n - 1] = wt_accel[n] + wt_vel[n - 1]
        sum_wt_accel += wt_accel[n - 1]; sum_wt_trans += wt_trans[n - 1]
    I_vel, I_accel, I_eta, I_trans = 1.0 / sum_wt_vel, 1.0 / sum_wt_accel, 1.0 / sum_wt_eta, 1.0 / sum_wt_trans
    for n in range(n_tot):
        wt_vel[n] *= I_vel
        wt_accel2[n] = wt_accel[n] * I_accel
        wt_trans[n] *= I_trans
        wt_accel[n] *= I_accel
        wt_eta[n] *= I_eta
    return dict(wt_vel=wt_vel, wt_eta=wt_eta, wt_accel=wt_accel, wt_trans=wt_trans, wt_accel2=wt_accel2)


def split_tile(dom_g, arrays_g, npi, npj, pi, pj):
    """Cut one rank's tile (with its halos) out of single-tile (global) inputs: what each PE of a
    LAYOUT=npi,npj run holds after the reference's halo updates (MOM_domains.F90:154-222)."""
    NI, NJ = dom_g.iec - dom_g.isc + 1, dom_g.jec - dom_g.jsc + 1
    ni, nj = NI // npi, NJ // npj
    halo, whalo = dom_g.isc - dom_g.isd, dom_g.isc - dom_g.isdw
    dom = make_domain(ni, nj, nk=dom_g.nk, halo=halo, whalo=whalo, cyclic_x=bool(dom_g.cyclic_x),
                      cyclic_y=bool(dom_g.cyclic_y), first_direction=dom_g.first_direction,
                      npi=npi, npj=npj, pi=pi, pj=pj)
    oi, oj = pi * ni, pj * nj
    out = {}
    for k, v in arrays_g.items():
        if not isinstance(v, np.ndarray) or k.startswith("wt_"):
            out[k] = v
            continue
        # identify stagger / domain from the global shape
        found = False
        for st in ("h", "u", "v", "q"):
            for wide in (True, False):
                ilo, ihi, jlo, jhi = fidx.extent(dom_g, st, wide)
                if v.shape[-2 if v.ndim == 2 or v.shape[-1] not in (4, 10) else -3:][:2] == (jhi - jlo + 1, ihi - ilo + 1) \
                        and _stagger_of(k) == st:
                    tl = fidx.extent(dom, st, wide)
                    j0, i0 = tl[2] + oj - jlo, tl[0] + oi - ilo
                    sl = (slice(j0, j0 + tl[3] - tl[2] + 1), slice(i0, i0 + tl[1] - tl[0] + 1))
                    if v.ndim == 3 and v.shape[-1] not in (4, 10):
                        sl = (slice(None),) + sl
                    out[k] = np.ascontiguousarray(v[sl])
                    found = True
                    break
            if found:
                break
        if not found:
            raise ValueError(f"split_tile: cannot place {k} {v.shape}")
    return dom, out


_U_KEYS = {"ubt", "uhbt0", "Datu", "BTCL_u", "f_4_u", "bt_rem_u", "BT_force_u", "Cor_ref_u", "IdxCu", "u_accel_bt",
           "ubtav", "uhbtav", "ubt_wtd"}
_V_KEYS = {"vbt", "vhbt0", "Datv", "BTCL_v", "f_4_v", "bt_rem_v", "BT_force_v", "Cor_ref_v", "IdyCv", "v_accel_bt",
           "vbtav", "vhbtav", "vbt_wtd"}


def _stagger_of(key):
    if key in _U_KEYS:
        return "u"
    if key in _V_KEYS:
        return "v"
    return "h"


# --------------------------------------------------------------------------------------------
# 3-D stages: grid metrics, vertical grid and a split-RK2-like state

def make_grid(dom, land_blocks=0, seed=SEED, dx0=2.5e4, dy0=2.5e4, f0=1.0e-4, beta=2.0e-11):
    """ocean_grid_type metrics (src/core/MOM_grid.F90:75-175) on G's memory domain for a beta-plane
    with a mildly non-uniform (periodic in i) mesh so every metric term is exercised."""
    ni, nj = dom.iec - dom.isc + 1, dom.jec - dom.jsc + 1
    mT, mU, mV, D = masks_and_depth(dom, False, land_blocks, seed)
    g = {}

    def pts(st):
        ilo, ihi, jlo, jhi = fidx.extent(dom, st)
        su = 0.5 if st in ("u", "q") else 0.0
        sv = 0.5 if st in ("v", "q") else 0.0
        ii = (np.arange(ilo, ihi + 1) - dom.isc + 0.5 + su)[None, :]
        jj = (np.arange(jlo, jhi + 1) - dom.jsc + 0.5 + sv)[:, None]
        return ii, jj

    def dxf(ii, jj):
        return dx0 * (1.0 - 0.3 * ((jj / nj) - 0.5) ** 2) * (1.0 + 0.0 * ii)

    def dyf(ii, jj):
        return dy0 * (1.0 + 0.05 * np.sin(2.0 * np.pi * ii / ni)) * (1.0 + 0.0 * jj)

    for st, sfx in (("h", "T"), ("u", "Cu"), ("v", "Cv"), ("q", "Bu")):
        ii, jj = pts(st)
        dx, dy = dxf(ii, jj), dyf(ii, jj)
        g["dx" + sfx], g["dy" + sfx] = dx, dy
        g["Idx" + sfx], g["Idy" + sfx] = 1.0 / dx, 1.0 / dy
        g["area" + sfx] = dx * dy
        g["Iarea" + sfx] = 1.0 / (dx * dy)
    mQ = fidx.new(dom, "q")
    mQ.s(mQ.ilo + 1, mQ.ihi - 1, mQ.jlo + 1, mQ.jhi - 1)[...] = (
        mT.s(mT.ilo, mT.ihi - 1, mT.jlo, mT.jhi - 1) * mT.s(mT.ilo + 1, mT.ihi, mT.jlo, mT.jhi - 1) *
        mT.s(mT.ilo, mT.ihi - 1, mT.jlo + 1, mT.jhi) * mT.s(mT.ilo + 1, mT.ihi, mT.jlo + 1, mT.jhi))
    fidx.fill_halo(dom, mQ, "q")
    fidx.fill_halo(dom, mU, "u")
    fidx.fill_halo(dom, mV, "v")
    if dom.cyclic_x:
        # the symmetric edge of the staggered masks
        mU.s(dom.isc - 1, dom.isc - 1, mU.jlo, mU.jhi)[...] = mU.s(dom.iec, dom.iec, mU.jlo, mU.jhi)
        mQ.s(dom.isc - 1, dom.isc - 1, mQ.jlo, mQ.jhi)[...] = mQ.s(dom.iec, dom.iec, mQ.jlo, mQ.jhi)
        fidx.fill_halo(dom, mU, "u"); fidx.fill_halo(dom, mQ, "q")
    g["mask2dT"], g["mask2dCu"], g["mask2dCv"], g["mask2dBu"] = mT.a, mU.a, mV.a, mQ.a
    g["dy_Cu"] = g["dyCu"] * mU.a
    g["dx_Cv"] = g["dxCv"] * mV.a
    g["bathyT"] = D.a
    ii, jj = pts("q")
    fq = f0 + beta * (jj * dy0) + 0.0 * ii
    g["CoriolisBu"] = fq
    g["Coriolis2Bu"] = fq * fq
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in g.items()}


def make_vgrid(Angstrom_H=1.0e-10, H_subroundoff=1.0e-30):
    return dict(Angstrom_H=Angstrom_H, H_subroundoff=H_subroundoff, Z_to_H=1.0, H_to_Z=1.0, g_Earth=9.8, Rho0=1035.0,
                H_to_RZ=1035.0, RZ_to_H=1.0 / 1035.0, H_to_m=1.0, m_to_H=1.0, Boussinesq=1)


def continuity_cs(nk, Angstrom_H=1.0e-10, **over):
    """continuity_PPM_init defaults (MOM_continuity_PPM.F90:2693-2747)."""
    cs = dict(upwind_1st=0, monotonic=0, simple_2nd=0, aggress_adjust=0, vol_CFL=0, better_iter=1, use_visc_rem_max=1,
              marginal_faces=1, tol_eta=0.5 * nk * Angstrom_H, tol_vel=3.0e8, CFL_limit_adjust=0.5)
    cs.update(over)
    return cs


def dyn_state(dom, grid, seed=SEED, vel=0.3, thin_layers=True, hnoise=0.2):
    """h, u, v, visc_rem_u/v on G's memory domain: Z*-like layers over the synthetic bathymetry with
    noise, some vanished layers over the seamount, velocities ~ vel*U*mask."""
    r = rng(seed + 101)
    nk = dom.nk
    ni, nj = dom.iec - dom.isc + 1, dom.jec - dom.jsc + 1

    def U3(st):
        f = fidx.new(dom, st, nk=nk)
        f.s(dom.isc - 1, dom.iec, dom.jsc - 1, dom.jec)[...] = r.uniform(-1.0, 1.0, size=(nk, nj + 1, ni + 1))
        return f

    def fill3(f, st):
        if dom.cyclic_x and st in ("u", "q"):
            f.s(dom.isc - 1, dom.isc - 1, f.jlo, f.jhi)[...] = f.s(dom.iec, dom.iec, f.jlo, f.jhi)
        if dom.cyclic_y and st in ("v", "q"):
            f.s(f.ilo, f.ihi, dom.jsc - 1, dom.jsc - 1)[...] = f.s(f.ilo, f.ihi, dom.jec, dom.jec)
        return fidx.fill_halo(dom, f, st)

    w = np.linspace(1.0, 3.0, nk); w /= w.sum()
    h = U3("h")
    D = grid["bathyT"]
    h.a[...] = D[None, :, :] * w[:, None, None] * (1.0 + hnoise * h.a)
    if thin_layers:
        # vanished layers where the water is shallower than the nominal interface depth
        zbot = np.cumsum(4000.0 * w)
        for k in range(nk):
            h.a[k][D < zbot[k] - 4000.0 * w[k] * 0.5] = 1.0e-10
    h.a[...] = np.maximum(h.a, 1.0e-10) * grid["mask2dT"][None] + 1.0e-10 * (1.0 - grid["mask2dT"][None])
    fill3(h, "h")
    u = U3("u"); u.a[...] = vel * u.a * grid["mask2dCu"][None]; fill3(u, "u")
    v = U3("v"); v.a[...] = vel * v.a * grid["mask2dCv"][None]; fill3(v, "v")
    zf = np.linspace(0.0, 1.0, nk)[:, None, None]
    vru = U3("u"); vru.a[...] = (1.0 - 0.6 * np.exp(-(1.0 - zf) / 0.1) * (1.0 + 0.2 * vru.a)) * grid["mask2dCu"][None]; fill3(vru, "u")
    vrv = U3("v"); vrv.a[...] = (1.0 - 0.6 * np.exp(-(1.0 - zf) / 0.1) * (1.0 + 0.2 * vrv.a)) * grid["mask2dCv"][None]; fill3(vrv, "v")
    return dict(h=h.a, u=u.a, v=v.a, visc_rem_u=vru.a, visc_rem_v=vrv.a)


def continuity_inputs(ni, nj, nk, halo=4, seed=SEED, land_blocks=0, with_uhbt=True, with_visc_rem=True, with_BT_cont=True,
                      with_cor=True, first_direction=0, cyclic_x=True, cyclic_y=False, dt=900.0, alias_h=False, cs_over=None, uhbt_noise=0.05):
    """Everything a continuity_PPM call needs (MOM_continuity_PPM.F90:86): returns dom, grid, vgrid, cs, args."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, cyclic_x=cyclic_x, cyclic_y=cyclic_y, first_direction=first_direction)
    grid = make_grid(dom, land_blocks, seed)
    gv = make_vgrid()
    cs = continuity_cs(nk, **(cs_over or {}))
    st = dyn_state(dom, grid, seed)
    r = rng(seed + 5)
    a = dict(u=st["u"], v=st["v"], hin=st["h"], dt=dt)
    a["h"] = a["hin"] if alias_h else st["h"].copy()
    a["uh"] = fidx.new(dom, "u", nk=nk).a
    a["vh"] = fidx.new(dom, "v", nk=nk).a
    if with_visc_rem:
        a["visc_rem_u"], a["visc_rem_v"] = st["visc_rem_u"], st["visc_rem_v"]
    if with_uhbt:
        # a target transport near the layer-summed first-guess transport (uhbt_noise ~ 1: far from it, so that the limits of
        # zonal_flux_adjust on the velocity correction bind)
        hu = 0.5 * (st["h"][:, :, :-1] + st["h"][:, :, 1:])
        uh0 = fidx.new(dom, "u")
        uh0.a[:, 1:-1] = (st["u"][:, :, 1:-1] * hu * grid["dy_Cu"][None, :, 1:-1]).sum(axis=0)
        uh0.a[...] = uh0.a * (1.0 + uhbt_noise * r.uniform(-1, 1, size=uh0.a.shape)) * grid["mask2dCu"]
        a["uhbt"] = _sym_u(dom, uh0).a
        hv = 0.5 * (st["h"][:, :-1, :] + st["h"][:, 1:, :])
        vh0 = fidx.new(dom, "v")
        vh0.a[1:-1, :] = (st["v"][:, 1:-1, :] * hv * grid["dx_Cv"][None, 1:-1, :]).sum(axis=0)
        vh0.a[...] = vh0.a * (1.0 + uhbt_noise * r.uniform(-1, 1, size=vh0.a.shape)) * grid["mask2dCv"]
        a["vhbt"] = _sym_v(dom, vh0).a
        if with_cor:
            a["u_cor"] = fidx.new(dom, "u", nk=nk).a
            a["v_cor"] = fidx.new(dom, "v", nk=nk).a
            a["du_cor"] = fidx.new(dom, "u").a
            a["dv_cor"] = fidx.new(dom, "v").a
    if with_BT_cont:
        b = {}
        for k in ("FA_u_EE", "FA_u_E0", "FA_u_W0", "FA_u_WW", "uBT_WW", "uBT_EE"):
            b[k] = fidx.new(dom, "u").a
        for k in ("FA_v_NN", "FA_v_N0", "FA_v_S0", "FA_v_SS", "vBT_SS", "vBT_NN"):
            b[k] = fidx.new(dom, "v").a
        b["h_u"] = fidx.new(dom, "u", nk=nk).a
        b["h_v"] = fidx.new(dom, "v", nk=nk).a
        a["BT_cont"] = b
    return dom, grid, gv, cs, a


def coriolisadv_cs(**over):
    """CoriolisAdv_init defaults (MOM_CoriolisAdv.F90:1077-1178)."""
    cs = dict(Coriolis_Scheme=1, KE_Scheme=10, PV_Adv_Scheme=21, no_slip=0, bound_Coriolis=0, Coriolis_En_Dis=0,
              F_eff_max_blend=4.0, wt_lin_blend=0.125)
    cs.update(over)
    return cs


def transports(dom, grid, st):
    """uh, vh consistent with (u, v, h): upwind face thickness times face length (what continuity returns)."""
    h, u, v = st["h"], st["u"], st["v"]
    nk = dom.nk
    uh = fidx.new(dom, "u", nk=nk)
    hu = np.where(u[:, :, 1:-1] > 0, h[:, :, :-1], h[:, :, 1:])
    uh.a[:, :, 1:-1] = u[:, :, 1:-1] * hu * grid["dy_Cu"][None, :, 1:-1]
    vh = fidx.new(dom, "v", nk=nk)
    hv = np.where(v[:, 1:-1, :] > 0, h[:, :-1, :], h[:, 1:, :])
    vh.a[:, 1:-1, :] = v[:, 1:-1, :] * hv * grid["dx_Cv"][None, 1:-1, :]
    return uh.a, vh.a


def coradcalc_inputs(ni, nj, nk, halo=4, seed=SEED, land_blocks=0, cyclic_x=True, cyclic_y=False, cs_over=None,
                     diags=False, por=False):
    """Everything a CorAdCalc call needs (MOM_CoriolisAdv.F90:125): returns dom, grid, vgrid, cs, args."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, cyclic_x=cyclic_x, cyclic_y=cyclic_y)
    grid = make_grid(dom, land_blocks, seed)
    gv = make_vgrid()
    cs = coriolisadv_cs(**(cs_over or {}))
    st = dyn_state(dom, grid, seed)
    uh, vh = transports(dom, grid, st)
    a = dict(u=st["u"], v=st["v"], h=st["h"], uh=uh, vh=vh)
    a["CAu"] = fidx.new(dom, "u", nk=nk).a
    a["CAv"] = fidx.new(dom, "v", nk=nk).a
    if diags:
        a["RV"] = fidx.new(dom, "q", nk=nk).a
        a["PV"] = fidx.new(dom, "q", nk=nk).a
        a["gradKEu"] = fidx.new(dom, "u", nk=nk).a
        a["gradKEv"] = fidx.new(dom, "v", nk=nk).a
    if por:
        r = rng(seed + 77)
        a["por_face_areaU"] = 1.0 - 0.3 * r.uniform(0, 1, size=a["u"].shape)
        a["por_face_areaV"] = 1.0 - 0.3 * r.uniform(0, 1, size=a["v"].shape)
    return dom, grid, gv, cs, a


def hor_visc_cs(dom, grid, dt=900.0, Laplacian=False, biharmonic=True, Kh=0.0, Kh_vel_scale=0.0, Ah=0.0, Ah_vel_scale=0.0,
                Ah_time_scale=0.0, Smagorinsky_Kh=False, Smag_Lap_const=0.15, Smagorinsky_Ah=True, Smag_bi_const=0.06,
                bound_Kh=True, better_bound_Kh=True, bound_Ah=True, better_bound_Ah=True, bound_Coriolis=False,
                bound_Cor_vel=6.0, bound_coef=0.8, no_slip=False, use_land_mask=False, add_LES_viscosity=False,
                Kh_bg_min=0.0, Re_Ah=0.0, backscatter_underbound=False, use_cont_thick=False):
    """hor_visc_init's static arrays (MOM_hor_visc.F90:2834-3120) in numpy, the way the Fortran host computes them
    once; defaults follow hor_visc_init's get_param defaults with BIHARMONIC + SMAGORINSKY_AH (benchmark-like)."""
    G = grid
    if not Laplacian:
        Smagorinsky_Kh = False; bound_Kh = False; better_bound_Kh = False
    if not biharmonic:
        Smagorinsky_Ah = False; bound_Ah = False; better_bound_Ah = False
    if not Smagorinsky_Ah:
        bound_Coriolis = False
    cs = dict(Laplacian=int(Laplacian), biharmonic=int(biharmonic), no_slip=int(no_slip), bound_Kh=int(bound_Kh),
              better_bound_Kh=int(better_bound_Kh), bound_Ah=int(bound_Ah), better_bound_Ah=int(better_bound_Ah),
              backscatter_underbound=int(backscatter_underbound), Smagorinsky_Kh=int(Smagorinsky_Kh),
              Smagorinsky_Ah=int(Smagorinsky_Ah), bound_Coriolis=int(bound_Coriolis), use_land_mask=int(use_land_mask),
              add_LES_viscosity=int(add_LES_viscosity), use_cont_thick=int(use_cont_thick), use_cont_thick_bug=0,
              unsupported=0, Kh_bg_min=float(Kh_bg_min), Re_Ah=float(Re_Ah))
    Idt = 1.0 / dt
    # faces / corners around an h point (array index shifts, see fidx): E, W faces; N, S faces; corners
    E = lambda U: U[:, 1:]; W = lambda U: U[:, :-1]        # noqa: E731
    N = lambda V: V[1:, :]; S = lambda V: V[:-1, :]        # noqa: E731
    dx2q = G["dxBu"] * G["dxBu"]; dy2q = G["dyBu"] * G["dyBu"]
    DX_dyBu = G["dxBu"] * G["IdyBu"]; DY_dxBu = G["dyBu"] * G["IdxBu"]
    dx2h = G["dxT"] * G["dxT"]; dy2h = G["dyT"] * G["dyT"]
    DX_dyT = G["dxT"] * G["IdyT"]; DY_dxT = G["dyT"] * G["IdxT"]
    cs.update(dx2q=dx2q, dy2q=dy2q, DX_dyBu=DX_dyBu, DY_dxBu=DY_dxBu, dx2h=dx2h, dy2h=dy2h, DX_dyT=DX_dyT, DY_dxT=DY_dxT)

    def reduce(red, d_, dC):
        with np.errstate(divide="ignore", invalid="ignore"):
            ok = (d_ > 0.0) & (d_ < dC) & (d_ < dC * red)
            return np.where(ok, d_ / dC, red)
    red = np.ones_like(dx2h)
    for d_, dC in ((E(G["dy_Cu"]), E(G["dyCu"])), (W(G["dy_Cu"]), W(G["dyCu"])), (N(G["dx_Cv"]), N(G["dxCv"])),
                   (S(G["dx_Cv"]), S(G["dxCv"]))):
        red = reduce(red, d_, dC)
    cs["reduction_xx"] = red
    redq = np.ones_like(dx2q)
    inner = redq[1:-1, 1:-1]
    # q(I,J): faces u(I,j), u(I,j+1), v(i,J), v(i+1,J)
    dyu, dyCu, dxv, dxCv = G["dy_Cu"], G["dyCu"], G["dx_Cv"], G["dxCv"]
    for d_, dC in ((dyu[:-1, 1:-1], dyCu[:-1, 1:-1]), (dyu[1:, 1:-1], dyCu[1:, 1:-1]),
                   (dxv[1:-1, :-1], dxCv[1:-1, :-1]), (dxv[1:-1, 1:], dxCv[1:-1, 1:])):
        inner = reduce(inner, d_, dC)
    redq[1:-1, 1:-1] = inner
    cs["reduction_xy"] = redq

    grid_sp_h2 = (2.0 * dx2h * dy2h) / (dx2h + dy2h)
    grid_sp_h3 = grid_sp_h2 * np.sqrt(grid_sp_h2)
    grid_sp_q2 = (2.0 * dx2q * dy2q) / (dx2q + dy2q)
    grid_sp_q3 = grid_sp_q2 * np.sqrt(grid_sp_q2)
    if Laplacian:
        Kh_Limit = 0.3 / (dt * 4.0)
        if Smagorinsky_Kh:
            cs["Laplac2_const_xx"] = Smag_Lap_const * grid_sp_h2
            cs["Laplac2_const_xy"] = Smag_Lap_const * grid_sp_q2
        Kh_bg_xx = np.maximum(Kh, Kh_vel_scale * np.sqrt(grid_sp_h2))
        Kh_bg_xy = np.maximum(Kh, Kh_vel_scale * np.sqrt(grid_sp_q2))
        if bound_Kh and not better_bound_Kh:
            cs["Kh_Max_xx"] = Kh_Limit * grid_sp_h2
            cs["Kh_Max_xy"] = Kh_Limit * grid_sp_q2
            Kh_bg_xx = np.minimum(Kh_bg_xx, cs["Kh_Max_xx"]); Kh_bg_xy = np.minimum(Kh_bg_xy, cs["Kh_Max_xy"])
        cs["Kh_bg_xx"], cs["Kh_bg_xy"] = Kh_bg_xx, Kh_bg_xy
    IdxCu, IdyCu, IdxCv, IdyCv = G["IdxCu"], G["IdyCu"], G["IdxCv"], G["IdyCv"]
    IareaCu, IareaCv = G["IareaCu"], G["IareaCv"]
    if biharmonic:
        cs["Idx2dyCu"] = (IdxCu * IdxCu) * IdyCu; cs["Idxdy2u"] = IdxCu * (IdyCu * IdyCu)
        cs["Idx2dyCv"] = (IdxCv * IdxCv) * IdyCv; cs["Idxdy2v"] = IdxCv * (IdyCv * IdyCv)
        Ah_Limit = 0.3 / (dt * 64.0)
        if Smagorinsky_Ah:
            cs["Biharm_const_xx"] = Smag_bi_const * (grid_sp_h2 * grid_sp_h2)
            cs["Biharm_const_xy"] = Smag_bi_const * (grid_sp_q2 * grid_sp_q2)
            if bound_Coriolis:
                BoundCorConst = 1.0 / (5.0 * (bound_Cor_vel * bound_Cor_vel))
                f = np.abs(G["CoriolisBu"])
                fmax = np.maximum(np.maximum(f[:-1, :-1], f[:-1, 1:]), np.maximum(f[1:, :-1], f[1:, 1:]))
                cs["Biharm_const2_xx"] = (grid_sp_h2 * grid_sp_h2 * grid_sp_h2) * (fmax * BoundCorConst)
                cs["Biharm_const2_xy"] = (grid_sp_q2 * grid_sp_q2 * grid_sp_q2) * (f * BoundCorConst)
        Ah_bg_xx = np.maximum(Ah, Ah_vel_scale * grid_sp_h2 * np.sqrt(grid_sp_h2))
        Ah_bg_xy = np.maximum(Ah, Ah_vel_scale * grid_sp_q2 * np.sqrt(grid_sp_q2))
        if Re_Ah > 0.0:
            cs["Re_Ah_const_xx"] = grid_sp_h3 / Re_Ah; cs["Re_Ah_const_xy"] = grid_sp_q3 / Re_Ah
        if Ah_time_scale > 0.0:
            Ah_bg_xx = np.maximum(Ah_bg_xx, (grid_sp_h2 * grid_sp_h2) / Ah_time_scale)
            Ah_bg_xy = np.maximum(Ah_bg_xy, (grid_sp_q2 * grid_sp_q2) / Ah_time_scale)
        if bound_Ah and not better_bound_Ah:
            cs["Ah_Max_xx"] = Ah_Limit * (grid_sp_h2 * grid_sp_h2); cs["Ah_Max_xy"] = Ah_Limit * (grid_sp_q2 * grid_sp_q2)
            Ah_bg_xx = np.minimum(Ah_bg_xx, cs["Ah_Max_xx"]); Ah_bg_xy = np.minimum(Ah_bg_xy, cs["Ah_Max_xy"])
        cs["Ah_bg_xx"], cs["Ah_bg_xy"] = Ah_bg_xx, Ah_bg_xy
    if Laplacian and better_bound_Kh:   # :3028-3048
        den = np.maximum(dy2h * DY_dxT * (E(IdyCu) + W(IdyCu)) * np.maximum(E(IdyCu) * E(IareaCu), W(IdyCu) * W(IareaCu)),
                         dx2h * DX_dyT * (N(IdxCv) + S(IdxCv)) * np.maximum(N(IdxCv) * N(IareaCv), S(IdxCv) * S(IareaCv)))
        cs["Kh_Max_xx"] = np.where(den > 0.0, bound_coef * 0.25 * Idt / np.where(den > 0, den, 1.0), 0.0)
        Kq = np.zeros_like(dx2q)
        q = (slice(1, -1), slice(1, -1))
        # u(I,j) -> [:-1, 1:-1]; u(I,j+1) -> [1:, 1:-1]; v(i,J) -> [1:-1, :-1]; v(i+1,J) -> [1:-1, 1:]
        uj, uj1 = (slice(None, -1), slice(1, -1)), (slice(1, None), slice(1, -1))
        vi, vi1 = (slice(1, -1), slice(None, -1)), (slice(1, -1), slice(1, None))
        den = np.maximum(dx2q[q] * DX_dyBu[q] * (IdxCu[uj1] + IdxCu[uj]) * np.maximum(IdxCu[uj] * IareaCu[uj], IdxCu[uj1] * IareaCu[uj1]),
                         dy2q[q] * DY_dxBu[q] * (IdyCv[vi1] + IdyCv[vi]) * np.maximum(IdyCv[vi] * IareaCv[vi], IdyCv[vi1] * IareaCv[vi1]))
        Kq[q] = np.where(den > 0.0, bound_coef * 0.25 * Idt / np.where(den > 0, den, 1.0), 0.0)
        cs["Kh_Max_xy"] = Kq
    if biharmonic and better_bound_Ah:  # :3056-3113
        Idxdy2u, Idx2dyCu, Idxdy2v, Idx2dyCv = cs["Idxdy2u"], cs["Idx2dyCu"], cs["Idxdy2v"], cs["Idx2dyCv"]
        nj, ni = dx2h.shape
        u0u = np.zeros_like(IdxCu); u0v = np.zeros_like(IdxCu); v0u = np.zeros_like(IdxCv); v0v = np.zeros_like(IdxCv)
        # u point (I,j), interior: h(i,j)=[r,c-1]... use explicit index arrays on the interior (c=1..ni-1, r=1..nj-2)
        r = slice(1, nj - 1)
        c = slice(1, ni)            # u columns with both neighbours
        hW = (r, slice(0, ni - 1)); hE = (r, slice(1, ni))         # h(i,j), h(i+1,j) for u col c
        uC = (r, c); uE = (r, slice(2, ni + 1)); uW = (r, slice(0, ni - 1))
        qN = (slice(2, nj), c); qS = (slice(1, nj - 1), c)          # q(I,J), q(I,J-1)
        uN = (slice(2, nj), c); uS = (slice(0, nj - 2), c)          # u(I,j+1), u(I,j-1)
        # v at (i,J),(i+1,J),(i,J-1),(i+1,J-1) around u(I,j)
        vNW = (slice(2, nj), slice(0, ni - 1)); vNE = (slice(2, nj), slice(1, ni))
        vSW = (slice(1, nj - 1), slice(0, ni - 1)); vSE = (slice(1, nj - 1), slice(1, ni))
        u0u[uC] = ((Idxdy2u[uC] * ((dy2h[hE] * DY_dxT[hE] * (IdyCu[uE] + IdyCu[uC])) + (dy2h[hW] * DY_dxT[hW] * (IdyCu[uC] + IdyCu[uW])))) +
                   (Idx2dyCu[uC] * ((dx2q[qN] * DX_dyBu[qN] * (IdxCu[uN] + IdxCu[uC])) + (dx2q[qS] * DX_dyBu[qS] * (IdxCu[uC] + IdxCu[uS])))))
        u0v[uC] = ((Idxdy2u[uC] * ((dy2h[hE] * DX_dyT[hE] * (IdxCv[vNE] + IdxCv[vSE])) + (dy2h[hW] * DX_dyT[hW] * (IdxCv[vNW] + IdxCv[vSW])))) +
                   (Idx2dyCu[uC] * ((dx2q[qN] * DY_dxBu[qN] * (IdyCv[vNE] + IdyCv[vNW])) + (dx2q[qS] * DY_dxBu[qS] * (IdyCv[vSE] + IdyCv[vSW])))))
        # v point (i,J): rows 1..nj-1, cols 1..ni-2
        rr = slice(1, nj); cc = slice(1, ni - 1)
        vC = (rr, cc); vNn = (slice(2, nj + 1), cc); vSs = (slice(0, nj - 1), cc); vEe = (rr, slice(2, ni)); vWw = (rr, slice(0, ni - 2))
        qE = (rr, slice(2, ni)); qW = (rr, slice(1, ni - 1))       # q(I,J), q(I-1,J)
        hN = (slice(1, nj), cc); hS = (slice(0, nj - 1), cc)       # h(i,j+1), h(i,j)
        # u at (I,j+1),(I,j),(I-1,j+1),(I-1,j) around v(i,J)
        uNE = (slice(1, nj), slice(2, ni)); uSE = (slice(0, nj - 1), slice(2, ni))
        uNW = (slice(1, nj), slice(1, ni - 1)); uSW = (slice(0, nj - 1), slice(1, ni - 1))
        v0u[vC] = ((Idxdy2v[vC] * ((dy2q[qE] * DX_dyBu[qE] * (IdxCu[uNE] + IdxCu[uSE])) + (dy2q[qW] * DX_dyBu[qW] * (IdxCu[uNW] + IdxCu[uSW])))) +
                   (Idx2dyCv[vC] * ((dx2h[hN] * DY_dxT[hN] * (IdyCu[uNE] + IdyCu[uNW])) + (dx2h[hS] * DY_dxT[hS] * (IdyCu[uSE] + IdyCu[uSW])))))
        v0v[vC] = ((Idxdy2v[vC] * ((dy2q[qE] * DY_dxBu[qE] * (IdyCv[vEe] + IdyCv[vC])) + (dy2q[qW] * DY_dxBu[qW] * (IdyCv[vC] + IdyCv[vWw])))) +
                   (Idx2dyCv[vC] * ((dx2h[hN] * DX_dyT[hN] * (IdxCv[vNn] + IdxCv[vC])) + (dx2h[hS] * DX_dyT[hS] * (IdxCv[vC] + IdxCv[vSs])))))
        den = np.maximum(
            dy2h * ((DY_dxT * ((E(IdyCu) * E(u0u)) + (W(IdyCu) * W(u0u)))) + (DX_dyT * ((N(IdxCv) * N(v0u)) + (S(IdxCv) * S(v0u))))) *
            np.maximum(E(IdyCu) * E(IareaCu), W(IdyCu) * W(IareaCu)),
            dx2h * ((DY_dxT * ((E(IdyCu) * E(u0v)) + (W(IdyCu) * W(u0v)))) + (DX_dyT * ((N(IdxCv) * N(v0v)) + (S(IdxCv) * S(v0v))))) *
            np.maximum(N(IdxCv) * N(IareaCv), S(IdxCv) * S(IareaCv)))
        cs["Ah_Max_xx"] = np.where(den > 0.0, bound_coef * 0.5 * Idt / np.where(den > 0, den, 1.0), 0.0)
        Aq = np.zeros_like(dx2q)
        q = (slice(1, -1), slice(1, -1))
        uj, uj1 = (slice(None, -1), slice(1, -1)), (slice(1, None), slice(1, -1))
        vi, vi1 = (slice(1, -1), slice(None, -1)), (slice(1, -1), slice(1, None))
        den = np.maximum(
            dx2q[q] * ((DX_dyBu[q] * ((u0u[uj1] * IdxCu[uj1]) + (u0u[uj] * IdxCu[uj]))) + (DY_dxBu[q] * ((v0u[vi1] * IdyCv[vi1]) + (v0u[vi] * IdyCv[vi])))) *
            np.maximum(IdxCu[uj] * IareaCu[uj], IdxCu[uj1] * IareaCu[uj1]),
            dy2q[q] * ((DX_dyBu[q] * ((u0v[uj1] * IdxCu[uj1]) + (u0v[uj] * IdxCu[uj]))) + (DY_dxBu[q] * ((v0v[vi1] * IdyCv[vi1]) + (v0v[vi] * IdyCv[vi])))) *
            np.maximum(IdyCv[vi] * IareaCv[vi], IdyCv[vi1] * IareaCv[vi1]))
        Aq[q] = np.where(den > 0.0, bound_coef * 0.5 * Idt / np.where(den > 0, den, 1.0), 0.0)
        cs["Ah_Max_xy"] = Aq
    return {k: (np.ascontiguousarray(v, dtype=np.float64) if isinstance(v, np.ndarray) else v) for k, v in cs.items()}


def hor_visc_inputs(ni, nj, nk, halo=4, seed=SEED, land_blocks=0, cyclic_x=True, cyclic_y=False, dt=900.0, cont_thick=False,
                    vel=0.3, **cs_kw):
    """Everything a horizontal_viscosity call needs (MOM_hor_visc.F90:266): returns dom, grid, vgrid, cs, args."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, cyclic_x=cyclic_x, cyclic_y=cyclic_y)
    grid = make_grid(dom, land_blocks, seed)
    gv = make_vgrid()
    cs = hor_visc_cs(dom, grid, dt=dt, use_cont_thick=cont_thick, **cs_kw)
    st = dyn_state(dom, grid, seed, vel=vel)
    a = dict(u=st["u"], v=st["v"], h=st["h"], dt=dt)
    a["diffu"] = fidx.new(dom, "u", nk=nk).a
    a["diffv"] = fidx.new(dom, "v", nk=nk).a
    if cont_thick:
        h = st["h"]
        hu = fidx.new(dom, "u", nk=nk); hu.a[:, :, 1:-1] = np.minimum(h[:, :, :-1], h[:, :, 1:]); hu.a[:, :, 0] = h[:, :, 0]; hu.a[:, :, -1] = h[:, :, -1]
        hv = fidx.new(dom, "v", nk=nk); hv.a[:, 1:-1, :] = np.minimum(h[:, :-1, :], h[:, 1:, :]); hv.a[:, 0, :] = h[:, 0, :]; hv.a[:, -1, :] = h[:, -1, :]
        a["hu_cont"], a["hv_cont"] = hu.a, hv.a
    return dom, grid, gv, cs, a


def btstep_inputs(ni, nj, nk, halo=4, whalo=6, seed=SEED, land_blocks=0, cyclic_x=True, cyclic_y=False, dt=900.0,
                  first_direction=0, with_uh0=True, with_etaav=True, with_bot=False, vel=0.2, hnoise=0.2, **cs_over):
    """Everything a btstep call needs (MOM_barotropic.F90:455): returns dom, grid, vgrid, cs, args.
    The barotropic_CS members (wide-halo copies of the metrics, linearised Coriolis thicknesses, frhatu/v ...) are
    built the way barotropic_init (:5301) / btcalc build them."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, whalo=whalo, cyclic_x=cyclic_x, cyclic_y=cyclic_y, first_direction=first_direction)
    domw = make_domain(ni, nj, nk=nk, halo=whalo, whalo=whalo, cyclic_x=cyclic_x, cyclic_y=cyclic_y)
    grid = make_grid(dom, land_blocks, seed)
    gw = make_grid(domw, land_blocks, seed)   # the same metrics on the wide-halo memory domain
    gv = make_vgrid()
    st = dyn_state(dom, grid, seed, vel=vel, thin_layers=False, hnoise=hnoise)
    r = rng(seed + 303)
    h = st["h"]
    D = grid["bathyT"]
    # btcalc-like weights (arithmetic mean thickness fractions)
    hu = np.zeros_like(st["u"]); hu[:, :, 1:-1] = 0.5 * (h[:, :, :-1] + h[:, :, 1:])
    hv = np.zeros_like(st["v"]); hv[:, 1:-1, :] = 0.5 * (h[:, :-1, :] + h[:, 1:, :])
    frhatu = hu * (grid["mask2dCu"] / (hu.sum(axis=0) + 1e-30))[None]
    frhatv = hv * (grid["mask2dCv"] / (hv.sum(axis=0) + 1e-30))[None]
    Du = np.zeros_like(grid["mask2dCu"]); Du[:, 1:-1] = 0.5 * (D[:, :-1] + D[:, 1:])
    Dv = np.zeros_like(grid["mask2dCv"]); Dv[1:-1, :] = 0.5 * (D[:-1, :] + D[1:, :])
    with np.errstate(divide="ignore"):
        IDatu = np.where(Du * grid["mask2dCu"] > 0, 1.0 / np.where(Du > 0, Du, 1.0), 0.0)
        IDatv = np.where(Dv * grid["mask2dCv"] > 0, 1.0 / np.where(Dv > 0, Dv, 1.0), 0.0)
    Dw = gw["bathyT"]
    Duw = np.zeros_like(gw["mask2dCu"]); Duw[:, 1:-1] = 0.5 * (Dw[:, :-1] + Dw[:, 1:])
    Dvw = np.zeros_like(gw["mask2dCv"]); Dvw[1:-1, :] = 0.5 * (Dw[:-1, :] + Dw[1:, :])
    aw = gw["areaT"]
    qD = np.zeros_like(gw["CoriolisBu"])
    den = ((aw[:-1, :-1] * Dw[:-1, :-1] + aw[1:, 1:] * Dw[1:, 1:]) + (aw[:-1, 1:] * Dw[:-1, 1:] + aw[1:, :-1] * Dw[1:, :-1]))
    qD[1:-1, 1:-1] = 0.25 * gw["CoriolisBu"][1:-1, 1:-1] * ((aw[:-1, :-1] + aw[1:, 1:]) + (aw[:-1, 1:] + aw[1:, :-1])) / np.maximum(den, 1e-30)
    cs = dict(Sadourny=1, BT_project_velocity=0, strong_drag=0, bound_BT_corr=0, BT_cont_bounds=1, wt_uv_bug=0, visc_rem_u_uh0=0,
              adjust_BT_cont=0, use_wide_halos=1, min_stencil=0, use_old_coriolis_bracket_bug=0, unsupported=0,
              dtbt=0.9 * 0.5 * 2.5e4 / np.sqrt(9.8 * 4000.0 * 2), bebt=0.1, vel_underflow=1e-30, maxCFL_BT_cont=0.25, G_extra=0.0,
              dt_bt_filter=-0.25)
    cs.update(cs_over)
    cs.update(IareaT=gw["IareaT"] * gw["mask2dT"], IareaT_OBCmask=gw["IareaT"] * gw["mask2dT"], bathyT=Dw, IdxCu=gw["IdxCu"],
              IdyCv=gw["IdyCv"], q_D=qD, D_u_Cor=Duw * gw["mask2dCu"], D_v_Cor=Dvw * gw["mask2dCv"],
              ua_polarity=np.ones_like(Dw), va_polarity=np.ones_like(Dw), OBCmask_u=np.ones_like(Duw), OBCmask_v=np.ones_like(Dvw),
              frhatu=frhatu, frhatv=frhatv, eta_cor=1e-3 * r.uniform(-1, 1, size=D.shape) * grid["mask2dT"],
              eta_cor_bound=np.full_like(D, 1e-7), IDatu=IDatu, IDatv=IDatv,
              ubtav=fidx.new(dom, "u").a, vbtav=fidx.new(dom, "v").a)
    cs = {k: (np.ascontiguousarray(v, dtype=np.float64) if isinstance(v, np.ndarray) else v) for k, v in cs.items()}

    def U3(stg, scale, mask):
        f = fidx.new(dom, stg, nk=nk)
        f.s(dom.isc - 1, dom.iec, dom.jsc - 1, dom.jec)[...] = scale * r.uniform(-1.0, 1.0, size=(nk, nj + 1, ni + 1))
        f.a *= mask[None]
        if dom.cyclic_x and stg == "u":
            f.s(dom.isc - 1, dom.isc - 1, f.jlo, f.jhi)[...] = f.s(dom.iec, dom.iec, f.jlo, f.jhi)
        if dom.cyclic_y and stg == "v":
            f.s(f.ilo, f.ihi, dom.jsc - 1, dom.jsc - 1)[...] = f.s(f.ilo, f.ihi, dom.jec, dom.jec)
        return fidx.fill_halo(dom, f, stg).a

    def U2(stg, scale, mask):
        f = fidx.new(dom, stg)
        f.s(dom.isc - 1, dom.iec, dom.jsc - 1, dom.jec)[...] = scale * r.uniform(-1.0, 1.0, size=(nj + 1, ni + 1))
        f.a *= mask
        if dom.cyclic_x and stg == "u":
            f.s(dom.isc - 1, dom.isc - 1, f.jlo, f.jhi)[...] = f.s(dom.iec, dom.iec, f.jlo, f.jhi)
        if dom.cyclic_y and stg == "v":
            f.s(f.ilo, f.ihi, dom.jsc - 1, dom.jsc - 1)[...] = f.s(f.ilo, f.ihi, dom.jec, dom.jec)
        return fidx.fill_halo(dom, f, stg).a
    mU, mV, mT = grid["mask2dCu"], grid["mask2dCv"], grid["mask2dT"]
    eta_in = U2("h", 0.1, mT)
    pb = fidx.new(dom, "h", nk=nk)
    pb.a[...] = (9.8 * (1.0 + 1e-3 * np.arange(nk) / nk))[:, None, None] * (1.0 + 1e-3 * U3("h", 1.0, np.ones_like(mT)))
    a = dict(U_in=st["u"], V_in=st["v"], eta_in=eta_in, dt=dt, bc_accel_u=U3("u", 1e-5, mU), bc_accel_v=U3("v", 1e-5, mV),
             taux=U2("u", 0.1, mU), tauy=U2("v", 0.1, mV), pbce=pb.a, eta_PF_in=eta_in + U2("h", 0.01, mT),
             U_Cor=st["u"] + U3("u", 0.01, mU), V_Cor=st["v"] + U3("v", 0.01, mV),
             accel_layer_u=fidx.new(dom, "u", nk=nk).a, accel_layer_v=fidx.new(dom, "v", nk=nk).a,
             eta_out=fidx.new(dom, "h").a, uhbtav=fidx.new(dom, "u").a, vhbtav=fidx.new(dom, "v").a,
             visc_rem_u=st["visc_rem_u"], visc_rem_v=st["visc_rem_v"])
    # a BT_cont like the one continuity leaves behind (face areas ~ dy*D with a 2% flare, |uBT| ~ 0.03)
    FAu = grid["dy_Cu"] * Du * (1.0 + 0.01 * U2("u", 1.0, np.ones_like(mU)))
    FAv = grid["dx_Cv"] * Dv * (1.0 + 0.01 * U2("v", 1.0, np.ones_like(mV)))
    a["BT_cont"] = dict(FA_u_EE=1.02 * FAu, FA_u_E0=FAu.copy(), FA_u_W0=FAu.copy(), FA_u_WW=1.03 * FAu,
                        uBT_WW=np.where(FAu > 0, 0.03, 0.0), uBT_EE=np.where(FAu > 0, -0.025, 0.0),
                        FA_v_NN=1.02 * FAv, FA_v_N0=FAv.copy(), FA_v_S0=FAv.copy(), FA_v_SS=1.03 * FAv,
                        vBT_SS=np.where(FAv > 0, 0.03, 0.0), vBT_NN=np.where(FAv > 0, -0.025, 0.0), h_u=None, h_v=None)
    if with_uh0:
        a["u_uh0"] = st["u"] + U3("u", 0.02, mU); a["v_vh0"] = st["v"] + U3("v", 0.02, mV)
        a["uh0"] = a["u_uh0"] * hu * grid["dy_Cu"][None]; a["vh0"] = a["v_vh0"] * hv * grid["dx_Cv"][None]
    if with_etaav:
        a["etaav"] = fidx.new(dom, "h").a
    if with_bot:
        a["taux_bot"] = U2("u", 0.01, mU); a["tauy_bot"] = U2("v", 0.01, mV)
    a = {k: (np.ascontiguousarray(v, dtype=np.float64) if isinstance(v, np.ndarray) else v) for k, v in a.items()}
    a["BT_cont"] = {k: (np.ascontiguousarray(v, dtype=np.float64) if isinstance(v, np.ndarray) else v) for k, v in a["BT_cont"].items()}
    a["_h"] = st["h"]   # not an argument of btstep (ignored by the marshalling); lets step_inputs share the state
    return dom, grid, gv, cs, a


def step_inputs(ni, nj, nk, halo=4, whalo=10, seed=SEED, land_blocks=0, dt=900.0, **bt_over):
    """Inputs of every implemented stage of one split-RK2 baroclinic step (MOM_dynamics_split_RK2.F90:294-1205), sharing
    one grid and one model state.  Returns dom, grid, gv and a dict of per-stage (cs, args)."""
    dom, grid, gv, cs_bt, a_bt = btstep_inputs(ni, nj, nk, halo=halo, whalo=whalo, seed=seed, land_blocks=land_blocks, dt=dt, **bt_over)
    h = a_bt["_h"]
    u, v, vru, vrv = a_bt["U_in"], a_bt["V_in"], a_bt["visc_rem_u"], a_bt["visc_rem_v"]
    new3 = lambda st: fidx.new(dom, st, nk=nk).a          # noqa: E731
    new2 = lambda st: fidx.new(dom, st).a                 # noqa: E731
    uh, vh = transports(dom, grid, dict(h=h, u=u, v=v))
    b = {k: (x.copy() if isinstance(x, np.ndarray) else x) for k, x in a_bt["BT_cont"].items()}
    b["h_u"], b["h_v"] = new3("u"), new3("v")
    cont = dict(u=u, v=v, hin=h, h=h.copy(), uh=new3("u"), vh=new3("v"), dt=dt, visc_rem_u=vru, visc_rem_v=vrv,
                uhbt=np.ascontiguousarray(uh.sum(axis=0)), vhbt=np.ascontiguousarray(vh.sum(axis=0)),
                u_cor=new3("u"), v_cor=new3("v"), BT_cont=b)
    corad = dict(u=u, v=v, h=h, uh=uh, vh=vh, CAu=new3("u"), CAv=new3("v"))
    hv = dict(u=u, v=v, h=h, diffu=new3("u"), diffv=new3("v"), dt=dt)
    btc = dict(h=h, h_u=b["h_u"], h_v=b["h_v"], frhatu=cs_bt["frhatu"], frhatv=cs_bt["frhatv"], bathyT=grid["bathyT"],
               hvel_scheme=4, may_use_default=0)
    stages = dict(continuity=(continuity_cs(nk), cont), coradcalc=(coriolisadv_cs(), corad),
                  horizontal_viscosity=(hor_visc_cs(dom, grid, dt=dt), hv), btstep=(cs_bt, a_bt), btcalc=(None, btc),
                  # eta consistent with the layer thicknesses, so the mass-source correction stays small (:5268-5292)
                  bt_mass_source=(None, dict(h=h, eta=np.ascontiguousarray((h.sum(axis=0) - grid["bathyT"]) * grid["mask2dT"] +
                                                                            1e-4 * a_bt["eta_in"]), eta_cor=cs_bt["eta_cor"])))
    # PressureForce (:503): T, S as in pressureforce_inputs on this state's h; pbce feeds btstep (:673)
    r = rng(seed + 404)
    zmid = -(np.cumsum(h, axis=0) - 0.5 * h)
    T = np.ascontiguousarray(20.0 * np.exp(zmid / 1000.0) + 0.05 * r.uniform(-1, 1, size=h.shape))
    S = np.ascontiguousarray(35.0 + 0.5 * np.exp(zmid / 500.0) + 0.01 * r.uniform(-1, 1, size=h.shape))
    pcs = dict(EOS_form=3, MassWghtInterp=0, use_SSH_in_Z0p=0, rho_ref_bug=0, unsupported=0, rho_ref=1035.0, GFS_scale=1.0, Z_ref=0.0,
               dZ_subroundoff=1e-30, Rho_T0_S0=1000.0, dRho_dT=-0.2, dRho_dS=0.8, dRho_dp=0.0, Rlay=None, g_prime=None)
    stages["pressure_force"] = (pcs, dict(h=h, T=T, S=S, PFu=new3("u"), PFv=new3("v"), pbce=new3("h"), eta=new2("h")))
    return dom, grid, gv, stages


def pressureforce_inputs(ni, nj, nk, halo=4, seed=SEED, land_blocks=0, cyclic_x=True, cyclic_y=False, eos="WRIGHT",
                         with_p_atm=False, with_pbce=True, with_eta=True, **cs_over):
    """Everything a PressureForce call needs (MOM_PressureForce.F90:40): returns dom, grid, vgrid, cs, args.
    T = 20 exp(z/1000) + noise, S = 35 + noise, Z*-like h with vanished layers over the seamount."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, cyclic_x=cyclic_x, cyclic_y=cyclic_y)
    grid = make_grid(dom, land_blocks, seed)
    gv = make_vgrid()
    st = dyn_state(dom, grid, seed)
    r = rng(seed + 404)
    h = st["h"]
    zmid = -(np.cumsum(h, axis=0) - 0.5 * h)
    T = 20.0 * np.exp(zmid / 1000.0) + 0.05 * r.uniform(-1, 1, size=h.shape)
    S = 35.0 + 0.5 * np.exp(zmid / 500.0) + 0.01 * r.uniform(-1, 1, size=h.shape)
    form = dict(NONE=0, LINEAR=1, WRIGHT=3)[eos]
    cs = dict(EOS_form=form, MassWghtInterp=0, use_SSH_in_Z0p=0, rho_ref_bug=0, unsupported=0, rho_ref=1035.0, GFS_scale=1.0,
              Z_ref=0.0, dZ_subroundoff=1e-30, Rho_T0_S0=1000.0, dRho_dT=-0.2, dRho_dS=0.8, dRho_dp=0.0,
              Rlay=np.ascontiguousarray(1026.0 + 2.0 * np.arange(nk) / max(nk - 1, 1)),
              g_prime=np.ascontiguousarray(np.concatenate(([9.8], np.full(nk, 9.8 * 2.0 / max(nk - 1, 1) / 1035.0)))))
    cs.update(cs_over)
    a = dict(h=h, T=np.ascontiguousarray(T) if form else None, S=np.ascontiguousarray(S) if form else None,
             PFu=fidx.new(dom, "u", nk=nk).a, PFv=fidx.new(dom, "v", nk=nk).a)
    if with_p_atm:
        a["p_atm"] = np.ascontiguousarray(1.0e5 + 500.0 * r.uniform(-1, 1, size=grid["bathyT"].shape))
    if with_pbce:
        a["pbce"] = fidx.new(dom, "h", nk=nk).a
    if with_eta:
        a["eta"] = fidx.new(dom, "h").a
    return dom, grid, gv, cs, a


def remap_inputs(ni, nj, nk, halo=4, seed=SEED, land_blocks=0, cyclic_x=True, cyclic_y=False, ntr=2, kind="zstar", **cs_over):
    """ALE remapping inputs (MOM_ALE.F90:760-925, :1089): h_old (Z*-like with vanished layers), h_new (kind 'zstar': the
    same column depth redistributed on a perturbed grid; 'uniform': equal layers), ntr tracers, u, v."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, cyclic_x=cyclic_x, cyclic_y=cyclic_y)
    grid = make_grid(dom, land_blocks, seed)
    st = dyn_state(dom, grid, seed)
    r = rng(seed + 505)
    h_old = st["h"]
    tot = h_old.sum(axis=0, keepdims=True)
    if kind == "uniform":
        w = np.ones_like(h_old)
    else:
        w = h_old * (1.0 + 0.3 * r.uniform(-1, 1, size=h_old.shape)) + 1.0e-3 * tot * (r.uniform(0, 1, size=h_old.shape) > 0.7)
    h_new = np.ascontiguousarray(w * (tot / w.sum(axis=0, keepdims=True)))
    zmid = -(np.cumsum(h_old, axis=0) - 0.5 * h_old)
    tr = [np.ascontiguousarray(20.0 * np.exp(zmid / 1000.0) + 0.05 * r.uniform(-1, 1, size=h_old.shape)),
          np.ascontiguousarray(35.0 + 0.5 * np.exp(zmid / 500.0) + 0.01 * r.uniform(-1, 1, size=h_old.shape))]
    for m in range(2, ntr):
        tr.append(np.ascontiguousarray(r.uniform(0, 1, size=h_old.shape) * 10.0 ** r.integers(-30, 1, size=h_old.shape)))
    cs = dict(remapping_scheme=4, boundary_extrapolation=0, force_bounds_in_subcell=0, force_bounds_in_target=1,
              om4_remap_via_sub_cells=1, answer_date=20190101, h_neglect=1.0e-30, h_neglect_edge=1.0e-30)   # OM4 defaults: PPM_H4
    cs.update(cs_over)
    return dom, grid, cs, dict(h_old=h_old, h_new=h_new, tr=tr[:ntr], u=st["u"], v=st["v"])


def advect_inputs(ni, nj, nk, halo=4, seed=SEED, land_blocks=0, cyclic_x=True, cyclic_y=False, ntr=2, dt=3600.0, dt_dyn=900.0,
                  cfl=0.7, scheme=0, **over):
    """advect_tracer inputs (MOM_tracer_advect.F90:53): transports accumulated over dt, each face moving up to cfl/4 of the
    upwind cell volume (cfl < 1: no cell is drained; cfl > 2: the flux limiter of :513-542 needs several passes and some
    cells are drained to the h_end floor), h_end consistent with them, T/S-like tracers plus random ones.  Returns dom, grid, gv, cs, args."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, cyclic_x=cyclic_x, cyclic_y=cyclic_y)
    grid = make_grid(dom, land_blocks, seed)
    gv = make_vgrid()
    st = dyn_state(dom, grid, seed)
    r = rng(seed + 606)
    h = st["h"]
    vol = h * grid["areaT"][None]
    # transports as a fraction of the upwind cell volume, so CFL is controlled everywhere (thin layers included)
    fu = cfl * st["u"] / max(np.abs(st["u"]).max(), 1e-30)
    fv = cfl * st["v"] / max(np.abs(st["v"]).max(), 1e-30)
    uhtr = fidx.new(dom, "u", nk=nk).a; vhtr = fidx.new(dom, "v", nk=nk).a
    uhtr[:, :, 1:-1] = fu[:, :, 1:-1] * np.where(fu[:, :, 1:-1] > 0, vol[:, :, :-1], vol[:, :, 1:]) * 0.25
    vhtr[:, 1:-1, :] = fv[:, 1:-1, :] * np.where(fv[:, 1:-1, :] > 0, vol[:, :-1, :], vol[:, 1:, :]) * 0.25
    uhtr *= grid["mask2dCu"][None]; vhtr *= grid["mask2dCv"][None]
    div = np.zeros_like(h)
    div[:, 1:-1, 1:-1] = (uhtr[:, 1:-1, 2:-1] - uhtr[:, 1:-1, 1:-2]) + (vhtr[:, 2:-1, 1:-1] - vhtr[:, 1:-2, 1:-1])
    h_end = np.ascontiguousarray(np.maximum(h - div * grid["IareaT"][None], 1.0e-10))
    zmid = -(np.cumsum(h, axis=0) - 0.5 * h)
    tr = [np.ascontiguousarray(20.0 * np.exp(zmid / 1000.0) + 2.0 * r.uniform(-1, 1, size=h.shape)),
          np.ascontiguousarray(35.0 + 0.5 * r.uniform(-1, 1, size=h.shape))]
    for m in range(2, ntr):
        tr.append(np.ascontiguousarray(r.uniform(0, 1, size=h.shape) * (r.uniform(0, 1, size=h.shape) > 0.5)))
    cs = dict(dt=dt_dyn, default_advect_scheme=scheme, useHuynhStencilBug=0)
    a = dict(h_end=h_end, uhtr=uhtr, vhtr=vhtr, dt=dt, tr=tr[:ntr])
    a.update(over)
    return dom, grid, gv, cs, a


def regrid_inputs(ni, nj, nk, halo=4, seed=SEED, land_blocks=0, cyclic_x=True, cyclic_y=False, eta_amp=0.5, **cs_over):
    """ALE_regrid inputs (MOM_ALE.F90:518): the Z*-like state of dyn_state with a sea-surface anomaly of up to eta_amp m
    added to the top layers, a stretched target resolution summing to the maximum depth, and the OM4-like time filter."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, cyclic_x=cyclic_x, cyclic_y=cyclic_y)
    grid = make_grid(dom, land_blocks, seed)
    gv = make_vgrid()
    st = dyn_state(dom, grid, seed)
    r = rng(seed + 707)
    h = st["h"].copy()
    h[0] += eta_amp * r.uniform(0, 1, size=h[0].shape) * grid["mask2dT"]
    w = np.linspace(1.0, 3.0, nk); w /= w.sum()
    cs = dict(regridding_scheme=2, nk=nk, min_thickness=1.0e-3, old_grid_weight=0.0, depth_of_time_filter_shallow=0.0,
              depth_of_time_filter_deep=0.0, Z_ref=0.0, coordinateResolution=np.ascontiguousarray(4000.0 * w))
    cs.update(cs_over)
    return dom, grid, gv, cs, dict(h=np.ascontiguousarray(h), h_new=np.zeros_like(h), dzRegrid=np.zeros((nk + 1,) + h.shape[1:]))


def vertvisc_cs(**over):
    """vertvisc_init defaults (MOM_vert_friction.F90:2929-3250) as OM4-like ALE runs resolve them."""
    cs = dict(bottomdraglaw=1, harmonic_visc=0, direct_stress=0, fixed_LOTW_ML=0, apply_LOTW_floor=0, dynamic_viscous_ML=0, nkml=0,
              answer_date=99991231, unsupported=0, Hbbl=10.0, Kv=1.0e-4, Kv_extra_bbl=0.0, Kvml_invZ2=0.0, Hmix=40.0, Hmix_stress=20.0,
              harm_BL_val=0.0, vonKar=0.41, vel_underflow=0.0, dZ_subroundoff=1.0e-30,
              CFL_based_trunc=1, CFL_trunc=0.5, maxvel=3.0e8)   # CFL_BASED_TRUNCATIONS, CFL_TRUNCATE, MAXVEL (:3389-3398)
    cs.update(over)
    return cs


def vertvisc_inputs(ni, nj, nk, halo=4, seed=SEED, land_blocks=0, cyclic_x=True, cyclic_y=False, dt=900.0, with_shear=True, with_Bu=False,
                    with_Ray=False, **cs_over):
    """Inputs of vertvisc_coef / vertvisc / vertvisc_remnant (MOM_vert_friction.F90:1357, :557, :1229): the dyn_state
    velocities and thicknesses, a bottom boundary layer of 2-20 m with a viscosity of 1e-3..1e-2 m2 s-1 (what
    set_viscous_BBL hands over), a shear-driven interface viscosity and a wind stress.  Returns dom, grid, gv, cs, coef
    args, solver args."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, cyclic_x=cyclic_x, cyclic_y=cyclic_y)
    grid = make_grid(dom, land_blocks, seed)
    gv = make_vgrid()
    st = dyn_state(dom, grid, seed)
    r = rng(seed + 808)
    new2 = lambda s, lo, hi: np.ascontiguousarray(r.uniform(lo, hi, size=fidx.new(dom, s).a.shape))   # noqa: E731
    coef = dict(u=st["u"], v=st["v"], h=st["h"], Kv_bbl_u=new2("u", 1e-3, 1e-2), Kv_bbl_v=new2("v", 1e-3, 1e-2),
                bbl_thick_u=new2("u", 2.0, 20.0), bbl_thick_v=new2("v", 2.0, 20.0), Kv_shear=None, Kv_shear_Bu=None,
                ustar=new2("h", 0.0, 0.02), dt=dt)
    if with_shear:
        ks = r.uniform(0, 1, size=(nk + 1,) + st["h"].shape[1:]) ** 8 * 0.05
        ks[0] = 0.0; ks[-1] = 0.0
        coef["Kv_shear"] = np.ascontiguousarray(ks)
    if with_Bu:
        kq = r.uniform(0, 1, size=(nk + 1,) + fidx.new(dom, "q").a.shape) ** 8 * 0.05
        coef["Kv_shear_Bu"] = np.ascontiguousarray(kq)
    sol = dict(u=st["u"].copy(), v=st["v"].copy(), h=st["h"], taux=new2("u", -0.2, 0.2), tauy=new2("v", -0.2, 0.2), Ray_u=None, Ray_v=None,
               dt=dt, taux_bot=fidx.new(dom, "u").a, tauy_bot=fidx.new(dom, "v").a)
    if with_Ray:
        sol["Ray_u"] = np.ascontiguousarray(r.uniform(0, 1e-5, size=st["u"].shape))
        sol["Ray_v"] = np.ascontiguousarray(r.uniform(0, 1e-5, size=st["v"].shape))
    return dom, grid, gv, vertvisc_cs(**cs_over), coef, sol


def step_dyn_inputs(ni, nj, nk, halo=4, whalo=10, seed=SEED, land_blocks=0, dt=900.0, vel=0.05, hnoise=1.0e-4, store_CAu=0, begw=0.0,
                    split_bottom_stress=0, calc_dtbt=0, **bt_over):
    """A full step_MOM_dyn_split_RK2 call (MOM_dynamics_split_RK2.F90:294): the shared state of step_inputs with a nearly
    level sea surface (so one step stays well inside CFL), the MOM_dyn_split_RK2_CS arrays as a previous step would have
    left them, visc% and forces%.  Returns dom, grid, gv, css (the stage control structures), cs, args."""
    dom, grid, gv, stages = step_inputs(ni, nj, nk, halo=halo, whalo=whalo, seed=seed, land_blocks=land_blocks, dt=dt, vel=vel, hnoise=hnoise,
                                        **bt_over)
    r = rng(seed + 909)
    cs_bt, a_bt = stages["btstep"]
    cont = stages["continuity"][1]; pf = stages["pressure_force"][1]
    u, v, h = cont["u"].copy(), cont["v"].copy(), cont["hin"].copy()
    new3 = lambda st: fidx.new(dom, st, nk=nk).a          # noqa: E731
    new2 = lambda st: fidx.new(dom, st).a                 # noqa: E731
    rnd2 = lambda st, lo, hi: np.ascontiguousarray(r.uniform(lo, hi, size=new2(st).shape))   # noqa: E731
    uh, vh = transports(dom, grid, dict(h=h, u=u, v=v))
    eta = np.ascontiguousarray((h.sum(axis=0) - grid["bathyT"]) * grid["mask2dT"])
    css = dict(continuity=stages["continuity"][0], coriolisadv=stages["coradcalc"][0], hor_visc=stages["horizontal_viscosity"][0],
               pressureforce=stages["pressure_force"][0], vertvisc=vertvisc_cs())
    cs = dict(be=0.6, begw=begw, split_bottom_stress=split_bottom_stress, store_CAu=store_CAu, CAu_pred_stored=0, visc_rem_dt_bug=1, hvel_scheme=4,
              unsupported=0, dtbt_use_bt_cont=0, BT_Nonlinear_continuity=0, dtbt_fraction=0.98, BT_Coriolis_scale=1.0, Z_ref=0.0, dtbt_max=0.0, CAu=new3("u"), CAv=new3("v"), CAu_pred=new3("u"), CAv_pred=new3("v"), PFu=new3("u"), PFv=new3("v"),
              diffu=new3("u"), diffv=new3("v"), visc_rem_u=new3("u"), visc_rem_v=new3("v"), u_accel_bt=new3("u"), v_accel_bt=new3("v"),
              u_av=u.copy(), v_av=v.copy(), h_av=h.copy(), pbce=new3("h"), eta=eta, eta_PF=new2("h"), uhbt=new2("u"), vhbt=new2("v"),
              taux_bot=new2("u"), tauy_bot=new2("v"), BT_cont=dict(cont["BT_cont"]), barotropic=dict(cs_bt))
    kvs = r.uniform(0, 1, size=(nk + 1,) + h.shape[1:]) ** 8 * 0.02
    kvs[0] = 0.0; kvs[-1] = 0.0
    a = dict(u_inst=u, v_inst=v, h=h, T=pf["T"], S=pf["S"], Kv_bbl_u=rnd2("u", 1e-3, 1e-2), Kv_bbl_v=rnd2("v", 1e-3, 1e-2),
             bbl_thick_u=rnd2("u", 2.0, 20.0), bbl_thick_v=rnd2("v", 2.0, 20.0), Kv_shear=np.ascontiguousarray(kvs), Kv_shear_Bu=None, Ray_u=None,
             Ray_v=None, taux=np.ascontiguousarray(a_bt["taux"]), tauy=np.ascontiguousarray(a_bt["tauy"]), ustar=rnd2("h", 0.0, 0.02), p_surf=None,
             dt=dt, uh=uh, vh=vh, uhtr=new3("u"), vhtr=new3("v"), eta_av=new2("h"), calc_dtbt=calc_dtbt)
    return dom, grid, gv, css, cs, a


# stagger / memory domain of every array argument of the step (for tile splitting and for resident uploads)
STEP_STAGGER = dict(
    CAu="u", CAv="v", CAu_pred="u", CAv_pred="v", PFu="u", PFv="v", diffu="u", diffv="v", visc_rem_u="u", visc_rem_v="v", u_accel_bt="u",
    v_accel_bt="v", u_av="u", v_av="v", h_av="h", pbce="h", eta="h", eta_PF="h", uhbt="u", vhbt="v", taux_bot="u", tauy_bot="v",
    u_inst="u", v_inst="v", h="h", T="h", S="h", Kv_bbl_u="u", Kv_bbl_v="v", bbl_thick_u="u", bbl_thick_v="v", Kv_shear="h", Kv_shear_Bu="q",
    Ray_u="u", Ray_v="v", taux="u", tauy="v", ustar="h", p_surf="h", uh="u", vh="v", uhtr="u", vhtr="v", eta_av="h",
    FA_u_EE="u", FA_u_E0="u", FA_u_W0="u", FA_u_WW="u", uBT_WW="u", uBT_EE="u", FA_v_NN="v", FA_v_N0="v", FA_v_S0="v", FA_v_SS="v",
    vBT_SS="v", vBT_NN="v", h_u="u", h_v="v",
    IareaT="h", IareaT_OBCmask="h", bathyT="h", IdxCu="u", IdyCv="v", q_D="q", D_u_Cor="u", D_v_Cor="v", ua_polarity="h", va_polarity="h",
    OBCmask_u="u", OBCmask_v="v", frhatu="u", frhatv="v", eta_cor="h", eta_cor_bound="h", IDatu="u", IDatv="v", ubtav="u", vbtav="v")
BT_WIDE = {"IareaT", "IareaT_OBCmask", "bathyT", "IdxCu", "IdyCv", "q_D", "D_u_Cor", "D_v_Cor", "ua_polarity", "va_polarity", "OBCmask_u", "OBCmask_v"}
GRID_STAGGER = dict(mask2dT="h", mask2dCu="u", mask2dCv="v", mask2dBu="q", dxT="h", dyT="h", IdxT="h", IdyT="h", areaT="h", IareaT="h",
                    dxCu="u", dyCu="u", IdxCu="u", IdyCu="u", dy_Cu="u", areaCu="u", IareaCu="u", dxCv="v", dyCv="v", IdxCv="v", IdyCv="v",
                    dx_Cv="v", areaCv="v", IareaCv="v", dxBu="q", dyBu="q", IdxBu="q", IdyBu="q", areaBu="q", IareaBu="q", bathyT="h",
                    CoriolisBu="q", Coriolis2Bu="q")


def _cut(dom_g, dom, arr, st, wide, oi, oj):
    gl, tl = fidx.extent(dom_g, st, wide), fidx.extent(dom, st, wide)
    j0, i0 = tl[2] + oj - gl[2], tl[0] + oi - gl[0]
    return np.ascontiguousarray(arr[..., j0:j0 + tl[3] - tl[2] + 1, i0:i0 + tl[1] - tl[0] + 1])


def split_step_tile(dom_g, grid, cs, a, npi, npj, pi, pj, hor_visc_cs_g=None):
    """One rank's tile (with halos) of global step_dyn_inputs: dom, grid, cs, args (and the hor_visc CS arrays)."""
    NI, NJ = dom_g.iec - dom_g.isc + 1, dom_g.jec - dom_g.jsc + 1
    ni, nj = NI // npi, NJ // npj
    halo, whalo = dom_g.isc - dom_g.isd, dom_g.isc - dom_g.isdw
    dom = make_domain(ni, nj, nk=dom_g.nk, halo=halo, whalo=whalo, cyclic_x=bool(dom_g.cyclic_x), cyclic_y=bool(dom_g.cyclic_y),
                      first_direction=dom_g.first_direction, npi=npi, npj=npj, pi=pi, pj=pj)
    oi, oj = pi * ni, pj * nj
    cut = lambda k, v, wide=False: (_cut(dom_g, dom, v, STEP_STAGGER[k], wide, oi, oj) if isinstance(v, np.ndarray) else v)   # noqa: E731
    g = {k: (_cut(dom_g, dom, v, GRID_STAGGER[k], False, oi, oj) if isinstance(v, np.ndarray) else v) for k, v in grid.items()}
    c = {k: cut(k, v) for k, v in cs.items() if k not in ("BT_cont", "barotropic")}
    c["BT_cont"] = {k: cut(k, v) for k, v in cs["BT_cont"].items()}
    c["barotropic"] = {k: cut(k, v, k in BT_WIDE) for k, v in cs["barotropic"].items()}
    t = {k: cut(k, v) for k, v in a.items()}
    hv = None
    if hor_visc_cs_g is not None:
        hv = {}
        for k, v in hor_visc_cs_g.items():
            if isinstance(v, np.ndarray):
                for st in ("h", "q", "u", "v"):
                    gl = fidx.extent(dom_g, st, False)
                    if v.shape == (gl[3] - gl[2] + 1, gl[1] - gl[0] + 1):
                        hv[k] = _cut(dom_g, dom, v, st, False, oi, oj)
                        break
                else:
                    raise ValueError(f"split_step_tile: cannot place hor_visc CS array {k} {v.shape}")
            else:
                hv[k] = v
    return dom, g, c, t, hv


def sum_output_cs(dom, depth_list, dt=900.0, do_APE_calc=True, use_temperature=True, Z_ref=0.0, **units):
    """Sum_output_CS as MOM_sum_output_init :147 and depth_list_setup :1161 leave it (src/diagnostics/MOM_sum_output.F90).
    depth_list = (depth, area, vol_below) from create_depth_list :1203 (host code of the reference; the tests take it from the
    oracle's restatement)."""
    nk = dom.nk
    depth, area, vol = (np.ascontiguousarray(x, dtype=np.float64) for x in depth_list)
    g_prime = np.ascontiguousarray(np.concatenate(([9.8], 9.8 * (0.5 + 0.02 * np.arange(1, nk + 1)) / 1035.0)))  # GV%g_prime(1:nk+1)
    cs = dict(do_APE_calc=int(do_APE_calc), use_temperature=int(use_temperature), dt_in_T=dt, DL_listsize=len(depth), DL_depth=depth,
              DL_area=area, DL_vol_below=vol, lH=np.full(nk, len(depth) - 1, dtype=np.int32), g_prime=g_prime, Z_ref=Z_ref,
              C_p=3991.86795711963, previous_calls=0, ntrunc=0)
    cs.update(units)
    return cs


def ale_chain_inputs(ni, nj, nk, seed=SEED, land_blocks=0, dtdia=7200.0, regrid_time_scale=3600.0, remap_aux_vars=1, store_CAu=1, with_Bu=True,
                     remapping_scheme=4, eta_amp=0.5):
    """ALE_regridding_and_remapping inputs (src/core/MOM.F90:1751): the state, visc% and MOM_dyn_split_RK2_CS of step_dyn_inputs
    with a sea-surface anomaly added to the top layer, three tracers (T, S and a passive one with tiny values), the Z* target grid
    of regrid_inputs and the OM4-like remapping control structures (PPM_H4 for tracers, the same scheme for velocities)."""
    dom, grid, gv, css, cs, a = step_dyn_inputs(ni, nj, nk, whalo=6, seed=seed, land_blocks=land_blocks, store_CAu=store_CAu)
    r = rng(seed + 1313)
    h = a["h"].copy()
    h[0] += eta_amp * r.uniform(0, 1, size=h[0].shape) * grid["mask2dT"]
    w = np.linspace(1.0, 3.0, nk); w /= w.sum()
    regridCS = dict(regridding_scheme=2, nk=nk, min_thickness=1.0e-3, old_grid_weight=0.0, depth_of_time_filter_shallow=0.0,
                    depth_of_time_filter_deep=0.0, Z_ref=0.0, coordinateResolution=np.ascontiguousarray(4000.0 * w))
    remapCS = dict(remapping_scheme=remapping_scheme, boundary_extrapolation=0, force_bounds_in_subcell=0, force_bounds_in_target=1,
                   om4_remap_via_sub_cells=1, answer_date=20190101, h_neglect=1.0e-30, h_neglect_edge=1.0e-30)
    ale = dict(regridCS=regridCS, remapCS=remapCS, vel_remapCS=dict(remapCS), regrid_time_scale=regrid_time_scale, remap_aux_vars=remap_aux_vars)
    passive = np.ascontiguousarray(r.uniform(0, 1, size=h.shape) * 10.0 ** r.integers(-30, 1, size=h.shape))
    shp1 = (nk + 1,) + h.shape[1:]
    kd = np.ascontiguousarray(r.uniform(0, 1, size=shp1) ** 6 * 1.0e-3)
    kvb = None
    if with_Bu:
        q = fidx.new(dom, "q", nk=nk + 1).a
        kvb = np.ascontiguousarray(r.uniform(0, 1, size=q.shape) ** 6 * 0.02)
    for k in ("diffu", "diffv", "CAu_pred", "CAv_pred"):   # as a previous step would have left them
        cs[k][...] = 1.0e-6 * r.standard_normal(cs[k].shape)
    args = dict(u=a["u_inst"].copy(), v=a["v_inst"].copy(), h=np.ascontiguousarray(h), tr=[a["T"].copy(), a["S"].copy(), passive],
                conc_underflow=np.array([0.0, 0.0, 1.0e-25]), iT=0, iS=1, dtdia=dtdia, Kd_shear=kd, Kv_shear=a["Kv_shear"].copy(), Kv_shear_Bu=kvb)
    return dom, grid, gv, ale, cs, args


def mle_cs_and_forcing(shp2, seed=SEED, eos="WRIGHT", **cs_over):
    """The OM4_025-like mixedlayer_restrat_CS (FOX_KEMPER_ML_RESTRAT_COEF = 1 with MLE_FRONT_LENGTH = 500 m, MLE_MLD_DECAY_TIME = 30 d,
    MLE_USE_PBL_MLD; mixedlayer_restrat_init, MOM_mixed_layer_restrat.F90:1618) and the 2-D inputs of a call on h-point arrays of
    shape shp2: a boundary-layer depth h_MLD of 20-150 m (zero on some columns), u* of 0-2 cm/s (zero on some), Rd_dx_h of 0.2-1.7."""
    r = rng(seed + 1414)
    h_MLD = (20.0 + 130.0 * r.uniform(0, 1, size=shp2)) * (r.uniform(0, 1, size=shp2) > 0.05)
    ustar = 0.02 * r.uniform(0, 1, size=shp2) * (r.uniform(0, 1, size=shp2) > 0.1)
    Rd = 0.2 + 1.5 * r.uniform(0, 1, size=shp2)
    form = dict(LINEAR=1, WRIGHT=3)[eos]
    cs = dict(ml_restrat_coef=1.0, ml_restrat_coef2=0.0, front_length=500.0, MLE_MLD_decay_time=2.592e6, MLE_MLD_decay_time2=0.0,
              MLE_MLD_stretch=1.0, MLE_tail_dh=0.0, ustar_min=2.0e-4 * 7.2921e-5 * (1.0e-10 + 1.0e-30), vonKar=0.41,
              MLE_density_diff=-9.0e9, MLE_use_PBL_MLD=1, use_Stanley_ML=0, use_Bodner=0, fl_from_file=0, EOS_form=form,
              Rho_T0_S0=1000.0, dRho_dT=-0.2, dRho_dS=0.8, dRho_dp=0.0,
              MLD_filtered=np.ascontiguousarray(60.0 * r.uniform(0, 1, size=shp2)),
              MLD_filtered_slow=np.ascontiguousarray(80.0 * r.uniform(0, 1, size=shp2)))
    cs.update(cs_over)
    return cs, dict(ustar=np.ascontiguousarray(ustar), h_MLD=np.ascontiguousarray(h_MLD), Rd_dx_h=np.ascontiguousarray(Rd))


def mle_inputs(ni, nj, nk, halo=4, seed=SEED, land_blocks=0, cyclic_x=True, cyclic_y=False, eos="WRIGHT", dt=900.0, front=2.0,
               **cs_over):
    """mixedlayer_restrat inputs (MOM_mixed_layer_restrat.F90:149): the Z*-like state of dyn_state, T with lateral fronts of
    `front` degC in the upper ocean, accumulated transports uhtr / vhtr, and mle_cs_and_forcing.  Returns dom, grid, gv, cs, args."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, cyclic_x=cyclic_x, cyclic_y=cyclic_y)
    grid = make_grid(dom, land_blocks, seed)
    gv = make_vgrid()
    st = dyn_state(dom, grid, seed)
    r = rng(seed + 1515)
    h = st["h"]
    shp2 = h.shape[1:]
    zmid = -(np.cumsum(h, axis=0) - 0.5 * h)
    jj, ii = np.meshgrid(np.arange(shp2[0]), np.arange(shp2[1]), indexing="ij")
    frontal = np.sin(2 * np.pi * ii / max(ni, 1) * 3.0) * np.cos(2 * np.pi * jj / max(nj, 1) * 2.0)
    T = 20.0 * np.exp(zmid / 1000.0) + front * frontal[None] * np.exp(zmid / 200.0) + 0.05 * r.uniform(-1, 1, size=h.shape)
    S = 35.0 + 0.5 * np.exp(zmid / 500.0) + 0.01 * r.uniform(-1, 1, size=h.shape)
    uhtr = np.ascontiguousarray(0.05 * st["u"] * dt * grid["dyCu"][None] * 10.0)
    vhtr = np.ascontiguousarray(0.05 * st["v"] * dt * grid["dxCv"][None] * 10.0)
    cs, f2 = mle_cs_and_forcing(shp2, seed, eos, **cs_over)
    a = dict(h=np.ascontiguousarray(h), uhtr=uhtr, vhtr=vhtr, T=np.ascontiguousarray(T), S=np.ascontiguousarray(S), dt=dt, **f2)
    return dom, grid, gv, cs, a


def hordiff_cs(**over):
    """tracer_hor_diff_init defaults (MOM_tracer_hor_diff.F90:1630-1778) with a KHTR that matters on a 25 km mesh."""
    cs = dict(KhTr=2000.0, KhTr_min=0.0, KhTr_max=0.0, KhTr_passivity_coeff=0.0, KhTr_passivity_min=0.5, KhTr_Slope_Cff=0.0, max_diff_CFL=-1.0,
              check_diffusive_CFL=0, use_neutral_diffusion=0, use_hor_bnd_diffusion=0, Diffuse_ML_interior=0, use_variable_mixing=0,
              Resoln_scaled_KhTr=0, use_MEKE_Kh=0, MEKE_KhTr_fac=1.0)
    cs.update(over)
    return cs


def hordiff_inputs(ni, nj, nk, halo=4, seed=SEED, land_blocks=0, cyclic_x=True, cyclic_y=False, ntr=3, dt=7200.0, with_df=False, **cs_over):
    """tracer_hordiff inputs (MOM_tracer_hor_diff.F90:119): the Z*-like state of dyn_state (vanished layers included), T- and S-like
    tracers with noise plus random / tiny-valued passive ones, VarMix's Res_fn_h and Rd_dx_h.  Returns dom, grid, gv, cs, args."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, cyclic_x=cyclic_x, cyclic_y=cyclic_y)
    grid = make_grid(dom, land_blocks, seed)
    gv = make_vgrid()
    st = dyn_state(dom, grid, seed)
    r = rng(seed + 1616)
    h = st["h"]
    zmid = -(np.cumsum(h, axis=0) - 0.5 * h)
    tr = [np.ascontiguousarray(20.0 * np.exp(zmid / 1000.0) + 2.0 * r.uniform(-1, 1, size=h.shape)),
          np.ascontiguousarray(35.0 + 0.5 * r.uniform(-1, 1, size=h.shape))]
    for m in range(2, ntr):
        tr.append(np.ascontiguousarray(r.uniform(0, 1, size=h.shape) * 10.0 ** r.integers(-30, 1, size=h.shape)))
    shp2 = h.shape[1:]
    a = dict(h=np.ascontiguousarray(h), dt=dt, tr=tr[:ntr], conc_underflow=np.array(([0.0, 0.0] + [1.0e-25] * ntr)[:ntr]),
             Res_fn_h=np.ascontiguousarray(r.uniform(0, 1, size=shp2)), Rd_dx_h=np.ascontiguousarray(0.2 + 1.5 * r.uniform(0, 1, size=shp2)))
    # VarMix%L2u / SN_u (calc_slope_functions: a squared length of O(dx^2) and an Eady growth rate of O(1e-6 s-1)) and MEKE%Kh
    for key, st in (("L2u", "u"), ("SN_u", "u"), ("L2v", "v"), ("SN_v", "v")):
        amp = 6.25e8 if key.startswith("L2") else 3.0e-6
        a[key] = np.ascontiguousarray(amp * r.uniform(0, 1, size=fidx.new(dom, st).a.shape))
    a["MEKE_Kh"] = np.ascontiguousarray(1500.0 * r.uniform(0, 1, size=shp2) ** 2)
    if with_df:
        a["df_x"] = [fidx.new(dom, "u", nk=nk, fill=7.0).a if m != 1 else None for m in range(ntr)]
        a["df_y"] = [fidx.new(dom, "v", nk=nk, fill=7.0).a if m != 0 else None for m in range(ntr)]
    return dom, grid, gv, hordiff_cs(**cs_over), a


def thickness_diffuse_cs(**over):
    """thickness_diffuse_init (MOM_thickness_diffuse.F90:2170-2476) as a GM-only ALE run resolves it: KHTH = 600 m2 s-1, KHTH_MAX_CFL = 0.8,
    KHTH_SLOPE_MAX = 0.01, KD_SMOOTH = 1e-6, the Wright equation of state."""
    cs = dict(Khth=600.0, Khth_Min=0.0, Khth_Max=0.0, max_Khth_CFL=0.8, slope_max=0.01, kappa_smooth=1.0e-6, dZ_subroundoff=1.0e-30,
              thickness_diffuse=1, read_khth=0, detangle_interfaces=0, interface_Kh=0, use_FGNV_streamfn=0, use_stanley_gm=0,
              use_GME_thickness_diffuse=0, find_work=0, use_variable_mixing=0, Resoln_scaled_KhTh=0, Depth_scaled_KhTh=0, use_stored_slopes=0,
              use_Visbeck=0, use_QG_Leith_GM=0, khth_struct=0, use_MEKE_Kh=0, EOS_form=3, Rho_T0_S0=1000.0, dRho_dT=-0.2, dRho_dS=0.8, dRho_dp=0.0,
              FGNV_scale=1.0, N2_floor=0.0, MEKE_KhTh_fac=1.0)
    cs.update(over)
    if cs["use_FGNV_streamfn"] and "N2_floor" not in over:
        cs["N2_floor"] = (1.0e-15 * 7.2921e-5) ** 2            # (FGNV_STRAT_FLOOR*OMEGA)**2, MOM_thickness_diffuse.F90:2340-2341
    return cs


def thickness_diffuse_inputs(ni, nj, nk, halo=4, seed=SEED, land_blocks=0, cyclic_x=True, cyclic_y=False, dt=900.0, front=2.0, with_p_surf=False,
                             with_GM=False, **cs_over):
    """thickness_diffuse inputs (MOM_thickness_diffuse.F90:134): the Z*-like state of dyn_state over the seamount (sloping interfaces near
    the bottom, vanished layers), T with lateral fronts and noise, S, accumulated transports, VarMix's Res_fn_u / Res_fn_v.  Returns dom,
    grid, gv, cs, args."""
    dom = make_domain(ni, nj, nk=nk, halo=halo, cyclic_x=cyclic_x, cyclic_y=cyclic_y)
    grid = make_grid(dom, land_blocks, seed)
    gv = make_vgrid()
    st = dyn_state(dom, grid, seed)
    r = rng(seed + 1717)
    h = st["h"]
    shp2 = h.shape[1:]
    zmid = -(np.cumsum(h, axis=0) - 0.5 * h)
    jj, ii = np.meshgrid(np.arange(shp2[0]), np.arange(shp2[1]), indexing="ij")
    frontal = np.sin(2 * np.pi * ii / max(ni, 1) * 2.0) * np.cos(2 * np.pi * jj / max(nj, 1) * 3.0)
    T = 20.0 * np.exp(zmid / 1000.0) + front * frontal[None] * np.exp(zmid / 800.0) + 0.02 * r.uniform(-1, 1, size=h.shape)
    S = 35.0 + 0.5 * np.exp(zmid / 500.0) + 0.01 * r.uniform(-1, 1, size=h.shape)
    a = dict(h=np.ascontiguousarray(h), uhtr=np.ascontiguousarray(0.05 * st["u"] * dt * grid["dyCu"][None] * 10.0),
             vhtr=np.ascontiguousarray(0.05 * st["v"] * dt * grid["dxCv"][None] * 10.0), T=np.ascontiguousarray(T), S=np.ascontiguousarray(S),
             p_surf=np.ascontiguousarray(1.0e5 + 500.0 * r.uniform(-1, 1, size=shp2)) if with_p_surf else None, dt=dt,
             Res_fn_u=np.ascontiguousarray(r.uniform(0, 1, size=fidx.new(dom, "u").a.shape)),
             Res_fn_v=np.ascontiguousarray(r.uniform(0, 1, size=fidx.new(dom, "v").a.shape)),
             uhGM=fidx.new(dom, "u", nk=nk, fill=7.0).a if with_GM else None, vhGM=fidx.new(dom, "v", nk=nk, fill=7.0).a if with_GM else None,
             # VarMix%slope_x / slope_y (calc_isoneutral_slopes: |slope| up to a few 1e-3, zero at the top and bottom interfaces), VarMix%cg1, MEKE%Kh
             slope_x=np.ascontiguousarray(3.0e-3 * r.uniform(-1, 1, size=(nk + 1,) + fidx.new(dom, "u").a.shape) ** 3),
             slope_y=np.ascontiguousarray(3.0e-3 * r.uniform(-1, 1, size=(nk + 1,) + fidx.new(dom, "v").a.shape) ** 3),
             cg1=np.ascontiguousarray(3.0 * r.uniform(0, 1, size=shp2) * (r.uniform(0, 1, size=shp2) > 0.1)),
             MEKE_Kh=np.ascontiguousarray(1500.0 * r.uniform(0, 1, size=shp2) ** 2))
    a["slope_x"][0] = 0.0; a["slope_x"][-1] = 0.0; a["slope_y"][0] = 0.0; a["slope_y"][-1] = 0.0
    return dom, grid, gv, thickness_diffuse_cs(**cs_over), a
