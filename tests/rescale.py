"""Power-of-2 dimensional rescaling of the inputs of the hot path: the reference's `dim` regression tests (.testing `test.dim.t/l/h/z`,
T_RESCALE_POWER etc.; src/framework/MOM_unit_scaling.F90) re-expressed on dict inputs.  Every quantity carries its dimension as exponents
of (T, L, H, Z); rescaling multiplies it by 2**(t*pT + l*pL + h*pH + z*pZ), which is exact in binary floating point, so a dimensionally
consistent routine must return bit-identical answers once they are scaled back."""
import numpy as np

#                 T   L   H   Z
NONDIM = (0, 0, 0, 0)
DIMS_GRID = dict(dxT=(0, 1, 0, 0), dyT=(0, 1, 0, 0), IdxT=(0, -1, 0, 0), IdyT=(0, -1, 0, 0), areaT=(0, 2, 0, 0), IareaT=(0, -2, 0, 0),
                 dxCu=(0, 1, 0, 0), dyCu=(0, 1, 0, 0), IdxCu=(0, -1, 0, 0), IdyCu=(0, -1, 0, 0), dy_Cu=(0, 1, 0, 0), areaCu=(0, 2, 0, 0),
                 IareaCu=(0, -2, 0, 0), dxCv=(0, 1, 0, 0), dyCv=(0, 1, 0, 0), IdxCv=(0, -1, 0, 0), IdyCv=(0, -1, 0, 0), dx_Cv=(0, 1, 0, 0),
                 areaCv=(0, 2, 0, 0), IareaCv=(0, -2, 0, 0), dxBu=(0, 1, 0, 0), dyBu=(0, 1, 0, 0), IdxBu=(0, -1, 0, 0), IdyBu=(0, -1, 0, 0),
                 areaBu=(0, 2, 0, 0), IareaBu=(0, -2, 0, 0), mask2dT=NONDIM, mask2dCu=NONDIM, mask2dCv=NONDIM, mask2dBu=NONDIM,
                 bathyT=(0, 0, 0, 1), CoriolisBu=(-1, 0, 0, 0), Coriolis2Bu=(-2, 0, 0, 0))
DIMS_GV = dict(Angstrom_H=(0, 0, 1, 0), H_subroundoff=(0, 0, 1, 0), Z_to_H=(0, 0, 1, -1), H_to_Z=(0, 0, -1, 1), g_Earth=(-2, 2, 0, -1),
               Rho0=NONDIM, H_to_RZ=(0, 0, -1, 1), RZ_to_H=(0, 0, 1, -1), H_to_m=(0, 0, -1, 0), m_to_H=(0, 0, 1, 0), Boussinesq=None)
VEL, THK, TRANSP, ACC = (-1, 1, 0, 0), (0, 0, 1, 0), (-1, 2, 1, 0), (-2, 1, 0, 0)


def factor(dim, p):
    return 2.0 ** sum(d * q for d, q in zip(dim, p))


def scale(d, dims, p, inverse=False):
    """Scale every entry of d that has a dimension in dims (None = leave alone); nested dicts and lists of arrays are followed."""
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out[k] = scale(v, dims, p, inverse)
            continue
        if v is None:
            out[k] = None
            continue
        dim = dims.get(k, "missing")
        if dim == "missing":
            raise KeyError(f"no dimension recorded for {k}")
        if dim is None or v is None:
            out[k] = v
            continue
        f = factor(dim, p)
        if inverse:
            f = 1.0 / f
        if isinstance(v, list):
            out[k] = [None if x is None else np.ascontiguousarray(x * f) for x in v]
        elif isinstance(v, np.ndarray):
            out[k] = np.ascontiguousarray(v * f)
        else:
            out[k] = v * f if isinstance(v, float) else v
    return out


def unit_scale(p):
    """mom6cu_unit_scale (US%...) for the powers p = (pT, pL, pH, pZ)."""
    T, L, Z = 2.0 ** p[0], 2.0 ** p[1], 2.0 ** p[3]
    return dict(m_to_L=L, L_to_m=1.0 / L, m_s_to_L_T=L / T, L_T_to_m_s=T / L, s_to_T=T, T_to_s=1.0 / T, m_to_Z=Z, Z_to_m=1.0 / Z,
                Z_to_L=L / Z, L_to_Z=Z / L)


TIME, STRESS, FACE_AREA, VOL = (1, 0, 0, 0), (-2, 1, 0, 1), (0, 1, 1, 0), (0, 2, 1, 0)
L2, L3, IL3, L4, L4T, L2T, HZT, ZL, HT = (0, 2, 0, 0), (0, 3, 0, 0), (0, -3, 0, 0), (0, 4, 0, 0), (-1, 4, 0, 0), (-1, 2, 0, 0), (-1, 0, 1, 1), (0, 0, 0, 1), (-1, 0, 1, 0)
BT_CONT = dict(FA_u_EE=FACE_AREA, FA_u_E0=FACE_AREA, FA_u_W0=FACE_AREA, FA_u_WW=FACE_AREA, uBT_WW=VEL, uBT_EE=VEL, FA_v_NN=FACE_AREA,
               FA_v_N0=FACE_AREA, FA_v_S0=FACE_AREA, FA_v_SS=FACE_AREA, vBT_SS=VEL, vBT_NN=VEL, h_u=THK, h_v=THK)
# continuity_PPM (MOM_continuity_PPM.F90:86-141, CS :35-67)
CONT = dict(u=VEL, v=VEL, hin=THK, h=THK, uh=TRANSP, vh=TRANSP, dt=TIME, visc_rem_u=NONDIM, visc_rem_v=NONDIM, uhbt=TRANSP, vhbt=TRANSP, u_cor=VEL,
            v_cor=VEL, du_cor=VEL, dv_cor=VEL, **BT_CONT)
CONT_CS = dict(tol_eta=THK, tol_vel=VEL, CFL_limit_adjust=NONDIM)
# CorAdCalc (MOM_CoriolisAdv.F90:125-144)
CORAD = dict(u=VEL, v=VEL, h=THK, uh=TRANSP, vh=TRANSP, CAu=ACC, CAv=ACC, RV=(-1, 0, 0, 0), PV=(-1, 0, -1, 0), gradKEu=ACC, gradKEv=ACC,
             por_face_areaU=NONDIM, por_face_areaV=NONDIM)
CORAD_CS = dict(F_eff_max_blend=NONDIM, wt_lin_blend=NONDIM)
# horizontal_viscosity (MOM_hor_visc.F90:266-305, hor_visc_CS :38-250)
HORVISC = dict(u=VEL, v=VEL, h=THK, dt=TIME, diffu=ACC, diffv=ACC, hu_cont=THK, hv_cont=THK)
HORVISC_CS = dict(dx2q=L2, dy2q=L2, dx2h=L2, dy2h=L2, DX_dyBu=NONDIM, DY_dxBu=NONDIM, DX_dyT=NONDIM, DY_dxT=NONDIM, reduction_xx=NONDIM,
                  reduction_xy=NONDIM, Idx2dyCu=IL3, Idxdy2u=IL3, Idx2dyCv=IL3, Idxdy2v=IL3, Biharm_const_xx=L4, Biharm_const_xy=L4, Ah_bg_xx=L4T,
                  Ah_bg_xy=L4T, Ah_Max_xx=L4T, Ah_Max_xy=L4T, Laplac2_const_xx=L2, Laplac2_const_xy=L2, Kh_bg_xx=L2T, Kh_bg_xy=L2T, Kh_Max_xx=L2T,
                  Kh_Max_xy=L2T, Kh_bg_min=L2T, Re_Ah=NONDIM, Re_Ah_const_xx=L3, Re_Ah_const_xy=L3)
# vertvisc_coef / vertvisc / vertvisc_remnant (MOM_vert_friction.F90:48-170, :557, :1229, :1357)
VERTVISC_CS = dict(Hbbl=ZL, Kv=HZT, Kv_extra_bbl=HZT, Kvml_invZ2=HZT, Hmix=ZL, Hmix_stress=THK, harm_BL_val=NONDIM, vonKar=NONDIM, vel_underflow=VEL,
                   dZ_subroundoff=ZL, maxvel=VEL, CFL_trunc=NONDIM)
VERTVISC_COEF = dict(u=VEL, v=VEL, h=THK, Kv_bbl_u=HZT, Kv_bbl_v=HZT, bbl_thick_u=ZL, bbl_thick_v=ZL, Kv_shear=HZT, Kv_shear_Bu=HZT, ustar=(-1, 0, 0, 1),
                     dt=TIME)
VERTVISC = dict(u=VEL, v=VEL, h=THK, taux=STRESS, tauy=STRESS, Ray_u=HT, Ray_v=HT, dt=TIME, taux_bot=STRESS, tauy_bot=STRESS)
VERTVISC_OUT = dict(a_u=HT, a_v=HT, h_u=THK, h_v=THK, visc_rem_u=NONDIM, visc_rem_v=NONDIM)
# btstep (MOM_barotropic.F90:455-529, barotropic_CS :112-364)
BTSTEP_CS = dict(dtbt=TIME, bebt=NONDIM, vel_underflow=VEL, maxCFL_BT_cont=NONDIM, G_extra=NONDIM, dt_bt_filter=None, IareaT=(0, -2, 0, 0),
                 IareaT_OBCmask=(0, -2, 0, 0), bathyT=ZL, IdxCu=(0, -1, 0, 0), IdyCv=(0, -1, 0, 0), q_D=(-1, 0, -1, 0), D_u_Cor=THK, D_v_Cor=THK,
                 ua_polarity=NONDIM, va_polarity=NONDIM, OBCmask_u=NONDIM, OBCmask_v=NONDIM, frhatu=NONDIM, frhatv=NONDIM, eta_cor=THK,
                 eta_cor_bound=THK, IDatu=(0, 0, -1, 0), IDatv=(0, 0, -1, 0), ubtav=VEL, vbtav=VEL)
BTSTEP = dict(U_in=VEL, V_in=VEL, eta_in=THK, dt=TIME, bc_accel_u=ACC, bc_accel_v=ACC, taux=STRESS, tauy=STRESS, pbce=(-2, 2, -1, 0), eta_PF_in=THK,
              U_Cor=VEL, V_Cor=VEL, accel_layer_u=ACC, accel_layer_v=ACC, eta_out=THK, uhbtav=TRANSP, vhbtav=TRANSP, visc_rem_u=NONDIM,
              visc_rem_v=NONDIM, u_uh0=VEL, v_vh0=VEL, uh0=TRANSP, vh0=TRANSP, etaav=THK, _h=THK, taux_bot=STRESS, tauy_bot=STRESS, **BT_CONT)
# advect_tracer (MOM_tracer_advect.F90:53-54)
ADVECT = dict(h_end=THK, uhtr=VOL, vhtr=VOL, dt=TIME, tr=NONDIM, conc_underflow=None, advect_scheme=None, x_first_in=None, max_iter_in=None,
              vol_prev=VOL, update_vol_prev=None, uhr_out=VOL, vhr_out=VOL)
ADVECT_CS = dict(dt=TIME, default_advect_scheme=None, useHuynhStencilBug=None)
# tracer_hordiff (MOM_tracer_hor_diff.F90:119, CS :40-106)
HORDIFF = dict(h=THK, dt=TIME, tr=NONDIM, conc_underflow=None, Res_fn_h=NONDIM, Rd_dx_h=NONDIM, df_x=TRANSP, df_y=TRANSP, L2u=L2, L2v=L2,
               SN_u=(-1, 0, 0, 0), SN_v=(-1, 0, 0, 0), MEKE_Kh=L2T)
HORDIFF_CS = dict(KhTr=L2T, KhTr_min=L2T, KhTr_max=L2T, KhTr_passivity_coeff=NONDIM, KhTr_passivity_min=NONDIM, KhTr_Slope_Cff=NONDIM,
                  max_diff_CFL=NONDIM, MEKE_KhTr_fac=NONDIM)
# mixedlayer_restrat (MOM_mixed_layer_restrat.F90:149, CS :42-115)
MLE = dict(h=THK, uhtr=VOL, vhtr=VOL, T=NONDIM, S=NONDIM, ustar=(-1, 0, 0, 1), dt=TIME, h_MLD=THK, Rd_dx_h=NONDIM)
MLE_CS = dict(ml_restrat_coef=NONDIM, ml_restrat_coef2=NONDIM, front_length=(0, 1, 0, 0), MLE_MLD_decay_time=TIME, MLE_MLD_decay_time2=TIME,
              MLE_MLD_stretch=NONDIM, MLE_tail_dh=NONDIM, ustar_min=HT, vonKar=NONDIM, MLE_density_diff=NONDIM, Rho_T0_S0=None, dRho_dT=None,
              dRho_dS=None, dRho_dp=None, MLD_filtered=THK, MLD_filtered_slow=THK)
# thickness_diffuse (MOM_thickness_diffuse.F90:134, CS :40-131).  The equation of state takes the pressure in Pa (no EOS%RL2_T2_to_Pa
# in the restatement), so only the H and Z rescalings leave the interface pressures -- (g_Earth*H_to_RZ)*h -- unchanged.
THICKDIFF = dict(h=THK, uhtr=VOL, vhtr=VOL, T=NONDIM, S=NONDIM, p_surf=NONDIM, dt=TIME, Res_fn_u=NONDIM, Res_fn_v=NONDIM, uhGM=TRANSP, vhGM=TRANSP,
                 slope_x=(0, -1, 0, 1), slope_y=(0, -1, 0, 1), cg1=VEL, MEKE_Kh=L2T)
THICKDIFF_CS = dict(Khth=L2T, Khth_Min=L2T, Khth_Max=L2T, max_Khth_CFL=NONDIM, slope_max=(0, -1, 0, 1), kappa_smooth=HZT, dZ_subroundoff=ZL,
                    Rho_T0_S0=None, dRho_dT=None, dRho_dS=None, dRho_dp=None, FGNV_scale=NONDIM, N2_floor=(-2, 2, 0, -2), MEKE_KhTh_fac=NONDIM)
# PressureForce_FV_Bouss (MOM_PressureForce_FV.F90:947-1090, CS :40-107).  Pressures are [R L2 T-2]; R is not rescaled, so the EOS conversion
# factor RL2_T2_to_Pa (MOM_EOS.F90:144) carries [T2 L-2] and dRho_dp of the linear EOS is given in mks (scaled inside by dRdp_scale, :1443)
PRES = (-2, 2, 0, 0)
PGF = dict(h=THK, T=NONDIM, S=NONDIM, PFu=ACC, PFv=ACC, p_atm=PRES, pbce=(-2, 2, -1, 0), eta=THK)
PGF_CS = dict(rho_ref=NONDIM, GFS_scale=NONDIM, Z_ref=ZL, dZ_subroundoff=ZL, Rho_T0_S0=NONDIM, dRho_dT=NONDIM, dRho_dS=NONDIM, dRho_dp=NONDIM, Rlay=NONDIM,
              g_prime=(-2, 2, 0, -1), h_nonvanished=THK, kg_m3_to_R=NONDIM, RL2_T2_to_Pa=(2, -2, 0, 0), C_to_degC=NONDIM, S_to_ppt=NONDIM)
# set_dtbt (MOM_barotropic.F90:3509-3633)
SET_DTBT = dict(pbce=(-2, 2, -1, 0), gtot_est=(-2, 2, -1, 0), have_gtot_est=None, eta=THK, SSH_add=ZL, frhatu=NONDIM, frhatv=NONDIM, bathyT=ZL, bebt=NONDIM,
                G_extra=NONDIM, dtbt_fraction=NONDIM, BT_Coriolis_scale=NONDIM, Z_ref=ZL, Nonlinear_continuity=None, **BT_CONT)
# ALE_regrid, Z* (MOM_regridding.F90:846-1367): the target resolution is in Z, the thicknesses in H
REGRID = dict(h=THK, h_new=THK, dzRegrid=THK)
REGRID_CS = dict(regridding_scheme=None, nk=None, min_thickness=THK, old_grid_weight=NONDIM, depth_of_time_filter_shallow=THK, depth_of_time_filter_deep=THK,
                 Z_ref=ZL, coordinateResolution=None)


def with_flags(dims, d):
    """dims plus `None` (leave alone) for every integer switch of the control structure d."""
    out = dict(dims)
    out.update({k: None for k, v in d.items() if isinstance(v, (int, np.integer)) and not isinstance(v, bool) and k not in dims})
    return out
