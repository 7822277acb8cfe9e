// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of /root/reference/src/parameterizations/lateral/MOM_thickness_diffuse.F90: thickness_diffuse :134-632 (the diffusivities
// :226-443, the transports applied to uhtr, vhtr and h :600-616) and thickness_diffuse_full :635-1670 (the density-gradient path with an
// equation of state: available volumes and pressures :862-883, slopes and the unlimited streamfunction :913-1100, the limited transports
// :1124-1176, the surface boundary condition :1532-1536; the v-direction twins :1236-1529); with find_eta_3d
// (src/core/MOM_interface_heights.F90:48-112, Boussinesq), vert_fill_TS (src/core/MOM_isopycnal_slopes.F90:612-700),
// thickness_to_dz (Boussinesq: dz = H_to_Z*h) and calculate_density_derivs of EOS_WRIGHT (MOM_EOS_Wright.F90:178-206) / EOS_LINEAR.
// PARITY: PINNED BY A REFERENCE RUN -- the reference's own MOM_thickness_diffuse.F90 (+ vert_fill_TS, the EOS derivative routines), executed
// by oracle/f90run, agrees bit for bit on 6 option sets (tests/test_reference_f90.py, thickness_diffuse/*).
#include "oracle.h"
#include "ogrid.hpp"
#include <cmath>
#include <vector>

using namespace orc;

namespace {
const double a0 = 7.057924e-4, a1 = 3.480336e-7, a2 = -1.112733e-7;
const double b0 = 5.790749e8, b1 = 3.516535e6, b2 = -4.002714e4, b3 = 2.084372e2, b4 = 5.944068e5, b5 = -9.643486e3;
const double c0 = 1.704853e5, c1 = 7.904722e2, c2 = -7.984422, c3 = 5.140652e-2, c4 = -2.302158e2, c5 = -3.079464;

// calculate_density_derivs_elem_buggy_Wright :178-206 / calculate_density_derivs_elem_linear
inline void density_derivs(const mom6cu_thickness_diffuse_cs* E, double T, double S, double pressure, double& drho_dT, double& drho_dS) {
  if (E->EOS_form == MOM6CU_EOS_LINEAR) { drho_dT = E->dRho_dT; drho_dS = E->dRho_dS; return; }
  const double al0 = (a0 + a1 * T) + a2 * S;
  const double p0 = (b0 + b4 * S) + T * (b1 + T * ((b2 + b3 * T)) + b5 * S);
  const double lambda = (c0 + c4 * S) + T * (c1 + T * ((c2 + c3 * T)) + c5 * S);
  double I_denom2 = 1.0 / (lambda + al0 * (pressure + p0));
  I_denom2 = I_denom2 * I_denom2;
  drho_dT = I_denom2 * (lambda * (b1 + T * (2.0 * b2 + 3.0 * b3 * T) + b5 * S) -
                        (pressure + p0) * ((pressure + p0) * a1 + (c1 + T * (c2 * 2.0 + c3 * 3.0 * T) + c5 * S)));
  drho_dS = I_denom2 * (lambda * (b4 + b5 * T) - (pressure + p0) * ((pressure + p0) * a2 + (c4 + c5 * T)));
}
}  // namespace

extern "C" int oracle_thickness_diffuse(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV, const mom6cu_unit_scale* US,
                                        const mom6cu_thickness_diffuse_cs* CS, const mom6cu_thickness_diffuse_args* a) {
  if (CS->read_khth || CS->detangle_interfaces || CS->interface_Kh || CS->use_stanley_gm || CS->use_GME_thickness_diffuse ||
      CS->find_work || CS->Depth_scaled_KhTh || CS->use_Visbeck || CS->use_QG_Leith_GM || CS->khth_struct || !GV->Boussinesq)
    return 3;
  if ((CS->use_stored_slopes && (!a->slope_x || !a->slope_y)) || (CS->use_MEKE_Kh && !a->MEKE_Kh)) return 2;
  if (CS->use_FGNV_streamfn && !a->cg1) return 2;  // "cg1 must be associated when using FGNV streamfunction."  :858
  if (CS->EOS_form != MOM6CU_EOS_LINEAR && CS->EOS_form != MOM6CU_EOS_WRIGHT) return 3;
  if (!CS->thickness_diffuse || !(CS->Khth > 0.0 || CS->use_variable_mixing)) return 0;  // :196-198
  if (d->nk < 2 || !(a->dt > 0.0)) return 2;
  const bool Resoln_scaled = CS->use_variable_mixing && CS->Resoln_scaled_KhTh;
  if (Resoln_scaled && (!a->Res_fn_u || !a->Res_fn_v)) return 2;
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, nz = G.ke;
  const double dt = a->dt;
  const V3 h = G.H3(a->h), uhtr = G.U3(a->uhtr), vhtr = G.V3_(a->vhtr), T_in = G.H3((double*)a->T), S_in = G.H3((double*)a->S);
  V2 p_surf, Res_fn_u, Res_fn_v;
  if (a->p_surf) p_surf = G.H((double*)a->p_surf);
  if (Resoln_scaled) { Res_fn_u = G.U((double*)a->Res_fn_u); Res_fn_v = G.V((double*)a->Res_fn_v); }
  const bool present_slope = CS->use_stored_slopes != 0, FGNV = CS->use_FGNV_streamfn != 0;
  V3 slope_x, slope_y;
  V2 cg1, MEKE_Kh;
  if (present_slope) { slope_x = G.U3((double*)a->slope_x, nz + 1); slope_y = G.V3_((double*)a->slope_y, nz + 1); }
  if (FGNV) cg1 = G.H((double*)a->cg1);
  if (CS->use_MEKE_Kh) MEKE_Kh = G.H((double*)a->MEKE_Kh);
  V3 uhGM, vhGM;
  if (a->uhGM) uhGM = G.U3(a->uhGM);
  if (a->vhGM) vhGM = G.V3_(a->vhGM);
  A3 e(G.isd, G.ied, G.jsd, G.jed, nz + 1), uhD(G.isd - 1, G.ied, G.jsd, G.jed, nz), vhD(G.isd, G.ied, G.jsd - 1, G.jed, nz);
  A3 KH_u(G.isd - 1, G.ied, G.jsd, G.jed, nz + 1), KH_v(G.isd, G.ied, G.jsd - 1, G.jed, nz + 1);
  A3 int_slope_u(G.isd - 1, G.ied, G.jsd, G.jed, nz + 1), int_slope_v(G.isd, G.ied, G.jsd - 1, G.jed, nz + 1);
  A2 KH_u_CFL = G.aU(), KH_v_CFL = G.aV(), Khth_loc_u = G.aU(), Khth_loc_v = G.aV();

  // ---- thickness_diffuse :226-443
  for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I)
    KH_u_CFL(I, j) = (0.25 * CS->max_Khth_CFL) / (dt * ((G.IdxCu(I, j) * G.IdxCu(I, j)) + (G.IdyCu(I, j) * G.IdyCu(I, j))));
  for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i)
    KH_v_CFL(i, J) = (0.25 * CS->max_Khth_CFL) / (dt * ((G.IdxCv(i, J) * G.IdxCv(i, J)) + (G.IdyCv(i, J) * G.IdyCv(i, J))));
  // find_eta(h, tv, G, GV, US, e, halo_size=1)
  for (int j = js - 1; j <= je + 1; ++j) for (int i = is - 1; i <= ie + 1; ++i) e(i, j, nz + 1) = -(G.bathyT(i, j) + 0.0);
  for (int j = js - 1; j <= je + 1; ++j) for (int k = nz; k >= 1; --k) for (int i = is - 1; i <= ie + 1; ++i)
    e(i, j, k) = e(i, j, k + 1) + h(i, j, k) * GV->H_to_Z;
  for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
    Khth_loc_u(I, j) = CS->Khth;
    if (CS->use_MEKE_Kh) Khth_loc_u(I, j) = Khth_loc_u(I, j) + CS->MEKE_KhTh_fac * std::sqrt(MEKE_Kh(I, j) * MEKE_Kh(I + 1, j));  // :281-284
    if (Resoln_scaled) Khth_loc_u(I, j) = Khth_loc_u(I, j) * Res_fn_u(I, j);
    if (CS->Khth_Max > 0) Khth_loc_u(I, j) = fmax2(CS->Khth_Min, fmin2(Khth_loc_u(I, j), CS->Khth_Max));
    else Khth_loc_u(I, j) = fmax2(CS->Khth_Min, Khth_loc_u(I, j));
    KH_u(I, j, 1) = fmin2(KH_u_CFL(I, j), Khth_loc_u(I, j));
    for (int K = 2; K <= nz + 1; ++K) KH_u(I, j, K) = KH_u(I, j, 1);
  }
  for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
    Khth_loc_v(i, J) = CS->Khth;
    if (CS->use_MEKE_Kh) Khth_loc_v(i, J) = Khth_loc_v(i, J) + CS->MEKE_KhTh_fac * std::sqrt(MEKE_Kh(i, J) * MEKE_Kh(i, J + 1));  // :381-384
    if (Resoln_scaled) Khth_loc_v(i, J) = Khth_loc_v(i, J) * Res_fn_v(i, J);
    if (CS->Khth_Max > 0) Khth_loc_v(i, J) = fmax2(CS->Khth_Min, fmin2(Khth_loc_v(i, J), CS->Khth_Max));
    else Khth_loc_v(i, J) = fmax2(CS->Khth_Min, Khth_loc_v(i, J));
    KH_v(i, J, 1) = fmin2(KH_v_CFL(i, J), Khth_loc_v(i, J));
    for (int K = 2; K <= nz + 1; ++K) KH_v(i, J, K) = KH_v(i, J, 1);
  }
  // int_slope_u = int_slope_v = 0 :470-473 (A3 is zero-initialised)

  // ---- thickness_diffuse_full :635-1670
  const double I4dt = 0.25 / dt;
  const double I_slope_max2 = 1.0 / (CS->slope_max * CS->slope_max);
  const double h_neglect = GV->H_subroundoff, h_neglect2 = h_neglect * h_neglect;
  const double dz_neglect = CS->dZ_subroundoff;
  const double G_rho0 = GV->g_Earth / GV->Rho0;
  const int nk_linear = 1;  // max(GV%nkml, 1)
  A3 T(G.isd, G.ied, G.jsd, G.jed, nz), S(G.isd, G.ied, G.jsd, G.jed, nz), h_avail(G.isd, G.ied, G.jsd, G.jed, nz), h_frac(G.isd, G.ied, G.jsd, G.jed, nz);
  A3 h_avail_rsum(G.isd, G.ied, G.jsd, G.jed, nz + 1), pres(G.isd, G.ied, G.jsd, G.jed, nz + 1);
  {  // vert_fill_TS(h, tv%T, tv%S, CS%kappa_smooth*dt, T, S, G, GV, US, halo=1, larger_h_denom=.true.)
    const int is1 = is - 1, ie1 = ie + 1, js1 = js - 1, je1 = je + 1;
    const double kappa_dt = CS->kappa_smooth * dt;
    const double kap_dt_x2 = (2.0 * kappa_dt) * (US->Z_to_m * GV->m_to_H);
    double h0 = h_neglect;
    h0 = 1.0e-16 * std::sqrt(0.5 * kap_dt_x2);
    if (kap_dt_x2 <= 0.0) {
      for (int k = 1; k <= nz; ++k) for (int j = js1; j <= je1; ++j) for (int i = is1; i <= ie1; ++i) { T(i, j, k) = T_in(i, j, k); S(i, j, k) = S_in(i, j, k); }
    } else {
      std::vector<double> ent((size_t)(G.ied + 2) * (nz + 2)), b1v(G.ied + 2), d1v(G.ied + 2), c1v((size_t)(G.ied + 2) * (nz + 1));
      auto ENT = [&](int i, int K) -> double& { return ent[(size_t)K * (G.ied + 2) + i]; };
      auto C1 = [&](int i, int k) -> double& { return c1v[(size_t)k * (G.ied + 2) + i]; };
      for (int j = js1; j <= je1; ++j) {
        for (int i = is1; i <= ie1; ++i) {
          ENT(i, 2) = kap_dt_x2 / ((h(i, j, 1) + h(i, j, 2)) + h0);
          const double h_tr = h(i, j, 1) + h_neglect;
          b1v[i] = 1.0 / (h_tr + ENT(i, 2));
          d1v[i] = b1v[i] * h_tr;
          T(i, j, 1) = (b1v[i] * h_tr) * T_in(i, j, 1);
          S(i, j, 1) = (b1v[i] * h_tr) * S_in(i, j, 1);
        }
        for (int k = 2; k <= nz - 1; ++k) for (int i = is1; i <= ie1; ++i) {
          ENT(i, k + 1) = kap_dt_x2 / ((h(i, j, k) + h(i, j, k + 1)) + h0);
          const double h_tr = h(i, j, k) + h_neglect;
          C1(i, k) = ENT(i, k) * b1v[i];
          b1v[i] = 1.0 / ((h_tr + d1v[i] * ENT(i, k)) + ENT(i, k + 1));
          d1v[i] = b1v[i] * (h_tr + d1v[i] * ENT(i, k));
          T(i, j, k) = b1v[i] * (h_tr * T_in(i, j, k) + ENT(i, k) * T(i, j, k - 1));
          S(i, j, k) = b1v[i] * (h_tr * S_in(i, j, k) + ENT(i, k) * S(i, j, k - 1));
        }
        for (int i = is1; i <= ie1; ++i) {
          C1(i, nz) = ENT(i, nz) * b1v[i];
          const double h_tr = h(i, j, nz) + h_neglect;
          b1v[i] = 1.0 / (h_tr + d1v[i] * ENT(i, nz));
          T(i, j, nz) = b1v[i] * (h_tr * T_in(i, j, nz) + ENT(i, nz) * T(i, j, nz - 1));
          S(i, j, nz) = b1v[i] * (h_tr * S_in(i, j, nz) + ENT(i, nz) * S(i, j, nz - 1));
        }
        for (int k = nz - 1; k >= 1; --k) for (int i = is1; i <= ie1; ++i) {
          T(i, j, k) = T(i, j, k) + C1(i, k + 1) * T(i, j, k + 1);
          S(i, j, k) = S(i, j, k) + C1(i, k + 1) * S(i, j, k + 1);
        }
      }
    }
  }
  // thickness_to_dz (Boussinesq): dz = GV%H_to_Z*h -- only used by the FGNV / non-Boussinesq branches
  for (int j = js - 1; j <= je + 1; ++j) for (int i = is - 1; i <= ie + 1; ++i) {  // :864-883
    h_avail_rsum(i, j, 1) = 0.0;
    pres(i, j, 1) = 0.0;
    if (a->p_surf) pres(i, j, 1) = p_surf(i, j);
    h_avail(i, j, 1) = fmax2(I4dt * G.areaT(i, j) * (h(i, j, 1) - GV->Angstrom_H), 0.0);
    h_avail_rsum(i, j, 2) = h_avail(i, j, 1);
    h_frac(i, j, 1) = 1.0;
    pres(i, j, 2) = pres(i, j, 1) + (GV->g_Earth * GV->H_to_RZ) * h(i, j, 1);
  }
  for (int j = js - 1; j <= je + 1; ++j) for (int k = 2; k <= nz; ++k) for (int i = is - 1; i <= ie + 1; ++i) {
    h_avail(i, j, k) = fmax2(I4dt * G.areaT(i, j) * (h(i, j, k) - GV->Angstrom_H), 0.0);
    h_avail_rsum(i, j, k + 1) = h_avail_rsum(i, j, k) + h_avail(i, j, k);
    h_frac(i, j, k) = 0.0;
    if (h_avail(i, j, k) > 0.0) h_frac(i, j, k) = h_avail(i, j, k) / h_avail_rsum(i, j, k + 1);
    pres(i, j, k + 1) = pres(i, j, k) + (GV->g_Earth * GV->H_to_RZ) * h(i, j, k);
  }
  A2 uhtot = G.aU(), vhtot = G.aV();
  std::vector<double> Sfn_unlim(nz + 2), slope2_Ratio(nz + 2), dzN2(nz + 2), c2_dz(nz + 2), c1s(nz + 2);
  const double dz_neglect2 = dz_neglect * dz_neglect, N2_floor = CS->N2_floor;
  // streamfn_solver :1674-1707 on Sfn_unlim(1:nz+1) with c2_h = c2_dz(1:nz), hN2 = dzN2(1:nz+1)
  auto streamfn_solver = [&]() {
    Sfn_unlim[1] = 0.;
    double b_denom = dzN2[2] + c2_dz[1];
    double beta = 1.0 / (b_denom + c2_dz[2]);
    double d1 = beta * b_denom;
    Sfn_unlim[2] = (beta * dzN2[2]) * Sfn_unlim[2];
    for (int K = 3; K <= nz; ++K) {
      c1s[K - 1] = beta * c2_dz[K - 1];
      b_denom = dzN2[K] + d1 * c2_dz[K - 1];
      beta = 1.0 / (b_denom + c2_dz[K]);
      d1 = beta * b_denom;
      Sfn_unlim[K] = beta * (dzN2[K] * Sfn_unlim[K] + c2_dz[K - 1] * Sfn_unlim[K - 1]);
    }
    c1s[nz] = beta * c2_dz[nz];
    Sfn_unlim[nz + 1] = 0.;
    for (int K = nz; K >= 2; --K) Sfn_unlim[K] = Sfn_unlim[K] + c1s[K] * Sfn_unlim[K + 1];
  };

  // ---- zonal fluxes :913-1229
  for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
    const int i = I;
    for (int K = nz; K >= 2; --K) {
      const int k = K;
      double drdiA = 0., drdiB = 0., drdkL = 0., drdkR = 0.;
      if ((k >= nk_linear) && (!present_slope || FGNV)) {  // calc_derivatives :921-922
        const double pres_u = 0.5 * (pres(i, j, K) + pres(i + 1, j, K));
        const double T_u = 0.25 * ((T(i, j, k) + T(i + 1, j, k)) + (T(i, j, k - 1) + T(i + 1, j, k - 1)));
        const double S_u = 0.25 * ((S(i, j, k) + S(i + 1, j, k)) + (S(i, j, k - 1) + S(i + 1, j, k - 1)));
        double drho_dT_u, drho_dS_u;
        density_derivs(CS, T_u, S_u, pres_u, drho_dT_u, drho_dS_u);
        drdiA = drho_dT_u * (T(i + 1, j, k - 1) - T(i, j, k - 1)) + drho_dS_u * (S(i + 1, j, k - 1) - S(i, j, k - 1));
        drdiB = drho_dT_u * (T(i + 1, j, k) - T(i, j, k)) + drho_dS_u * (S(i + 1, j, k) - S(i, j, k));
        drdkL = (drho_dT_u * (T(i, j, k) - T(i, j, k - 1)) + drho_dS_u * (S(i, j, k) - S(i, j, k - 1)));
        drdkR = (drho_dT_u * (T(i + 1, j, k) - T(i + 1, j, k - 1)) + drho_dS_u * (S(i + 1, j, k) - S(i + 1, j, k - 1)));
      }
      if (k > nk_linear) {
        double drdz = 0., hg2A = 0., hg2B = 0., haA = 0., haB = 0.;
        if (FGNV || !present_slope) {  // :980-1022
          const double hg2L = h(i, j, k - 1) * h(i, j, k) + h_neglect2;
          const double hg2R = h(i + 1, j, k - 1) * h(i + 1, j, k) + h_neglect2;
          const double haL = 0.5 * (h(i, j, k - 1) + h(i, j, k)) + h_neglect;
          const double haR = 0.5 * (h(i + 1, j, k - 1) + h(i + 1, j, k)) + h_neglect;
          const double dzaL = haL * GV->H_to_Z, dzaR = haR * GV->H_to_Z;
          const double wtL = hg2L * (haR * dzaR), wtR = hg2R * (haL * dzaL);
          drdz = ((wtL * drdkL) + (wtR * drdkR)) / ((dzaL * wtL) + (dzaR * wtR));
          hg2A = h(i, j, k - 1) * h(i + 1, j, k - 1) + h_neglect2;
          hg2B = h(i, j, k) * h(i + 1, j, k) + h_neglect2;
          haA = 0.5 * (h(i, j, k - 1) + h(i + 1, j, k - 1)) + h_neglect;
          haB = 0.5 * (h(i, j, k) + h(i + 1, j, k)) + h_neglect;
          const double N2_unlim = drdz * G_rho0;
          const double dzL1 = GV->H_to_Z * h(i, j, k - 1), dzR1 = GV->H_to_Z * h(i + 1, j, k - 1);  // thickness_to_dz (Boussinesq)
          const double dzL0 = GV->H_to_Z * h(i, j, k), dzR0 = GV->H_to_Z * h(i + 1, j, k);
          const double dzg2A = dzL1 * dzR1 + dz_neglect2, dzg2B = dzL0 * dzR0 + dz_neglect2;
          const double dzaA = 0.5 * (dzL1 + dzR1) + dz_neglect, dzaB = 0.5 * (dzL0 + dzR0) + dz_neglect;
          dzN2[K] = (0.5 * (dzg2A / dzaA + dzg2B / dzaB)) * fmax2(N2_unlim, N2_floor);
        }
        double Slope;
        if (present_slope) {
          Slope = slope_x(I, j, k);
          slope2_Ratio[K] = (Slope * Slope) * I_slope_max2;
        } else {
          const double wtA = hg2A * haB, wtB = hg2B * haA;
          const double drdx = ((wtA * drdiA + wtB * drdiB) / (wtA + wtB) - drdz * (e(i, j, K) - e(i + 1, j, K))) * G.IdxCu(I, j);
          const double mag_grad2 = (US->Z_to_L * drdx) * (US->Z_to_L * drdx) + drdz * drdz;
          if (mag_grad2 > 0.0) {
            Slope = drdx / std::sqrt(mag_grad2);
            slope2_Ratio[K] = (Slope * Slope) * I_slope_max2;
          } else {
            Slope = 0.0;
            slope2_Ratio[K] = 1.0e20;
          }
        }
        Slope = (1.0 - int_slope_u(I, j, K)) * Slope + int_slope_u(I, j, K) * ((e(i + 1, j, K) - e(i, j, K)) * G.IdxCu(I, j));
        slope2_Ratio[K] = (1.0 - int_slope_u(I, j, K)) * slope2_Ratio[K];
        Sfn_unlim[K] = -(KH_u(I, j, K) * G.dy_Cu(I, j)) * Slope;
        if (Sfn_unlim[K] > 0.0) {
          if (e(i, j, K) < e(i + 1, j, nz + 1)) Sfn_unlim[K] = 0.0;
          else if (e(i + 1, j, nz + 1) > e(i, j, K + 1))
            Sfn_unlim[K] = Sfn_unlim[K] * ((e(i, j, K) - e(i + 1, j, nz + 1)) / ((e(i, j, K) - e(i, j, K + 1)) + dz_neglect));
        } else {
          if (e(i + 1, j, K) < e(i, j, nz + 1)) Sfn_unlim[K] = 0.0;
          else if (e(i, j, nz + 1) > e(i + 1, j, K + 1))
            Sfn_unlim[K] = Sfn_unlim[K] * ((e(i + 1, j, K) - e(i, j, nz + 1)) / ((e(i + 1, j, K) - e(i + 1, j, K + 1)) + dz_neglect));
        }
      } else {
        dzN2[K] = N2_floor * dz_neglect;
        Sfn_unlim[K] = 0.;
      }
    }
    if (FGNV) {  // :1103-1122
      dzN2[1] = 0.; dzN2[nz + 1] = 0.;
      if (G.mask2dCu(I, j) > 0.) {
        for (int k = 1; k <= nz; ++k) {
          const double dzL = GV->H_to_Z * h(i, j, k), dzR = GV->H_to_Z * h(i + 1, j, k);
          const double dz_harm = fmax2(dz_neglect, 2. * dzL * dzR / ((dzL + dzR) + dz_neglect));
          const double cg = 0.5 * (cg1(i, j) + cg1(i + 1, j));
          c2_dz[k] = CS->FGNV_scale * (cg * cg) / dz_harm;
        }
        for (int K = 2; K <= nz; ++K) Sfn_unlim[K] = (1. + CS->FGNV_scale) * Sfn_unlim[K];
        streamfn_solver();
      } else {
        for (int K = 2; K <= nz; ++K) Sfn_unlim[K] = 0.;
      }
    }
    uhtot(I, j) = 0.0;
    for (int K = nz; K >= 2; --K) {  // :1124-1176
      const int k = K;
      const double Z_to_H = GV->Z_to_H;
      if (k > nk_linear) {
        double Sfn_safe;
        if (uhtot(I, j) <= 0.0) Sfn_safe = uhtot(I, j) * (1.0 - h_frac(i, j, k));
        else Sfn_safe = uhtot(I, j) * (1.0 - h_frac(i + 1, j, k));
        const double Sfn_est = (Z_to_H * Sfn_unlim[K] + slope2_Ratio[K] * Sfn_safe) / (1.0 + slope2_Ratio[K]);
        const double Sfn_in_H = fmin2(fmax2(Sfn_est, -h_avail_rsum(i, j, K)), h_avail_rsum(i + 1, j, K));
        uhD(I, j, k) = fmax2(fmin2((Sfn_in_H - uhtot(I, j)), h_avail(i, j, k)), -h_avail(i + 1, j, k));
      } else {
        if (uhtot(I, j) <= 0.0) uhD(I, j, k) = -uhtot(I, j) * h_frac(i, j, k);
        else uhD(I, j, k) = -uhtot(I, j) * h_frac(i + 1, j, k);
      }
      uhtot(I, j) = uhtot(I, j) + uhD(I, j, k);
    }
  }
  // ---- meridional fluxes :1236-1529
  for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
    const int j = J;
    for (int K = nz; K >= 2; --K) {
      const int k = K;
      double drdjA = 0., drdjB = 0., drdkL = 0., drdkR = 0.;
      if ((k >= nk_linear) && (!present_slope || FGNV)) {  // calc_derivatives :1243-1244
        const double pres_v = 0.5 * (pres(i, j, K) + pres(i, j + 1, K));
        const double T_v = 0.25 * ((T(i, j, k) + T(i, j + 1, k)) + (T(i, j, k - 1) + T(i, j + 1, k - 1)));
        const double S_v = 0.25 * ((S(i, j, k) + S(i, j + 1, k)) + (S(i, j, k - 1) + S(i, j + 1, k - 1)));
        double drho_dT_v, drho_dS_v;
        density_derivs(CS, T_v, S_v, pres_v, drho_dT_v, drho_dS_v);
        drdjA = drho_dT_v * (T(i, j + 1, k - 1) - T(i, j, k - 1)) + drho_dS_v * (S(i, j + 1, k - 1) - S(i, j, k - 1));
        drdjB = drho_dT_v * (T(i, j + 1, k) - T(i, j, k)) + drho_dS_v * (S(i, j + 1, k) - S(i, j, k));
        drdkL = (drho_dT_v * (T(i, j, k) - T(i, j, k - 1)) + drho_dS_v * (S(i, j, k) - S(i, j, k - 1)));
        drdkR = (drho_dT_v * (T(i, j + 1, k) - T(i, j + 1, k - 1)) + drho_dS_v * (S(i, j + 1, k) - S(i, j + 1, k - 1)));
      }
      if (k > nk_linear) {
        double drdz = 0., hg2A = 0., hg2B = 0., haA = 0., haB = 0.;
        if (FGNV || !present_slope) {  // :1300-1344
          const double hg2L = h(i, j, k - 1) * h(i, j, k) + h_neglect2;
          const double hg2R = h(i, j + 1, k - 1) * h(i, j + 1, k) + h_neglect2;
          const double haL = 0.5 * (h(i, j, k - 1) + h(i, j, k)) + h_neglect;
          const double haR = 0.5 * (h(i, j + 1, k - 1) + h(i, j + 1, k)) + h_neglect;
          const double dzaL = haL * GV->H_to_Z, dzaR = haR * GV->H_to_Z;
          const double wtL = hg2L * (haR * dzaR), wtR = hg2R * (haL * dzaL);
          drdz = ((wtL * drdkL) + (wtR * drdkR)) / ((dzaL * wtL) + (dzaR * wtR));
          hg2A = h(i, j, k - 1) * h(i, j + 1, k - 1) + h_neglect2;
          hg2B = h(i, j, k) * h(i, j + 1, k) + h_neglect2;
          haA = 0.5 * (h(i, j, k - 1) + h(i, j + 1, k - 1)) + h_neglect;
          haB = 0.5 * (h(i, j, k) + h(i, j + 1, k)) + h_neglect;
          const double N2_unlim = drdz * G_rho0;
          const double dzL1 = GV->H_to_Z * h(i, j, k - 1), dzR1 = GV->H_to_Z * h(i, j + 1, k - 1);
          const double dzL0 = GV->H_to_Z * h(i, j, k), dzR0 = GV->H_to_Z * h(i, j + 1, k);
          const double dzg2A = dzL1 * dzR1 + dz_neglect2, dzg2B = dzL0 * dzR0 + dz_neglect2;
          const double dzaA = 0.5 * (dzL1 + dzR1) + dz_neglect, dzaB = 0.5 * (dzL0 + dzR0) + dz_neglect;
          dzN2[K] = (0.5 * (dzg2A / dzaA + dzg2B / dzaB)) * fmax2(N2_unlim, N2_floor);
        }
        double Slope;
        if (present_slope) {
          Slope = slope_y(i, J, k);
          slope2_Ratio[K] = (Slope * Slope) * I_slope_max2;
        } else {
          const double wtA = hg2A * haB, wtB = hg2B * haA;
          const double drdy = ((wtA * drdjA + wtB * drdjB) / (wtA + wtB) - drdz * (e(i, j, K) - e(i, j + 1, K))) * G.IdyCv(i, J);
          const double mag_grad2 = (US->Z_to_L * drdy) * (US->Z_to_L * drdy) + drdz * drdz;
          if (mag_grad2 > 0.0) {
            Slope = drdy / std::sqrt(mag_grad2);
            slope2_Ratio[K] = (Slope * Slope) * I_slope_max2;
          } else {
            Slope = 0.0;
            slope2_Ratio[K] = 1.0e20;
          }
        }
        Slope = (1.0 - int_slope_v(i, J, K)) * Slope + int_slope_v(i, J, K) * ((e(i, j + 1, K) - e(i, j, K)) * G.IdyCv(i, J));
        slope2_Ratio[K] = (1.0 - int_slope_v(i, J, K)) * slope2_Ratio[K];
        Sfn_unlim[K] = -((KH_v(i, J, K) * G.dx_Cv(i, J)) * Slope);
        if (Sfn_unlim[K] > 0.0) {
          if (e(i, j, K) < e(i, j + 1, nz + 1)) Sfn_unlim[K] = 0.0;
          else if (e(i, j + 1, nz + 1) > e(i, j, K + 1))
            Sfn_unlim[K] = Sfn_unlim[K] * ((e(i, j, K) - e(i, j + 1, nz + 1)) / ((e(i, j, K) - e(i, j, K + 1)) + dz_neglect));
        } else {
          if (e(i, j + 1, K) < e(i, j, nz + 1)) Sfn_unlim[K] = 0.0;
          else if (e(i, j, nz + 1) > e(i, j + 1, K + 1))
            Sfn_unlim[K] = Sfn_unlim[K] * ((e(i, j + 1, K) - e(i, j, nz + 1)) / ((e(i, j + 1, K) - e(i, j + 1, K + 1)) + dz_neglect));
        }
      } else {
        dzN2[K] = N2_floor * dz_neglect;
        Sfn_unlim[K] = 0.;
      }
    }
    if (FGNV) {  // :1423-1442
      dzN2[1] = 0.; dzN2[nz + 1] = 0.;
      if (G.mask2dCv(i, J) > 0.) {
        for (int k = 1; k <= nz; ++k) {
          const double dzL = GV->H_to_Z * h(i, j, k), dzR = GV->H_to_Z * h(i, j + 1, k);
          const double dz_harm = fmax2(dz_neglect, 2. * dzL * dzR / ((dzL + dzR) + dz_neglect));
          const double cg = 0.5 * (cg1(i, j) + cg1(i, j + 1));
          c2_dz[k] = CS->FGNV_scale * (cg * cg) / dz_harm;
        }
        for (int K = 2; K <= nz; ++K) Sfn_unlim[K] = (1. + CS->FGNV_scale) * Sfn_unlim[K];
        streamfn_solver();
      } else {
        for (int K = 2; K <= nz; ++K) Sfn_unlim[K] = 0.;
      }
    }
    vhtot(i, J) = 0.0;
    for (int K = nz; K >= 2; --K) {
      const int k = K;
      const double Z_to_H = GV->Z_to_H;
      if (k > nk_linear) {
        double Sfn_safe;
        if (vhtot(i, J) <= 0.0) Sfn_safe = vhtot(i, J) * (1.0 - h_frac(i, j, k));
        else Sfn_safe = vhtot(i, J) * (1.0 - h_frac(i, j + 1, k));
        const double Sfn_est = (Z_to_H * Sfn_unlim[K] + slope2_Ratio[K] * Sfn_safe) / (1.0 + slope2_Ratio[K]);
        const double Sfn_in_H = fmin2(fmax2(Sfn_est, -h_avail_rsum(i, j, K)), h_avail_rsum(i, j + 1, K));
        vhD(i, J, k) = fmax2(fmin2((Sfn_in_H - vhtot(i, J)), h_avail(i, j, k)), -h_avail(i, j + 1, k));
      } else {
        if (vhtot(i, J) <= 0.0) vhD(i, J, k) = -vhtot(i, J) * h_frac(i, j, k);
        else vhD(i, J, k) = -vhtot(i, J) * h_frac(i, j + 1, k);
      }
      vhtot(i, J) = vhtot(i, J) + vhD(i, J, k);
    }
  }
  // In layer 1, enforce the boundary conditions that Sfn(z=0) = 0.0  :1532-1536
  for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) uhD(I, j, 1) = -uhtot(I, j);
  for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) vhD(i, J, 1) = -vhtot(i, J);

  // ---- thickness_diffuse :600-616
  for (int k = 1; k <= nz; ++k) {
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
      uhtr(I, j, k) = uhtr(I, j, k) + uhD(I, j, k) * dt;
      if (a->uhGM) uhGM(I, j, k) = uhD(I, j, k);
    }
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
      vhtr(i, J, k) = vhtr(i, J, k) + vhD(i, J, k) * dt;
      if (a->vhGM) vhGM(i, J, k) = vhD(i, J, k);
    }
    for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      h(i, j, k) = h(i, j, k) - dt * G.IareaT(i, j) * ((uhD(i, j, k) - uhD(i - 1, j, k)) + (vhD(i, j, k) - vhD(i, j - 1, k)));
      if (h(i, j, k) < GV->Angstrom_H) h(i, j, k) = GV->Angstrom_H;
    }
  }
  return 0;
}

// The calculate_density_derivs restatement this file uses, for the known-answer test of the equation of state.
extern "C" void oracle_thickdiff_eos_derivs(int form, const double* lin4, double T, double S, double p, double* drho_dT, double* drho_dS) {
  mom6cu_thickness_diffuse_cs E = {};
  E.EOS_form = form;
  if (lin4) { E.dRho_dT = lin4[1]; E.dRho_dS = lin4[2]; }
  density_derivs(&E, T, S, p, *drho_dT, *drho_dS);
}
