// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of /root/reference/src/tracer/MOM_tracer_hor_diff.F90: tracer_hordiff :119-640 -- the online diffusivities
// (:203-340), the diffusive-CFL iteration count (:354-372) and the along-surface diffusion loop (:537-604); the neutral /
// boundary / epipycnal branches are outside the frozen option set.  Single tile: do_group_pass is the periodic wrap of
// oracle_fill_halo_2d and max_across_PEs the identity.
// PARITY: PINNED BY A REFERENCE RUN -- the reference's own MOM_tracer_hor_diff.F90, executed by oracle/f90run, agrees bit for bit on 6
// option sets (tests/test_reference_f90.py, tracer_hordiff/*); also the rotation test and the conservation / maximum-principle properties.
#include "oracle.h"
#include "ogrid.hpp"
#include <cfloat>
#include <cmath>
#include <vector>

using namespace orc;

extern "C" int oracle_tracer_hordiff(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV, const mom6cu_tracer_hor_diff_cs* CS,
                                     const mom6cu_tracer_hordiff_args* a, int* num_itts_out) {
  if (CS->use_neutral_diffusion || CS->use_hor_bnd_diffusion || CS->Diffuse_ML_interior) return 3;
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, nz = G.ke;
  const int ntr = a->ntr;
  if (num_itts_out) *num_itts_out = 0;
  if (ntr == 0 || (CS->KhTr <= 0.0 && !CS->use_variable_mixing)) return 0;  // :153
  const double dt = a->dt;
  const double Idt = 1.0 / dt;
  const double h_neglect = GV->H_subroundoff;
  const V3 h = G.H3((double*)a->h);
  std::vector<V3> Tr(ntr), dfx(ntr), dfy(ntr);
  std::vector<bool> has_dfx(ntr, false), has_dfy(ntr, false);
  for (int m = 0; m < ntr; ++m) {
    Tr[m] = G.H3(a->tr[m]);
    if (a->df_x && a->df_x[m]) { dfx[m] = G.U3(a->df_x[m]); has_dfx[m] = true; }
    if (a->df_y && a->df_y[m]) { dfy[m] = G.V3_(a->df_y[m]); has_dfy[m] = true; }
  }
  const bool use_VarMix = CS->use_variable_mixing != 0;
  const bool Resoln_scaled = use_VarMix && CS->Resoln_scaled_KhTr;
  if (Resoln_scaled && !a->Res_fn_h) return 2;
  if (use_VarMix && CS->KhTr_passivity_coeff > 0. && !a->Rd_dx_h) return 2;
  const bool use_Eady = use_VarMix && CS->KhTr_Slope_Cff > 0.;  // :167
  const bool use_MEKE = use_VarMix && CS->use_MEKE_Kh;          // allocated(MEKE%Kh), read only inside the use_VarMix branch :207-208
  if ((use_Eady && (!a->L2u || !a->SN_u || !a->L2v || !a->SN_v)) || (use_MEKE && !a->MEKE_Kh)) return 2;
  V2 L2u, SN_u, L2v, SN_v, MEKE_Kh;
  if (use_Eady) { L2u = G.U((double*)a->L2u); SN_u = G.U((double*)a->SN_u); L2v = G.V((double*)a->L2v); SN_v = G.V((double*)a->SN_v); }
  if (use_MEKE) MEKE_Kh = G.H((double*)a->MEKE_Kh);
  V2 Res_fn_h, Rd_dx_h;
  if (a->Res_fn_h) Res_fn_h = G.H((double*)a->Res_fn_h);
  if (a->Rd_dx_h) Rd_dx_h = G.H((double*)a->Rd_dx_h);
  A2 khdt_x = G.aU(), khdt_y = G.aV(), Coef_x = G.aU(), Coef_y = G.aV(), Ihdxdy = G.aH(), dTr = G.aH(), CFL = G.aH();

  if (use_VarMix) {  // :204-246
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
      const int i = I;
      double Kh_loc = CS->KhTr;
      if (use_Eady) Kh_loc = Kh_loc + CS->KhTr_Slope_Cff * L2u(I, j) * SN_u(I, j);
      if (use_MEKE) Kh_loc = Kh_loc + CS->MEKE_KhTr_fac * std::sqrt(MEKE_Kh(i, j) * MEKE_Kh(i + 1, j));
      if (CS->KhTr_max > 0.) Kh_loc = fmin2(Kh_loc, CS->KhTr_max);
      if (Resoln_scaled) Kh_loc = Kh_loc * 0.5 * (Res_fn_h(i, j) + Res_fn_h(i + 1, j));
      double Kh_u = fmax2(Kh_loc, CS->KhTr_min);
      if (CS->KhTr_passivity_coeff > 0.) {
        const double Rd_dx = 0.5 * (Rd_dx_h(i, j) + Rd_dx_h(i + 1, j));
        Kh_loc = Kh_u * fmax2(CS->KhTr_passivity_min, CS->KhTr_passivity_coeff * Rd_dx);
        if (CS->KhTr_max > 0.) Kh_loc = fmin2(Kh_loc, CS->KhTr_max);
        Kh_u = fmax2(Kh_loc, CS->KhTr_min);
      }
      khdt_x(I, j) = dt * (Kh_u * (G.dy_Cu(I, j) * G.IdxCu(I, j)));
    }
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
      const int j = J;
      double Kh_loc = CS->KhTr;
      if (use_Eady) Kh_loc = Kh_loc + CS->KhTr_Slope_Cff * L2v(i, J) * SN_v(i, J);
      if (use_MEKE) Kh_loc = Kh_loc + CS->MEKE_KhTr_fac * std::sqrt(MEKE_Kh(i, j) * MEKE_Kh(i, j + 1));
      if (CS->KhTr_max > 0.) Kh_loc = fmin2(Kh_loc, CS->KhTr_max);
      if (Resoln_scaled) Kh_loc = Kh_loc * 0.5 * (Res_fn_h(i, j) + Res_fn_h(i, j + 1));
      double Kh_v = fmax2(Kh_loc, CS->KhTr_min);
      if (CS->KhTr_passivity_coeff > 0.) {
        const double Rd_dx = 0.5 * (Rd_dx_h(i, j) + Rd_dx_h(i, j + 1));
        Kh_loc = Kh_v * fmax2(CS->KhTr_passivity_min, CS->KhTr_passivity_coeff * Rd_dx);
        if (CS->KhTr_max > 0.) Kh_loc = fmin2(Kh_loc, CS->KhTr_max);
        Kh_v = fmax2(Kh_loc, CS->KhTr_min);
      }
      khdt_y(i, J) = dt * (Kh_v * (G.dx_Cv(i, J) * G.IdyCv(i, J)));
    }
  } else {  // :266-291, a simple constant diffusivity
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) khdt_x(I, j) = dt * (CS->KhTr * (G.dy_Cu(I, j) * G.IdxCu(I, j)));
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) khdt_y(i, J) = dt * (CS->KhTr * (G.dx_Cv(i, J) * G.IdyCv(i, J)));
  }
  if (CS->max_diff_CFL > 0.0) {  // :293-327
    for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) {
      const double khdt_max = 0.125 * CS->max_diff_CFL * fmin2(G.areaT(I, j), G.areaT(I + 1, j));
      khdt_x(I, j) = fmin2(khdt_x(I, j), khdt_max);
    }
    for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
      const double khdt_max = 0.125 * CS->max_diff_CFL * fmin2(G.areaT(i, J), G.areaT(i, J + 1));
      khdt_y(i, J) = fmin2(khdt_y(i, J), khdt_max);
    }
  }
  int num_itts;
  double I_numitts;
  if (CS->check_diffusive_CFL) {  // :354-366
    double max_CFL = 0.0;
    for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) {
      CFL(i, j) = 2.0 * ((khdt_x(i - 1, j) + khdt_x(i, j)) + (khdt_y(i, j - 1) + khdt_y(i, j))) * G.IareaT(i, j);
      if (max_CFL < CFL(i, j)) max_CFL = CFL(i, j);
    }
    num_itts = std::max(1, (int)std::ceil(max_CFL - 4.0 * (DBL_EPSILON * 1.0)));  // EPSILON(max_CFL) = 2**-52
    I_numitts = 1.0 / ((double)num_itts);
  } else if (CS->max_diff_CFL > 0.0) {
    num_itts = std::max(1, (int)std::ceil(CS->max_diff_CFL - 4.0 * DBL_EPSILON));
    I_numitts = 1.0 / ((double)num_itts);
  } else { num_itts = 1; I_numitts = 1.0; }
  if (num_itts_out) *num_itts_out = num_itts;

  for (int m = 0; m < ntr; ++m) {  // :374-390
    if (has_dfx[m]) for (int k = 1; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int I = is - 1; I <= ie; ++I) dfx[m](I, j, k) = 0.0;
    if (has_dfy[m]) for (int k = 1; k <= nz; ++k) for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) dfy[m](i, J, k) = 0.0;
  }

  for (int itt = 1; itt <= num_itts; ++itt) {  // :540-603
    for (int m = 0; m < ntr; ++m) for (int k = 1; k <= nz; ++k) oracle_fill_halo_2d(d, &Tr[m](Tr[m].ilo, Tr[m].jlo, k), 0, 0);
    for (int k = 1; k <= nz; ++k) {
      const double scale = I_numitts;
      for (int J = js - 1; J <= je; ++J) for (int i = is; i <= ie; ++i) {
        const int j = J;
        Coef_y(i, J) = ((scale * khdt_y(i, J)) * 2.0 * (h(i, j, k) * h(i, j + 1, k))) / (h(i, j, k) + h(i, j + 1, k) + h_neglect);
      }
      for (int j = js; j <= je; ++j) {
        for (int I = is - 1; I <= ie; ++I) {
          const int i = I;
          Coef_x(I, j) = ((scale * khdt_x(I, j)) * 2.0 * (h(i, j, k) * h(i + 1, j, k))) / (h(i, j, k) + h(i + 1, j, k) + h_neglect);
        }
        for (int i = is; i <= ie; ++i) Ihdxdy(i, j) = G.IareaT(i, j) / (h(i, j, k) + h_neglect);
      }
      for (int m = 0; m < ntr; ++m) {
        const V3& T = Tr[m];
        for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i)
          dTr(i, j) = Ihdxdy(i, j) * (((Coef_x(i - 1, j) * (T(i - 1, j, k) - T(i, j, k))) - (Coef_x(i, j) * (T(i, j, k) - T(i + 1, j, k)))) +
                                      ((Coef_y(i, j - 1) * (T(i, j - 1, k) - T(i, j, k))) - (Coef_y(i, j) * (T(i, j, k) - T(i, j + 1, k)))));
        if (has_dfx[m]) for (int j = js; j <= je; ++j) for (int I = G.IscB; I <= G.IecB; ++I)
          dfx[m](I, j, k) = dfx[m](I, j, k) + Coef_x(I, j) * (T(I, j, k) - T(I + 1, j, k)) * Idt;
        if (has_dfy[m]) for (int J = G.JscB; J <= G.JecB; ++J) for (int i = is; i <= ie; ++i)
          dfy[m](i, J, k) = dfy[m](i, J, k) + Coef_y(i, J) * (T(i, J, k) - T(i, J + 1, k)) * Idt;
        for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i) T(i, j, k) = T(i, j, k) + dTr(i, j);
      }
    }
    for (int m = 0; m < ntr; ++m) if (a->conc_underflow && a->conc_underflow[m] > 0.0) {
      const V3& T = Tr[m];
      for (int k = 1; k <= nz; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i)
        if (std::fabs(T(i, j, k)) < a->conc_underflow[m]) T(i, j, k) = 0.0;
    }
  }
  return 0;
}
