// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of CorAdCalc and gradKE, /root/reference/src/core/MOM_CoriolisAdv.F90:125-965, :969-1051.
// Same loop nests, index ranges and parenthesisation as the Fortran; one OpenMP loop over k exactly where
// the reference has `!$OMP parallel do` (:281).  Frozen options: no OBCs, no Stokes vortex force (Waves
// absent), no AD%rv_x_u/rv_x_v diagnostics.
#include "oracle.h"
#include "ogrid.hpp"
#include <cmath>
#include <omp.h>

using namespace orc;

namespace {

// gradKE :969-1051
void gradKE(const OGrid& G, const mom6cu_coriolisadv_cs* CS, const V3& u, const V3& v, const V2& KE, const V2& KEx,
            const V2& KEy, int k) {
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  if (CS->KE_Scheme == MOM6CU_KE_ARAKAWA) {
    for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i)
      KE(i, j) = (((G.areaCu(i, j) * (u(i, j, k) * u(i, j, k))) + (G.areaCu(i - 1, j) * (u(i - 1, j, k) * u(i - 1, j, k)))) +
                  ((G.areaCv(i, j) * (v(i, j, k) * v(i, j, k))) + (G.areaCv(i, j - 1) * (v(i, j - 1, k) * v(i, j - 1, k))))) *
                 0.25 * G.IareaT(i, j);
  } else if (CS->KE_Scheme == MOM6CU_KE_SIMPLE_GUDONOV) {
    for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) {
      const double up = 0.5 * (u(i - 1, j, k) + std::fabs(u(i - 1, j, k))), up2 = up * up;
      const double um = 0.5 * (u(i, j, k) - std::fabs(u(i, j, k))), um2 = um * um;
      const double vp = 0.5 * (v(i, j - 1, k) + std::fabs(v(i, j - 1, k))), vp2 = vp * vp;
      const double vm = 0.5 * (v(i, j, k) - std::fabs(v(i, j, k))), vm2 = vm * vm;
      KE(i, j) = (fmax2(up2, um2) + fmax2(vp2, vm2)) * 0.5;
    }
  } else if (CS->KE_Scheme == MOM6CU_KE_GUDONOV) {
    for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) {
      const double up = 0.5 * (u(i - 1, j, k) + std::fabs(u(i - 1, j, k))), up2a = up * up * G.areaCu(i - 1, j);
      const double um = 0.5 * (u(i, j, k) - std::fabs(u(i, j, k))), um2a = um * um * G.areaCu(i, j);
      const double vp = 0.5 * (v(i, j - 1, k) + std::fabs(v(i, j - 1, k))), vp2a = vp * vp * G.areaCv(i, j - 1);
      const double vm = 0.5 * (v(i, j, k) - std::fabs(v(i, j, k))), vm2a = vm * vm * G.areaCv(i, j);
      KE(i, j) = (fmax2(um2a, up2a) + fmax2(vm2a, vp2a)) * 0.5 * G.IareaT(i, j);
    }
  }
  for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) KEx(I, j) = (KE(I + 1, j) - KE(I, j)) * G.IdxCu(I, j);
  for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) KEy(i, J) = (KE(i, J + 1) - KE(i, J)) * G.IdyCv(i, J);
}

}  // namespace

extern "C" int oracle_coradcalc(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV,
                                const mom6cu_unit_scale* US, const mom6cu_coriolisadv_cs* CS,
                                const mom6cu_coradcalc_args* A, int nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  const int nz = G.ke;
  const V3 u = G.U3(A->u), v = G.V3_(A->v), h = G.H3(A->h), uh = G.U3(A->uh), vh = G.V3_(A->vh);
  const V3 CAu = G.U3(A->CAu), CAv = G.V3_(A->CAv);
  V3 porU, porV, RV, PV, gKEu, gKEv;
  if (A->por_face_areaU) porU = G.U3(A->por_face_areaU);
  if (A->por_face_areaV) porV = G.V3_(A->por_face_areaV);
  if (A->RV) RV = G.Q3(A->RV);
  if (A->PV) PV = G.Q3(A->PV);
  if (A->gradKEu) gKEu = G.U3(A->gradKEu);
  if (A->gradKEv) gKEv = G.V3_(A->gradKEv);
  const double m_to_L = US ? US->m_to_L : 1.0, m_s_to_L_T = US ? US->m_s_to_L_T : 1.0;
  const double C1_12 = 1.0 / 12.0, C1_24 = 1.0 / 24.0;
  const double vol_neglect = GV->H_subroundoff * ((1e-4 * m_to_L) * (1e-4 * m_to_L));  // :241
  const double eps_vel = 1.0e-10 * m_s_to_L_T;
  const double h_tiny = GV->Angstrom_H;

  A2 Area_h = G.aH(), Area_q = G.aQ();
  for (int j = Jsq - 1; j <= Jeq + 2; ++j) for (int i = Isq - 1; i <= Ieq + 2; ++i)
    Area_h(i, j) = G.mask2dT(i, j) * G.areaT(i, j);
  for (int J = Jsq - 1; J <= Jeq + 1; ++J) for (int I = Isq - 1; I <= Ieq + 1; ++I)
    Area_q(I, J) = (Area_h(I, J) + Area_h(I + 1, J + 1)) + (Area_h(I + 1, J) + Area_h(I, J + 1));

#pragma omp parallel for
  for (int k = 1; k <= nz; ++k) {
    A2 q = G.aQ(), Ih_q = G.aQ(), dvdx = G.aQ(), dudy = G.aQ(), rel_vort = G.aQ(), abs_vort = G.aQ();
    A2 a = G.aU(), b = G.aU(), c = G.aU(), dd = G.aU(), hArea_u = G.aU(), KEx = G.aU(), uh_center = G.aU();
    A2 hArea_v = G.aV(), KEy = G.aV(), vh_center = G.aV();
    A2 KE = G.aH(), uh_min = G.aH(), uh_max = G.aH(), vh_min = G.aH(), vh_max = G.aH(), ep_u = G.aH(), ep_v = G.aH();

    // :314-324
    for (int J = Jsq - 1; J <= Jeq + 1; ++J) for (int I = Isq - 1; I <= Ieq + 1; ++I) {
      dvdx(I, J) = (v(I + 1, J, k) * G.dyCv(I + 1, J)) - (v(I, J, k) * G.dyCv(I, J));
      dudy(I, J) = (u(I, J + 1, k) * G.dxCu(I, J + 1)) - (u(I, J, k) * G.dxCu(I, J));
    }
    for (int J = Jsq - 1; J <= Jeq + 1; ++J) for (int i = Isq - 1; i <= Ieq + 2; ++i)
      hArea_v(i, J) = 0.5 * ((Area_h(i, J) * h(i, J, k)) + (Area_h(i, J + 1) * h(i, J + 1, k)));
    for (int j = Jsq - 1; j <= Jeq + 2; ++j) for (int I = Isq - 1; I <= Ieq + 1; ++I)
      hArea_u(I, j) = 0.5 * ((Area_h(I, j) * h(I, j, k)) + (Area_h(I + 1, j) * h(I + 1, j, k)));

    if (CS->Coriolis_En_Dis) {  // :326-333
      for (int j = Jsq; j <= Jeq + 1; ++j) for (int I = is - 1; I <= ie; ++I) {
        const double por = A->por_face_areaU ? porU(I, j, k) : 1.0;
        uh_center(I, j) = 0.5 * ((G.dy_Cu(I, j) * por) * u(I, j, k)) * (h(I, j, k) + h(I + 1, j, k));
      }
      for (int J = js - 1; J <= je; ++J) for (int i = Isq; i <= Ieq + 1; ++i) {
        const double por = A->por_face_areaV ? porV(i, J, k) : 1.0;
        vh_center(i, J) = 0.5 * ((G.dx_Cv(i, J) * por) * v(i, J, k)) * (h(i, J, k) + h(i, J + 1, k));
      }
    }

    // :459-491
    if (CS->no_slip) {
      for (int J = Jsq - 1; J <= Jeq + 1; ++J) for (int I = Isq - 1; I <= Ieq + 1; ++I)
        rel_vort(I, J) = (2.0 - G.mask2dBu(I, J)) * (dvdx(I, J) - dudy(I, J)) * G.IareaBu(I, J);
    } else {
      for (int J = Jsq - 1; J <= Jeq + 1; ++J) for (int I = Isq - 1; I <= Ieq + 1; ++I)
        rel_vort(I, J) = G.mask2dBu(I, J) * (dvdx(I, J) - dudy(I, J)) * G.IareaBu(I, J);
    }
    for (int J = Jsq - 1; J <= Jeq + 1; ++J) for (int I = Isq - 1; I <= Ieq + 1; ++I)
      abs_vort(I, J) = G.CoriolisBu(I, J) + rel_vort(I, J);
    for (int J = Jsq - 1; J <= Jeq + 1; ++J) for (int I = Isq - 1; I <= Ieq + 1; ++I) {
      const double hArea_q = (hArea_u(I, J) + hArea_u(I, J + 1)) + (hArea_v(I, J) + hArea_v(I + 1, J));
      Ih_q(I, J) = Area_q(I, J) / (hArea_q + vol_neglect);
      q(I, J) = abs_vort(I, J) * Ih_q(I, J);
    }
    if (A->RV) for (int J = Jsq - 1; J <= Jeq + 1; ++J) for (int I = Isq - 1; I <= Ieq + 1; ++I) RV(I, J, k) = rel_vort(I, J);
    if (A->PV) for (int J = Jsq - 1; J <= Jeq + 1; ++J) for (int I = Isq - 1; I <= Ieq + 1; ++I) PV(I, J, k) = q(I, J);

    // :523-588
    if (CS->Coriolis_Scheme == MOM6CU_ARAKAWA_HSU90) {
      for (int j = Jsq; j <= Jeq + 1; ++j) {
        for (int I = is - 1; I <= Ieq; ++I) {
          a(I, j) = (q(I, j) + (q(I + 1, j) + q(I, j - 1))) * C1_12;
          dd(I, j) = ((q(I, j) + q(I + 1, j - 1)) + q(I, j - 1)) * C1_12;
        }
        for (int I = Isq; I <= Ieq; ++I) {
          b(I, j) = (q(I, j) + (q(I - 1, j) + q(I, j - 1))) * C1_12;
          c(I, j) = ((q(I, j) + q(I - 1, j - 1)) + q(I, j - 1)) * C1_12;
        }
      }
    } else if (CS->Coriolis_Scheme == MOM6CU_ARAKAWA_LAMB81) {
      for (int j = Jsq; j <= Jeq + 1; ++j) for (int I = Isq; I <= Ieq + 1; ++I) {
        a(I - 1, j) = (2.0 * (q(I, j) + q(I - 1, j - 1)) + (q(I - 1, j) + q(I, j - 1))) * C1_24;
        dd(I - 1, j) = ((q(I, j) + q(I - 1, j - 1)) + 2.0 * (q(I - 1, j) + q(I, j - 1))) * C1_24;
        b(I, j) = ((q(I, j) + q(I - 1, j - 1)) + 2.0 * (q(I - 1, j) + q(I, j - 1))) * C1_24;
        c(I, j) = (2.0 * (q(I, j) + q(I - 1, j - 1)) + (q(I - 1, j) + q(I, j - 1))) * C1_24;
        ep_u(I, j) = ((q(I, j) - q(I - 1, j - 1)) + (q(I - 1, j) - q(I, j - 1))) * C1_24;
        ep_v(I, j) = (-(q(I, j) - q(I - 1, j - 1)) + (q(I - 1, j) - q(I, j - 1))) * C1_24;
      }
    } else if (CS->Coriolis_Scheme == MOM6CU_AL_BLEND) {
      double Fe_m2 = CS->F_eff_max_blend - 2.0;
      double rat_lin = 1.5 * Fe_m2 / fmax2(CS->wt_lin_blend, 1.0e-16);
      if (CS->F_eff_max_blend <= 2.0) { Fe_m2 = -1.; rat_lin = -1.0; }
      for (int j = Jsq; j <= Jeq + 1; ++j) for (int I = Isq; I <= Ieq + 1; ++I) {
        const double min_Ihq = min4(Ih_q(I - 1, j - 1), Ih_q(I, j - 1), Ih_q(I - 1, j), Ih_q(I, j));
        const double max_Ihq = max4(Ih_q(I - 1, j - 1), Ih_q(I, j - 1), Ih_q(I - 1, j), Ih_q(I, j));
        double rat_m1 = 1.0e15;
        if (max_Ihq < 1.0e15 * min_Ihq) rat_m1 = max_Ihq / min_Ihq - 1.0;
        double AL_wt, Sad_wt;
        if (rat_m1 <= Fe_m2) AL_wt = 1.0;
        else if (rat_m1 < 1.5 * Fe_m2) AL_wt = 3.0 * Fe_m2 / rat_m1 - 2.0;
        else AL_wt = 0.0;
        if (rat_m1 <= 1.5 * Fe_m2) Sad_wt = 0.0;
        else if (rat_m1 <= rat_lin) Sad_wt = 1.0 - (1.5 * Fe_m2) / rat_m1;
        else if (rat_m1 < 2.0 * rat_lin) Sad_wt = 1.0 - (CS->wt_lin_blend / rat_lin) * (rat_m1 - 2.0 * rat_lin);
        else Sad_wt = 1.0;
        a(I - 1, j) = Sad_wt * 0.25 * q(I - 1, j) + (1.0 - Sad_wt) *
                      (((2.0 - AL_wt) * q(I - 1, j) + AL_wt * q(I, j - 1)) + 2.0 * (q(I, j) + q(I - 1, j - 1))) * C1_24;
        dd(I - 1, j) = Sad_wt * 0.25 * q(I - 1, j - 1) + (1.0 - Sad_wt) *
                       (((2.0 - AL_wt) * q(I - 1, j - 1) + AL_wt * q(I, j)) + 2.0 * (q(I - 1, j) + q(I, j - 1))) * C1_24;
        b(I, j) = Sad_wt * 0.25 * q(I, j) + (1.0 - Sad_wt) *
                  (((2.0 - AL_wt) * q(I, j) + AL_wt * q(I - 1, j - 1)) + 2.0 * (q(I - 1, j) + q(I, j - 1))) * C1_24;
        c(I, j) = Sad_wt * 0.25 * q(I, j - 1) + (1.0 - Sad_wt) *
                  (((2.0 - AL_wt) * q(I, j - 1) + AL_wt * q(I - 1, j)) + 2.0 * (q(I, j) + q(I - 1, j - 1))) * C1_24;
        ep_u(I, j) = AL_wt * ((q(I, j) - q(I - 1, j - 1)) + (q(I - 1, j) - q(I, j - 1))) * C1_24;
        ep_v(I, j) = AL_wt * (-(q(I, j) - q(I - 1, j - 1)) + (q(I - 1, j) - q(I, j - 1))) * C1_24;
      }
    }

    if (CS->Coriolis_En_Dis) {  // :590-636
      const double c1 = 1.0 - 1.5 * 0.5, c2 = 1.0 - 0.5, c3 = 2.0, slope = 0.5;
      for (int j = Jsq; j <= Jeq + 1; ++j) for (int I = is - 1; I <= ie; ++I) {
        double uhc = uh_center(I, j), uhm = uh(I, j, k);
        if (G.dy_Cu(I, j) == 0.0) uhc = uhm;
        if (std::fabs(uhc) < 0.1 * std::fabs(uhm)) uhm = 10.0 * uhc;
        else if (std::fabs(uhc) > c1 * std::fabs(uhm)) {
          if (std::fabs(uhc) < c2 * std::fabs(uhm)) uhc = (3.0 * uhc + (1.0 - c2 * 3.0) * uhm);
          else if (std::fabs(uhc) <= c3 * std::fabs(uhm)) uhc = uhm;
          else uhc = slope * uhc + (1.0 - c3 * slope) * uhm;
        }
        if (uhc > uhm) { uh_min(I, j) = uhm; uh_max(I, j) = uhc; }
        else { uh_max(I, j) = uhm; uh_min(I, j) = uhc; }
      }
      for (int J = js - 1; J <= je; ++J) for (int i = Isq; i <= Ieq + 1; ++i) {
        double vhc = vh_center(i, J), vhm = vh(i, J, k);
        if (G.dx_Cv(i, J) == 0.0) vhc = vhm;
        if (std::fabs(vhc) < 0.1 * std::fabs(vhm)) vhm = 10.0 * vhc;
        else if (std::fabs(vhc) > c1 * std::fabs(vhm)) {
          if (std::fabs(vhc) < c2 * std::fabs(vhm)) vhc = (3.0 * vhc + (1.0 - c2 * 3.0) * vhm);
          else if (std::fabs(vhc) <= c3 * std::fabs(vhm)) vhc = vhm;
          else vhc = slope * vhc + (1.0 - c3 * slope) * vhm;
        }
        if (vhc > vhm) { vh_min(i, J) = vhm; vh_max(i, J) = vhc; }
        else { vh_max(i, J) = vhm; vh_min(i, J) = vhc; }
      }
    }

    gradKE(G, CS, u, v, KE, KEx, KEy, k);  // :639

    // ---- zonal acceleration :644-758
    if (CS->Coriolis_Scheme == MOM6CU_SADOURNY75_ENERGY) {
      if (CS->Coriolis_En_Dis) {
        for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) {
          double temp1, temp2;
          if (q(I, j) * u(I, j, k) == 0.0)
            temp1 = q(I, j) * ((vh_max(I, j) + vh_max(I + 1, j)) + (vh_min(I, j) + vh_min(I + 1, j))) * 0.5;
          else if (q(I, j) * u(I, j, k) < 0.0) temp1 = q(I, j) * (vh_max(I, j) + vh_max(I + 1, j));
          else temp1 = q(I, j) * (vh_min(I, j) + vh_min(I + 1, j));
          if (q(I, j - 1) * u(I, j, k) == 0.0)
            temp2 = q(I, j - 1) * ((vh_max(I, j - 1) + vh_max(I + 1, j - 1)) + (vh_min(I, j - 1) + vh_min(I + 1, j - 1))) * 0.5;
          else if (q(I, j - 1) * u(I, j, k) < 0.0) temp2 = q(I, j - 1) * (vh_max(I, j - 1) + vh_max(I + 1, j - 1));
          else temp2 = q(I, j - 1) * (vh_min(I, j - 1) + vh_min(I + 1, j - 1));
          CAu(I, j, k) = 0.25 * G.IdxCu(I, j) * (temp1 + temp2);
        }
      } else {
        for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I)
          CAu(I, j, k) = 0.25 * ((q(I, j) * (vh(I + 1, j, k) + vh(I, j, k))) +
                                 (q(I, j - 1) * (vh(I, j - 1, k) + vh(I + 1, j - 1, k)))) * G.IdxCu(I, j);
      }
    } else if (CS->Coriolis_Scheme == MOM6CU_SADOURNY75_ENSTRO) {
      for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I)
        CAu(I, j, k) = 0.125 * (G.IdxCu(I, j) * (q(I, j) + q(I, j - 1))) *
                       ((vh(I + 1, j, k) + vh(I, j, k)) + (vh(I, j - 1, k) + vh(I + 1, j - 1, k)));
    } else if (CS->Coriolis_Scheme == MOM6CU_ARAKAWA_HSU90 || CS->Coriolis_Scheme == MOM6CU_ARAKAWA_LAMB81 ||
               CS->Coriolis_Scheme == MOM6CU_AL_BLEND) {
      for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I)
        CAu(I, j, k) = (((a(I, j) * vh(I + 1, j, k)) + (c(I, j) * vh(I, j - 1, k))) +
                        ((b(I, j) * vh(I, j, k)) + (dd(I, j) * vh(I + 1, j - 1, k)))) * G.IdxCu(I, j);
    } else if (CS->Coriolis_Scheme == MOM6CU_ROBUST_ENSTRO) {
      for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) {
        const int i = I;
        double Heff1 = std::fabs(vh(i, j, k) * G.IdxCv(i, j)) / (eps_vel + std::fabs(v(i, j, k)));
        Heff1 = fmax2(Heff1, fmin2(h(i, j, k), h(i, j + 1, k)));
        Heff1 = fmin2(Heff1, fmax2(h(i, j, k), h(i, j + 1, k)));
        double Heff2 = std::fabs(vh(i, j - 1, k) * G.IdxCv(i, j - 1)) / (eps_vel + std::fabs(v(i, j - 1, k)));
        Heff2 = fmax2(Heff2, fmin2(h(i, j - 1, k), h(i, j, k)));
        Heff2 = fmin2(Heff2, fmax2(h(i, j - 1, k), h(i, j, k)));
        double Heff3 = std::fabs(vh(i + 1, j, k) * G.IdxCv(i + 1, j)) / (eps_vel + std::fabs(v(i + 1, j, k)));
        Heff3 = fmax2(Heff3, fmin2(h(i + 1, j, k), h(i + 1, j + 1, k)));
        Heff3 = fmin2(Heff3, fmax2(h(i + 1, j, k), h(i + 1, j + 1, k)));
        double Heff4 = std::fabs(vh(i + 1, j - 1, k) * G.IdxCv(i + 1, j - 1)) / (eps_vel + std::fabs(v(i + 1, j - 1, k)));
        Heff4 = fmax2(Heff4, fmin2(h(i + 1, j - 1, k), h(i + 1, j, k)));
        Heff4 = fmin2(Heff4, fmax2(h(i + 1, j - 1, k), h(i + 1, j, k)));
        if (CS->PV_Adv_Scheme == MOM6CU_PV_ADV_CENTERED) {
          CAu(I, j, k) = 0.5 * (abs_vort(I, j) + abs_vort(I, j - 1)) *
                         ((vh(i, j, k) + vh(i + 1, j - 1, k)) + (vh(i, j - 1, k) + vh(i + 1, j, k))) /
                         (h_tiny + ((Heff1 + Heff4) + (Heff2 + Heff3))) * G.IdxCu(I, j);
        } else if (CS->PV_Adv_Scheme == MOM6CU_PV_ADV_UPWIND1) {
          const double VHeff = ((vh(i, j, k) + vh(i + 1, j - 1, k)) + (vh(i, j - 1, k) + vh(i + 1, j, k)));
          const double QVHeff = 0.5 * (((abs_vort(I, j) + abs_vort(I, j - 1)) * VHeff) -
                                       ((abs_vort(I, j) - abs_vort(I, j - 1)) * std::fabs(VHeff)));
          CAu(I, j, k) = (QVHeff / (h_tiny + ((Heff1 + Heff4) + (Heff2 + Heff3)))) * G.IdxCu(I, j);
        }
      }
    }
    if (CS->Coriolis_Scheme == MOM6CU_ARAKAWA_LAMB81 || CS->Coriolis_Scheme == MOM6CU_AL_BLEND)
      for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I)
        CAu(I, j, k) = CAu(I, j, k) + ((ep_u(I, j) * uh(I - 1, j, k)) - (ep_u(I + 1, j) * uh(I + 1, j, k))) * G.IdxCu(I, j);
    if (CS->bound_Coriolis)
      for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) {
        const double fv1 = abs_vort(I, j) * v(I + 1, j, k), fv2 = abs_vort(I, j) * v(I, j, k);
        const double fv3 = abs_vort(I, j - 1) * v(I + 1, j - 1, k), fv4 = abs_vort(I, j - 1) * v(I, j - 1, k);
        const double max_fv = max4(fv1, fv2, fv3, fv4), min_fv = min4(fv1, fv2, fv3, fv4);
        CAu(I, j, k) = fmin2(CAu(I, j, k), max_fv);
        CAu(I, j, k) = fmax2(CAu(I, j, k), min_fv);
      }
    for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) CAu(I, j, k) = CAu(I, j, k) - KEx(I, j);
    if (A->gradKEu) for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I) gKEu(I, j, k) = -KEx(I, j);

    // ---- meridional acceleration :763-881
    if (CS->Coriolis_Scheme == MOM6CU_SADOURNY75_ENERGY) {
      if (CS->Coriolis_En_Dis) {
        for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) {
          double temp1, temp2;
          if (q(i - 1, J) * v(i, J, k) == 0.0)
            temp1 = q(i - 1, J) * ((uh_max(i - 1, J) + uh_max(i - 1, J + 1)) + (uh_min(i - 1, J) + uh_min(i - 1, J + 1))) * 0.5;
          else if (q(i - 1, J) * v(i, J, k) > 0.0) temp1 = q(i - 1, J) * (uh_max(i - 1, J) + uh_max(i - 1, J + 1));
          else temp1 = q(i - 1, J) * (uh_min(i - 1, J) + uh_min(i - 1, J + 1));
          if (q(i, J) * v(i, J, k) == 0.0)
            temp2 = q(i, J) * ((uh_max(i, J) + uh_max(i, J + 1)) + (uh_min(i, J) + uh_min(i, J + 1))) * 0.5;
          else if (q(i, J) * v(i, J, k) > 0.0) temp2 = q(i, J) * (uh_max(i, J) + uh_max(i, J + 1));
          else temp2 = q(i, J) * (uh_min(i, J) + uh_min(i, J + 1));
          CAv(i, J, k) = -0.25 * G.IdyCv(i, J) * (temp1 + temp2);
        }
      } else {
        for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i)
          CAv(i, J, k) = -0.25 * ((q(i - 1, J) * (uh(i - 1, J, k) + uh(i - 1, J + 1, k))) +
                                  (q(i, J) * (uh(i, J, k) + uh(i, J + 1, k)))) * G.IdyCv(i, J);
      }
    } else if (CS->Coriolis_Scheme == MOM6CU_SADOURNY75_ENSTRO) {
      for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i)
        CAv(i, J, k) = -0.125 * (G.IdyCv(i, J) * (q(i - 1, J) + q(i, J))) *
                       ((uh(i - 1, J, k) + uh(i - 1, J + 1, k)) + (uh(i, J, k) + uh(i, J + 1, k)));
    } else if (CS->Coriolis_Scheme == MOM6CU_ARAKAWA_HSU90 || CS->Coriolis_Scheme == MOM6CU_ARAKAWA_LAMB81 ||
               CS->Coriolis_Scheme == MOM6CU_AL_BLEND) {
      for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i)
        CAv(i, J, k) = -(((a(i - 1, J) * uh(i - 1, J, k)) + (c(i, J + 1) * uh(i, J + 1, k))) +
                         ((b(i, J) * uh(i, J, k)) + (dd(i - 1, J + 1) * uh(i - 1, J + 1, k)))) * G.IdyCv(i, J);
    } else if (CS->Coriolis_Scheme == MOM6CU_ROBUST_ENSTRO) {
      for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) {
        const int I = i, j = J;
        double Heff1 = std::fabs(uh(I, j, k) * G.IdyCu(I, j)) / (eps_vel + std::fabs(u(I, j, k)));
        Heff1 = fmax2(Heff1, fmin2(h(i, j, k), h(i + 1, j, k)));
        Heff1 = fmin2(Heff1, fmax2(h(i, j, k), h(i + 1, j, k)));
        double Heff2 = std::fabs(uh(I - 1, j, k) * G.IdyCu(I - 1, j)) / (eps_vel + std::fabs(u(I - 1, j, k)));
        Heff2 = fmax2(Heff2, fmin2(h(i - 1, j, k), h(i, j, k)));
        Heff2 = fmin2(Heff2, fmax2(h(i - 1, j, k), h(i, j, k)));
        double Heff3 = std::fabs(uh(I, j + 1, k) * G.IdyCu(I, j + 1)) / (eps_vel + std::fabs(u(I, j + 1, k)));
        Heff3 = fmax2(Heff3, fmin2(h(i, j + 1, k), h(i + 1, j + 1, k)));
        Heff3 = fmin2(Heff3, fmax2(h(i, j + 1, k), h(i + 1, j + 1, k)));
        double Heff4 = std::fabs(uh(I - 1, j + 1, k) * G.IdyCu(I - 1, j + 1)) / (eps_vel + std::fabs(u(I - 1, j + 1, k)));
        Heff4 = fmax2(Heff4, fmin2(h(i - 1, j + 1, k), h(i, j + 1, k)));
        Heff4 = fmin2(Heff4, fmax2(h(i - 1, j + 1, k), h(i, j + 1, k)));
        if (CS->PV_Adv_Scheme == MOM6CU_PV_ADV_CENTERED) {
          CAv(i, J, k) = -0.5 * (abs_vort(I, J) + abs_vort(I - 1, J)) *
                         ((uh(I, j, k) + uh(I - 1, j + 1, k)) + (uh(I - 1, j, k) + uh(I, j + 1, k))) /
                         (h_tiny + ((Heff1 + Heff4) + (Heff2 + Heff3))) * G.IdyCv(i, J);
        } else if (CS->PV_Adv_Scheme == MOM6CU_PV_ADV_UPWIND1) {
          const double UHeff = ((uh(I, j, k) + uh(I - 1, j + 1, k)) + (uh(I - 1, j, k) + uh(I, j + 1, k)));
          const double QUHeff = 0.5 * (((abs_vort(I, J) + abs_vort(I - 1, J)) * UHeff) -
                                       ((abs_vort(I, J) - abs_vort(I - 1, J)) * std::fabs(UHeff)));
          CAv(i, J, k) = -QUHeff / (h_tiny + ((Heff1 + Heff4) + (Heff2 + Heff3))) * G.IdyCv(i, J);
        }
      }
    }
    if (CS->Coriolis_Scheme == MOM6CU_ARAKAWA_LAMB81 || CS->Coriolis_Scheme == MOM6CU_AL_BLEND)
      for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i)
        CAv(i, J, k) = CAv(i, J, k) + ((ep_v(i, J) * vh(i, J - 1, k)) - (ep_v(i, J + 1) * vh(i, J + 1, k))) * G.IdyCv(i, J);
    if (CS->bound_Coriolis)
      for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) {
        const double fu1 = -abs_vort(i, J) * u(i, J + 1, k), fu2 = -abs_vort(i, J) * u(i, J, k);
        const double fu3 = -abs_vort(i - 1, J) * u(i - 1, J + 1, k), fu4 = -abs_vort(i - 1, J) * u(i - 1, J, k);
        const double max_fu = max4(fu1, fu2, fu3, fu4), min_fu = min4(fu1, fu2, fu3, fu4);
        CAv(i, J, k) = fmin2(CAv(i, J, k), max_fu);
        CAv(i, J, k) = fmax2(CAv(i, J, k), min_fu);
      }
    for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) CAv(i, J, k) = CAv(i, J, k) - KEy(i, J);
    if (A->gradKEv) for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i) gKEv(i, J, k) = -KEy(i, J);
  }
  return 0;
}
