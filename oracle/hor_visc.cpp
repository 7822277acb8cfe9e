// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/oracle.h).
// Restatement of horizontal_viscosity, /root/reference/src/parameterizations/lateral/MOM_hor_visc.F90:266-2317,
// for the frozen option set: Laplacian and/or biharmonic, constant/MICOM background + Smagorinsky
// coefficients, BOUND_KH/BOUND_AH, BETTER_BOUND_KH/BETTER_BOUND_AH, BOUND_CORIOLIS, RE_AH, no-slip / free-slip,
// USE_LAND_MASK, USE_CONT_THICKNESS.  Not restated (rejected by the product): Leith/Leith+E, GME, MEKE
// viscosities/backscatter, anisotropic viscosity, ZB2020, resolution-function scaling, OBCs, FrictWork diagnostics.
// Same loop nests, ranges and parenthesisation as the Fortran; OpenMP over k where the reference has it (:669).
#include "oracle.h"
#include "ogrid.hpp"
#include <cmath>
#include <omp.h>

using namespace orc;

extern "C" int oracle_horizontal_viscosity(const mom6cu_domain* d, const mom6cu_grid* Gp, const mom6cu_vgrid* GV,
                                           const mom6cu_hor_visc_cs* CS, const mom6cu_hor_visc_args* A, int nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  if (CS->unsupported) return 3;  // Leith / GME / MEKE / anisotropic / ZB2020 / resolution-scaled viscosities, OBCs: not restated
  const OGrid G(d, Gp);
  const int is = G.isc, ie = G.iec, js = G.jsc, je = G.jec, Isq = G.IscB, Ieq = G.IecB, Jsq = G.JscB, Jeq = G.JecB;
  const int nz = G.ke;
  if (!(CS->Laplacian || CS->biharmonic)) return 0;  // :507
  const V3 u = G.U3(A->u), v = G.V3_(A->v), h = G.H3(A->h), diffu = G.U3(A->diffu), diffv = G.V3_(A->diffv);
  const bool use_cont_huv = CS->use_cont_thick && A->hu_cont && A->hv_cont;  // :523
  V3 hu_cont, hv_cont;
  if (use_cont_huv) { hu_cont = G.U3(A->hu_cont); hv_cont = G.V3_(A->hv_cont); }
  // control-structure arrays
  const V2 dx2h = G.H(CS->dx2h), dy2h = G.H(CS->dy2h), DX_dyT = G.H(CS->DX_dyT), DY_dxT = G.H(CS->DY_dxT),
           reduction_xx = G.H(CS->reduction_xx), Kh_bg_xx = G.H(CS->Kh_bg_xx), Ah_bg_xx = G.H(CS->Ah_bg_xx),
           Kh_Max_xx = G.H(CS->Kh_Max_xx), Ah_Max_xx = G.H(CS->Ah_Max_xx), Laplac2_const_xx = G.H(CS->Laplac2_const_xx),
           Biharm_const_xx = G.H(CS->Biharm_const_xx), Biharm_const2_xx = G.H(CS->Biharm_const2_xx),
           Re_Ah_const_xx = G.H(CS->Re_Ah_const_xx);
  const V2 dx2q = G.Q(CS->dx2q), dy2q = G.Q(CS->dy2q), DX_dyBu = G.Q(CS->DX_dyBu), DY_dxBu = G.Q(CS->DY_dxBu),
           reduction_xy = G.Q(CS->reduction_xy), Kh_bg_xy = G.Q(CS->Kh_bg_xy), Ah_bg_xy = G.Q(CS->Ah_bg_xy),
           Kh_Max_xy = G.Q(CS->Kh_Max_xy), Ah_Max_xy = G.Q(CS->Ah_Max_xy), Laplac2_const_xy = G.Q(CS->Laplac2_const_xy),
           Biharm_const_xy = G.Q(CS->Biharm_const_xy), Biharm_const2_xy = G.Q(CS->Biharm_const2_xy),
           Re_Ah_const_xy = G.Q(CS->Re_Ah_const_xy);
  const V2 Idx2dyCu = G.U(CS->Idx2dyCu), Idxdy2u = G.U(CS->Idxdy2u), Idx2dyCv = G.V(CS->Idx2dyCv), Idxdy2v = G.V(CS->Idxdy2v);

  const double h_neglect = GV->H_subroundoff;
  const double h_neglect3 = h_neglect * h_neglect * h_neglect;
  // :541-556 (no Leith): halo sizes of the thickness-point viscosities and of the vorticity-point strains
  const int js_Kh = Jsq, je_Kh = je + 1, is_Kh = Isq, ie_Kh = ie + 1;
  const int js_vort = js - 2, je_vort = Jeq + 1, is_vort = is - 2, ie_vort = Ieq + 1;
  const bool legacy_bound = (CS->Smagorinsky_Kh) && (CS->bound_Kh && !CS->better_bound_Kh);  // :557

#pragma omp parallel for
  for (int k = 1; k <= nz; ++k) {
    A2 dudx = G.aH(), dvdy = G.aH(), sh_xx = G.aH(), str_xx = G.aH();
    A2 dvdx = G.aQ(), dudy = G.aQ(), sh_xy = G.aQ(), str_xy = G.aQ(), hq = G.aQ(), dDel2vdx = G.aQ(), dDel2udy = G.aQ();
    A2 Ah = G.aQ(), Kh = G.aQ(), Shear_mag = G.aQ(), hrat_min = G.aQ(), visc_bound_rem = G.aQ();
    A2 Del2u = G.aU(), h_u = G.aU(), Del2v = G.aV(), h_v = G.aV();

    // Calculate horizontal tension :720-726
    for (int j = Jsq - 1; j <= Jeq + 2; ++j) for (int i = Isq - 1; i <= Ieq + 2; ++i) {
      dudx(i, j) = DY_dxT(i, j) * ((G.IdyCu(i, j) * u(i, j, k)) - (G.IdyCu(i - 1, j) * u(i - 1, j, k)));
      dvdy(i, j) = DX_dyT(i, j) * ((G.IdxCv(i, j) * v(i, j, k)) - (G.IdxCv(i, j - 1) * v(i, j - 1, k)));
      sh_xx(i, j) = dudx(i, j) - dvdy(i, j);
    }
    // Components for the shearing strain :729-732
    for (int J = js_vort; J <= je_vort; ++J) for (int I = is_vort; I <= ie_vort; ++I) {
      dvdx(I, J) = DY_dxBu(I, J) * ((v(I + 1, J, k) * G.IdyCv(I + 1, J)) - (v(I, J, k) * G.IdyCv(I, J)));
      dudy(I, J) = DX_dyBu(I, J) * ((u(I, J + 1, k) * G.IdxCu(I, J + 1)) - (u(I, J, k) * G.IdxCu(I, J)));
    }
    // Interpolate the thicknesses to velocity points :764-785
    if (use_cont_huv) {
      for (int j = js - 2; j <= je + 2; ++j) for (int I = Isq - 1; I <= Ieq + 1; ++I) h_u(I, j) = hu_cont(I, j, k);
      for (int J = Jsq - 1; J <= Jeq + 1; ++J) for (int i = is - 2; i <= ie + 2; ++i) h_v(i, J) = hv_cont(i, J, k);
    } else if (CS->use_land_mask) {
      for (int j = js - 2; j <= je + 2; ++j) for (int I = is - 2; I <= Ieq + 1; ++I)
        h_u(I, j) = 0.5 * (G.mask2dT(I, j) * h(I, j, k) + G.mask2dT(I + 1, j) * h(I + 1, j, k));
      for (int J = js - 2; J <= Jeq + 1; ++J) for (int i = is - 2; i <= ie + 2; ++i)
        h_v(i, J) = 0.5 * (G.mask2dT(i, J) * h(i, J, k) + G.mask2dT(i, J + 1) * h(i, J + 1, k));
    } else {
      for (int j = js - 2; j <= je + 2; ++j) for (int I = is - 2; I <= Ieq + 1; ++I)
        h_u(I, j) = 0.5 * (h(I, j, k) + h(I + 1, j, k));
      for (int J = js - 2; J <= Jeq + 1; ++J) for (int i = is - 2; i <= ie + 2; ++i)
        h_v(i, J) = 0.5 * (h(i, J, k) + h(i, J + 1, k));
    }
    // Shearing strain :909-919
    if (CS->no_slip) {
      for (int J = js - 2; J <= Jeq + 1; ++J) for (int I = is - 2; I <= Ieq + 1; ++I)
        sh_xy(I, J) = (2.0 - G.mask2dBu(I, J)) * (dvdx(I, J) + dudy(I, J));
    } else {
      for (int J = js - 2; J <= Jeq + 1; ++J) for (int I = is - 2; I <= Ieq + 1; ++I)
        sh_xy(I, J) = G.mask2dBu(I, J) * (dvdx(I, J) + dudy(I, J));
    }
    // Del2u, Del2v :936-944
    if (CS->biharmonic) {
      for (int j = js - 1; j <= Jeq + 1; ++j) for (int I = Isq - 1; I <= Ieq + 1; ++I)
        Del2u(I, j) = Idx2dyCu(I, j) * ((dx2q(I, j) * sh_xy(I, j)) - (dx2q(I, j - 1) * sh_xy(I, j - 1))) +
                      Idxdy2u(I, j) * ((dy2h(I + 1, j) * sh_xx(I + 1, j)) - (dy2h(I, j) * sh_xx(I, j)));
      for (int J = Jsq - 1; J <= Jeq + 1; ++J) for (int i = is - 1; i <= Ieq + 1; ++i)
        Del2v(i, J) = Idxdy2v(i, J) * ((dy2q(i, J) * sh_xy(i, J)) - (dy2q(i - 1, J) * sh_xy(i - 1, J))) -
                      Idx2dyCv(i, J) * ((dx2h(i, J + 1) * sh_xx(i, J + 1)) - (dx2h(i, J) * sh_xx(i, J)));
    }
    // :1114-1122
    if (CS->Smagorinsky_Kh || CS->Smagorinsky_Ah) {
      for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) {
        const double sh_xx_sq = sh_xx(i, j) * sh_xx(i, j);
        const double sh_xy_sq = 0.25 * (((sh_xy(i - 1, j - 1) * sh_xy(i - 1, j - 1)) + (sh_xy(i, j) * sh_xy(i, j))) +
                                        ((sh_xy(i - 1, j) * sh_xy(i - 1, j)) + (sh_xy(i, j - 1) * sh_xy(i, j - 1))));
        Shear_mag(i, j) = std::sqrt(sh_xx_sq + sh_xy_sq);
      }
    }
    // :1124-1129
    if (CS->better_bound_Ah || CS->better_bound_Kh) {
      for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) {
        const double h_min = min4(h_u(i, j), h_u(i - 1, j), h_v(i, j), h_v(i, j - 1));
        hrat_min(i, j) = fmin2(1.0, h_min / (h(i, j, k) + h_neglect));
      }
    }
    if (CS->Laplacian) {  // :1131-1278
      for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) Kh(i, j) = Kh_bg_xx(i, j);
      for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) {
        if (CS->add_LES_viscosity) {
          if (CS->Smagorinsky_Kh) Kh(i, j) = Kh(i, j) + Laplac2_const_xx(i, j) * Shear_mag(i, j);
        } else {
          if (CS->Smagorinsky_Kh) Kh(i, j) = fmax2(Kh(i, j), Laplac2_const_xx(i, j) * Shear_mag(i, j));
        }
      }
      if (legacy_bound)
        for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) Kh(i, j) = fmin2(Kh(i, j), Kh_Max_xx(i, j));
      for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) Kh(i, j) = fmax2(Kh(i, j), CS->Kh_bg_min);
      if (CS->better_bound_Kh && CS->better_bound_Ah) {
        for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) {
          visc_bound_rem(i, j) = 1.0;
          const double Kh_max_here = hrat_min(i, j) * Kh_Max_xx(i, j);
          if (Kh(i, j) >= Kh_max_here) {
            visc_bound_rem(i, j) = 0.0;
            Kh(i, j) = Kh_max_here;
          } else if ((Kh(i, j) > 0.0) || (CS->backscatter_underbound && (Kh_max_here > 0.0))) {
            visc_bound_rem(i, j) = 1.0 - Kh(i, j) / Kh_max_here;
          }
        }
      } else if (CS->better_bound_Kh) {
        for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i)
          Kh(i, j) = fmin2(Kh(i, j), hrat_min(i, j) * Kh_Max_xx(i, j));
      }
      for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) str_xx(i, j) = -Kh(i, j) * sh_xx(i, j);
    } else {
      for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) str_xx(i, j) = 0.0;
    }
    if (CS->biharmonic) {  // :1293-1458
      for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) Ah(i, j) = Ah_bg_xx(i, j);
      if (CS->Smagorinsky_Ah) {
        if (CS->bound_Coriolis) {
          for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) {
            const double AhSm = Shear_mag(i, j) * (Biharm_const_xx(i, j) + Biharm_const2_xx(i, j) * Shear_mag(i, j));
            Ah(i, j) = fmax2(Ah(i, j), AhSm);
          }
        } else {
          for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) {
            const double AhSm = Biharm_const_xx(i, j) * Shear_mag(i, j);
            Ah(i, j) = fmax2(Ah(i, j), AhSm);
          }
        }
        if (CS->bound_Ah && !CS->better_bound_Ah)
          for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) Ah(i, j) = fmin2(Ah(i, j), Ah_Max_xx(i, j));
      }
      if (CS->Re_Ah > 0.0) {
        for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i) {
          const double su = u(i, j, k) + u(i - 1, j, k), sv = v(i, j, k) + v(i, j - 1, k);
          const double KE = 0.125 * ((su * su) + (sv * sv));
          Ah(i, j) = std::sqrt(KE) * Re_Ah_const_xx(i, j);
        }
      }
      if (CS->better_bound_Ah) {
        if (CS->better_bound_Kh) {
          for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i)
            Ah(i, j) = fmin2(Ah(i, j), visc_bound_rem(i, j) * hrat_min(i, j) * Ah_Max_xx(i, j));
        } else {
          for (int j = js_Kh; j <= je_Kh; ++j) for (int i = is_Kh; i <= ie_Kh; ++i)
            Ah(i, j) = fmin2(Ah(i, j), hrat_min(i, j) * Ah_Max_xx(i, j));
        }
      }
      for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i) {
        const double d_del2u = (G.IdyCu(i, j) * Del2u(i, j)) - (G.IdyCu(i - 1, j) * Del2u(i - 1, j));
        const double d_del2v = (G.IdxCv(i, j) * Del2v(i, j)) - (G.IdxCv(i, j - 1) * Del2v(i, j - 1));
        const double d_str = Ah(i, j) * ((DY_dxT(i, j) * d_del2u) - (DX_dyT(i, j) * d_del2v));
        str_xx(i, j) = str_xx(i, j) + d_str;
      }
      // Gradient of Laplacian :1490-1495
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) {
        dDel2vdx(I, J) = DY_dxBu(I, J) * ((Del2v(I + 1, J) * G.IdyCv(I + 1, J)) - (Del2v(I, J) * G.IdyCv(I, J)));
        dDel2udy(I, J) = DX_dyBu(I, J) * ((Del2u(I, J + 1) * G.IdxCu(I, J + 1)) - (Del2u(I, J) * G.IdxCu(I, J)));
      }
    }
    // :1521-1528
    if (CS->Smagorinsky_Kh || CS->Smagorinsky_Ah) {
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) {
        const double sh_xy_sq = sh_xy(I, J) * sh_xy(I, J);
        const double sh_xx_sq = 0.25 * (((sh_xx(I, J) * sh_xx(I, J)) + (sh_xx(I + 1, J + 1) * sh_xx(I + 1, J + 1))) +
                                        ((sh_xx(I, J + 1) * sh_xx(I, J + 1)) + (sh_xx(I + 1, J) * sh_xx(I + 1, J))));
        Shear_mag(I, J) = std::sqrt(sh_xy_sq + sh_xx_sq);
      }
    }
    // :1530-1535
    for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) {
      const double h2uq = 4.0 * (h_u(I, J) * h_u(I, J + 1));
      const double h2vq = 4.0 * (h_v(I, J) * h_v(I + 1, J));
      hq(I, J) = (2.0 * (h2uq * h2vq)) / (h_neglect3 + (h2uq + h2vq) * ((h_u(I, J) + h_u(I, J + 1)) + (h_v(I, J) + h_v(I + 1, J))));
    }
    if (CS->better_bound_Ah || CS->better_bound_Kh) {
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) {
        const double h_min = min4(h_u(I, J), h_u(I, J + 1), h_v(I, J), h_v(I + 1, J));
        hrat_min(I, J) = fmin2(1.0, h_min / (hq(I, J) + h_neglect));
      }
    }
    if (CS->no_slip) {  // :1545-1567
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) {
        if (CS->no_slip && (G.mask2dBu(I, J) < 0.5)) {
          if ((G.mask2dCu(I, J) + G.mask2dCu(I, J + 1)) + (G.mask2dCv(I, J) + G.mask2dCv(I + 1, J)) > 0.0) {
            const double hu = G.mask2dCu(I, J) * h_u(I, J) + G.mask2dCu(I, J + 1) * h_u(I, J + 1);
            const double hv = G.mask2dCv(I, J) * h_v(I, J) + G.mask2dCv(I + 1, J) * h_v(I + 1, J);
            if ((G.mask2dCu(I, J) + G.mask2dCu(I, J + 1)) * (G.mask2dCv(I, J) + G.mask2dCv(I + 1, J)) == 0.0) {
              hq(I, J) = hu + hv;
              hrat_min(I, J) = 1.0;
            } else {
              hq(I, J) = 2.0 * (hu * hv) / ((hu + hv) + h_neglect);
              hrat_min(I, J) = fmin2(1.0, fmin2(hu, hv) / (hq(I, J) + h_neglect));
            }
          }
        }
      }
    }
    if (CS->Laplacian) {  // :1574-1726
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) Kh(I, J) = Kh_bg_xy(I, J);
      if (CS->Smagorinsky_Kh) {
        if (CS->add_LES_viscosity) {
          for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I)
            Kh(I, J) = Kh(I, J) + Laplac2_const_xy(I, J) * Shear_mag(I, J);
        } else {
          for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I)
            Kh(I, J) = fmax2(Kh(I, J), Laplac2_const_xy(I, J) * Shear_mag(I, J));
        }
      }
      if (legacy_bound)
        for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) Kh(I, J) = fmin2(Kh(I, J), Kh_Max_xy(I, J));
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) Kh(I, J) = fmax2(Kh(I, J), CS->Kh_bg_min);
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) {
        if (CS->better_bound_Kh && CS->better_bound_Ah) {
          visc_bound_rem(I, J) = 1.0;
          const double Kh_max_here = hrat_min(I, J) * Kh_Max_xy(I, J);
          if (Kh(I, J) >= Kh_max_here) {
            visc_bound_rem(I, J) = 0.0;
            Kh(I, J) = Kh_max_here;
          } else if ((Kh(I, J) > 0.0) || (CS->backscatter_underbound && (Kh_max_here > 0.0))) {
            visc_bound_rem(I, J) = 1.0 - Kh(I, J) / Kh_max_here;
          }
        } else if (CS->better_bound_Kh) {
          Kh(I, J) = fmin2(Kh(I, J), hrat_min(I, J) * Kh_Max_xy(I, J));
        }
      }
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) str_xy(I, J) = -Kh(I, J) * sh_xy(I, J);
    } else {
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) str_xy(I, J) = 0.;
    }
    if (CS->biharmonic) {  // :1737-1836
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) Ah(I, J) = Ah_bg_xy(I, J);
      if (CS->Smagorinsky_Ah) {
        if (CS->bound_Coriolis) {
          for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) {
            const double AhSm = Shear_mag(I, J) * (Biharm_const_xy(I, J) + Biharm_const2_xy(I, J) * Shear_mag(I, J));
            Ah(I, J) = fmax2(Ah(I, J), AhSm);
          }
        } else {
          for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) {
            const double AhSm = Biharm_const_xy(I, J) * Shear_mag(I, J);
            Ah(I, J) = fmax2(Ah(I, J), AhSm);
          }
        }
        if (CS->bound_Ah && !CS->better_bound_Ah)
          for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) Ah(I, J) = fmin2(Ah(I, J), Ah_Max_xy(I, J));
      }
      if (CS->Re_Ah > 0.0) {
        for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) {
          const double su = u(I, J, k) + u(I, J + 1, k), sv = v(I, J, k) + v(I + 1, J, k);
          const double KE = 0.125 * ((su * su) + (sv * sv));
          Ah(I, J) = std::sqrt(KE) * Re_Ah_const_xy(I, J);
        }
      }
      if (CS->better_bound_Ah) {
        if (CS->better_bound_Kh) {
          for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I)
            Ah(I, J) = fmin2(Ah(I, J), visc_bound_rem(I, J) * hrat_min(I, J) * Ah_Max_xy(I, J));
        } else {
          for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I)
            Ah(I, J) = fmin2(Ah(I, J), hrat_min(I, J) * Ah_Max_xy(I, J));
        }
      }
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I) {
        const double d_str = Ah(I, J) * (dDel2vdx(I, J) + dDel2udy(I, J));
        str_xy(I, J) = str_xy(I, J) + d_str;
      }
    }
    // :1911-1924
    for (int j = Jsq; j <= Jeq + 1; ++j) for (int i = Isq; i <= Ieq + 1; ++i)
      str_xx(i, j) = str_xx(i, j) * (h(i, j, k) * reduction_xx(i, j));
    if (CS->no_slip) {
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I)
        str_xy(I, J) = str_xy(I, J) * (hq(I, J) * reduction_xy(I, J));
    } else {
      for (int J = js - 1; J <= Jeq; ++J) for (int I = is - 1; I <= Ieq; ++I)
        str_xy(I, J) = str_xy(I, J) * (hq(I, J) * G.mask2dBu(I, J) * reduction_xy(I, J));
    }
    // :1929-1954
    for (int j = js; j <= je; ++j) for (int I = Isq; I <= Ieq; ++I)
      diffu(I, j, k) = ((G.IdxCu(I, j) * ((dx2q(I, j - 1) * str_xy(I, j - 1)) - (dx2q(I, j) * str_xy(I, j))) +
                         G.IdyCu(I, j) * ((dy2h(I, j) * str_xx(I, j)) - (dy2h(I + 1, j) * str_xx(I + 1, j)))) *
                        G.IareaCu(I, j)) / (h_u(I, j) + h_neglect);
    for (int J = Jsq; J <= Jeq; ++J) for (int i = is; i <= ie; ++i)
      diffv(i, J, k) = ((G.IdyCv(i, J) * ((dy2q(i - 1, J) * str_xy(i - 1, J)) - (dy2q(i, J) * str_xy(i, J))) -
                         G.IdxCv(i, J) * ((dx2h(i, J) * str_xx(i, J)) - (dx2h(i, J + 1) * str_xx(i, J + 1)))) *
                        G.IareaCv(i, J)) / (h_v(i, J) + h_neglect);
  }
  return 0;
}
