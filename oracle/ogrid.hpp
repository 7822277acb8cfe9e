// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/oracle.h).
// Views of the ocean_grid_type metrics (src/core/MOM_grid.F90:75-175) with the reference's index bounds,
// plus the hor_index_type shorthands every routine starts with (is,ie,...,Isq,Ieq,...).
#pragma once
#include "farray.hpp"
#include "../include/mom6cu.h"

namespace orc {

struct OGrid {
  int isc, iec, jsc, jec, isd, ied, jsd, jed, IscB, IecB, JscB, JecB, IsdB, IedB, JsdB, JedB, ke;
  V2 mask2dT, mask2dCu, mask2dCv, mask2dBu;
  V2 dxT, dyT, IdxT, IdyT, areaT, IareaT;
  V2 dxCu, dyCu, IdxCu, IdyCu, dy_Cu, areaCu, IareaCu;
  V2 dxCv, dyCv, IdxCv, IdyCv, dx_Cv, areaCv, IareaCv;
  V2 dxBu, dyBu, IdxBu, IdyBu, areaBu, IareaBu;
  V2 bathyT, CoriolisBu, Coriolis2Bu;

  OGrid(const mom6cu_domain* d, const mom6cu_grid* G) {
    isc = d->isc; iec = d->iec; jsc = d->jsc; jec = d->jec; isd = d->isd; ied = d->ied; jsd = d->jsd; jed = d->jed;
    // symmetric memory: IscB = isc-1, IsdB = isd-1 (MOM_hor_index.F90)
    IscB = isc - 1; IecB = iec; JscB = jsc - 1; JecB = jec; IsdB = isd - 1; IedB = ied; JsdB = jsd - 1; JedB = jed;
    ke = d->nk;
    if (!G) return;
    mask2dT = H(G->mask2dT); mask2dCu = U(G->mask2dCu); mask2dCv = V(G->mask2dCv); mask2dBu = Q(G->mask2dBu);
    dxT = H(G->dxT); dyT = H(G->dyT); IdxT = H(G->IdxT); IdyT = H(G->IdyT); areaT = H(G->areaT); IareaT = H(G->IareaT);
    dxCu = U(G->dxCu); dyCu = U(G->dyCu); IdxCu = U(G->IdxCu); IdyCu = U(G->IdyCu); dy_Cu = U(G->dy_Cu);
    areaCu = U(G->areaCu); IareaCu = U(G->IareaCu);
    dxCv = V(G->dxCv); dyCv = V(G->dyCv); IdxCv = V(G->IdxCv); IdyCv = V(G->IdyCv); dx_Cv = V(G->dx_Cv);
    areaCv = V(G->areaCv); IareaCv = V(G->IareaCv);
    dxBu = Q(G->dxBu); dyBu = Q(G->dyBu); IdxBu = Q(G->IdxBu); IdyBu = Q(G->IdyBu); areaBu = Q(G->areaBu);
    IareaBu = Q(G->IareaBu);
    bathyT = H(G->bathyT); CoriolisBu = Q(G->CoriolisBu); Coriolis2Bu = Q(G->Coriolis2Bu);
  }
  V2 H(const double* p) const { return V2((double*)p, isd, ied, jsd, jed); }
  V2 U(const double* p) const { return V2((double*)p, isd - 1, ied, jsd, jed); }
  V2 V(const double* p) const { return V2((double*)p, isd, ied, jsd - 1, jed); }
  V2 Q(const double* p) const { return V2((double*)p, isd - 1, ied, jsd - 1, jed); }
  V3 H3(const double* p, int nk = -1) const { return V3((double*)p, isd, ied, jsd, jed, nk < 0 ? ke : nk); }
  V3 U3(const double* p, int nk = -1) const { return V3((double*)p, isd - 1, ied, jsd, jed, nk < 0 ? ke : nk); }
  V3 V3_(const double* p, int nk = -1) const { return V3((double*)p, isd, ied, jsd - 1, jed, nk < 0 ? ke : nk); }
  V3 Q3(const double* p, int nk = -1) const { return V3((double*)p, isd - 1, ied, jsd - 1, jed, nk < 0 ? ke : nk); }
  // scratch arrays with the SZI_/SZIB_ extents
  A2 aH() const { return A2(isd, ied, jsd, jed); }
  A2 aU() const { return A2(isd - 1, ied, jsd, jed); }
  A2 aV() const { return A2(isd, ied, jsd - 1, jed); }
  A2 aQ() const { return A2(isd - 1, ied, jsd - 1, jed); }
};

// column reconstructions of oracle/remap.cpp used outside it (1-based arrays, element 0 unused)
void ale_plm_edge_values_column(int nk, const double* h, const double* Q, bool bdry_extrap, double h_neglect, double* Q_t, double* Q_b);
void ale_ppm_edge_values_column(int nk, const double* h, const double* Q, bool bdry_extrap, double h_neglect, double h_neglect_edge,
                                double* Q_t, double* Q_b);

static inline double max4(double a, double b, double c, double d) { return fmax2(fmax2(fmax2(a, b), c), d); }
static inline double min4(double a, double b, double c, double d) { return fmin2(fmin2(fmin2(a, b), c), d); }

}  // namespace orc
