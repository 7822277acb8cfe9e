// Resident grid metrics (ocean_grid_type, src/core/MOM_grid.F90:75-175) and vertical-grid scalars.
#include "ctx.h"

extern "C" int mom6cu_set_grid(mom6cu_ctx* c, const mom6cu_grid* G) {
  if (!c || !G) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  struct F { const char* name; const double* src; int st; const double** dst; };
  GridDev& D = c->grid;
  const F f[MOM6CU_GRID_NFIELDS] = {
      {"mask2dT", G->mask2dT, ST_H, &D.mask2dT}, {"mask2dCu", G->mask2dCu, ST_U, &D.mask2dCu},
      {"mask2dCv", G->mask2dCv, ST_V, &D.mask2dCv}, {"mask2dBu", G->mask2dBu, ST_Q, &D.mask2dBu},
      {"dxT", G->dxT, ST_H, &D.dxT}, {"dyT", G->dyT, ST_H, &D.dyT}, {"IdxT", G->IdxT, ST_H, &D.IdxT},
      {"IdyT", G->IdyT, ST_H, &D.IdyT}, {"areaT", G->areaT, ST_H, &D.areaT}, {"IareaT", G->IareaT, ST_H, &D.IareaT},
      {"dxCu", G->dxCu, ST_U, &D.dxCu}, {"dyCu", G->dyCu, ST_U, &D.dyCu}, {"IdxCu", G->IdxCu, ST_U, &D.IdxCu},
      {"IdyCu", G->IdyCu, ST_U, &D.IdyCu}, {"dy_Cu", G->dy_Cu, ST_U, &D.dy_Cu}, {"areaCu", G->areaCu, ST_U, &D.areaCu},
      {"IareaCu", G->IareaCu, ST_U, &D.IareaCu},
      {"dxCv", G->dxCv, ST_V, &D.dxCv}, {"dyCv", G->dyCv, ST_V, &D.dyCv}, {"IdxCv", G->IdxCv, ST_V, &D.IdxCv},
      {"IdyCv", G->IdyCv, ST_V, &D.IdyCv}, {"dx_Cv", G->dx_Cv, ST_V, &D.dx_Cv}, {"areaCv", G->areaCv, ST_V, &D.areaCv},
      {"IareaCv", G->IareaCv, ST_V, &D.IareaCv},
      {"dxBu", G->dxBu, ST_Q, &D.dxBu}, {"dyBu", G->dyBu, ST_Q, &D.dyBu}, {"IdxBu", G->IdxBu, ST_Q, &D.IdxBu},
      {"IdyBu", G->IdyBu, ST_Q, &D.IdyBu}, {"areaBu", G->areaBu, ST_Q, &D.areaBu}, {"IareaBu", G->IareaBu, ST_Q, &D.IareaBu},
      {"bathyT", G->bathyT, ST_H, &D.bathyT}, {"CoriolisBu", G->CoriolisBu, ST_Q, &D.CoriolisBu},
      {"Coriolis2Bu", G->Coriolis2Bu, ST_Q, &D.Coriolis2Bu}};
  for (int m = 0; m < MOM6CU_GRID_NFIELDS; ++m) {
    if (!f[m].src) return c->fail(MOM6CU_ERR_BAD_ARG, "mom6cu_set_grid: G%%%s is null", f[m].name);
    double* p = c->plane2(std::string("G.") + f[m].name);
    if (!p) return MOM6CU_ERR_CUDA;
    int rc = m6_up(c, f[m].src, f[m].st, 0, 1, p);
    if (rc) return rc;
    *f[m].dst = p;
  }
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  c->have_grid = true;
  return 0;
}

extern "C" int mom6cu_set_vgrid(mom6cu_ctx* c, const mom6cu_vgrid* GV) {
  if (!c || !GV) return MOM6CU_ERR_BAD_ARG;
  c->vgrid = *GV;
  c->have_vgrid = true;
  return 0;
}

extern "C" int mom6cu_set_unit_scale(mom6cu_ctx* c, const mom6cu_unit_scale* US) {
  if (!c || !US) return MOM6CU_ERR_BAD_ARG;
  c->US = *US;
  return 0;
}
