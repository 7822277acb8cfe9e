// Multi-rank halo exchange over NCCL point-to-point (placeholder until the NCCL layer lands).
#include "ctx.h"
int m6_halo_nccl(mom6cu_ctx* c, double* const*, const int*, int, int, int, int) {
  return c->fail(MOM6CU_ERR_NCCL, "multi-rank halo exchange requested but no communicator is attached");
}
