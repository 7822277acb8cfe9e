// Halo updates: the device-side stand-in for the reference's group passes
// (pass_var / pass_vector / do_group_pass, src/framework/MOM_domains.F90 ->
// config_src/infra/FMS2/MOM_domain_infra.F90:171-216, :1141-1200).
//
// Single rank: reentrant directions wrap inside the tile, closed edges are left
// untouched (what mpp_update_domains does at a non-periodic domain edge).
// Multi rank: pack -> ncclSend/ncclRecv to the <=8 neighbours inside one NCCL group on
// the side stream -> unpack (halo_nccl.cu).
#include "ctx.h"

using m6::Geom;

namespace {

// Generic wrap of one unified plane (any stagger) over `nk` levels.
// su/sv = 1 if the field is staggered in i / j (u,q / v,q points).
__global__ void wrap_x_kernel(const Geom G, double* f, int su, int sv, int ilo_mem, int ihi_mem, int jlo_mem,
                              int jhi_mem, int nk) {
  const int ni = G.iec - G.isc + 1;
  const int nwest = (G.isc - 1 - su) - ilo_mem + 1;  // halo columns i <= isc-1-su
  const int neast = ihi_mem - G.iec;                 // halo columns i >= iec+1
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = jlo_mem + blockIdx.y;
  if (t >= nwest + neast || j > jhi_mem) return;
  int i, s;
  if (t < nwest) { i = ilo_mem + t; s = i + ni; }
  else { i = G.iec + 1 + (t - nwest); s = i - ni; }
  for (int k = blockIdx.z; k < nk; k += gridDim.z) {
    double* p = f + (long long)k * G.plane;
    p[G.idx(i, j)] = p[G.idx(s, j)];
  }
}

__global__ void wrap_y_kernel(const Geom G, double* f, int su, int sv, int ilo_mem, int ihi_mem, int jlo_mem,
                              int jhi_mem, int nk) {
  const int nj = G.jec - G.jsc + 1;
  const int nsouth = (G.jsc - 1 - sv) - jlo_mem + 1;
  const int nnorth = jhi_mem - G.jec;
  const int i = ilo_mem + blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (i > ihi_mem || t >= nsouth + nnorth) return;
  int j, s;
  if (t < nsouth) { j = jlo_mem + t; s = j + nj; }
  else { j = G.jec + 1 + (t - nsouth); s = j - nj; }
  for (int k = blockIdx.z; k < nk; k += gridDim.z) {
    double* p = f + (long long)k * G.plane;
    p[G.idx(i, j)] = p[G.idx(i, s)];
  }
}

}  // namespace

int m6_halo_nccl(mom6cu_ctx* c, double* const* fields, const int* staggers, int nfields, int wide, int nk,
                 int halo);  // halo_nccl.cu

// Update the halos of `nfields` unified planes (nk levels each) out to the full
// memory domain (wide=1: the barotropic wide-halo domain; wide=0: G's).
int m6_halo_update(mom6cu_ctx* c, double* const* fields, const int* staggers, int nfields, int wide, int nk) {
  const mom6cu_domain& d = c->dom;
  if (d.npi * d.npj > 1) {  // one message set per group of at most 8 fields
    for (int f0 = 0; f0 < nfields; f0 += 8) {
      const int rc = m6_halo_nccl(c, fields + f0, staggers + f0, nfields - f0 < 8 ? nfields - f0 : 8, wide, nk, -1);
      if (rc) return rc;
    }
    return 0;
  }
  const Geom& G = c->g;
  for (int f = 0; f < nfields; ++f) {
    int ilo, ihi, jlo, jhi;
    m6_extent(c, staggers[f], wide, &ilo, &ihi, &jlo, &jhi);
    const int su = (staggers[f] == ST_U || staggers[f] == ST_Q) ? 1 : 0;
    const int sv = (staggers[f] == ST_V || staggers[f] == ST_Q) ? 1 : 0;
    const int kz = nk < 64 ? nk : 64;
    if (d.cyclic_x) {
      const int nh = ((G.isc - 1 - su) - ilo + 1) + (ihi - G.iec);
      if (nh > 0) {
        dim3 grid((nh + 31) / 32, jhi - jlo + 1, kz);
        M6_LAUNCH(c, wrap_x_kernel, grid, 32, 0, G, fields[f], su, sv, ilo, ihi, jlo, jhi, nk);
      }
    }
    if (d.cyclic_y) {
      const int nh = ((G.jsc - 1 - sv) - jlo + 1) + (jhi - G.jec);
      if (nh > 0) {
        dim3 grid((ihi - ilo + 128) / 128, nh, kz);
        M6_LAUNCH(c, wrap_y_kernel, grid, 128, 0, G, fields[f], su, sv, ilo, ihi, jlo, jhi, nk);
      }
    }
  }
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

// do_group_pass(CS%pass_eta_ubt, CS%BT_Domain), MOM_barotropic.F90:2495-2512
int m6_bt_halo_exchange(mom6cu_ctx* c, double* eta, double* ubt, double* vbt) {
  double* f[3] = {eta, ubt, vbt};
  const int st[3] = {ST_H, ST_U, ST_V};
  return m6_halo_update(c, f, st, 3, 1, 1);
}

// pass_var / pass_vector / do_group_pass for callers that chain resident entries (include/mom6cu.h)
extern "C" int mom6cu_do_group_pass(mom6cu_ctx* c, int nfields, double* const* fields, const int* stagger, int nk) {
  if (!c || nfields < 0 || nk < 1 || (nfields > 0 && (!fields || !stagger))) return MOM6CU_ERR_BAD_ARG;
  if (nfields == 0) return 0;
  M6_CUDA(c, cudaSetDevice(c->device));
  Stager S(c, "pass.");
  std::vector<double*> dev(nfields);
  std::vector<int> st(nfields);
  int rc;
  for (int f = 0; f < nfields; ++f) {
    if (!fields[f] || stagger[f] < 0 || stagger[f] > 3) return c->fail(MOM6CU_ERR_BAD_ARG, "do_group_pass: field %d is null or has an invalid stagger", f);
    st[f] = stagger[f];
    char nm[32];
    snprintf(nm, sizeof nm, "f%d", f);
    if ((rc = S.io(fields[f], st[f], 0, nk, nm, &dev[f]))) return rc;
  }
  if ((rc = S.begin())) return rc;
  if ((rc = m6_halo_update(c, dev.data(), st.data(), nfields, 0, nk))) return rc;
  return S.finish();
}
