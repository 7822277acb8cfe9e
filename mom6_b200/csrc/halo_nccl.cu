// Multi-rank halo exchange: pass_var / pass_vector / do_group_pass of the reference
// (src/framework/MOM_domains.F90 -> config_src/infra/FMS2/MOM_domain_infra.F90:171-216,
// :1141-1200, FMS mpp_update_domains / mpp_do_group_update over MPI) re-done as
//   ONE pack kernel -> ncclSend/ncclRecv to the <=8 neighbours in one NCCL group -> ONE unpack kernel
// per group pass, all fields and all k levels of the group in the same messages.
// The 2-D (i,j) tile decomposition, reentrant wrap and closed edges follow the reference (one rectangular tile per rank, all k local).
// One deliberate difference from FMS's symmetric-memory updates: the shared edge point of a staggered field (u at I = isc-1, v at
// J = jsc-1) is NOT refreshed from the neighbour's iec / jec value.  Both tiles compute that point from the same inputs with the same
// instruction sequence, so the two copies are bit-identical by construction -- which the layout tests (1 vs 2, 4, 8 tiles, bit for bit)
// verify on every field of the step; a host that writes the edge on one side only must exchange it itself.
#include "ctx.h"
#include <nccl.h>
#include <vector>
#include <algorithm>

using m6::Geom;

// the 8 neighbour directions in a fixed order shared by all ranks
static const int DIRS[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, -1}, {1, -1}, {-1, 1}};

// Host-only planning (no CUDA): boxes exchanged with the neighbour in direction `dir` for a field
// of the given stagger.  Returns the peer rank or -1 (closed edge / no neighbour).
// send box = part of my computational domain the peer needs; recv box = my halo it fills.
extern "C" int mom6cu_halo_plan(const mom6cu_domain* d, int stagger, int wide, int halo, int dir, int* send_box,
                                int* recv_box) {
  if (!d || dir < 0 || dir > 7) return -1;
  const int dx = DIRS[dir][0], dy = DIRS[dir][1];
  const int su = (stagger == ST_U || stagger == ST_Q) ? 1 : 0;
  const int sv = (stagger == ST_V || stagger == ST_Q) ? 1 : 0;
  const int npi = d->npi < 1 ? 1 : d->npi, npj = d->npj < 1 ? 1 : d->npj;
  int qi = d->pi + dx, qj = d->pj + dy;
  if (qi < 0 || qi >= npi) { if (!d->cyclic_x) return -1; qi = (qi + npi) % npi; }
  if (qj < 0 || qj >= npj) { if (!d->cyclic_y) return -1; qj = (qj + npj) % npj; }
  const int hwx = halo >= 0 ? halo : (wide ? d->isc - d->isdw : d->isc - d->isd);
  const int hwy = halo >= 0 ? halo : (wide ? d->jsc - d->jsdw : d->jsc - d->jsd);
  if ((dx != 0 && hwx <= 0) || (dy != 0 && hwy <= 0)) return -1;
  int si0, si1, ri0, ri1, sj0, sj1, rj0, rj1;
  if (dx > 0) { si0 = d->iec - hwx + 1 - su; si1 = d->iec - su; ri0 = d->iec + 1; ri1 = d->iec + hwx; }
  else if (dx < 0) { si0 = d->isc; si1 = d->isc + hwx - 1; ri0 = d->isc - hwx - su; ri1 = d->isc - 1 - su; }
  else { si0 = ri0 = d->isc - su; si1 = ri1 = d->iec; }
  if (dy > 0) { sj0 = d->jec - hwy + 1 - sv; sj1 = d->jec - sv; rj0 = d->jec + 1; rj1 = d->jec + hwy; }
  else if (dy < 0) { sj0 = d->jsc; sj1 = d->jsc + hwy - 1; rj0 = d->jsc - hwy - sv; rj1 = d->jsc - 1 - sv; }
  else { sj0 = rj0 = d->jsc - sv; sj1 = rj1 = d->jec; }
  if (send_box) { send_box[0] = si0; send_box[1] = si1; send_box[2] = sj0; send_box[3] = sj1; }
  if (recv_box) { recv_box[0] = ri0; recv_box[1] = ri1; recv_box[2] = rj0; recv_box[3] = rj1; }
  return qj * npi + qi;
}

namespace {

constexpr int MAXF = 8;
struct Box { int i0, ni, j0, nj; long long off; };  // off: offset (doubles) in the direction's buffer
struct PackPlan {
  double* f[MAXF];
  Box b[8][MAXF];
  double* buf[8];
  int active[8];
  int nf, nk;
};

template <bool PACK>
__global__ void halo_pack_kernel(const Geom G, const PackPlan P) {
  const int dir = blockIdx.z, fi = blockIdx.y;
  if (!P.active[dir] || fi >= P.nf) return;
  const Box b = P.b[dir][fi];
  const long long n2 = (long long)b.ni * b.nj, n = n2 * P.nk;
  double* fld = P.f[fi];
  double* buf = P.buf[dir] + b.off;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e / n2);
    const int r = (int)(e - (long long)k * n2);
    const int jj = r / b.ni, ii = r - jj * b.ni;
    const long long g = (long long)k * G.plane + G.idx(b.i0 + ii, b.j0 + jj);
    if (PACK) buf[e] = fld[g];
    else fld[g] = buf[e];
  }
}

}  // namespace

extern "C" int mom6cu_comm_unique_id(void* out, int nbytes) {
  if (!out || nbytes < (int)sizeof(ncclUniqueId)) return -(int)sizeof(ncclUniqueId);
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return MOM6CU_ERR_NCCL;
  memcpy(out, &id, sizeof(id));
  return 0;
}

extern "C" int mom6cu_comm_init(mom6cu_ctx* c, const void* id_bytes, int nbytes, int rank, int nranks) {
  if (!c || !id_bytes || nbytes < (int)sizeof(ncclUniqueId)) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  if (c->dom.npi * c->dom.npj != nranks || c->dom.pj * c->dom.npi + c->dom.pi != rank)
    return c->fail(MOM6CU_ERR_BAD_ARG, "comm_init: layout %dx%d / tile (%d,%d) does not match rank %d of %d", c->dom.npi,
                   c->dom.npj, c->dom.pi, c->dom.pj, rank, nranks);
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm;
  ncclResult_t r = ncclCommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) return c->fail(MOM6CU_ERR_NCCL, "ncclCommInitRank: %s", ncclGetErrorString(r));
  c->comm = comm;
  c->rank = rank;
  c->nranks = nranks;
  return 0;
}

int m6_allreduce_max_int(mom6cu_ctx* c, int* v) {
  if (c->nranks <= 1) return 0;
  if (!c->comm) return c->fail(MOM6CU_ERR_NCCL, "multi-rank reduction requested but no communicator is attached");
  int* d = (int*)c->buf("comm.flag", 2);
  int* h = (int*)c->host_scratch("comm.flag", 2);
  if (!d || !h) return MOM6CU_ERR_CUDA;
  h[0] = *v;
  M6_CUDA(c, cudaMemcpyAsync(d, h, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  ncclResult_t r = ncclAllReduce(d, d, 1, ncclInt, ncclMax, (ncclComm_t)c->comm, c->stream);
  if (r != ncclSuccess) return c->fail(MOM6CU_ERR_NCCL, "ncclAllReduce: %s", ncclGetErrorString(r));
  M6_CUDA(c, cudaMemcpyAsync(h, d, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  *v = h[0];
  return 0;
}

int m6_allreduce_min_double(mom6cu_ctx* c, double* v) {
  if (c->nranks <= 1) return 0;
  if (!c->comm) return c->fail(MOM6CU_ERR_NCCL, "multi-rank reduction requested but no communicator is attached");
  double* d = c->buf("comm.dmin", 2);
  double* h = c->host_scratch("comm.dmin", 2);
  if (!d || !h) return MOM6CU_ERR_CUDA;
  h[0] = *v;
  M6_CUDA(c, cudaMemcpyAsync(d, h, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  ncclResult_t r = ncclAllReduce(d, d, 1, ncclDouble, ncclMin, (ncclComm_t)c->comm, c->stream);
  if (r != ncclSuccess) return c->fail(MOM6CU_ERR_NCCL, "ncclAllReduce: %s", ncclGetErrorString(r));
  M6_CUDA(c, cudaMemcpyAsync(h, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  *v = h[0];
  return 0;
}

// sum_across_PEs of n 64-bit integers (the extended-fixed-point sums, MOM_coms.F90:214) / max_across_PEs of n reals; in place
// on host values, identity on one rank.
static int m6_allreduce_host(mom6cu_ctx* c, void* v, int n, ncclDataType_t type, ncclRedOp_t op) {
  if (c->nranks <= 1 || n <= 0) return 0;
  if (!c->comm) return c->fail(MOM6CU_ERR_NCCL, "multi-rank reduction requested but no communicator is attached");
  double* d = c->buf("comm.vec", (size_t)n);
  double* h = c->host_scratch("comm.vec", (size_t)n);
  if (!d || !h) return MOM6CU_ERR_CUDA;
  memcpy(h, v, (size_t)n * 8);
  M6_CUDA(c, cudaMemcpyAsync(d, h, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  ncclResult_t r = ncclAllReduce(d, d, n, type, op, (ncclComm_t)c->comm, c->stream);
  if (r != ncclSuccess) return c->fail(MOM6CU_ERR_NCCL, "ncclAllReduce: %s", ncclGetErrorString(r));
  M6_CUDA(c, cudaMemcpyAsync(h, d, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(v, h, (size_t)n * 8);
  return 0;
}
int m6_allreduce_sum_i64(mom6cu_ctx* c, long long* v, int n) { return m6_allreduce_host(c, v, n, ncclInt64, ncclSum); }
int m6_allreduce_max_doubles(mom6cu_ctx* c, double* v, int n) { return m6_allreduce_host(c, v, n, ncclDouble, ncclMax); }
int m6_allreduce_min_doubles(mom6cu_ctx* c, double* v, int n) { return m6_allreduce_host(c, v, n, ncclDouble, ncclMin); }

extern "C" int mom6cu_comm_destroy(mom6cu_ctx* c) {
  if (c && c->comm) { ncclCommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
  return 0;
}

int m6_halo_nccl(mom6cu_ctx* c, double* const* fields, const int* staggers, int nfields, int wide, int nk, int halo) {
  if (!c->comm) return c->fail(MOM6CU_ERR_NCCL, "multi-rank halo exchange requested but no communicator is attached");
  if (nfields > MAXF) return c->fail(MOM6CU_ERR_BAD_ARG, "halo group of %d fields exceeds %d", nfields, MAXF);
  {  // a halo wider than the memory halo of the fields would index outside their planes
    const mom6cu_domain& d = c->dom;
    const int hmax = wide ? std::min(d.isc - d.isdw, d.jsc - d.jsdw) : std::min(d.isc - d.isd, d.jsc - d.jsd);
    if (halo > hmax) return c->fail(MOM6CU_ERR_BAD_ARG, "halo exchange of width %d requested on fields with a memory halo of %d", halo, hmax);
  }
  const Geom& G = c->g;
  PackPlan S = {}, R = {};
  S.nf = R.nf = nfields;
  S.nk = R.nk = nk;
  int peer[8];
  long long cnt[8];
  long long total = 0;
  for (int f = 0; f < nfields; ++f) S.f[f] = R.f[f] = fields[f];
  for (int dir = 0; dir < 8; ++dir) {
    cnt[dir] = 0;
    peer[dir] = -1;
    for (int f = 0; f < nfields; ++f) {
      int sb[4], rb[4];
      const int p = mom6cu_halo_plan(&c->dom, staggers[f], wide, halo, dir, sb, rb);
      peer[dir] = p;
      if (p < 0) break;
      S.b[dir][f] = {sb[0], sb[1] - sb[0] + 1, sb[2], sb[3] - sb[2] + 1, cnt[dir]};
      R.b[dir][f] = {rb[0], rb[1] - rb[0] + 1, rb[2], rb[3] - rb[2] + 1, cnt[dir]};
      cnt[dir] += (long long)S.b[dir][f].ni * S.b[dir][f].nj * nk;
    }
    S.active[dir] = R.active[dir] = (peer[dir] >= 0 && cnt[dir] > 0) ? 1 : 0;
    total += cnt[dir];
  }
  if (total == 0) return 0;
  double* sbuf = c->buf("halo.send", (size_t)total);
  double* rbuf = c->buf("halo.recv", (size_t)total);
  if (!sbuf || !rbuf) return MOM6CU_ERR_CUDA;
  long long off = 0;
  long long maxn = 0;
  for (int dir = 0; dir < 8; ++dir) {
    S.buf[dir] = sbuf + off;
    R.buf[dir] = rbuf + off;
    off += cnt[dir];
    if (cnt[dir] > maxn) maxn = cnt[dir];
  }
  const int bx = (int)std::min<long long>((maxn / nfields + 255) / 256 + 1, 1024);
  dim3 grid(bx, nfields, 8);
  M6_LAUNCH(c, halo_pack_kernel<true>, grid, 256, 0, G, S);
  ncclComm_t comm = (ncclComm_t)c->comm;
  // A message sent towards direction d arrives at the peer from its direction -d: post the
  // receives in the order the peers post their sends (same DIRS order on every rank).
  static const int OPP[8] = {1, 0, 3, 2, 5, 4, 7, 6};
  ncclResult_t post = ncclSuccess;   // the first failed post; the group is still closed so that the communicator stays usable
  ncclGroupStart();
  for (int dir = 0; dir < 8; ++dir) {
    if (!S.active[dir]) continue;
    if (peer[dir] == c->rank) continue;  // self-neighbour (reentrant with one tile in that direction)
    const ncclResult_t q = ncclSend(S.buf[dir], (size_t)cnt[dir], ncclDouble, peer[dir], comm, c->stream);
    if (q != ncclSuccess && post == ncclSuccess) post = q;
  }
  for (int dir = 0; dir < 8; ++dir) {
    const int rd = OPP[dir];  // what the peer sent towards `dir` lands in my halo on side -dir
    if (!R.active[rd]) continue;
    if (peer[rd] == c->rank) continue;
    const ncclResult_t q = ncclRecv(R.buf[rd], (size_t)cnt[rd], ncclDouble, peer[rd], comm, c->stream);
    if (q != ncclSuccess && post == ncclSuccess) post = q;
  }
  ncclResult_t r = ncclGroupEnd();
  if (post != ncclSuccess) return c->fail(MOM6CU_ERR_NCCL, "halo exchange: ncclSend/ncclRecv: %s", ncclGetErrorString(post));
  if (r != ncclSuccess) return c->fail(MOM6CU_ERR_NCCL, "halo exchange: %s", ncclGetErrorString(r));
  for (int dir = 0; dir < 8; ++dir) {
    if (!S.active[dir] || peer[dir] != c->rank) continue;
    // my own message towards `dir` fills my halo on the opposite side
    M6_CUDA(c, cudaMemcpyAsync(R.buf[OPP[dir]], S.buf[dir], (size_t)cnt[dir] * sizeof(double), cudaMemcpyDeviceToDevice,
                               c->stream));
  }
  M6_LAUNCH(c, halo_pack_kernel<false>, grid, 256, 0, G, R);
  M6_CUDA(c, cudaGetLastError());
  return 0;
}
