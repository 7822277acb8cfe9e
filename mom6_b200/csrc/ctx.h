// Host-side context of the C ABI (one per rank / GPU).
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <set>
#include <string>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../include/mom6cu.h"
#include "common.cuh"
#include "stage.h"

enum { ST_H = 0, ST_U = 1, ST_V = 2, ST_Q = 3 };

struct mom6cu_ctx {
  mom6cu_domain dom;
  m6::Geom g;
  int device = 0;
  cudaStream_t stream = nullptr;  // compute stream
  cudaStream_t side = nullptr;    // halo-exchange stream
  cudaStream_t copy = nullptr;    // staging copies that overlap the stage kernels (Stager::defer / early)
  cudaStream_t xfer = nullptr;    // stream m6_up / m6_down enqueue on (null: the compute stream)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_side = nullptr, ev_copy = nullptr;
  double last_ms = 0.0;   // device time of the most recent compute entry / repetition
  double total_ms = 0.0;  // summed over the repetitions of the most recent resident call
  long long launches = 0;
  int last_iterations = 0;  // passes made by the most recent iterative entry (advect_tracer)
  double stage_ms[8] = {0};  // device time per stage inside the most recent step (mom6cu_last_step_stage_ms)
  std::vector<cudaEvent_t> stage_ev;  // event pool of the in-step stage timer
  int warnings = 0;
  char err[1024] = {0};
  std::map<std::string, double*> bufs;
  std::map<std::string, size_t> buf_sz;
  std::set<const void*> polarity_checked;  // resident ua_polarity / va_polarity arrays already verified to be all +1 (btstep)
  std::map<std::string, std::pair<double*, size_t>> pinned;  // page-locked host scratch
  void* comm = nullptr;  // ncclComm_t when multi-rank
  // resident grid metrics and resolved control structures
  GridDev grid = {};
  bool have_grid = false;
  mom6cu_vgrid vgrid = {};
  bool have_vgrid = false;
  mom6cu_continuity_cs cont_cs = {};
  bool have_cont_cs = false;
  mom6cu_unit_scale US = {1., 1., 1., 1., 1., 1., 1., 1., 1., 1.};
  mom6cu_coriolisadv_cs corad_cs = {};
  bool have_corad_cs = false;
  mom6cu_pressureforce_cs pgf_cs = {};
  bool have_pgf_cs = false;
  const double *pgf_Rlay = nullptr, *pgf_gprime = nullptr;  // device copies of GV%Rlay, GV%g_prime
  mom6cu_vertvisc_cs vv_cs = {};
  bool have_vv_cs = false;
  mom6cu_hor_visc_cs hv_cs = {};      // as given (host pointers)
  mom6cu_hor_visc_cs hv_cs_dev = {};  // same flags, pointers to the resident planes
  bool have_hv_cs = false;
  int rank = 0, nranks = 1;

  // named, persistent, zero-initialised device buffer of n doubles
  double* buf(const std::string& name, size_t n);
  double* plane2(const std::string& name) { return buf(name, (size_t)g.plane); }
  double* plane3(const std::string& name) { return buf(name, (size_t)g.plane * g.nk); }
  double* plane3k(const std::string& name, int nk) { return buf(name, (size_t)g.plane * nk); }
  // named, persistent, page-locked host scratch of n doubles
  double* host_scratch(const std::string& name, size_t n);
  int fail(int code, const char* fmt, ...);
  // true if p points at the start of one of this context's resident buffers
  bool is_plane(const void* p) const {
    for (const auto& kv : bufs) if ((const void*)kv.second == p) return true;
    return false;
  }
};

#define M6_CUDA(ctx, call)                                                              \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return (ctx)->fail(MOM6CU_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,   \
                         cudaGetErrorString(e_));                                       \
  } while (0)

#define M6_LAUNCH(ctx, kern, grid, block, smem, ...)                                    \
  do {                                                                                  \
    kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                      \
    (ctx)->launches++;                                                                  \
  } while (0)

// extents of a Fortran array of the given stagger on G (wide=0) or the wide domain
void m6_extent(const mom6cu_ctx* c, int stagger, int wide, int* ilo, int* ihi, int* jlo, int* jhi);
bool m6_is_device_ptr(const void* p);
// halo update of nfields unified planes (nk levels each): cyclic wrap on one tile, NCCL between tiles (halo.cu)
int m6_halo_update(mom6cu_ctx* c, double* const* fields, const int* staggers, int nfields, int wide, int nk);
// max over ranks of one host int (sum_across_PEs of a flag, MOM_coms.F90); identity on one rank
int m6_allreduce_max_int(mom6cu_ctx* c, int* v);
int m6_allreduce_min_double(mom6cu_ctx* c, double* v);  // min_across_PEs
int m6_allreduce_sum_i64(mom6cu_ctx* c, long long* v, int n);      // sum_across_PEs of the EFP integers
int m6_allreduce_max_doubles(mom6cu_ctx* c, double* v, int n);   // max_across_PEs
int m6_allreduce_min_doubles(mom6cu_ctx* c, double* v, int n);
// copy a Fortran-shaped (host or device) array into / out of unified planes
int m6_up(mom6cu_ctx* c, const double* src, int stagger, int wide, int nk, double* dst);
int m6_down(mom6cu_ctx* c, const double* src_plane, int stagger, int wide, int nk, double* dst);
// array-of-structs (nm reals per point, m fastest) -> nm separate planes
int m6_up_aos(mom6cu_ctx* c, const double* src, int nm, int stagger, int wide, double* const* dst);

// Staging of one C-ABI call: Fortran-shaped (host or device) arrays -> named resident planes, and the
// outputs back.  in*: intent(in) (null stays null); io*: intent(out)/(inout) -- uploaded first so that points
// the stage does not touch keep the caller's values, downloaded by finish().  begin()/finish() bracket the
// stage kernels with CUDA events on the compute stream (mom6cu_last_kernel_ms).
struct Stager {
  mom6cu_ctx* c;
  std::string pfx;
  struct Out { const double* dev; double* host; int st, wide, nk; };
  std::vector<Out> outs;
  // Only page-locked host arrays take the overlapped path: copies from pageable memory block the host thread anyway.
  static bool pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
  }
  Stager(mom6cu_ctx* c_, const char* prefix) : c(c_), pfx(prefix) {}
  ~Stager() { c->xfer = nullptr; }
  int in(const double* src, int st, int wide, int nk, const char* name, const double** dst) {
    *dst = nullptr;
    if (!src) return 0;
    if (c->is_plane(src)) { *dst = src; return 0; }  // already resident: use in place
    cudaStream_t keep = c->xfer;
    if (keep && !pinned(src)) c->xfer = nullptr;  // pageable sources stay on the compute stream (see defer_begin)
    double* p = c->buf(pfx + name, (size_t)c->g.plane * nk);  // a new buffer is zero-filled on the stream of its upload
    int rc = MOM6CU_ERR_CUDA;
    if (p) { *dst = p; rc = m6_up(c, src, st, wide, nk, p); }
    c->xfer = keep;
    return rc;
  }
  int io(double* src, int st, int wide, int nk, const char* name, double** dst) {
    const double* p = nullptr;
    int rc = in(src, st, wide, nk, name, &p);
    *dst = (double*)p;
    if (!rc && p && p != src) outs.push_back({p, src, st, wide, nk});  // resident outputs stay on the device
    return rc;
  }
  int in3(const double* s, int st, const char* n, const double** d) { return in(s, st, 0, c->g.nk, n, d); }
  int in2(const double* s, int st, const char* n, const double** d) { return in(s, st, 0, 1, n, d); }
  int io3(double* s, int st, const char* n, double** d) { return io(s, st, 0, c->g.nk, n, d); }
  int io2(double* s, int st, const char* n, double** d) { return io(s, st, 0, 1, n, d); }
  // Overlapped staging (host-array callers only; resident planes never get here).  defer_begin(): the uploads registered from
  // now on go to the copy stream, queued behind the uploads already issued on the compute stream (one H2D stream at a time, so
  // the first group arrives at full PCIe speed); defer_end() closes the group and wait_deferred() makes the compute stream wait
  // for it -- call it just before the first kernel that reads a deferred field.
  static bool overlap_on() { static int v = -1; if (v < 0) { const char* e = getenv("MOM6CU_NO_OVERLAP"); v = (e && atoi(e) != 0) ? 0 : 1; } return v == 1; }
  bool overlap = false;
  int defer_begin() {
    overlap = overlap_on();
    if (!overlap) return 0;
    M6_CUDA(c, cudaEventRecord(c->ev_copy, c->stream));
    M6_CUDA(c, cudaStreamWaitEvent(c->copy, c->ev_copy, 0));
    c->xfer = c->copy;
    return 0;
  }
  int defer_end() { if (!overlap) return 0; c->xfer = nullptr; M6_CUDA(c, cudaEventRecord(c->ev_copy, c->copy)); deferred = true; return 0; }
  int wait_deferred() { if (deferred) { M6_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); deferred = false; } return 0; }
  // early(dev): the output staged from `dev` is final -- download it now on the copy stream, under the kernels still to come.
  int early(const double* dev) {
    if (!overlap_on()) return 0;
    for (Out& o : outs) if (o.dev == dev && o.host && pinned(o.host)) {
      M6_CUDA(c, cudaEventRecord(c->ev_copy, c->stream));
      M6_CUDA(c, cudaStreamWaitEvent(c->copy, c->ev_copy, 0));
      c->xfer = c->copy;
      const int rc = m6_down(c, o.dev, o.st, o.wide, o.nk, o.host);
      c->xfer = nullptr;
      o.host = nullptr; used_copy = true;
      return rc;
    }
    return 0;
  }
  bool deferred = false, used_copy = false;
  int begin() { M6_CUDA(c, cudaEventRecord(c->ev0, c->stream)); return 0; }
  int finish() {
    M6_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    for (const Out& o : outs) { if (!o.host) continue; int rc = m6_down(c, o.dev, o.st, o.wide, o.nk, o.host); if (rc) return rc; }
    M6_CUDA(c, cudaStreamSynchronize(c->stream));
    if (used_copy || deferred) M6_CUDA(c, cudaStreamSynchronize(c->copy));
    float ms = 0.f;
    M6_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_ms = ms; c->total_ms = ms;
    return 0;
  }
};
