// Context, device-memory and staging plumbing behind the C ABI.
#include "ctx.h"
#include <cstring>
#include <cstdarg>
#include <algorithm>

double* mom6cu_ctx::buf(const std::string& name, size_t n) {
  auto it = bufs.find(name);
  if (it != bufs.end() && buf_sz[name] >= n) return it->second;
  if (it != bufs.end()) { cudaFree(it->second); bufs.erase(it); }
  double* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(double)) != cudaSuccess) {
    fail(MOM6CU_ERR_CUDA, "cudaMalloc of %zu doubles for '%s' failed", n, name.c_str());
    return nullptr;
  }
  // zero-filled on the stream the first upload will use, so that the fill can never land after it
  cudaMemsetAsync(p, 0, n * sizeof(double), xfer ? xfer : stream);
  bufs[name] = p;
  buf_sz[name] = n;
  return p;
}

double* mom6cu_ctx::host_scratch(const std::string& name, size_t n) {
  auto it = pinned.find(name);
  if (it != pinned.end() && it->second.second >= n) return it->second.first;
  if (it != pinned.end()) { cudaFreeHost(it->second.first); pinned.erase(it); }
  double* p = nullptr;
  if (cudaHostAlloc(&p, n * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
    fail(MOM6CU_ERR_CUDA, "cudaHostAlloc of %zu doubles for '%s' failed", n, name.c_str());
    return nullptr;
  }
  pinned[name] = {p, n};
  return p;
}

int mom6cu_ctx::fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err, sizeof(err), fmt, ap);
  va_end(ap);
  return code;
}

void m6_extent(const mom6cu_ctx* c, int stagger, int wide, int* ilo, int* ihi, int* jlo, int* jhi) {
  const mom6cu_domain& d = c->dom;
  const int su = (stagger == ST_U || stagger == ST_Q) ? 1 : 0;
  const int sv = (stagger == ST_V || stagger == ST_Q) ? 1 : 0;
  *ilo = (wide ? d.isdw : d.isd) - su;
  *ihi = wide ? d.iedw : d.ied;
  *jlo = (wide ? d.jsdw : d.jsd) - sv;
  *jhi = wide ? d.jedw : d.jed;
}

bool m6_is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int m6_up(mom6cu_ctx* c, const double* src, int stagger, int wide, int nk, double* dst) {
  if (!src || !dst) return c->fail(MOM6CU_ERR_BAD_ARG, "m6_up: null pointer");
  int ilo, ihi, jlo, jhi;
  m6_extent(c, stagger, wide, &ilo, &ihi, &jlo, &jhi);
  const size_t ni = ihi - ilo + 1, nj = jhi - jlo + 1;
  cudaMemcpy3DParms p = {};
  p.srcPtr = make_cudaPitchedPtr((void*)src, ni * sizeof(double), ni, nj);
  p.dstPtr = make_cudaPitchedPtr((void*)(dst + c->g.idx(ilo, jlo)), (size_t)c->g.pitch * sizeof(double),
                                 c->g.pitch, c->g.rows);
  p.extent = make_cudaExtent(ni * sizeof(double), nj, nk);
  p.kind = cudaMemcpyDefault;
  M6_CUDA(c, cudaMemcpy3DAsync(&p, c->xfer ? c->xfer : c->stream));
  return 0;
}

int m6_down(mom6cu_ctx* c, const double* src_plane, int stagger, int wide, int nk, double* dst) {
  if (!src_plane || !dst) return c->fail(MOM6CU_ERR_BAD_ARG, "m6_down: null pointer");
  int ilo, ihi, jlo, jhi;
  m6_extent(c, stagger, wide, &ilo, &ihi, &jlo, &jhi);
  const size_t ni = ihi - ilo + 1, nj = jhi - jlo + 1;
  cudaMemcpy3DParms p = {};
  p.srcPtr = make_cudaPitchedPtr((void*)(src_plane + c->g.idx(ilo, jlo)), (size_t)c->g.pitch * sizeof(double),
                                 c->g.pitch, c->g.rows);
  p.dstPtr = make_cudaPitchedPtr((void*)dst, ni * sizeof(double), ni, nj);
  p.extent = make_cudaExtent(ni * sizeof(double), nj, nk);
  p.kind = cudaMemcpyDefault;
  M6_CUDA(c, cudaMemcpy3DAsync(&p, c->xfer ? c->xfer : c->stream));
  return 0;
}

namespace {
struct PlanePtrs { double* p[10]; };

__global__ void aos_to_planes_kernel(const double* __restrict__ src, int nm, int ni, int nj, long long off0,
                                     int pitch, PlanePtrs dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= ni || j >= nj) return;
  const double* s = src + ((size_t)j * ni + i) * nm;
  const long long o = off0 + (long long)j * pitch + i;
  for (int m = 0; m < nm; ++m) dst.p[m][o] = s[m];
}
}  // namespace

int m6_up_aos(mom6cu_ctx* c, const double* src, int nm, int stagger, int wide, double* const* dst) {
  if (!src) return c->fail(MOM6CU_ERR_BAD_ARG, "m6_up_aos: null pointer");
  if (nm > 10) return c->fail(MOM6CU_ERR_BAD_ARG, "m6_up_aos: nm > 10");
  int ilo, ihi, jlo, jhi;
  m6_extent(c, stagger, wide, &ilo, &ihi, &jlo, &jhi);
  const int ni = ihi - ilo + 1, nj = jhi - jlo + 1;
  const double* dsrc = src;
  if (!m6_is_device_ptr(src)) {
    double* stage = c->buf("__aos_stage", (size_t)10 * c->g.plane);
    if (!stage) return MOM6CU_ERR_CUDA;
    M6_CUDA(c, cudaMemcpyAsync(stage, src, (size_t)nm * ni * nj * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    dsrc = stage;
  }
  PlanePtrs pp;
  for (int m = 0; m < 10; ++m) pp.p[m] = (m < nm) ? dst[m] : nullptr;
  dim3 block(128), grid((ni + 127) / 128, nj);
  M6_LAUNCH(c, aos_to_planes_kernel, grid, block, 0, dsrc, nm, ni, nj, c->g.idx(ilo, jlo), c->g.pitch, pp);
  M6_CUDA(c, cudaGetLastError());
  return 0;
}

extern "C" {

int mom6cu_build_arch(void) { return 100; }

// sizeof of every struct that crosses the boundary, by name: lets a binding (ctypes, ISO_C_BINDING, ...) verify its mirror of
// include/mom6cu.h at start-up (c_sizeof / storage_size against this) instead of discovering a layout slip as a wild pointer.
long long mom6cu_sizeof(const char* name) {
  if (!name) return -1;
  if (!strcmp(name, "mom6cu_advect_tracer_args")) return (long long)sizeof(mom6cu_advect_tracer_args);
  if (!strcmp(name, "mom6cu_ale_args")) return (long long)sizeof(mom6cu_ale_args);
  if (!strcmp(name, "mom6cu_ale_cs")) return (long long)sizeof(mom6cu_ale_cs);
  if (!strcmp(name, "mom6cu_barotropic_cs")) return (long long)sizeof(mom6cu_barotropic_cs);
  if (!strcmp(name, "mom6cu_bt_cont")) return (long long)sizeof(mom6cu_bt_cont);
  if (!strcmp(name, "mom6cu_bt_timeloop_args")) return (long long)sizeof(mom6cu_bt_timeloop_args);
  if (!strcmp(name, "mom6cu_btcalc_args")) return (long long)sizeof(mom6cu_btcalc_args);
  if (!strcmp(name, "mom6cu_btstep_args")) return (long long)sizeof(mom6cu_btstep_args);
  if (!strcmp(name, "mom6cu_continuity_args")) return (long long)sizeof(mom6cu_continuity_args);
  if (!strcmp(name, "mom6cu_continuity_cs")) return (long long)sizeof(mom6cu_continuity_cs);
  if (!strcmp(name, "mom6cu_coradcalc_args")) return (long long)sizeof(mom6cu_coradcalc_args);
  if (!strcmp(name, "mom6cu_coriolisadv_cs")) return (long long)sizeof(mom6cu_coriolisadv_cs);
  if (!strcmp(name, "mom6cu_domain")) return (long long)sizeof(mom6cu_domain);
  if (!strcmp(name, "mom6cu_dyn_split_rk2_cs")) return (long long)sizeof(mom6cu_dyn_split_rk2_cs);
  if (!strcmp(name, "mom6cu_efp")) return (long long)sizeof(mom6cu_efp);
  if (!strcmp(name, "mom6cu_energy_out")) return (long long)sizeof(mom6cu_energy_out);
  if (!strcmp(name, "mom6cu_grid")) return (long long)sizeof(mom6cu_grid);
  if (!strcmp(name, "mom6cu_hor_visc_args")) return (long long)sizeof(mom6cu_hor_visc_args);
  if (!strcmp(name, "mom6cu_hor_visc_cs")) return (long long)sizeof(mom6cu_hor_visc_cs);
  if (!strcmp(name, "mom6cu_mle_cs")) return (long long)sizeof(mom6cu_mle_cs);
  if (!strcmp(name, "mom6cu_pressureforce_args")) return (long long)sizeof(mom6cu_pressureforce_args);
  if (!strcmp(name, "mom6cu_pressureforce_cs")) return (long long)sizeof(mom6cu_pressureforce_cs);
  if (!strcmp(name, "mom6cu_regridding_cs")) return (long long)sizeof(mom6cu_regridding_cs);
  if (!strcmp(name, "mom6cu_remapping_cs")) return (long long)sizeof(mom6cu_remapping_cs);
  if (!strcmp(name, "mom6cu_set_dtbt_args")) return (long long)sizeof(mom6cu_set_dtbt_args);
  if (!strcmp(name, "mom6cu_step_dyn_args")) return (long long)sizeof(mom6cu_step_dyn_args);
  if (!strcmp(name, "mom6cu_sum_output_cs")) return (long long)sizeof(mom6cu_sum_output_cs);
  if (!strcmp(name, "mom6cu_thickness_diffuse_args")) return (long long)sizeof(mom6cu_thickness_diffuse_args);
  if (!strcmp(name, "mom6cu_thickness_diffuse_cs")) return (long long)sizeof(mom6cu_thickness_diffuse_cs);
  if (!strcmp(name, "mom6cu_tracer_advect_cs")) return (long long)sizeof(mom6cu_tracer_advect_cs);
  if (!strcmp(name, "mom6cu_tracer_hor_diff_cs")) return (long long)sizeof(mom6cu_tracer_hor_diff_cs);
  if (!strcmp(name, "mom6cu_tracer_hordiff_args")) return (long long)sizeof(mom6cu_tracer_hordiff_args);
  if (!strcmp(name, "mom6cu_unit_scale")) return (long long)sizeof(mom6cu_unit_scale);
  if (!strcmp(name, "mom6cu_vertvisc_args")) return (long long)sizeof(mom6cu_vertvisc_args);
  if (!strcmp(name, "mom6cu_vertvisc_coef_args")) return (long long)sizeof(mom6cu_vertvisc_coef_args);
  if (!strcmp(name, "mom6cu_vertvisc_cs")) return (long long)sizeof(mom6cu_vertvisc_cs);
  if (!strcmp(name, "mom6cu_vgrid")) return (long long)sizeof(mom6cu_vgrid);
  return -1;
}

int mom6cu_create(mom6cu_ctx** out, const mom6cu_domain* dom, int device) {
  if (!out || !dom) return MOM6CU_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return MOM6CU_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) return MOM6CU_ERR_BAD_ARG;
  if (dom->iec < dom->isc || dom->jec < dom->jsc || dom->nk < 1) return MOM6CU_ERR_BAD_ARG;
  if (dom->isd > dom->isc || dom->ied < dom->iec || dom->jsd > dom->jsc || dom->jed < dom->jec)
    return MOM6CU_ERR_BAD_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return MOM6CU_ERR_CUDA;
  mom6cu_ctx* c = new mom6cu_ctx();
  c->dom = *dom;
  c->device = device;
  mom6cu_domain& d = c->dom;
  // a degenerate wide domain means "same as G"
  if (d.iedw < d.isdw) { d.isdw = d.isd; d.iedw = d.ied; d.jsdw = d.jsd; d.jedw = d.jed; }
  if (d.npi < 1) d.npi = 1;
  if (d.npj < 1) d.npj = 1;
  m6::Geom& g = c->g;
  const int ilo = std::min(d.isd, d.isdw) - 1, ihi = std::max(d.ied, d.iedw);
  const int jlo = std::min(d.jsd, d.jsdw) - 1, jhi = std::max(d.jed, d.jedw);
  // lead padding so that column isc sits on a 128-byte boundary
  const int lead = (16 - ((d.isc - ilo) % 16)) % 16;
  g.i0 = ilo - lead;
  g.j0 = jlo;
  g.nx = ihi - g.i0 + 1;
  g.ny = jhi - g.j0 + 1;
  g.pitch = ((g.nx + 15) / 16) * 16;
  g.rows = g.ny;
  g.nk = d.nk;
  g.plane = (long long)g.pitch * g.rows;
  g.isc = d.isc; g.iec = d.iec; g.jsc = d.jsc; g.jec = d.jec;
  g.isd = d.isd; g.ied = d.ied; g.jsd = d.jsd; g.jed = d.jed;
  g.isdw = d.isdw; g.iedw = d.iedw; g.jsdw = d.jsdw; g.jedw = d.jedw;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return MOM6CU_ERR_CUDA;
  }
  *out = c;
  return 0;
}

int mom6cu_destroy(mom6cu_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& kv : c->bufs) cudaFree(kv.second);
  for (auto& kv : c->pinned) cudaFreeHost(kv.second.first);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev_side) cudaEventDestroy(c->ev_side);
  for (cudaEvent_t e : c->stage_ev) cudaEventDestroy(e);
  if (c->ev_copy) cudaEventDestroy(c->ev_copy);
  if (c->copy) cudaStreamDestroy(c->copy);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->side) cudaStreamDestroy(c->side);
  delete c;
  return 0;
}

int mom6cu_last_error(const mom6cu_ctx* c, char* b, size_t len) {
  if (!b || len == 0) return MOM6CU_ERR_BAD_ARG;
  if (!c) { b[0] = 0; return MOM6CU_ERR_BAD_ARG; }
  strncpy(b, c->err, len - 1);
  b[len - 1] = 0;
  return 0;
}

long long mom6cu_launch_count(const mom6cu_ctx* c) { return c ? c->launches : 0; }
double mom6cu_last_kernel_ms(const mom6cu_ctx* c) { return c ? c->last_ms : 0.0; }
int mom6cu_last_iterations(const mom6cu_ctx* c) { return c ? c->last_iterations : 0; }
int mom6cu_last_step_stage_ms(const mom6cu_ctx* c, double* ms, int n) {
  if (!c || !ms) return 0;
  const int m = n < MOM6CU_NSTAGES ? n : MOM6CU_NSTAGES;
  for (int i = 0; i < m; ++i) ms[i] = c->stage_ms[i];
  return m;
}
double mom6cu_total_kernel_ms(const mom6cu_ctx* c) { return c ? c->total_ms : 0.0; }

double* mom6cu_plane_alloc(mom6cu_ctx* c, const char* name, int nk) {
  if (!c || !name || nk < 1) return nullptr;
  cudaSetDevice(c->device);
  // A name already in use with a smaller size would be freed and reallocated by buf(), leaving the pointer handed out earlier dangling:
  // refuse instead (the caller keeps its plane; a differently sized field needs its own name).
  const std::string key = std::string("user.") + name;
  auto it = c->buf_sz.find(key);
  if (it != c->buf_sz.end() && it->second < (size_t)c->g.plane * nk) {
    c->fail(MOM6CU_ERR_BAD_ARG, "plane_alloc: '%s' already exists with fewer levels; resident planes are never reallocated", name);
    return nullptr;
  }
  return c->buf(key, (size_t)c->g.plane * nk);
}

int mom6cu_plane_upload(mom6cu_ctx* c, double* plane, const double* host, int stagger, int wide, int nk) {
  if (!c || !plane || !host) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  int rc = m6_up(c, host, stagger, wide, nk, plane);
  if (rc) return rc;
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int mom6cu_plane_download(mom6cu_ctx* c, const double* plane, double* host, int stagger, int wide, int nk) {
  if (!c || !plane || !host) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  int rc = m6_down(c, plane, stagger, wide, nk, host);
  if (rc) return rc;
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// A resident field set to zero on the device (the reference's  CS%uhtr(:,:,:) = 0.0  after tracer advection, MOM.F90:1540-1541, for a host
// that keeps the accumulated transports resident).
int mom6cu_plane_zero(mom6cu_ctx* c, double* plane, int nk) {
  if (!c || !plane || nk < 1) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaSetDevice(c->device));
  M6_CUDA(c, cudaMemsetAsync(plane, 0, (size_t)c->g.plane * nk * sizeof(double), c->stream));
  return 0;
}

int mom6cu_sync(mom6cu_ctx* c) {
  if (!c) return MOM6CU_ERR_BAD_ARG;
  M6_CUDA(c, cudaStreamSynchronize(c->stream));
  M6_CUDA(c, cudaStreamSynchronize(c->side));
  return 0;
}

}  // extern "C"
