// gb25_api.cu — the C ABI of libgb25cuda (include/gb25cuda.h): handle, device memory, parent-shaped
// transfers, and the stage sequencing of the Oceananigans HydrostaticFreeSurfaceModel time step
// (QuasiAdamsBashforth2; SURVEY.md A.4, stage order /root/reference/src/precompile.jl:31-42).
// There is no CPU path: without a CUDA device gb25_create fails with GB25_ERR_NO_DEVICE.
#include <algorithm>
#include <chrono>
#include <mutex>
#include <vector>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <utility>

#include "gb25_internal.h"

static thread_local std::string g_create_error;

#define CK(h, call)                                                                                   \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      char buf_[512];                                                                                 \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      (h)->err = buf_; (h)->sticky = GB25_ERR_CUDA;                                                   \
      return GB25_ERR_CUDA;                                                                           \
    }                                                                                                 \
  } while (0)

#define REQUIRE(h)                                                       \
  do {                                                                   \
    if (!(h)) return GB25_ERR_INVALID;                                   \
    if ((h)->sticky) return (h)->sticky;                                 \
    cudaError_t e0_ = cudaSetDevice((h)->device);                        \
    if (e0_ != cudaSuccess) { (h)->err = cudaGetErrorString(e0_); (h)->sticky = GB25_ERR_CUDA; return GB25_ERR_CUDA; } \
  } while (0)

// Anything that touches the state outside the step path discards what the last step left for the next one: the AB2
// epilogue's speculation and the "final since" events that let downloads start before the step has finished.
static inline void invalidate(Handle* h) { h->spec.valid = false; h->early_valid = false; }

static int check_async(Handle* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    h->err = std::string(what) + ": " + cudaGetErrorString(e);
    h->sticky = GB25_ERR_CUDA;
    return GB25_ERR_CUDA;
  }
  return GB25_OK;
}

// ------------------------------------------------------------------ stage timers (StageScope: gb25_internal.h)
static void drain_timers(Handle* h) {
  for (auto& s : h->timers) {
    for (size_t q = 0; q < s.used; q++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, s.ev[q].first, s.ev[q].second) == cudaSuccess) { s.total_ms += ms; s.calls++; }
    }
    s.used = 0;
  }
}

// ------------------------------------------------------------------ field table
struct FieldInfo { int lx, ly, lz; bool three_d; };
static const FieldInfo kFieldInfo[GB25_FIELD_COUNT] = {
    {1, 0, 0, true},  {0, 1, 0, true},  {0, 0, 1, true},  {0, 0, 0, true},  {0, 0, 0, true},  {0, 0, 0, true},   // u v w T S p
    {1, 0, 0, true},  {0, 1, 0, true},  {0, 0, 0, true},  {0, 0, 0, true},                                        // Gn
    {1, 0, 0, true},  {0, 1, 0, true},  {0, 0, 0, true},  {0, 0, 0, true},                                        // G-
    {0, 0, 1, false}, {1, 0, 0, false}, {0, 1, 0, false},                                                         // eta U V
    {0, 0, 1, false}, {1, 0, 0, false}, {0, 1, 0, false},                                                         // filtered
    {1, 0, 0, false}, {0, 1, 0, false}, {1, 0, 0, false}, {0, 1, 0, false}};                                      // Gn.U Gn.V G-.U G-.V

static void parent_shape(const Handle* h, int field, int shape[3]) {
  const FieldInfo& fi = kFieldInfo[field];
  const gb25_config& c = h->cfg;
  shape[0] = c.Nx + 2 * c.Hx;
  shape[1] = c.Ny + 2 * c.Hy + ((fi.ly && h->g.wall_n) ? 1 : 0);   // Bounded Face-y: N+1 points on the tile that owns the wall
  shape[2] = fi.three_d ? c.Nz + 2 * c.Hz + fi.lz : 1;
}

// Every allocation gb25_create makes goes through here.  With GB25_GUARD=1 (debug aid; compute-sanitizer is closed on the
// shared pool) the array sits between two 64 KiB guard zones filled with a byte pattern, and gb25_check_guards counts the
// guard bytes a kernel has overwritten: an out-of-bounds store within 64 KiB of any array is caught, at full speed, on every
// kernel generation.  Guarded allocations cannot be exported to other processes (IPC handles map whole allocations).
#define GB25_GUARD_BYTES ((size_t)64 << 10)
#define GB25_GUARD_BYTE 0xA5
static cudaError_t galloc(Handle* h, void** out, size_t bytes) {
  if (!h->guard) {
    const cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaSuccess) h->allocs.push_back(*out);
    return e;
  }
  char* base = nullptr;
  const size_t rounded = (bytes + 255) & ~(size_t)255;
  const cudaError_t e = cudaMalloc(&base, rounded + 2 * GB25_GUARD_BYTES);
  if (e != cudaSuccess) return e;
  cudaMemset(base, GB25_GUARD_BYTE, rounded + 2 * GB25_GUARD_BYTES);
  cudaMemset(base + GB25_GUARD_BYTES, 0, bytes);
  h->allocs.push_back(base);
  h->guards.push_back({base, GB25_GUARD_BYTES + bytes, rounded - bytes + GB25_GUARD_BYTES});
  *out = base + GB25_GUARD_BYTES;
  return cudaSuccess;
}
extern "C" int gb25_check_guards(gb25_handle* h, long* corrupted_bytes) {
  REQUIRE(h);
  if (!corrupted_bytes) return GB25_ERR_INVALID;
  if (!h->guard) { h->err = "gb25_check_guards: the handle was not created with GB25_GUARD=1"; return GB25_ERR_INVALID; }
  CK(h, cudaStreamSynchronize(h->stream));
  std::vector<unsigned char> buf;
  long bad = 0;
  for (const auto& gz : h->guards) {
    const char* zone[2] = {gz.base, gz.base + gz.lo_end};
    const size_t len[2] = {GB25_GUARD_BYTES, gz.hi_len};
    for (int z = 0; z < 2; z++) {
      buf.resize(len[z]);
      CK(h, cudaMemcpy(buf.data(), zone[z], len[z], cudaMemcpyDeviceToHost));
      for (unsigned char b : buf) bad += b != GB25_GUARD_BYTE;
    }
  }
  *corrupted_bytes = bad;
  return GB25_OK;
}

template <class T>
static int upload(Handle* h, const T* host, size_t n, const T** out) {
  T* d = nullptr;
  CK(h, galloc(h, (void**)&d, n * sizeof(T)));
  CK(h, cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
  *out = d;
  return GB25_OK;
}

extern "C" int gb25_abi_version(void) { return GB25_ABI_VERSION; }
extern "C" int gb25_real_bytes(void) { return (int)sizeof(real); }

extern "C" const char* gb25_last_error(const gb25_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
extern "C" int gb25_clear_error(gb25_handle* h) {
  if (!h) return GB25_ERR_INVALID;
  h->err.clear(); h->sticky = 0; cudaGetLastError();
  return GB25_OK;
}

// ------------------------------------------------------------------ create / destroy
static int build_immersed_products(Handle* h, const gb25_grid* grid) {
  const gb25_config& c = h->cfg;
  const int PX = c.Nx + 2 * c.Hx, PY = c.Ny + 2 * c.Hy + 1, PZ = c.Nz + 2 * c.Hz + 1;
  const int n2 = PX * PY;
  auto zf = [&](int k) { return grid->z_f[k + c.Hz - 1]; };
  auto zc = [&](int k) { return grid->z_c[k + c.Hz - 1]; };
  (void)PZ;
  std::vector<short> kb(n2, 0), kbe(n2, 0);
  std::vector<real> Hcc(n2), Hfc(n2), Hcf(n2);
  const real ztop = zf(c.Nz + 1), zbot = zf(1);
  for (int J = 0; J < PY; J++)
    for (int I = 0; I < PX; I++) {
      const int q = I + PX * J;
      int k0 = 0;
      if (c.immersed && grid->bottom_height) {
        const real bh = std::min(std::max(grid->bottom_height[q], zbot), ztop);
        for (int k = 1; k <= c.Nz; k++) if (zc(k) <= bh) k0 = k;
      }
      kb[q] = (short)k0;
      Hcc[q] = ztop - zf(k0 + 1);
      const int j = J - c.Hy + 1;
      const bool yout = (h->g.wall_s && j < 1) || (h->g.wall_n && j > c.Ny);
      kbe[q] = yout ? (short)GB25_BIG : (short)k0;
    }
  for (int J = 0; J < PY; J++)
    for (int I = 0; I < PX; I++) {
      const int q = I + PX * J, qw = std::max(I - 1, 0) + PX * J, qs = I + PX * std::max(J - 1, 0);
      Hfc[q] = std::min(Hcc[qw], Hcc[q]);
      Hcf[q] = std::min(Hcc[qs], Hcc[q]);
    }
  // order-reduction thresholds: buffer B allowed iff k > threshold (window maxima of the column rule)
  auto at = [&](const std::vector<short>& a, int I, int J) -> int {
    if (I < 0 || I >= PX || J < 0 || J >= PY) return GB25_BIG;  // beyond the stored halo: treat as solid
    return a[I + PX * J];
  };
  std::vector<short> fx3(n2), fx2(n2), fy3(n2), fy2(n2), cx3(n2), cx2(n2), cy3(n2), cy2(n2), knear(n2), ksolid(n2);
  for (int J = 0; J < PY; J++)
    for (int I = 0; I < PX; I++) {
      const int q = I + PX * J;
      int m;
      auto facemax = [&](int B, int dx, int dy) { int r = 0; for (int s = -B; s <= B - 1; s++) r = std::max(r, at(kbe, I + dx * s, J + dy * s)); return r; };
      auto nodemax = [&](int B, int dx, int dy) {
        int r = 0;
        for (int s = -B + 1; s <= B; s++) r = std::max(r, std::min(at(kbe, I + dx * (s - 1), J + dy * (s - 1)), at(kbe, I + dx * s, J + dy * s)));
        return r;
      };
      fx3[q] = (short)facemax(3, 1, 0); fx2[q] = (short)facemax(2, 1, 0);
      fy3[q] = (short)facemax(3, 0, 1); fy2[q] = (short)facemax(2, 0, 1);
      cx3[q] = (short)nodemax(3, 1, 0); cx2[q] = (short)nodemax(2, 1, 0);
      cy3[q] = (short)nodemax(3, 0, 1); cy2[q] = (short)nodemax(2, 0, 1);
      m = 0;
      for (int dj = -4; dj <= 4; dj++) for (int di = -4; di <= 4; di++) m = std::max(m, at(kbe, I + di, J + dj));
      knear[q] = (short)m;
      int mn = GB25_BIG;
      for (int dj = -4; dj <= 4; dj++) for (int di = -4; di <= 4; di++) {
        const int II = I + di, JJ = J + dj;
        mn = std::min(mn, (II < 0 || II >= PX || JJ < 0 || JJ >= PY) ? 0 : (int)kb[II + PX * JJ]);
      }
      ksolid[q] = (short)mn;
    }
  int rc;
  if ((rc = upload(h, kb.data(), n2, &h->g.kb))) return rc;
  if ((rc = upload(h, fx3.data(), n2, &h->g.fx3))) return rc;
  if ((rc = upload(h, fx2.data(), n2, &h->g.fx2))) return rc;
  if ((rc = upload(h, fy3.data(), n2, &h->g.fy3))) return rc;
  if ((rc = upload(h, fy2.data(), n2, &h->g.fy2))) return rc;
  if ((rc = upload(h, cx3.data(), n2, &h->g.cx3))) return rc;
  if ((rc = upload(h, cx2.data(), n2, &h->g.cx2))) return rc;
  if ((rc = upload(h, cy3.data(), n2, &h->g.cy3))) return rc;
  if ((rc = upload(h, cy2.data(), n2, &h->g.cy2))) return rc;
  if ((rc = upload(h, knear.data(), n2, &h->g.knear))) return rc;
  if ((rc = upload(h, ksolid.data(), n2, &h->g.ksolid))) return rc;
  // pair-based level ranges and the list of generic cells (see DevGrid)
  std::vector<short> kgen2(n2, 0), kzero2(n2, 0);
  std::vector<int> glist;
  for (int J = 0; J < PY; J++)
    for (int I = 0; I < PX; I++) {
      const int q = I + PX * J;
      const int i = I - c.Hx + 1;
      const int Ip = (i >= 1) ? I - ((i - 1) & 1) : I;          // first column of the aligned pair
      const int q0 = Ip + PX * J, q1 = std::min(Ip + 1, PX - 1) + PX * J;
      const int kg = std::min((int)std::max(knear[q0], knear[q1]), c.Nz);
      int kz = c.cond_diff ? std::min((int)ksolid[q0], (int)ksolid[q1]) - 3 : 0;
      kz = std::max(std::min(kz, kg), 0);
      kgen2[q] = (short)kg; kzero2[q] = (short)kz;
    }
  for (int j = 1; j <= c.Ny; j++)
    for (int k = 1; k <= c.Nz; k++)
      for (int i = 1; i <= c.Nx; i++) {
        const int q = (i + c.Hx - 1) + PX * (j + c.Hy - 1);
        if (k > kzero2[q] && k <= kgen2[q]) glist.push_back(q + n2 * (k + c.Hz - 1));
      }
  if ((rc = upload(h, kgen2.data(), n2, &h->g.kgen2))) return rc;
  if ((rc = upload(h, kzero2.data(), n2, &h->g.kzero2))) return rc;
  h->g.nglist = (int)glist.size();
  if (glist.empty()) glist.push_back(0);
  if ((rc = upload(h, glist.data(), glist.size(), &h->g.glist))) return rc;
  if ((rc = upload(h, Hfc.data(), n2, &h->g.Hfc))) return rc;
  if ((rc = upload(h, Hcf.data(), n2, &h->g.Hcf))) return rc;
  return GB25_OK;
}

// A partitioned handle waits for its neighbours inside the stream (stream memory operations have no time-out of their own):
// instead of blocking for ever on a neighbour that died, poll the stream and give up after GB25_SYNC_TIMEOUT_S seconds
// (default 600; 0 = block).  Returns cudaErrorNotReady on a time-out.
static double sync_limit_seconds() { const char* e = getenv("GB25_SYNC_TIMEOUT_S"); return e ? atof(e) : 600.0; }
static cudaError_t stream_drain(Handle* h, cudaStream_t s) {
  const double limit = sync_limit_seconds();
  if (!h->ex.on || limit <= 0.0) return cudaStreamSynchronize(s);
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    const cudaError_t e = cudaStreamQuery(s);
    if (e != cudaErrorNotReady) return e;
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit) return cudaErrorNotReady;
    std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
}

extern "C" int gb25_destroy(gb25_handle* h) {
  if (!h) return GB25_ERR_INVALID;
  cudaSetDevice(h->device);
  if (h->stream && h->ex.on) {
    // a stream that still waits for a neighbour tile would block cudaStreamSynchronize / cudaFree for ever: give up after the
    // time-out and leave the device memory of this handle to the end of the process
    bool stuck = stream_drain(h, h->stream) == cudaErrorNotReady;
    if (!stuck && h->stream2) stuck = stream_drain(h, h->stream2) == cudaErrorNotReady;
    if (stuck) {
      fprintf(stderr, "gb25_destroy: a stream of this tile still waits for a neighbour tile; its device memory is not released\n");
      cudaGetLastError();
      return GB25_ERR_COMM;
    }
  }
  if (h->stream) cudaStreamSynchronize(h->stream);
  exchange_close(h);
  tma_free(h);
  baro_plan_free(h);
  for (void* p : h->allocs) cudaFree(p);
  for (auto& s : h->timers) for (auto& e : s.ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  if (h->loop_start) cudaEventDestroy(h->loop_start);
  if (h->loop_stop) cudaEventDestroy(h->loop_stop);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->stream_d2h) { cudaStreamSynchronize(h->stream_d2h); cudaStreamDestroy(h->stream_d2h); }
  if (h->ev_ts_final) cudaEventDestroy(h->ev_ts_final);
  if (h->ev_uv_final) cudaEventDestroy(h->ev_uv_final);
  delete h;
  return GB25_OK;
}

// Lazy module loading (the CUDA 12 default) loads a module / kernel at its first use, and that load may synchronise the
// context.  A partitioned step enqueues waits for the neighbour tiles; a first use behind such a wait blocks the host until
// the neighbour has signalled — which never happens when ONE host thread drives all tiles (gb25_exchange_connect_local),
// because the neighbour's step is enqueued by the same thread afterwards.  So every kernel of the library is loaded on a
// device when the first handle is created there (cudaFuncGetAttributes is the documented way to preload under the runtime API).
static int kernel_tables(KernelTable t[5]) {
  t[0] = kernel_table_core(); t[1] = kernel_table_tend_v2(); t[2] = kernel_table_tend_tma(); t[3] = kernel_table_exchange();
  t[4] = kernel_table_baro();
  return 5;
}
extern "C" int gb25_kernel_table_size(void) {
  KernelTable t[5];
  const int nt = kernel_tables(t);
  int n = 0;
  for (int q = 0; q < nt; q++) n += t[q].n;
  return n;
}
static cudaError_t preload_kernels(int device) {
  static std::mutex mu;
  static std::vector<int> done;
  std::lock_guard<std::mutex> lock(mu);
  for (int d : done) if (d == device) return cudaSuccess;
  KernelTable t[5];
  const int nt = kernel_tables(t);
  for (int q = 0; q < nt; q++)
    for (int k = 0; k < t[q].n; k++) {
      cudaFuncAttributes a;
      const cudaError_t e = cudaFuncGetAttributes(&a, t[q].fn[k]);
      if (e != cudaSuccess) return e;
    }
  done.push_back(device);
  return cudaSuccess;
}

extern "C" int gb25_create(const gb25_config* cfg, const gb25_grid* grid, gb25_handle** out) {
  if (out) *out = nullptr;
  if (!cfg || !grid || !out) { g_create_error = "gb25_create: null argument"; return GB25_ERR_INVALID; }
  if (cfg->Nx < 8 || cfg->Ny < 8 || cfg->Nz < 1 || cfg->Hx < 4 || cfg->Hy < 4 || cfg->Hz < 4 || cfg->nsubsteps < 1 ||
      cfg->nsubsteps > 256 || cfg->Nz > 32000) {
    g_create_error = "gb25_create: unsupported sizes (need Nx,Ny >= 8, halo >= 4, 1 <= nsubsteps <= 256)";
    return GB25_ERR_INVALID;
  }
  if (cfg->closure < 0 || cfg->closure > 2 || (cfg->closure && (cfg->kappa < 0.f || cfg->nu < 0.f))) {
    g_create_error = "gb25_create: closure must be 0 (nothing), 1 (explicit) or 2 (vertically implicit) with kappa, nu >= 0";
    return GB25_ERR_INVALID;
  }
  if (cfg->topo_y == GB25_TOPO_FOLD && (cfg->Nx % 2)) { g_create_error = "gb25_create: tripolar fold needs even Nx"; return GB25_ERR_INVALID; }
  if (!grid->dx_cc || !grid->z_f || !grid->z_c || !grid->dz_c || !grid->dz_f || !grid->avg_weights || !grid->f_ff) {
    g_create_error = "gb25_create: missing grid array"; return GB25_ERR_INVALID;
  }
  if (cfg->Rx < 1 || cfg->Ry < 1 || cfg->rx < 0 || cfg->rx >= cfg->Rx || cfg->ry < 0 || cfg->ry >= cfg->Ry) {
    g_create_error = "gb25_create: bad partition (Rx, Ry, rx, ry)"; return GB25_ERR_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("gb25_create: no CUDA device available (libgb25cuda has no CPU path): ") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    cudaGetLastError();
    return GB25_ERR_NO_DEVICE;
  }
  Handle* h = new Handle();
  h->timers.reserve(64);
  { const char* gd = getenv("GB25_GUARD"); h->guard = gd && gd[0] == '1'; }
  h->cfg = *cfg;
  if (cfg->device >= 0) h->device = cfg->device; else cudaGetDevice(&h->device);
  if ((e = cudaSetDevice(h->device)) != cudaSuccess) {
    g_create_error = std::string("gb25_create: cudaSetDevice: ") + cudaGetErrorString(e);
    delete h; return GB25_ERR_NO_DEVICE;
  }
  if ((e = preload_kernels(h->device)) != cudaSuccess) {
    g_create_error = std::string("gb25_create: loading the kernels (is this an sm_100a device?): ") + cudaGetErrorString(e);
    cudaGetLastError();
    delete h; return GB25_ERR_CUDA;
  }
#define CKC(call)                                                                         \
  do { int rc_ = (call); if (rc_ != GB25_OK) { g_create_error = h->err; gb25_destroy(h); return rc_; } } while (0)
  auto ckcuda = [&](cudaError_t ce, const char* what) -> int {
    if (ce != cudaSuccess) { h->err = std::string(what) + ": " + cudaGetErrorString(ce); return GB25_ERR_CUDA; }
    return GB25_OK;
  };
  CKC(ckcuda(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate"));
  CKC(ckcuda(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking), "cudaStreamCreate"));
  CKC(ckcuda(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming), "cudaEventCreate"));
  CKC(ckcuda(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming), "cudaEventCreate"));
  CKC(ckcuda(cudaStreamCreateWithFlags(&h->stream_d2h, cudaStreamNonBlocking), "cudaStreamCreate"));
  CKC(ckcuda(cudaEventCreateWithFlags(&h->ev_ts_final, cudaEventDisableTiming), "cudaEventCreate"));
  CKC(ckcuda(cudaEventCreateWithFlags(&h->ev_uv_final, cudaEventDisableTiming), "cudaEventCreate"));
  CKC(ckcuda(cudaEventCreate(&h->loop_start), "cudaEventCreate"));
  CKC(ckcuda(cudaEventCreate(&h->loop_stop), "cudaEventCreate"));
  DevGrid& g = h->g;
  g.Nx = cfg->Nx; g.Ny = cfg->Ny; g.Nz = cfg->Nz; g.Hx = cfg->Hx; g.Hy = cfg->Hy; g.Hz = cfg->Hz;
  g.PX = g.Nx + 2 * g.Hx; g.PY = g.Ny + 2 * g.Hy + 1; g.PZ = g.Nz + 2 * g.Hz + 1;
  if ((double)g.PX * g.PY * g.PZ > 2.0e9) { g_create_error = "gb25_create: tile too large for 32-bit plane offsets"; gb25_destroy(h); return GB25_ERR_INVALID; }
  g.n2 = g.PX * g.PY;
  g.topo_y = cfg->topo_y; g.immersed = cfg->immersed && grid->bottom_height; g.coriolis_scheme = cfg->coriolis_scheme;
  g.fold_variant = cfg->fold_variant; g.south_inactive = cfg->south_inactive; g.cond_diff = cfg->cond_diff; g.eos_r0 = cfg->eos_r0;
  g.g = cfg->g; g.rho0 = cfg->rho0; g.eps = cfg->weno_eps;
  g.wall_s = (cfg->ry == 0) && (cfg->topo_y == GB25_TOPO_BOUNDED || cfg->south_inactive);
  g.wall_n = (cfg->ry == cfg->Ry - 1) && (cfg->topo_y == GB25_TOPO_BOUNDED);
  h->cfg.immersed = g.immersed;
  const size_t n2 = g.n2, n3 = n2 * g.PZ;
  const real* src2[13] = {grid->dx_cc, grid->dx_fc, grid->dx_cf, grid->dx_ff, grid->dy_cc, grid->dy_fc, grid->dy_cf, grid->dy_ff,
                           grid->az_cc, grid->az_fc, grid->az_cf, grid->az_ff, grid->f_ff};
  const real** dst2[13] = {&g.dxcc, &g.dxfc, &g.dxcf, &g.dxff, &g.dycc, &g.dyfc, &g.dycf, &g.dyff, &g.azcc, &g.azfc, &g.azcf, &g.azff, &g.fff};
  for (int a = 0; a < 13; a++) {
    if (!src2[a]) { g_create_error = "gb25_create: missing metric array"; gb25_destroy(h); return GB25_ERR_INVALID; }
    CKC(upload(h, src2[a], n2, dst2[a]));
  }
  const real* srcz[4] = {grid->z_f, grid->z_c, grid->dz_c, grid->dz_f};
  const real** dstz[4] = {&g.zf, &g.zc, &g.dzc, &g.dzf};
  for (int a = 0; a < 4; a++) CKC(upload(h, srcz[a], (size_t)g.PZ, dstz[a]));
  CKC(build_immersed_products(h, grid));
  {
    const DevGrid* gd = nullptr;
    CKC(upload(h, &h->g, 1, &gd));
    h->g_dev = const_cast<DevGrid*>(gd);
  }
  h->weights.assign(grid->avg_weights, grid->avg_weights + cfg->nsubsteps);
  // fields
  for (int fidx = 0; fidx < GB25_FIELD_COUNT; fidx++) {
    const size_t n = kFieldInfo[fidx].three_d ? n3 : n2;
    real* d = nullptr;
    cudaError_t ce = galloc(h, (void**)&d, n * sizeof(real));
    if (ce != cudaSuccess) { g_create_error = std::string("gb25_create: cudaMalloc field: ") + cudaGetErrorString(ce); gb25_destroy(h); return GB25_ERR_ALLOC; }
    CKC(ckcuda(cudaMemsetAsync(d, 0, n * sizeof(real), h->stream), "cudaMemset"));
    h->field_ptr[fidx] = d;
  }
  for (int q = 0; q < 4; q++) {
    static const int ids[4] = {GB25_U, GB25_V, GB25_T, GB25_S};
    h->state_buf[0][q] = h->field_ptr[ids[q]];
    cudaError_t ce = galloc(h, (void**)&h->state_buf[1][q], n3 * sizeof(real));
    if (ce != cudaSuccess) { g_create_error = std::string("gb25_create: cudaMalloc state buffer: ") + cudaGetErrorString(ce); gb25_destroy(h); return GB25_ERR_ALLOC; }
    CKC(ckcuda(cudaMemsetAsync(h->state_buf[1][q], 0, n3 * sizeof(real), h->stream), "cudaMemset"));
  }
  for (real** sp : {&h->zeta, &h->dxU, &h->dyV}) {
    cudaError_t ce = galloc(h, (void**)sp, n3 * sizeof(real));
    if (ce != cudaSuccess) { g_create_error = std::string("gb25_create: cudaMalloc scratch: ") + cudaGetErrorString(ce); gb25_destroy(h); return GB25_ERR_ALLOC; }
    CKC(ckcuda(cudaMemsetAsync(*sp, 0, n3 * sizeof(real), h->stream), "cudaMemset"));
  }
  for (real** sp : {&h->us2, &h->vs2, &h->corr_u, &h->corr_v, &h->carry[0], &h->carry[1], &h->carry[2], &h->carry[3],
                     &h->spec2d[0], &h->spec2d[1], &h->spec2d[2], &h->spec2d[3]}) {
    cudaError_t ce = galloc(h, (void**)sp, n2 * sizeof(real));
    if (ce != cudaSuccess) { g_create_error = std::string("gb25_create: cudaMalloc scratch: ") + cudaGetErrorString(ce); gb25_destroy(h); return GB25_ERR_ALLOC; }
    CKC(ckcuda(cudaMemsetAsync(*sp, 0, n2 * sizeof(real), h->stream), "cudaMemset"));
  }
  {
    const char* e = getenv("GB25_FUSED");
    h->use_fused = !(e && e[0] == '0');
#ifdef GB25_F64
    h->use_fused = false;     // Float64 build: the operator-per-kernel generation only
#endif
    const char* t = getenv("GB25_TMA");
    h->use_tma = !(t && t[0] == '0');
    const char* sp = getenv("GB25_SPECULATE");
    h->use_spec = !(sp && sp[0] == '0');
    const char* ov = getenv("GB25_OVERLAP");
    h->use_overlap = !(ov && ov[0] == '0');
    const char* zf = getenv("GB25_ZHALO_FOLD");
    h->use_zfold = !(zf && zf[0] == '0');
    const char* pk = getenv("GB25_PACKED");
    h->use_packed = !(pk && pk[0] == '0');
    const char* tt = getenv("GB25_TMA_TRACER");
    h->use_tma_tracer = !(tt && tt[0] == '0');
    const char* bp = getenv("GB25_BARO_PERSISTENT");
    h->use_baro_persistent = !(bp && bp[0] == '0');
  }
  DevFields& f = h->f;
  f.u = h->field_ptr[GB25_U]; f.v = h->field_ptr[GB25_V]; f.w = h->field_ptr[GB25_W];
  f.T = h->field_ptr[GB25_T]; f.S = h->field_ptr[GB25_S]; f.p = h->field_ptr[GB25_P];
  for (int q = 0; q < 4; q++) { f.gn[q] = h->field_ptr[GB25_GN_U + q]; f.gm[q] = h->field_ptr[GB25_GM_U + q]; }
  f.eta = h->field_ptr[GB25_ETA]; f.bu = h->field_ptr[GB25_BARO_U]; f.bv = h->field_ptr[GB25_BARO_V];
  f.feta = h->field_ptr[GB25_FILT_ETA]; f.fu = h->field_ptr[GB25_FILT_U]; f.fv = h->field_ptr[GB25_FILT_V];
  f.gU = h->field_ptr[GB25_GN_BARO_U]; f.gV = h->field_ptr[GB25_GN_BARO_V];
  f.gmU = h->field_ptr[GB25_GM_BARO_U]; f.gmV = h->field_ptr[GB25_GM_BARO_V];
  CKC(ckcuda(cudaStreamSynchronize(h->stream), "create sync"));
#undef CKC
  *out = h;
  return GB25_OK;
}

// ------------------------------------------------------------------ transfers (parent shape <-> internal layout)
extern "C" int gb25_field_shape(const gb25_handle* h, int field, int shape[3]) {
  if (!h || field < 0 || field >= GB25_FIELD_COUNT || !shape) return GB25_ERR_INVALID;
  parent_shape(h, field, shape);
  return GB25_OK;
}
static void interior_shape(const Handle* h, int field, int shape[3]) {
  const FieldInfo& fi = kFieldInfo[field];
  const gb25_config& c = h->cfg;
  shape[0] = c.Nx;
  shape[1] = c.Ny + ((fi.ly && h->g.wall_n) ? 1 : 0);
  shape[2] = fi.three_d ? c.Nz + fi.lz : 1;
}
extern "C" int gb25_interior_shape(const gb25_handle* h, int field, int shape[3]) {
  if (!h || field < 0 || field >= GB25_FIELD_COUNT || !shape) return GB25_ERR_INVALID;
  interior_shape(h, field, shape);
  return GB25_OK;
}
// one pitched 3-D copy between a host array (parent or interior shape) and the internal (PX, PY, PZ) layout
static int copy_field(Handle* h, int field, real* host, bool to_device, bool interior = false, bool sync = true) {
  if (field < 0 || field >= GB25_FIELD_COUNT || !host) { h->err = "bad field id or null buffer"; return GB25_ERR_INVALID; }
  int s[3];
  if (interior) interior_shape(h, field, s); else parent_shape(h, field, s);
  const DevGrid& g = h->g;
  const bool three_d = kFieldInfo[field].three_d;
  cudaMemcpy3DParms p = {};
  p.extent = make_cudaExtent((size_t)s[0] * sizeof(real), s[1], s[2]);
  cudaPitchedPtr hp = make_cudaPitchedPtr(host, (size_t)s[0] * sizeof(real), s[0], s[1]);
  cudaPitchedPtr dp = make_cudaPitchedPtr(h->field_ptr[field], (size_t)g.PX * sizeof(real), g.PX, g.PY);
  const cudaPos dpos = interior ? make_cudaPos((size_t)g.Hx * sizeof(real), g.Hy, three_d ? g.Hz : 0) : make_cudaPos(0, 0, 0);
  if (to_device) { p.srcPtr = hp; p.dstPtr = dp; p.dstPos = dpos; p.kind = cudaMemcpyHostToDevice; }
  else { p.srcPtr = dp; p.srcPos = dpos; p.dstPtr = hp; p.kind = cudaMemcpyDeviceToHost; }
  // Interior-shaped downloads of the prognostic fields right after a step need not wait for the step's tail: the interiors
  // of T, S are final once the AB2 stage is done, those of u, v, eta, U, V once the corrector is (the tendency kernels, 70 %
  // of the step, only read them), so the copy runs on its own stream behind the matching event, under those kernels.
  if (!to_device && interior && h->early_valid) {
    cudaEvent_t ev = nullptr;
    if (field == GB25_T || field == GB25_S) ev = h->ev_ts_final;
    else if (field == GB25_U || field == GB25_V || field == GB25_ETA || field == GB25_BARO_U || field == GB25_BARO_V) ev = h->ev_uv_final;
    if (ev) {
      CK(h, cudaStreamWaitEvent(h->stream_d2h, ev, 0));
      CK(h, cudaMemcpy3DAsync(&p, h->stream_d2h));
      h->d2h_pending = true;
      if (sync) { CK(h, cudaStreamSynchronize(h->stream_d2h)); h->d2h_pending = false; }
      return GB25_OK;
    }
  }
  CK(h, cudaMemcpy3DAsync(&p, h->stream));
  if (to_device) {
    invalidate(h);
    // the other half of a double-buffered field receives the same parent, so that halo cells no fill ever writes
    // (y-z corners, the rows behind an impenetrable wall) hold the uploaded values whichever buffer is current
    // (an interior upload touches no halo cell, and the other half's interior is rewritten before it is read)
    if (!interior)
      for (int q = 0; q < 4; q++)
        if (h->field_ptr[field] == h->state_buf[h->parity][q]) {
          p.dstPtr = make_cudaPitchedPtr(h->state_buf[1 - h->parity][q], (size_t)g.PX * sizeof(real), g.PX, g.PY);
          CK(h, cudaMemcpy3DAsync(&p, h->stream));
        }
  }
  if (sync) CK(h, cudaStreamSynchronize(h->stream));
  return GB25_OK;
}
extern "C" int gb25_set_field(gb25_handle* h, int field, const real* host_parent) {
  REQUIRE(h);
  return copy_field(h, field, const_cast<real*>(host_parent), true);
}
extern "C" int gb25_get_field(gb25_handle* h, int field, real* host_parent) {
  REQUIRE(h);
  return copy_field(h, field, host_parent, false);
}
extern "C" int gb25_set_interior(gb25_handle* h, int field, const real* host_interior) {
  REQUIRE(h);
  return copy_field(h, field, const_cast<real*>(host_interior), true, true);
}
extern "C" int gb25_get_interior(gb25_handle* h, int field, real* host_interior) {
  REQUIRE(h);
  return copy_field(h, field, host_interior, false, true);
}
extern "C" int gb25_set_fields(gb25_handle* h, int n, const int* fields, const real* const* host, int interior) {
  REQUIRE(h);
  if (n < 0 || (n && (!fields || !host))) { h->err = "gb25_set_fields: null argument"; return GB25_ERR_INVALID; }
  for (int q = 0; q < n; q++) {
    const int rc = copy_field(h, fields[q], const_cast<real*>(host[q]), true, interior != 0, false);
    if (rc != GB25_OK) return rc;
  }
  CK(h, cudaStreamSynchronize(h->stream));     // the host buffers are borrowed for the duration of the call only
  return GB25_OK;
}
extern "C" int gb25_get_fields(gb25_handle* h, int n, const int* fields, real* const* host, int interior) {
  REQUIRE(h);
  if (n < 0 || (n && (!fields || !host))) { h->err = "gb25_get_fields: null argument"; return GB25_ERR_INVALID; }
  for (int q = 0; q < n; q++) {
    const int rc = copy_field(h, fields[q], host[q], false, interior != 0, false);
    if (rc != GB25_OK) return rc;
  }
  if (h->d2h_pending) { CK(h, cudaStreamSynchronize(h->stream_d2h)); h->d2h_pending = false; }
  CK(h, cudaStreamSynchronize(h->stream));
  return GB25_OK;
}
extern "C" int gb25_set_clock(gb25_handle* h, double time, long iteration, float last_dt) {
  if (!h) return GB25_ERR_INVALID;
  h->time = time; h->iteration = iteration; h->last_dt = last_dt;
  return GB25_OK;
}
extern "C" int gb25_get_clock(const gb25_handle* h, double* time, long* iteration, float* last_dt) {
  if (!h) return GB25_ERR_INVALID;
  if (time) *time = h->time;
  if (iteration) *iteration = h->iteration;
  if (last_dt) *last_dt = h->last_dt;
  return GB25_OK;
}
static int sync_with_timeout(Handle* h) {
  const cudaError_t e = stream_drain(h, h->stream);
  if (e == cudaErrorNotReady) {
    h->err = "gb25_synchronize: timed out waiting for the stream (a neighbouring tile has not arrived)";
    h->sticky = GB25_ERR_COMM;
    return GB25_ERR_COMM;
  }
  CK(h, e);
  return GB25_OK;
}
extern "C" int gb25_synchronize(gb25_handle* h) {
  REQUIRE(h);
  { const int rc = sync_with_timeout(h); if (rc != GB25_OK) return rc; }
  if (exchange_check_timeout(h)) { h->err = "halo exchange timed out waiting for a neighbour tile"; h->sticky = GB25_ERR_COMM; return GB25_ERR_COMM; }
  if (baro_check_timeout(h)) { h->err = "split-explicit substeps timed out waiting for a neighbouring band or tile"; h->sticky = GB25_ERR_COMM; return GB25_ERR_COMM; }
  return check_async(h, "gb25_synchronize");
}

// ------------------------------------------------------------------ stages
static void fill_prognostic(Handle* h) {
  StageScope t(h, "fill_halo_regions");
  DevFields& f = h->f;
  HaloSpec s3[4] = {{f.u, 1, 0, 0, -1.f}, {f.v, 0, 1, 0, -1.f}, {f.T, 0, 0, 0, 1.f}, {f.S, 0, 0, 0, 1.f}};
  launch_fill_halo(h, s3, 4, true);
  HaloSpec s2[3] = {{f.eta, 0, 0, 1, 1.f}, {f.bu, 1, 0, 0, -1.f}, {f.bv, 0, 1, 0, -1.f}};
  launch_fill_halo(h, s2, 3, false);
}
// fused step path: ONE batch for the 3-D prognostic fields and the five 2-D fields whose halos the reference fills at
// three different points of the step (GU,GV before the substeps, U,V after them, eta,U,V in update_state!).  Nothing reads
// those halos in between (the substep kernels take GU,GV at interior points and start from the halos of the previous
// step's fill), and the later fill overwrites every halo cell of the earlier ones, so the state after the step is the same —
// with one exchange sequence between tiles instead of four.
static void fill_prognostic_fused(Handle* h, bool zdone_uv, bool zdone_ts, bool ts_done) {
  StageScope t(h, "fill_halo_regions");
  DevFields& f = h->f;
  const int zu = zdone_uv ? 1 : 0, zv = zdone_uv ? 2 : 0, zt = zdone_ts ? 1 : 0;
  if (ts_done) {     // T, S were filled ahead of the barotropic solve (fill_tracers_early)
    HaloSpec s[7] = {{f.u, 1, 0, 0, -1.f, 0, zu}, {f.v, 0, 1, 0, -1.f, 0, zv},
                     {f.eta, 0, 0, 1, 1.f, 1}, {f.bu, 1, 0, 0, -1.f, 1}, {f.bv, 0, 1, 0, -1.f, 1},
                     {f.gU, 1, 0, 0, -1.f, 1}, {f.gV, 0, 1, 0, -1.f, 1}};
    launch_fill_halo(h, s, 7, true);
    return;
  }
  HaloSpec s[9] = {{f.u, 1, 0, 0, -1.f, 0, zu}, {f.v, 0, 1, 0, -1.f, 0, zv}, {f.T, 0, 0, 0, 1.f, 0, zt}, {f.S, 0, 0, 0, 1.f, 0, zt},
                   {f.eta, 0, 0, 1, 1.f, 1}, {f.bu, 1, 0, 0, -1.f, 1}, {f.bv, 0, 1, 0, -1.f, 1},
                   {f.gU, 1, 0, 0, -1.f, 1}, {f.gV, 0, 1, 0, -1.f, 1}};
  launch_fill_halo(h, s, 9, true);
}
struct StreamSwap {   // launch on another stream for the lifetime of the object
  Handle* h; cudaStream_t saved;
  StreamSwap(Handle* h_, cudaStream_t s) : h(h_), saved(h_->stream) { h->stream = s; }
  ~StreamSwap() { h->stream = saved; }
};
// T and S are final the moment the AB2 epilogue of the previous step has written them (masked, z halos included): their
// horizontal halo fill / tile exchange and the hydrostatic pressure scan depend on nothing the barotropic solve or the
// corrector produce, so they run on the second stream underneath those two stages.  Between tiles this is the second lane
// of the exchange (own sequence numbers, flag words and inbox slots), so that the strips of T and S cross NVLink while the
// persistent substep kernel is latency-bound, and the exchange on the critical path carries two 3-D fields instead of four.
static void fill_tracers_early(Handle* h, bool zdone_ts) {
  DevFields& f = h->f;
  cudaEventRecord(h->ev_fork, h->stream);
  cudaStreamWaitEvent(h->stream2, h->ev_fork, 0);
  StreamSwap sw(h, h->stream2);
  h->ex.lane = 1;
  {
    StageScope t(h, "fill_halo_regions_TS");
    const int zt = zdone_ts ? 1 : 0;
    HaloSpec s3[2] = {{f.T, 0, 0, 0, 1.f, 0, zt}, {f.S, 0, 0, 0, 1.f, 0, zt}};
    launch_fill_halo(h, s3, 2, true);
  }
  h->ex.lane = 0;
  { StageScope t(h, "update_hydrostatic_pressure"); launch_compute_p(h); }
  cudaEventRecord(h->ev_join, h->stream2);
}
static void stage_mask(Handle* h) { StageScope t(h, "mask_immersed_fields"); launch_mask(h, false); }
static void stage_aux(Handle* h) {
  if (h->use_fused && h->g.Nx % 2 == 0) { StageScope t(h, "compute_w_from_continuity"); launch_aux_columns(h); }
  else { StageScope t(h, "compute_w_from_continuity"); launch_compute_w(h); }
  { StageScope t(h, "update_hydrostatic_pressure"); launch_compute_p(h); }
}
static void stage_tend(Handle* h, const Ab2Spec* spec = nullptr) {
  { StageScope t(h, "momentum_tendencies"); launch_momentum_tendency(h, spec); }
  { StageScope t(h, "tracer_tendencies"); launch_tracer_tendency(h, spec); }
  if (h->cfg.closure == 1) { StageScope t(h, "vertical_diffusion"); launch_vdiff_explicit(h); }
  if (h->has_bflux) { StageScope t(h, "boundary_tendencies"); launch_boundary_tendencies(h); }
}
static void stage_update_state(Handle* h) {
  stage_mask(h);
  fill_prognostic(h);
  stage_aux(h);
  stage_tend(h);
}
static void stage_ab2(Handle* h, float dt, float chi) {
  DevFields& f = h->f;
  { StageScope t(h, "ab2_step_fields"); launch_ab2_columns(h, dt, chi); }
  if (h->cfg.closure == 2) { StageScope t(h, "vertical_diffusion"); launch_implicit_columns(h, dt, false); }
  {
    StageScope t(h, "split_explicit_free_surface");
    HaloSpec sg[2] = {{f.gU, 1, 0, 0, -1.f}, {f.gV, 0, 1, 0, -1.f}};
    launch_fill_halo(h, sg, 2, false);
    launch_barotropic(h, dt);
    launch_mask(h, true);
    HaloSpec sb[2] = {{f.bu, 1, 0, 0, -1.f}, {f.bv, 0, 1, 0, -1.f}};
    launch_fill_halo(h, sb, 2, false);
  }
}
static void stage_correct(Handle* h) { StageScope t(h, "correct_velocities_and_cache"); launch_correct_cache(h); }
static void stage_initialize(Handle* h) {
  StageScope t(h, "initialize");
  launch_barotropic_mode(h);
  HaloSpec sb[2] = {{h->f.bu, 1, 0, 0, -1.f}, {h->f.bv, 0, 1, 0, -1.f}};
  launch_fill_halo(h, sb, 2, false);
}
static void swap_state_buffers(Handle* h) {
  static const int ids[4] = {GB25_U, GB25_V, GB25_T, GB25_S};
  h->parity ^= 1;
  DevFields& f = h->f;
  f.u = h->state_buf[h->parity][0]; f.v = h->state_buf[h->parity][1]; f.T = h->state_buf[h->parity][2]; f.S = h->state_buf[h->parity][3];
  for (int q = 0; q < 4; q++) h->field_ptr[ids[q]] = h->state_buf[h->parity][q];
}
// fused step path: identical results, fewer passes over the 3-D state (see gb25_kernels.cu "Fused step path")
static void one_time_step_fused(Handle* h, float dt, float chi) {
  DevFields& f = h->f;
  bool zdone_ts = false, ts_early = false;
  if (h->spec.valid && h->spec.dt == dt && h->spec.chi == chi) {
    zdone_ts = h->spec.zhalo;
    ts_early = h->use_overlap && h->cfg.closure != 2;
    // the tendency kernels of the previous step already wrote u*, v*, T', S' (masked) into the other state buffers and
    // GU, GV, sum dz u*, sum dz v* into their 2-D arrays: the AB2 stage is a pointer swap
    StageScope t(h, "ab2_step_fields");
    swap_state_buffers(h);
    launch_commit_spec(h);
  } else {
    StageScope t(h, "ab2_step_fields"); launch_ab2_fused(h, dt, chi);
  }
  h->spec.valid = false;
  if (ts_early) fill_tracers_early(h, zdone_ts);
  if (h->cfg.closure == 2) { StageScope t(h, "vertical_diffusion"); launch_implicit_columns(h, dt, true); }
  cudaEventRecord(h->ev_ts_final, h->stream);      // the interiors of T, S are final for this step from here on
  { StageScope t(h, "split_explicit_free_surface"); launch_barotropic(h, dt); }
  h->time += (double)dt; h->iteration += 1; h->last_dt = dt;
  bool zdone_uv;
  { StageScope t(h, "correct_velocities_and_cache"); zdone_uv = launch_correct_fused(h); }
  cudaEventRecord(h->ev_uv_final, h->stream);      // ... and those of u, v, eta, U, V from here on
  // G- <- Gn: swap the buffers; the tendency kernels below overwrite the whole interior of the new Gn,
  // and the halos of both are identically zero
  for (int q = 0; q < 4; q++) {
    std::swap(f.gn[q], f.gm[q]);
    std::swap(h->field_ptr[GB25_GN_U + q], h->field_ptr[GB25_GM_U + q]);
  }
  fill_prognostic_fused(h, zdone_uv, zdone_ts && h->cfg.closure != 2, ts_early);   // (the implicit solve rewrites T, S after the epilogue)
  if (ts_early) {
    { StageScope t(h, "compute_w_from_continuity"); launch_aux_columns(h); }
    cudaStreamWaitEvent(h->stream, h->ev_join, 0);
  } else {
    stage_aux(h);
  }
  if (spec_possible(h)) {
    const Ab2Spec sp = {dt, 1.5f + h->cfg.chi, 0.5f + h->cfg.chi, (h->use_zfold && h->g.Nz >= h->g.Hz) ? 1 : 0};
    stage_tend(h, &sp);
    h->spec.valid = true; h->spec.dt = dt; h->spec.chi = h->cfg.chi; h->spec.zhalo = sp.zhalo != 0;
  } else {
    stage_tend(h);
  }
  h->early_valid = true;
}
static void one_time_step(Handle* h, float dt, bool euler) {
  euler = euler || (dt != h->last_dt);
  const float chi = euler ? -0.5f : h->cfg.chi;
  if (h->use_fused) { one_time_step_fused(h, dt, chi); return; }
  stage_ab2(h, dt, chi);
  h->time += (double)dt; h->iteration += 1; h->last_dt = dt;
  stage_correct(h);
  stage_update_state(h);
}

extern "C" int gb25_initialize(gb25_handle* h) { REQUIRE(h); invalidate(h); stage_initialize(h); return check_async(h, "gb25_initialize"); }
extern "C" int gb25_update_state(gb25_handle* h) { REQUIRE(h); invalidate(h); stage_update_state(h); return check_async(h, "gb25_update_state"); }
extern "C" int gb25_first_time_step(gb25_handle* h, float dt) {
  REQUIRE(h);
  if (dt <= 0.f) dt = h->last_dt;
  invalidate(h);
  stage_initialize(h);
  stage_update_state(h);
  one_time_step(h, dt, true);
  return check_async(h, "gb25_first_time_step");
}
extern "C" int gb25_time_step(gb25_handle* h, float dt) {
  REQUIRE(h);
  if (dt <= 0.f) dt = h->last_dt;
  one_time_step(h, dt, false);
  return check_async(h, "gb25_time_step");
}
extern "C" int gb25_loop(gb25_handle* h, float dt, int nsteps) {
  REQUIRE(h);
  if (nsteps < 0) { h->err = "gb25_loop: negative step count"; return GB25_ERR_INVALID; }
  if (dt <= 0.f) dt = h->last_dt;
  CK(h, cudaEventRecord(h->loop_start, h->stream));
  for (int n = 0; n < nsteps; n++) one_time_step(h, dt, false);
  CK(h, cudaEventRecord(h->loop_stop, h->stream));
  h->loop_timed = true;
  return check_async(h, "gb25_loop");
}
extern "C" int gb25_mask_immersed_fields(gb25_handle* h) { REQUIRE(h); invalidate(h); stage_mask(h); return check_async(h, "gb25_mask_immersed_fields"); }
extern "C" int gb25_fill_halo_regions(gb25_handle* h) { REQUIRE(h); invalidate(h); fill_prognostic(h); return check_async(h, "gb25_fill_halo_regions"); }
extern "C" int gb25_compute_auxiliaries(gb25_handle* h) { REQUIRE(h); invalidate(h); stage_aux(h); return check_async(h, "gb25_compute_auxiliaries"); }
extern "C" int gb25_compute_tendencies(gb25_handle* h) { REQUIRE(h); invalidate(h); stage_tend(h); return check_async(h, "gb25_compute_tendencies"); }
extern "C" int gb25_compute_momentum_tendencies(gb25_handle* h) {
  REQUIRE(h);
  invalidate(h);
  { StageScope t(h, "momentum_tendencies"); launch_momentum_tendency(h); }
  return check_async(h, "gb25_compute_momentum_tendencies");
}
extern "C" int gb25_compute_tracer_tendencies(gb25_handle* h) {
  REQUIRE(h);
  invalidate(h);
  { StageScope t(h, "tracer_tendencies"); launch_tracer_tendency(h); }
  return check_async(h, "gb25_compute_tracer_tendencies");
}
extern "C" int gb25_compute_boundary_tendencies(gb25_handle* h) {
  REQUIRE(h);
  invalidate(h);
  { StageScope t(h, "boundary_tendencies"); launch_boundary_tendencies(h); }
  return check_async(h, "gb25_compute_boundary_tendencies");
}
extern "C" int gb25_set_flux_boundary_condition(gb25_handle* h, int field, int side, const real* flux) {
  REQUIRE(h);
  int q = -1;
  if (field == GB25_U) q = 0; else if (field == GB25_V) q = 1; else if (field == GB25_T) q = 2; else if (field == GB25_S) q = 3;
  if (q < 0 || side < 0 || side > 1) { h->err = "gb25_set_flux_boundary_condition: field must be GB25_U/V/T/S, side 0 (bottom) or 1 (top)"; return GB25_ERR_INVALID; }
  invalidate(h);
  CK(h, cudaStreamSynchronize(h->stream));
  if (!flux) {
    h->bflux[q][side] = nullptr;      // (the allocation stays in h->allocs until gb25_destroy)
  } else {
    int s[3];
    parent_shape(h, field == GB25_V ? GB25_BARO_V : (field == GB25_U ? GB25_BARO_U : GB25_ETA), s);   // 2-D parent of that staggering
    real* d = nullptr;
    CK(h, galloc(h, (void**)&d, (size_t)h->g.n2 * sizeof(real)));
    CK(h, cudaMemset(d, 0, (size_t)h->g.n2 * sizeof(real)));
    CK(h, cudaMemcpy2D(d, (size_t)h->g.PX * sizeof(real), flux, (size_t)s[0] * sizeof(real), (size_t)s[0] * sizeof(real), s[1], cudaMemcpyHostToDevice));
    h->bflux[q][side] = d;
  }
  h->has_bflux = false;
  for (int a = 0; a < 4; a++) for (int b = 0; b < 2; b++) h->has_bflux |= h->bflux[a][b] != nullptr;
  return GB25_OK;
}
extern "C" int gb25_ab2_step(gb25_handle* h, float dt, float chi) {
  REQUIRE(h);
  invalidate(h);
  if (dt <= 0.f) dt = h->last_dt;
  stage_ab2(h, dt, chi);
  return check_async(h, "gb25_ab2_step");
}
extern "C" int gb25_correct_velocities_and_cache_previous_tendencies(gb25_handle* h) {
  REQUIRE(h);
  invalidate(h);
  stage_correct(h);
  return check_async(h, "gb25_correct_velocities_and_cache_previous_tendencies");
}

// ------------------------------------------------------------------ measurement
extern "C" int gb25_last_loop_seconds(gb25_handle* h, double* seconds) {
  REQUIRE(h);
  if (!seconds || !h->loop_timed) { h->err = "gb25_last_loop_seconds: no loop has been timed"; return GB25_ERR_INVALID; }
  CK(h, cudaEventSynchronize(h->loop_stop));
  float ms = 0.f;
  CK(h, cudaEventElapsedTime(&ms, h->loop_start, h->loop_stop));
  *seconds = (double)ms * 1e-3;
  return GB25_OK;
}
extern "C" int gb25_kernel_launch_count(const gb25_handle* h, long* launches) {
  if (!h || !launches) return GB25_ERR_INVALID;
  *launches = h->launches;
  return GB25_OK;
}
extern "C" int gb25_enable_stage_timers(gb25_handle* h, int enable) {
  REQUIRE(h);
  CK(h, cudaStreamSynchronize(h->stream));
  drain_timers(h);
  h->timers_on = enable != 0;
  if (enable) for (auto& s : h->timers) { s.total_ms = 0.f; s.calls = 0; }
  return GB25_OK;
}
extern "C" int gb25_get_stage_times(gb25_handle* h, const char** names, float* ms, long* calls, int cap) {
  REQUIRE(h);
  CK(h, cudaStreamSynchronize(h->stream));
  drain_timers(h);
  int n = 0;
  for (auto& s : h->timers) {
    if (n >= cap) break;
    if (names) names[n] = s.name;
    if (ms) ms[n] = s.total_ms;
    if (calls) calls[n] = s.calls;
    n++;
  }
  return n;
}

// ------------------------------------------------------------------ multi-GPU exchange (see gb25_exchange.cu)
