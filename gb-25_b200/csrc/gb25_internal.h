// gb25_internal.h — host-side handle and launcher declarations shared by the .cu files.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/gb25cuda.h"
#include "gb25_device.cuh"

// flat = 1: a 2-D field riding in a batch of 3-D fields; zdone = 1: the z halos of the interior rows were already written
// by the kernel that produced the field (corrector, tracer epilogue); zdone = 2: except the wall row Ny+1 of a Face-y field
struct HaloSpec { real* a; int lx, ly, lz; real sign; int flat; int zdone; };

// ---- multi-GPU halo exchange over peer-mapped (CUDA IPC) memory, one process per GPU (gb25_exchange.cu)
// exported allocations: both halves of the double-buffered 3-D state (all tiles flip in lockstep, so a tile's current
// buffer always faces its neighbours' current buffers) and the 2-D fields
// EX_XBOX: the column inbox (west / east strips arrive packed, see k_push_cols_packed)
enum ExField { EX_U = 0, EX_V, EX_T, EX_S, EX_ETA, EX_BU, EX_BV, EX_GU, EX_GV, EX_U2, EX_V2, EX_T2, EX_S2, EX_XBOX, EX_NF };
enum ExSlot { SLOT_W = 0, SLOT_E, SLOT_S, SLOT_N, SLOT_FOLD, SLOT_FOLD2, EX_NSLOT };
#define EX_LANE_FLAGS 16   // flag words of lane 1 start here (lane 0: words 0 .. EX_NSLOT-1, time-out word EX_NSLOT)
struct ExPeer { real* fld[EX_NF]; int* flags; int rank; };
struct Exchange {
  bool on = false;
  int nranks = 1, rank = 0;
  int* flags = nullptr;          // local inbox: EX_NSLOT sequence numbers written by the neighbours + 1 error word
  real* xbox = nullptr;         // local column inbox: [seq parity][from west, from east][field slot][plane][row][Hx]
  size_t xbox_stride = 0;        // floats per (parity, direction) box
  ExPeer to[EX_NSLOT];           // the tile lying in that direction (destination of my pushes); rank < 0: none
  int from_mask_y = 0, from_mask_x = 0, from_mask_fold = 0;   // slots I receive on in each phase
  int seqs[2] = {0, 0}, xseqs[2] = {0, 0};   // sequence numbers per lane (all phases / column phases)
  int lane = 0;                              // 0: the step's main exchange; 1: the early T, S exchange on the second stream
  std::vector<void*> opened;     // pointers returned by cudaIpcOpenMemHandle
};

struct StageTimer {
  const char* name;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  size_t used = 0;
  float total_ms = 0.f;
  long calls = 0;
};

struct gb25_handle {
  gb25_config cfg;
  DevGrid g;
  DevGrid* g_dev = nullptr;   // copy of g in device memory (for non-inlined device functions)
  DevFields f;
  cudaStream_t stream = nullptr;
  int device = 0;
  std::vector<float> weights;
  std::vector<void*> allocs;
  // GB25_GUARD=1: every allocation of gb25_create sits between two guard zones (gb25_check_guards)
  bool guard = false;
  struct GuardZone { char* base; size_t lo_end, hi_len; };   // zones: [base, base + 64 KiB) and [base + lo_end, base + lo_end + hi_len)
  std::vector<GuardZone> guards;
  real* field_ptr[GB25_FIELD_COUNT];
  // scratch 3-D arrays shared by the v2 kernels: vorticity (F,F,C), delta_x(Ax u) and delta_y(Ay v) at (C,C,C)
  real *zeta = nullptr, *dxU = nullptr, *dyV = nullptr;
  real *us2 = nullptr, *vs2 = nullptr;   // 2-D: column sums of the AB2-updated, masked velocities (fused path)
  real *corr_u = nullptr, *corr_v = nullptr;   // 2-D: unmasked barotropic transports for the streamed corrector
  real* carry[4] = {nullptr, nullptr, nullptr, nullptr};   // 2-D: vertical flux through the top face of the topmost generic cell (u, v, T, S)
  // Double-buffered prognostic 3-D state.  The tendency kernels that end a step can apply the AB2 update of the NEXT step in
  // their epilogue (same dt, chi = cfg.chi): they read the state from state_buf[parity] and write the updated, masked state
  // into state_buf[1 - parity] (neighbouring tiles still read the old one), together with the column sums the barotropic
  // solve and the corrector need.  The next gb25_time_step with matching (dt, chi) then starts with a pointer swap instead of
  // the AB2 pass; anything else (another dt, an upload, an operator-level call) discards the speculation.
  real* state_buf[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};   // [parity][u, v, T, S]
  int parity = 0;
  real* spec2d[4] = {nullptr, nullptr, nullptr, nullptr};   // speculative GU, GV, sum dz u*, sum dz v* (committed by launch_commit_spec)
  struct { bool valid = false; float dt = 0.f, chi = 0.f; bool zhalo = false; } spec;
  bool use_spec = true;
  bool use_overlap = true;             // T, S halo fill / exchange + pressure scan on the second stream, under the substeps
  bool use_zfold = true;               // corrector / tracer epilogue also write the z halos of the fields they produce
  // flux boundary conditions (row A7): 2-D device arrays [u, v, T, S][bottom, top], nullptr = no-flux
  real* bflux[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  bool has_bflux = false;
  // clock (model.clock)
  double time = 0.0;
  long iteration = 0;
  float last_dt = 0.f;
  Exchange ex;
  // errors
  std::string err;
  int sticky = 0;
  // measurement
  long launches = 0;
  cudaEvent_t loop_start = nullptr, loop_stop = nullptr;
  bool loop_timed = false;
  bool timers_on = false;
  std::vector<StageTimer> timers;
  bool use_fused = true;
  bool use_tma = true;
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // downloads that overlap the tail of a step: copy stream, "interior final since" events of the last fused step
  cudaStream_t stream_d2h = nullptr;
  cudaEvent_t ev_ts_final = nullptr, ev_uv_final = nullptr;
  bool early_valid = false, d2h_pending = false;
  bool use_tma_tracer = true;
  bool use_packed = true;              // FP32x2 (FFMA2) momentum kernels
  void* tma = nullptr;   // TMA tensor maps (gb25_tend_tma.cu)
  bool use_baro_persistent = true;     // all split-explicit substeps in one persistent kernel (gb25_baro.cu)
  void* baro_plan = nullptr;   // persistent split-explicit kernel: band decomposition, flags (gb25_baro.cu)

  inline void count_launch() { launches++; }
};
typedef gb25_handle Handle;

// RAII CUDA-event timer on the handle's stream (active only while gb25_enable_stage_timers is on).
// Names starting with "kernel:" time a single kernel launch.
#include <cstring>
struct StageScope {
  Handle* h; StageTimer* t = nullptr; size_t slot = 0;
  StageScope(Handle* h_, const char* name) : h(h_) {
    if (!h->timers_on) return;
    for (auto& s : h->timers) if (s.name == name || !strcmp(s.name, name)) { t = &s; break; }
    if (!t) { h->timers.push_back(StageTimer{name}); t = &h->timers.back(); }
    if (t->used == t->ev.size()) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      t->ev.push_back({a, b});
    }
    slot = t->used++;
    cudaEventRecord(t->ev[slot].first, h->stream);
  }
  ~StageScope() { if (t) cudaEventRecord(t->ev[slot].second, h->stream); }
};

struct Ab2Spec { float dt, c1, c2; int zhalo; };   // zhalo: the tracer epilogue also writes the (mirror) z halos of T', S'   // AB2 epilogue of the tendency kernels: psi' = psi + dt (c1 Gn - c2 G-)
bool spec_possible(Handle* h);          // gb25_kernels.cu: the configuration runs the kernels that have the epilogue

// stage launchers (gb25_kernels.cu)
void launch_fill_halo(Handle* h, const HaloSpec* specs, int n, bool three_d);
void launch_halo_south_north(Handle* h, const HaloSpec* specs, int n, bool three_d, int mode_s, int mode_n);
void launch_halo_bottom_top(Handle* h, const HaloSpec* specs, int n);
void launch_halo_periodic_x(Handle* h, const HaloSpec* specs, int n, bool three_d);
void launch_fill_halo_dist(Handle* h, const HaloSpec* specs, int n, bool three_d);   // gb25_exchange.cu
void exchange_baro_eta(Handle* h);
void exchange_baro_uv(Handle* h);
int exchange_check_timeout(Handle* h);
void exchange_close(Handle* h);
void exchange_table(Handle* h, real* tab[EX_NF]);   // the allocations behind ExField, in order
void launch_mask(Handle* h, bool uv_only);
void launch_compute_w(Handle* h);
void launch_compute_p(Handle* h);
void launch_tracer_tendency(Handle* h, const Ab2Spec* spec = nullptr);
void launch_tracer_tendency_v1(Handle* h);
void launch_tracer_tendency_v2(Handle* h);   // gb25_tend_v2.cu
void launch_momentum_tendency_v1(Handle* h);
void launch_momentum_tendency_v2(Handle* h);  // gb25_tend_v2.cu
void launch_aux_columns(Handle* h);
void launch_generic_list(Handle* h, bool momentum, bool tracers);   // gb25_tend_v2.cu
void launch_momentum_tendency_tma(Handle* h, const Ab2Spec* spec = nullptr);   // gb25_tend_tma.cu
bool tma_available(Handle* h);
void launch_tracer_tendency_tma(Handle* h, const Ab2Spec* spec = nullptr);
void tma_free(Handle* h);           // gb25_tend_v2.cu: w + zeta + flux divergences in one column pass
void launch_momentum_tendency(Handle* h, const Ab2Spec* spec = nullptr);
void launch_ab2_columns(Handle* h, float dt, float chi);
void launch_barotropic(Handle* h, float dt);
void launch_correct_cache(Handle* h);
void launch_barotropic_mode(Handle* h);
void launch_ab2_fused(Handle* h, float dt, float chi);
void launch_commit_spec(Handle* h);
bool launch_correct_fused(Handle* h);   // true: it also wrote the z halos of u, v
void launch_vdiff_explicit(Handle* h);
void launch_boundary_tendencies(Handle* h);
void launch_implicit_columns(Handle* h, float dt, bool with_sums);
bool launch_barotropic_persistent(Handle* h, float dt);   // gb25_baro.cu; false: not applicable, use the substep kernels
void baro_plan_free(Handle* h);
void baro_plan_prepare(Handle* h);     // builds the plan of the persistent substep kernel now (allocations, attributes)
int baro_check_timeout(Handle* h);
// Every __global__ of a translation unit, for preload_kernels() (gb25_api.cu): lazy module loading (the CUDA 12 default)
// loads a kernel at its first use and that load may synchronise the context, which must not happen while a stream of the
// device waits for a neighbour tile.  tests/test_abi.py compares the total with the entry points of the built cubins.
struct KernelTable { const void* const* fn; int n; };
KernelTable kernel_table_core();       // gb25_kernels.cu
KernelTable kernel_table_tend_v2();    // gb25_tend_v2.cu
KernelTable kernel_table_tend_tma();   // gb25_tend_tma.cu
KernelTable kernel_table_exchange();   // gb25_exchange.cu
KernelTable kernel_table_baro();       // gb25_baro.cu
