// gb25_internal.h — host-side handle and launcher declarations shared by the .cu files.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/gb25cuda.h"
#include "gb25_device.cuh"

struct HaloSpec { float* a; int lx, ly, lz; float sign; int flat; };   // flat = 1: a 2-D field riding in a batch of 3-D fields

// ---- multi-GPU halo exchange over peer-mapped (CUDA IPC) memory, one process per GPU (gb25_exchange.cu)
enum ExField { EX_U = 0, EX_V, EX_T, EX_S, EX_ETA, EX_BU, EX_BV, EX_GU, EX_GV, EX_NF };
enum ExSlot { SLOT_W = 0, SLOT_E, SLOT_S, SLOT_N, SLOT_FOLD, SLOT_FOLD2, EX_NSLOT };
struct ExPeer { float* fld[EX_NF]; int* flags; int rank; };
struct Exchange {
  bool on = false;
  int nranks = 1, rank = 0;
  int* flags = nullptr;          // local inbox: EX_NSLOT sequence numbers written by the neighbours + 1 error word
  ExPeer to[EX_NSLOT];           // the tile lying in that direction (destination of my pushes); rank < 0: none
  int from_mask_y = 0, from_mask_x = 0, from_mask_fold = 0;   // slots I receive on in each phase
  int seq = 0;
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
  float* field_ptr[GB25_FIELD_COUNT];
  // scratch 3-D arrays shared by the v2 kernels: vorticity (F,F,C), delta_x(Ax u) and delta_y(Ay v) at (C,C,C)
  float *zeta = nullptr, *dxU = nullptr, *dyV = nullptr;
  float *us2 = nullptr, *vs2 = nullptr;   // 2-D: column sums of the AB2-updated, masked velocities (fused path)
  float *corr_u = nullptr, *corr_v = nullptr;   // 2-D: unmasked barotropic transports for the streamed corrector
  float* carry[4] = {nullptr, nullptr, nullptr, nullptr};   // 2-D: vertical flux through the top face of the topmost generic cell (u, v, T, S)
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
  bool use_overlap = false;            // run the T,S halo fill + hydrostatic pressure on a second stream during the substeps
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
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
void launch_mask(Handle* h, bool uv_only);
void launch_compute_w(Handle* h);
void launch_compute_p(Handle* h);
void launch_tracer_tendency(Handle* h);
void launch_tracer_tendency_v1(Handle* h);
void launch_tracer_tendency_v2(Handle* h);   // gb25_tend_v2.cu
void launch_momentum_tendency_v1(Handle* h);
void launch_momentum_tendency_v2(Handle* h);  // gb25_tend_v2.cu
void launch_aux_columns(Handle* h);
void launch_generic_list(Handle* h, bool momentum, bool tracers);   // gb25_tend_v2.cu
void launch_momentum_tendency_tma(Handle* h);   // gb25_tend_tma.cu
bool tma_available(Handle* h);
void launch_tracer_tendency_tma(Handle* h);
void tma_free(Handle* h);           // gb25_tend_v2.cu: w + zeta + flux divergences in one column pass
void launch_momentum_tendency(Handle* h);
void launch_ab2_columns(Handle* h, float dt, float chi);
void launch_barotropic(Handle* h, float dt);
void launch_correct_cache(Handle* h);
void launch_barotropic_mode(Handle* h);
void launch_ab2_fused(Handle* h, float dt, float chi);
void launch_correct_fused(Handle* h);
void launch_vdiff_explicit(Handle* h);
void launch_implicit_columns(Handle* h, float dt, bool with_sums);
bool launch_barotropic_persistent(Handle* h, float dt);   // gb25_baro.cu; false: not applicable, use the substep kernels
void baro_plan_free(Handle* h);
int baro_check_timeout(Handle* h);
