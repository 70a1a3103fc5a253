// gb25_exchange.cu — multi-GPU halo exchange, one process per GPU (row (e) of SURVEY.md §8).
//
// Replaces the XLA collective-permute traffic of Distributed(ReactantState(); partition=Partition(Rx,Ry,1))
// (/root/reference/sharding/sharded_baroclinic_instability_simulation_run.jl:65-72).  Design for one NVSwitch box:
//   * every rank exports the device allocations of its exchanged fields (u, v, T, S, eta, U, V, GU, GV) and a
//     small flag array as CUDA IPC handles; neighbours map them, so a neighbour's halo cells are ordinary global
//     addresses reached over NVLink;
//   * PUSH model: the owner of the data writes the edge strips straight into the neighbours' halo cells with
//     coalesced peer stores (no packing, no staging buffers, no NCCL call on the data path), then publishes a
//     per-sender sequence number in the neighbour's flag array (system-scope fence + store);
//   * the receiver's stream waits for the sequence numbers of exactly the tiles it receives from; the wait is a
//     one-warp kernel spinning (bounded) on LOCAL memory — every rank owns its GPU, so no two spinning kernels
//     ever share one;
//   * two phases per fill, mirroring the single-tile order: y (south/north strips, tripolar fold with
//     the x-mirrored partner, wall conditions on the boundary tiles), then x over the full parent extent so that
//     corners are carried along.  No acknowledgements are needed: every site exchanges with the same neighbours
//     in both directions, so a tile can only re-write a neighbour's halo after it has received something the
//     neighbour sent after consuming the previous contents.
#include <cstring>
#include <cstdlib>

#include <cuda.h>   // CUstream / CUdeviceptr types; entry points come from cudaGetDriverEntryPoint

#include "gb25_internal.h"

struct ExBlob {
  cudaIpcMemHandle_t fld[EX_NF];
  cudaIpcMemHandle_t flags;
  int rank, Nx, Ny, Nz, Rx, Ry;
};

// --------------------------------------------------------------------------------- device side
struct PushField { const real* src; real* dst; int lx, ly, lz; real sign; int flat; };
struct PushBatch { PushField f[9]; int n; };

// rows [srow, srow+nrows) of src -> rows [drow, ...) of dst; interior columns; planes [p0, p0+np)
__global__ void k_push_rows(DevGrid g, PushBatch pb, int srow, int drow, int nrows, int p0, int np, int three_d) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x + g.Hx;
  if (I >= g.Nx + g.Hx) return;
  const int r = blockIdx.y % nrows, pl = blockIdx.y / nrows;
  if (pl >= np) return;
  const PushField pf = pb.f[blockIdx.z];
  if (pf.flat) { if (pl > 0) return; three_d = 0; }
  const size_t po = three_d ? (size_t)g.n2 * (p0 + pl) : 0;
  pf.dst[po + (size_t)g.PX * (drow + r) + I] = pf.src[po + (size_t)g.PX * (srow + r) + I];
}
// columns [scol, scol+ncols) -> [dcol, ...); all rows; all planes
__global__ void k_push_cols(DevGrid g, PushBatch pb, int scol, int dcol, int ncols, int three_d, int row0, int nrows_) {
  const int cI = threadIdx.x;
  const int J = row0 + blockIdx.x * blockDim.y + threadIdx.y;
  if (cI >= ncols || J >= row0 + nrows_) return;
  const PushField pf = pb.f[blockIdx.z];
  if (pf.flat) { if (blockIdx.y > 0) return; three_d = 0; }
  const size_t po = (three_d ? (size_t)g.n2 * blockIdx.y : 0) + (size_t)g.PX * J;
  pf.dst[po + dcol + cI] = pf.src[po + scol + cI];
}
// West / east strips travel PACKED.  Written straight into the neighbour's halo columns they are 32-byte pieces at a pitch
// of PX floats, and NVLink moves those at ~50 GB/s (measured: 0.22 ms per step for 11 MB at 1440 x 600 x 50 tiles, six
// times the cost of the row strips).  Here a warp gathers four rows x Hx columns and stores them as one contiguous 128-byte
// line into the neighbour's column inbox; after the handshake the receiver scatters its inbox into its own halo columns
// (a local copy).  The inbox has two halves, used alternately by sequence parity: a neighbour that runs ahead may already
// push the strips of the NEXT fill while this tile has not unpacked the current one.
__global__ void k_push_cols_packed(DevGrid g, PushBatch pb, int scol, int three_d, real* __restrict__ box, int slot0) {
  const int c = threadIdx.x;                                  // 0 .. Hx-1
  const int J = blockIdx.x * blockDim.y + threadIdx.y;        // storage row
  if (c >= g.Hx || J >= g.PY) return;
  const PushField pf = pb.f[blockIdx.z];
  if (pf.flat) { if (blockIdx.y > 0) return; three_d = 0; }
  const size_t po = (three_d ? (size_t)g.n2 * blockIdx.y : 0) + (size_t)g.PX * J;
  box[(((size_t)(blockIdx.z + slot0) * gridDim.y + blockIdx.y) * g.PY + J) * g.Hx + c] = pf.src[po + scol + c];
}
// blockIdx.z = 2 * field slot + direction (0: from the west tile -> columns [0, Hx), 1: from the east tile -> [Nx+Hx, PX))
__global__ void k_unpack_cols(DevGrid g, PushBatch pb, int three_d, const real* __restrict__ box_w, const real* __restrict__ box_e, int slot0) {
  const int c = threadIdx.x;
  const int J = blockIdx.x * blockDim.y + threadIdx.y;
  if (c >= g.Hx || J >= g.PY) return;
  const int q = blockIdx.z >> 1, dir = blockIdx.z & 1;
  const PushField pf = pb.f[q];
  if (pf.flat) { if (blockIdx.y > 0) return; three_d = 0; }
  const size_t po = (three_d ? (size_t)g.n2 * blockIdx.y : 0) + (size_t)g.PX * J;
  const real* box = dir ? box_e : box_w;
  pf.dst[po + (dir ? g.Nx + g.Hx : 0) + c] = box[(((size_t)(q + slot0) * gridDim.y + blockIdx.y) * g.PY + J) * g.Hx + c];
}
// tripolar fold: my top rows -> the partner's north halo rows, x-mirrored, sign-flipped for vectors.
// `second` selects the Face-x column whose partner lives one tile further (see the header comment of
// fold_index_maps in grids.py / SURVEY A.5); quirk_pos: that element keeps |sign| (global wrap past Nx).
__global__ void k_push_fold(DevGrid g, PushBatch pb, int three_d, int nrows, int second, int quirk_pos) {
  const int is = blockIdx.x * blockDim.x + threadIdx.x + 1;   // source column (local, 1-based)
  if (is > g.Nx) return;
  const PushField pf = pb.f[blockIdx.z];
  if (pf.flat) three_d = 0;
  const int nk = three_d ? g.Nz + pf.lz : 1;
  const int k = blockIdx.y + 1;
  if (k > nk) return;
  int id; real sg = pf.sign;
  if (pf.lx == 0) { if (second) return; id = g.Nx - is + 1; }
  else if (is == 1) { if (!second) return; id = 1; if (quirk_pos) sg = rabs(sg); }
  else { if (second) return; id = g.Nx + 2 - is; }
  const size_t po = three_d ? (size_t)g.n2 * (k + g.Hz - 1) : 0;
  for (int m = 1; m <= nrows; m++) {
    const int js = pf.ly == 0 ? g.Ny - m : g.Ny - m + 1;
    pf.dst[po + (size_t)g.PX * (g.Ny + m + g.Hy - 1) + id + g.Hx - 1] = sg * pf.src[po + (size_t)g.PX * (js + g.Hy - 1) + is + g.Hx - 1];
  }
}
struct SignalSet { int* dst[EX_NSLOT]; int n; };
__global__ void k_signal(SignalSet s, int val) {
  __threadfence_system();
  if ((int)threadIdx.x < s.n) *((volatile int*)s.dst[threadIdx.x]) = val;
  __threadfence_system();
}
__global__ void k_wait(volatile int* flags, int mask, int val) {
  const int t = threadIdx.x;
  if (t < EX_NSLOT && ((mask >> t) & 1)) {
    long spins = 0;
    while (flags[t] < val) {
      __nanosleep(200);
      if (++spins > 20000000L) { flags[EX_NSLOT] = 1; break; }   // ~ several seconds: record a timeout, do not hang
    }
  }
  __threadfence_system();
}

// --------------------------------------------------------------------------------- host side
void exchange_table(Handle* h, real* tab[EX_NF]) {
  const DevFields& f = h->f;
  real* t[EX_NF] = {h->state_buf[0][0], h->state_buf[0][1], h->state_buf[0][2], h->state_buf[0][3], f.eta, f.bu, f.bv, f.gU, f.gV,
                     h->state_buf[1][0], h->state_buf[1][1], h->state_buf[1][2], h->state_buf[1][3], h->ex.xbox};
  for (int q = 0; q < EX_NF; q++) tab[q] = t[q];
}
static int ex_field_id(Handle* h, const real* a) {
  real* tab[EX_NF];
  exchange_table(h, tab);
  for (int q = 0; q < EX_NF; q++) if (q != EX_XBOX && tab[q] == a) return q;
  return -1;
}
static const ExSlot kOpposite[EX_NSLOT] = {SLOT_E, SLOT_W, SLOT_N, SLOT_S, SLOT_FOLD, SLOT_FOLD2};

static PushBatch make_push(Handle* h, const HaloSpec* specs, int n, int slot, bool only_face_x = false) {
  PushBatch pb; pb.n = 0;
  for (int q = 0; q < n; q++) {
    const int id = ex_field_id(h, specs[q].a);
    if (id < 0) continue;
    if (only_face_x && specs[q].lx == 0) continue;
    pb.f[pb.n++] = PushField{specs[q].a, h->ex.to[slot].fld[id], specs[q].lx, specs[q].ly, specs[q].lz, specs[q].sign, specs[q].flat};
  }
  return pb;
}
// Stream memory operations: the flag write / wait is executed by the GPU front end in stream order (no kernel
// launch, no SM).  The write carries the default memory barrier, so the peer stores of the preceding push kernel
// are visible at the destination before the flag.  GB25_STREAM_MEMOPS=0 (or a driver without the entry points)
// falls back to the one-warp signal / wait kernels, whose wait has a time-out.
typedef CUresult (*PFN_writeValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*PFN_waitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static PFN_writeValue32 g_write32 = nullptr;
static PFN_waitValue32 g_wait32 = nullptr;
static int g_memops = -1;
static bool memops_available() {
  if (g_memops < 0) {
    g_memops = 0;
    const char* e = getenv("GB25_STREAM_MEMOPS");
    if (!(e && e[0] == '0')) {
      void *pw = nullptr, *pq = nullptr;
      cudaDriverEntryPointQueryResult q1, q2;
      if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &pw, cudaEnableDefault, &q1) == cudaSuccess && q1 == cudaDriverEntryPointSuccess &&
          cudaGetDriverEntryPoint("cuStreamWaitValue32", &pq, cudaEnableDefault, &q2) == cudaSuccess && q2 == cudaDriverEntryPointSuccess) {
        g_write32 = (PFN_writeValue32)pw; g_wait32 = (PFN_waitValue32)pq; g_memops = 1;
      }
    }
  }
  return g_memops == 1;
}
static void signal_slots(Handle* h, int slot_mask) {
  Exchange& X = h->ex;
  SignalSet s; s.n = 0;
  for (int sl = 0; sl < EX_NSLOT; sl++)
    if (((slot_mask >> sl) & 1) && X.to[sl].rank >= 0 && X.to[sl].rank != X.rank) s.dst[s.n++] = X.to[sl].flags + EX_LANE_FLAGS * X.lane + kOpposite[sl];
  if (!s.n) return;
  if (memops_available()) {
    bool ok = true;
    for (int q = 0; q < s.n; q++) ok &= g_write32((CUstream)h->stream, (CUdeviceptr)s.dst[q], (cuuint32_t)X.seqs[X.lane], CU_STREAM_WRITE_VALUE_DEFAULT) == CUDA_SUCCESS;
    if (ok) return;
    g_memops = 0;   // (not supported on this address / driver: use the kernels from now on)
  }
  k_signal<<<1, 32, 0, h->stream>>>(s, X.seqs[X.lane]); h->count_launch();
}
static void wait_slots(Handle* h, int slot_mask) {
  Exchange& X = h->ex;
  int m = 0;
  for (int sl = 0; sl < EX_NSLOT; sl++)
    if (((slot_mask >> sl) & 1) && X.to[sl].rank >= 0 && X.to[sl].rank != X.rank) m |= 1 << sl;
  if (!m) return;
  if (memops_available()) {
    bool ok = true;
    for (int sl = 0; sl < EX_NSLOT; sl++)
      if ((m >> sl) & 1) ok &= g_wait32((CUstream)h->stream, (CUdeviceptr)(X.flags + EX_LANE_FLAGS * X.lane + sl), (cuuint32_t)X.seqs[X.lane], CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS;
    if (ok) return;
    g_memops = 0;
  }
  k_wait<<<1, 32, 0, h->stream>>>(X.flags + EX_LANE_FLAGS * X.lane, m, X.seqs[X.lane]); h->count_launch();
}

void launch_fill_halo_dist(Handle* h, const HaloSpec* specs, int n, bool three_d) {
  Exchange& X = h->ex;
  const DevGrid& g = h->g;
  const gb25_config& c = h->cfg;
  const bool top = c.ry == c.Ry - 1, bottom = c.ry == 0;
  const bool fold = c.topo_y == GB25_TOPO_FOLD && top;
  const int np = three_d ? g.PZ : 1;
  // ---- wall conditions of the boundary tiles, then z halos of the interior columns (local)
  { StageScope ts(h, "exchange:local_fills");
  launch_halo_south_north(h, specs, n, three_d, bottom ? 1 : 0, (top && c.topo_y == GB25_TOPO_BOUNDED) ? 1 : ((fold && c.Rx == 1) ? 2 : 0));
  if (three_d) launch_halo_bottom_top(h, specs, n); }
  // ---- phase Y: strips to the north / south tiles (all planes), fold rows to the mirrored partner
  int mask_y = 0;
  dim3 b(128);
  StageScope* tsy = new StageScope(h, "exchange:push_y");
  if (!top) {
    PushBatch pb = make_push(h, specs, n, SLOT_N);
    dim3 gr((g.Nx + 127) / 128, g.Hy * np, pb.n);
    if (pb.n) { k_push_rows<<<gr, b, 0, h->stream>>>(g, pb, g.Ny, 0, g.Hy, 0, np, three_d); h->count_launch(); }   // rows Ny-Hy+1..Ny -> 1-Hy..0
    mask_y |= 1 << SLOT_N;
  }
  if (!bottom) {
    PushBatch pb = make_push(h, specs, n, SLOT_S);
    dim3 gr((g.Nx + 127) / 128, g.Hy * np, pb.n);
    if (pb.n) { k_push_rows<<<gr, b, 0, h->stream>>>(g, pb, g.Hy, g.Ny + g.Hy, g.Hy, 0, np, three_d); h->count_launch(); }  // rows 1..Hy -> Ny+1..Ny+Hy
    mask_y |= 1 << SLOT_S;
  }
  if (fold && c.Rx > 1) {
    int maxlz = 0;
    for (int q = 0; q < n; q++) if (!specs[q].flat) maxlz = max(maxlz, specs[q].lz);
    const int nk = three_d ? g.Nz + maxlz : 1;
    PushBatch pb = make_push(h, specs, n, SLOT_FOLD);
    dim3 gr((g.Nx + 127) / 128, nk, pb.n);
    if (pb.n) { k_push_fold<<<gr, b, 0, h->stream>>>(g, pb, three_d, g.Hy, 0, 0); h->count_launch(); }
    PushBatch p2 = make_push(h, specs, n, SLOT_FOLD2, true);
    dim3 g2(1, nk, p2.n);
    if (p2.n) { k_push_fold<<<g2, b, 0, h->stream>>>(g, p2, three_d, g.Hy, 1, c.rx == 0 ? 1 : 0); h->count_launch(); }
    mask_y |= (1 << SLOT_FOLD) | (1 << SLOT_FOLD2);
  }
  delete tsy;
  if (mask_y) { StageScope ts(h, "exchange:handshake_y"); X.seqs[X.lane]++; signal_slots(h, mask_y); wait_slots(h, mask_y); }
  // ---- phase X: west / east strips over the full parent extent (carries the y and z halos into the corners)
  if (c.Rx == 1) { launch_halo_periodic_x(h, specs, n, three_d); return; }
  {
    dim3 bc(g.Hx, 32), gc((g.PY + 31) / 32, np, 0);
    // the second lane (T, S ahead of the barotropic solve, on the second stream) has its own sequence numbers and flag words
    // and uses field slots 7, 8 of the column inbox; the first lane then carries at most seven fields (slots 0 .. 6)
    const int slot0 = X.lane ? 7 : 0;
    X.seqs[X.lane]++;
    X.xseqs[X.lane]++;     // (the parity of the column phases, not of all phases: a fill with a row phase advances seq twice)
    real* const mybox = X.xbox + (size_t)(X.xseqs[X.lane] & 1) * 2 * X.xbox_stride;
    {
      StageScope ts(h, "exchange:push_x");
      const size_t off = (size_t)(X.xseqs[X.lane] & 1) * 2 * X.xbox_stride;
      PushBatch pe = make_push(h, specs, n, SLOT_E);   // my last Hx interior columns -> the east tile's inbox "from the west"
      gc.z = pe.n;
      if (pe.n) { k_push_cols_packed<<<gc, bc, 0, h->stream>>>(g, pe, g.Nx, three_d, X.to[SLOT_E].fld[EX_XBOX] + off, slot0); h->count_launch(); }
      PushBatch pw = make_push(h, specs, n, SLOT_W);   // my first Hx interior columns -> the west tile's inbox "from the east"
      gc.z = pw.n;
      if (pw.n) { k_push_cols_packed<<<gc, bc, 0, h->stream>>>(g, pw, g.Hx, three_d, X.to[SLOT_W].fld[EX_XBOX] + off + X.xbox_stride, slot0); h->count_launch(); }
    }
    {
      StageScope ts(h, "exchange:handshake_x");
      const int mx = (1 << SLOT_W) | (1 << SLOT_E);
      signal_slots(h, mx); wait_slots(h, mx);
    }
    {
      StageScope ts(h, "exchange:unpack_x");
      PushBatch pu; pu.n = 0;
      for (int q = 0; q < n; q++)
        if (ex_field_id(h, specs[q].a) >= 0) pu.f[pu.n++] = PushField{nullptr, specs[q].a, specs[q].lx, specs[q].ly, specs[q].lz, specs[q].sign, specs[q].flat};
      gc.z = 2 * pu.n;
      if (pu.n) { k_unpack_cols<<<gc, bc, 0, h->stream>>>(g, pu, three_d, mybox, mybox + X.xbox_stride, slot0); h->count_launch(); }
    }
  }
}

// one-cell refresh of the barotropic halos between the substep kernels (no corners involved): the substep kernels
// push their edges themselves (BaroPeers in gb25_kernels.cu); here only the flags are published and awaited
void exchange_baro_eta(Handle* h) {   // after the eta kernel: the U,V kernel reads eta(i-1), eta(j-1)
  Exchange& X = h->ex;
  const gb25_config& c = h->cfg;
  int mask_out = 0, mask_in = 0;
  if (c.Rx > 1) { mask_out |= 1 << SLOT_E; mask_in |= 1 << SLOT_W; }
  if (c.ry < c.Ry - 1) mask_out |= 1 << SLOT_N;
  if (c.ry > 0) mask_in |= 1 << SLOT_S;
  // Fold partners acknowledge each other here: k_baro_uv(m+1) of the partner stores -V into MY row Ny+1, which my
  // k_baro_eta(m+1) (just finished) had to read with the substep-m value.  The partner is not my x neighbour when
  // Rx >= 4, so without this handshake nothing orders its next store after my read.
  if (c.topo_y == GB25_TOPO_FOLD && c.ry == c.Ry - 1 && c.Rx > 1) { mask_out |= 1 << SLOT_FOLD; mask_in |= 1 << SLOT_FOLD; }
  if (mask_out | mask_in) { X.seqs[X.lane]++; signal_slots(h, mask_out); wait_slots(h, mask_in); }
}
void exchange_baro_uv(Handle* h) {    // after the U,V kernel: the eta kernel reads U(i+1), V(j+1)
  Exchange& X = h->ex;
  const gb25_config& c = h->cfg;
  int mask_out = 0, mask_in = 0;
  if (c.Rx > 1) { mask_out |= 1 << SLOT_W; mask_in |= 1 << SLOT_E; }
  if (c.ry > 0) mask_out |= 1 << SLOT_S;
  if (c.ry < c.Ry - 1) mask_in |= 1 << SLOT_N;
  if (c.topo_y == GB25_TOPO_FOLD && c.ry == c.Ry - 1 && c.Rx > 1) { mask_out |= 1 << SLOT_FOLD; mask_in |= 1 << SLOT_FOLD; }
  if (mask_out | mask_in) { X.seqs[X.lane]++; signal_slots(h, mask_out); wait_slots(h, mask_in); }
}

// --------------------------------------------------------------------------------- C ABI
extern "C" int gb25_exchange_blob_size(void) { return (int)sizeof(ExBlob); }

// the windows a tile offers to its neighbours besides its fields: the flag / pair-inbox buffer and the column inbox
static int exchange_alloc_windows(Handle* h) {
  Exchange& X = h->ex;
  if (!X.flags) {
    // a dedicated 2 MiB allocation so that the IPC handle maps exactly this buffer
    if (cudaMalloc(&X.flags, 2 << 20) != cudaSuccess) { h->err = "exchange: cudaMalloc flags"; return GB25_ERR_ALLOC; }
    cudaMemset(X.flags, 0, 2 << 20);
  }
  if (!X.xbox) {
    X.xbox_stride = (size_t)9 * h->g.PZ * h->g.PY * h->g.Hx;
    if (cudaMalloc(&X.xbox, 4 * X.xbox_stride * sizeof(real)) != cudaSuccess) { h->err = "exchange: cudaMalloc column inbox"; return GB25_ERR_ALLOC; }
    cudaMemset(X.xbox, 0, 4 * X.xbox_stride * sizeof(real));
  }
  return GB25_OK;
}
// ranks of the tiles this tile exchanges with, by slot (-1: none)
static void exchange_neighbours(const gb25_config& c, int want[EX_NSLOT]) {
  auto rk = [&](int rx, int ry) { return ((rx % c.Rx) + c.Rx) % c.Rx + c.Rx * ry; };
  for (int s = 0; s < EX_NSLOT; s++) want[s] = -1;
  if (c.Rx > 1) { want[SLOT_W] = rk(c.rx - 1, c.ry); want[SLOT_E] = rk(c.rx + 1, c.ry); }
  if (c.ry > 0) want[SLOT_S] = rk(c.rx, c.ry - 1);
  if (c.ry < c.Ry - 1) want[SLOT_N] = rk(c.rx, c.ry + 1);
  if (c.topo_y == GB25_TOPO_FOLD && c.ry == c.Ry - 1 && c.Rx > 1) {
    want[SLOT_FOLD] = rk(c.Rx - 1 - c.rx, c.ry);
    want[SLOT_FOLD2] = rk(c.Rx - c.rx, c.ry);
  }
}
static int exchange_finish_connect(Handle* h) {
  Exchange& X = h->ex;
  // A reconnect restarts every sequence number, so the flag words and the pair inbox must not keep values of the previous
  // session (a stale number would satisfy a wait).  The caller synchronises all ranks before reconnecting and puts a
  // barrier after it (distributed.connect), so no neighbour writes into this buffer while it is cleared.
  if (cudaMemset(X.flags, 0, 2 << 20) != cudaSuccess) { h->err = "exchange connect: cudaMemset flags"; return GB25_ERR_CUDA; }
  X.seqs[0] = X.seqs[1] = 0; X.xseqs[0] = X.xseqs[1] = 0; X.lane = 0;
  X.on = true;
  baro_plan_free(h);   // the persistent substep kernel restarts its sequence numbers on the (zeroed) shared flag buffer
  baro_plan_prepare(h);   // ... and its allocations are made now, not behind the first wait for a neighbour (see preload_kernels)
  return GB25_OK;
}

extern "C" int gb25_exchange_export(gb25_handle* h, void* blob) {
  if (!h || !blob) return GB25_ERR_INVALID;
  if (h->guard) { h->err = "gb25_exchange_export: guarded allocations (GB25_GUARD=1) cannot be exported"; return GB25_ERR_INVALID; }
  cudaSetDevice(h->device);
  ExBlob b; memset(&b, 0, sizeof b);
  Exchange& X = h->ex;
  { const int rc = exchange_alloc_windows(h); if (rc != GB25_OK) return rc; }
  real* tab[EX_NF];
  exchange_table(h, tab);
  for (int q = 0; q < EX_NF; q++) {
    cudaError_t e = cudaIpcGetMemHandle(&b.fld[q], tab[q]);
    if (e != cudaSuccess) { h->err = std::string("gb25_exchange_export: cudaIpcGetMemHandle: ") + cudaGetErrorString(e); return GB25_ERR_COMM; }
  }
  cudaError_t e = cudaIpcGetMemHandle(&b.flags, X.flags);
  if (e != cudaSuccess) { h->err = std::string("gb25_exchange_export: cudaIpcGetMemHandle(flags): ") + cudaGetErrorString(e); return GB25_ERR_COMM; }
  const gb25_config& c = h->cfg;
  b.rank = c.rx + c.Rx * c.ry; b.Nx = c.Nx; b.Ny = c.Ny; b.Nz = c.Nz; b.Rx = c.Rx; b.Ry = c.Ry;
  memcpy(blob, &b, sizeof b);
  return GB25_OK;
}

extern "C" int gb25_exchange_connect(gb25_handle* h, const void* blobs, int nranks) {
  if (!h || !blobs) return GB25_ERR_INVALID;
  cudaSetDevice(h->device);
  const gb25_config& c = h->cfg;
  Exchange& X = h->ex;
  if (nranks != c.Rx * c.Ry) { h->err = "gb25_exchange_connect: nranks != Rx*Ry"; return GB25_ERR_INVALID; }
  if (!X.flags) { h->err = "gb25_exchange_connect: call gb25_exchange_export first"; return GB25_ERR_INVALID; }
  if (c.fold_variant == 1 && c.Rx > 1 && c.topo_y == GB25_TOPO_FOLD) { h->err = "gb25_exchange_connect: fold_variant 1 is only implemented for Rx == 1"; return GB25_ERR_INVALID; }
  const ExBlob* B = (const ExBlob*)blobs;
  X.nranks = nranks; X.rank = c.rx + c.Rx * c.ry;
  for (int r = 0; r < nranks; r++)
    if (B[r].rank != r || B[r].Nx != c.Nx || B[r].Ny != c.Ny || B[r].Nz != c.Nz || B[r].Rx != c.Rx || B[r].Ry != c.Ry) {
      h->err = "gb25_exchange_connect: blobs must be ordered by rank and describe equal tiles"; return GB25_ERR_INVALID;
    }
  int want[EX_NSLOT];
  exchange_neighbours(c, want);
  // map every distinct peer once
  std::vector<ExPeer> mapped(nranks);
  std::vector<char> have(nranks, 0);
  real* mine[EX_NF];
  exchange_table(h, mine);
  for (int s = 0; s < EX_NSLOT; s++) {
    const int r = want[s];
    X.to[s].rank = r;
    if (r < 0) continue;
    if (!have[r]) {
      ExPeer p; p.rank = r;
      if (r == X.rank) {
        for (int q = 0; q < EX_NF; q++) p.fld[q] = mine[q];
        p.flags = X.flags;
      } else {
        for (int q = 0; q < EX_NF; q++) {
          void* ptr = nullptr;
          cudaError_t e = cudaIpcOpenMemHandle(&ptr, B[r].fld[q], cudaIpcMemLazyEnablePeerAccess);
          if (e != cudaSuccess) { h->err = std::string("gb25_exchange_connect: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); return GB25_ERR_COMM; }
          X.opened.push_back(ptr); p.fld[q] = (real*)ptr;
        }
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, B[r].flags, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { h->err = std::string("gb25_exchange_connect: cudaIpcOpenMemHandle(flags): ") + cudaGetErrorString(e); return GB25_ERR_COMM; }
        X.opened.push_back(ptr); p.flags = (int*)ptr;
      }
      mapped[r] = p; have[r] = 1;
    }
    X.to[s] = mapped[r];
  }
  return exchange_finish_connect(h);
}

// Single process, one handle per device (the reference drives all GPUs of a node from one process:
// /root/reference/sharding/sharded_baroclinic_instability_simulation_run.jl:49, `single_gpu_per_process=false`).  No IPC:
// after cudaDeviceEnablePeerAccess the neighbours' allocations are ordinary device pointers.  handles[r] is the tile of
// rank r = rx + Rx*ry.  Every tile must live on its own device: the persistent substep kernels of neighbouring tiles wait
// for one another and must be co-resident.
extern "C" int gb25_exchange_connect_local(gb25_handle** hs, int n) {
  if (!hs || n < 1) return GB25_ERR_INVALID;
  for (int r = 0; r < n; r++) if (!hs[r]) return GB25_ERR_INVALID;
  Handle* h0 = hs[0];
  const gb25_config& c0 = h0->cfg;
  if (n != c0.Rx * c0.Ry) { h0->err = "gb25_exchange_connect_local: n != Rx*Ry"; return GB25_ERR_INVALID; }
  for (int r = 0; r < n; r++) {
    const gb25_config& c = hs[r]->cfg;
    if (c.rx + c.Rx * c.ry != r || c.Nx != c0.Nx || c.Ny != c0.Ny || c.Nz != c0.Nz || c.Rx != c0.Rx || c.Ry != c0.Ry) {
      h0->err = "gb25_exchange_connect_local: handles must be ordered by rank and describe equal tiles"; return GB25_ERR_INVALID;
    }
    if (c.fold_variant == 1 && c.Rx > 1 && c.topo_y == GB25_TOPO_FOLD) { h0->err = "gb25_exchange_connect_local: fold_variant 1 is only implemented for Rx == 1"; return GB25_ERR_INVALID; }
    for (int q = 0; q < r; q++)
      if (hs[q]->device == hs[r]->device) { h0->err = "gb25_exchange_connect_local: every tile needs its own device"; return GB25_ERR_INVALID; }
  }
  for (int r = 0; r < n; r++) {
    cudaSetDevice(hs[r]->device);
    cudaStreamSynchronize(hs[r]->stream);
    const int rc = exchange_alloc_windows(hs[r]);
    if (rc != GB25_OK) { h0->err = hs[r]->err; return rc; }
    for (int q = 0; q < n; q++) {
      if (q == r) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, hs[r]->device, hs[q]->device);
      if (!can) { h0->err = "gb25_exchange_connect_local: devices cannot access each other's memory"; return GB25_ERR_COMM; }
      const cudaError_t e = cudaDeviceEnablePeerAccess(hs[q]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { h0->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return GB25_ERR_COMM; }
      cudaGetLastError();
    }
  }
  for (int r = 0; r < n; r++) {
    Handle* h = hs[r];
    Exchange& X = h->ex;
    X.nranks = n; X.rank = r;
    int want[EX_NSLOT];
    exchange_neighbours(h->cfg, want);
    for (int s = 0; s < EX_NSLOT; s++) {
      X.to[s].rank = want[s];
      if (want[s] < 0) continue;
      Handle* p = hs[want[s]];
      exchange_table(p, X.to[s].fld);
      X.to[s].flags = p->ex.flags;
    }
    cudaSetDevice(h->device);
    const int rc = exchange_finish_connect(h);
    if (rc != GB25_OK) { h0->err = h->err; return rc; }
  }
  return GB25_OK;
}
// One loop over all tiles of a single-process partition: every step is enqueued on every device before the next one, so no
// device's launch queue fills up with work that waits for a neighbour whose step has not been enqueued yet.
extern "C" int gb25_loop_all(gb25_handle** hs, int n, float dt, int nsteps) {
  if (!hs || n < 1 || nsteps < 0) return GB25_ERR_INVALID;
  for (int s = 0; s < nsteps; s++)
    for (int r = 0; r < n; r++) {
      const int rc = gb25_time_step(hs[r], dt);
      if (rc != GB25_OK) return rc;
    }
  return GB25_OK;
}

int exchange_check_timeout(Handle* h) {
  if (!h->ex.on) return 0;
  int flag = 0;
  cudaMemcpy(&flag, h->ex.flags + EX_NSLOT, sizeof(int), cudaMemcpyDeviceToHost);
  return flag;
}
void exchange_close(Handle* h) {
  for (void* p : h->ex.opened) cudaIpcCloseMemHandle(p);
  h->ex.opened.clear();
  if (h->ex.flags) { cudaFree(h->ex.flags); h->ex.flags = nullptr; }
  if (h->ex.xbox) { cudaFree(h->ex.xbox); h->ex.xbox = nullptr; }
  h->ex.on = false;
}

// ---------------------------------------------------------------- kernel table (preload_kernels, gb25_api.cu)
KernelTable kernel_table_exchange() {
  static const void* const k[] = {
    (const void*)k_push_rows, (const void*)k_push_cols, (const void*)k_push_cols_packed, (const void*)k_unpack_cols,
    (const void*)k_push_fold, (const void*)k_signal, (const void*)k_wait,
  };
  return {k, (int)(sizeof k / sizeof k[0])};
}
