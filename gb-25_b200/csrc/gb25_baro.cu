// gb25_baro.cu — the split-explicit free-surface substeps (row A10) as ONE persistent kernel.
//
// The per-substep kernels (k_baro_eta / k_baro_uv, gb25_kernels.cu) stream ~17 L2-resident 2-D arrays per substep:
// 43 launches, 0.44 ms per step at 1440 x 600, L2-bandwidth and launch bound; between tiles they need 42 flag
// handshakes executed by the stream front end (1 ms per step on two GPUs).  Here the 2-D state never leaves the SM:
//   * one CTA per SM (cooperative launch: all CTAs are co-resident), CTA b owns a band of whole rows;
//   * eta, U, V of the band and as many of the nine constant arrays as fit live in shared memory (8 arrays of
//     28.8 KB at 1440 x 600 on 148 SMs), the three running averages live in registers, the remaining constants
//     stream from L2; a thread slot is a quad of four x-adjacent points (128-bit shared-memory accesses);
//   * neighbouring bands exchange ONE row per half-substep through a small buffer of (value, sequence number) pairs,
//     each pair written with one 64-bit store and polled by the consumer (the "LL" idea of NCCL): no fences, no flag
//     round trips, no grid-wide barrier.  Every column is its own producer/consumer chain (eta(j) needs V(j+1),
//     V(j) needs eta(j-1)), so one buffer per direction is enough and the exact sequence number is the guard;
//   * on a partitioned grid the same pairs are stored straight into the neighbouring tile's peer-mapped inbox over
//     NVLink by the thread that computed the value: band b talks to band b of the east / west tile, the top /
//     bottom bands to the north / south tile, the top band to the fold partner (a ring of four row buffers there,
//     because the partner's eta phase is not part of my own dependency chain);
//   * the quads a neighbour waits for are computed first in every phase, the rest of the band hides the latency.
// The arithmetic is, operation by operation, that of k_baro_eta / k_baro_uv / k_baro_finish (the FMUL / FFMA sequence
// of their SASS is pinned with intrinsics; the parity test compares the two paths bit for bit).  The three divisions per
// point are the fast path of `/` without its range-check branch (div_nr, gb25_device.cuh): areas and spacings are normal numbers.
#include <cstdlib>

#include "gb25_internal.h"

#ifndef BP_MAXT
#define BP_MAXT 512                   // threads per CTA (upper bound; the launch picks the count that balances the band)
#endif
#ifndef BP_MAXQ
#define BP_MAXQ 4                     // quads (4 points) per thread: a band holds at most 4 * BP_MAXQ * BP_MAXT = 8192 points
#endif
#define BP_NCONST 9
#define BP_SPIN_LIMIT 4000000
#define BP_INBOX_OFFSET 8192          // byte offset of the pair inbox inside the peer-mapped flag buffer
#define BPF_ERR EX_NSLOT              // time-out word shared with the stream-ordered exchange

typedef unsigned long long u64;

struct BaroArgs {
  int nb, rows_base, rows_extra;   // band b owns rows_base + (b < rows_extra) rows
  int cap;                         // floats per shared-memory array (max rows per band * Nx)
  unsigned magic;                  // ceil(2^32 / (Nx/4)): q / (Nx/4) == __umulhi(q, magic) for quad index q < 2^16
  int nsub;
  float dtau;
  int bflags;                      // bit 0 wrap in x locally, 1 south wall, 2 north wall, 3 local north fold
  int seq0;                        // sequence numbers of this launch: seq0 (initial values), seq0 + 1 .. seq0 + nsub (substeps)
  int* err;                        // time-out word
  u64 *ll_eta, *ll_v;              // local pair buffers [nb][Nx]: last eta row / first V row of every band
  u64* inbox;                      // my pair inbox (written by the neighbouring tiles), layout below
  u64 *in_E, *in_W, *in_N, *in_S, *in_F;   // inboxes of the neighbouring tiles (nullptr: no such neighbour)
  float wgt[64];
};
// inbox layout (pairs): [W: eta(0,j)] [E: U(Nx+1,j)] [S: eta(i,0)] [N: V(i,Ny+1)] [F: 4 x V(i,Ny+1) through the fold]
__host__ __device__ __forceinline__ int inbox_W(int Ny) { return 0; }
__host__ __device__ __forceinline__ int inbox_E(int Ny) { return (Ny + 3) & ~3; }
__host__ __device__ __forceinline__ int inbox_S(int Ny) { return 2 * ((Ny + 3) & ~3); }
__host__ __device__ __forceinline__ int inbox_N(int Nx, int Ny) { return inbox_S(Ny) + Nx; }
__host__ __device__ __forceinline__ int inbox_F(int Nx, int Ny, int r) { return inbox_S(Ny) + (2 + r) * Nx; }
__host__ __device__ __forceinline__ int inbox_pairs(int Nx, int Ny) { return inbox_S(Ny) + 6 * Nx; }

// ---- (value, sequence) pairs
__device__ __forceinline__ u64 ll_pack(float v, int seq) { return ((u64)(unsigned)seq << 32) | (u64)__float_as_uint(v); }
__device__ __forceinline__ void ll_store1(u64* p, float v, int seq) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(ll_pack(v, seq)) : "memory");
}
__device__ __forceinline__ void ll_store4(u64* p, float v0, float v1, float v2, float v3, int seq) {   // p 16-byte aligned
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(ll_pack(v0, seq)), "l"(ll_pack(v1, seq)) : "memory");
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p + 2), "l"(ll_pack(v2, seq)), "l"(ll_pack(v3, seq)) : "memory");
}
struct LLGuard { int* s_abort; int* err; };
// consumers poll until the pair carries the expected sequence number (bounded: a lost neighbour raises the
// time-out word instead of hanging the GPU).  Not inlined: the polls are rare, the kernel must stay i-cache sized.
__device__ __noinline__ float ll_load1(const u64* p, int seq, int* s_abort, int* err) {
  int spins = 0;
  for (;;) {
    u64 w;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if ((int)(w >> 32) == seq) return __uint_as_float((unsigned)w);
    if (++spins > BP_SPIN_LIMIT || *(volatile int*)s_abort) {
      *(volatile int*)s_abort = 1; *(volatile int*)err = 1;
      return 0.f;
    }
  }
}
__device__ __noinline__ float4 ll_load4(const u64* p, int seq, int* s_abort, int* err) {   // p 16-byte aligned
  int spins = 0;
  for (;;) {
    u64 a, b, c, d;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 2) : "memory");
    if ((int)(a >> 32) == seq && (int)(b >> 32) == seq && (int)(c >> 32) == seq && (int)(d >> 32) == seq)
      return make_float4(__uint_as_float((unsigned)a), __uint_as_float((unsigned)b), __uint_as_float((unsigned)c), __uint_as_float((unsigned)d));
    if (++spins > BP_SPIN_LIMIT || *(volatile int*)s_abort) {
      *(volatile int*)s_abort = 1; *(volatile int*)err = 1;
      return make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

struct BandCtx {
  int Nx, NQ, Ny, PX, R, j0, b, cap, bflags;
  bool top, bot;
  float dtau, grav;
  float *s_eta, *s_U, *s_V;
  const float* s_c;
  const float* gc[BP_NCONST];
  LLGuard G;
};
enum { C_DYFC = 0, C_DXCF, C_AZCC, C_DXFC, C_DYCF, C_HFC, C_HCF, C_GU, C_GV };
__device__ __forceinline__ void f4_to(float (&d)[4], float4 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
__device__ __forceinline__ float4 to_f4(const float (&d)[4]) { return make_float4(d[0], d[1], d[2], d[3]); }
__device__ __forceinline__ void lds4(float (&d)[4], const float* p) { f4_to(d, *reinterpret_cast<const float4*>(p)); }
template <int NCS, int C>
__device__ __forceinline__ float cst1(const BandCtx& B, int lp, int q) {
  if (C < NCS) return B.s_c[C * B.cap + lp];
  return __ldg(B.gc[C] + q);
}
template <int NCS, int C>
__device__ __forceinline__ void cst4(const BandCtx& B, int lp, int q, float (&d)[4]) {
  if (C < NCS) lds4(d, B.s_c + C * B.cap + lp);
  else f4_to(d, __ldg(reinterpret_cast<const float4*>(B.gc[C] + q)));
}

// eta -= dtau * div(dy U, dx V) / Az for the quad at row jr, quad column iq.  XP: the tile has east / west neighbours.
template <int NCS, bool XP>
__device__ __forceinline__ void eta_quad(const BandCtx& B, const DevGrid& g, const BaroArgs& a, int jr, int iq, int seq) {
  const int Nx = B.Nx, PX = B.PX, ir = 4 * iq, p = jr * Nx + ir, j = B.j0 + jr;
  const int q2 = id2(g, ir + 1, j);
  float U[4], dy[4], V[4], dx[4], az[4], e[4], dU[4], dV[4], en[4], UE, dyE;
  lds4(U, B.s_U + p);
  cst4<NCS, C_DYFC>(B, p, q2, dy);
  if (iq < B.NQ - 1) { dyE = cst1<NCS, C_DYFC>(B, p + 4, q2 + 4); UE = B.s_U[p + 4]; }
  else if (!XP) { dyE = cst1<NCS, C_DYFC>(B, p + 4 - Nx, q2 + 4 - Nx); UE = B.s_U[p + 4 - Nx]; }
  else {   // east tile: U(Nx+1, j) of the previous substep
    dyE = __ldg(g.dyfc + q2 + 4);
    UE = ll_load1(a.inbox + inbox_E(B.Ny) + (j - 1), seq - 1, B.G.s_abort, B.G.err);
  }
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const float t = __fmul_rn(dy[c], U[c]);
    dU[c] = __fmaf_rn(c < 3 ? dy[c < 3 ? c + 1 : 3] : dyE, c < 3 ? U[c < 3 ? c + 1 : 3] : UE, -t);
  }
  lds4(V, B.s_V + p);
  cst4<NCS, C_DXCF>(B, p, q2, dx);
  const bool south_wall = j == 1 && (B.bflags & 2), north_edge = j == B.Ny && (B.bflags & 12);
  if (north_edge && (B.bflags & 4)) {
#pragma unroll
    for (int c = 0; c < 4; c++) dV[c] = -__fmul_rn(dx[c], V[c]);
  } else {
    float VN[4], dxN[4];
    if (north_edge) {   // folded row Ny+1: V[i,Ny+1] = -V[Nx-i+1,Ny]  (row Ny is the last row of this band)
      float m[4];
      lds4(m, B.s_V + jr * Nx + (Nx - ir - 4));
#pragma unroll
      for (int c = 0; c < 4; c++) VN[c] = -m[3 - c];
      f4_to(dxN, __ldg(reinterpret_cast<const float4*>(g.dxcf + q2 + PX)));
    } else if (jr < B.R - 1) {
      lds4(VN, B.s_V + p + Nx);
      cst4<NCS, C_DXCF>(B, p + Nx, q2 + PX, dxN);
    } else {            // the row above belongs to the next band, the north tile or the fold partner
      f4_to(dxN, __ldg(reinterpret_cast<const float4*>(g.dxcf + q2 + PX)));
      const u64* src = !B.top ? a.ll_v + (size_t)(B.b + 1) * Nx + ir
                              : a.inbox + (a.in_F ? inbox_F(Nx, B.Ny, (seq - 1) & 3) : inbox_N(Nx, B.Ny)) + ir;
      f4_to(VN, ll_load4(src, seq - 1, B.G.s_abort, B.G.err));
    }
    if (south_wall) {
#pragma unroll
      for (int c = 0; c < 4; c++) dV[c] = __fmul_rn(dxN[c], VN[c]);
    } else {
#pragma unroll
      for (int c = 0; c < 4; c++) dV[c] = __fmaf_rn(dxN[c], VN[c], -__fmul_rn(dx[c], V[c]));
    }
  }
  cst4<NCS, C_AZCC>(B, p, q2, az);
  lds4(e, B.s_eta + p);
#pragma unroll
  for (int c = 0; c < 4; c++) en[c] = __fsub_rn(e[c], div_nr(__fmul_rn(__fadd_rn(dU[c], dV[c]), B.dtau), az[c]));
  *reinterpret_cast<float4*>(B.s_eta + p) = to_f4(en);
  if (jr == B.R - 1 && !B.top) ll_store4(a.ll_eta + (size_t)B.b * Nx + ir, en[0], en[1], en[2], en[3], seq);   // -> the band above
  if (XP && iq == B.NQ - 1) ll_store1(a.in_E + inbox_W(B.Ny) + (j - 1), en[3], seq);                           // -> east tile: eta(0, j)
  if (a.in_N && j == B.Ny) ll_store4(a.in_N + inbox_S(B.Ny) + ir, en[0], en[1], en[2], en[3], seq);            // -> north tile: eta(i, 0)
}

// U += dtau (-g H d_x eta + GU), V likewise, running averages
template <int NCS, bool XP>
__device__ __forceinline__ void uv_quad(const BandCtx& B, const DevGrid& g, const BaroArgs& a, int jr, int iq, int seq,
                                        float wgt, float (&ae)[4], float (&au)[4], float (&av)[4]) {
  const int Nx = B.Nx, ir = 4 * iq, p = jr * Nx + ir, j = B.j0 + jr;
  const int q2 = id2(g, ir + 1, j);
  float e[4], eW, dxf[4], dyc[4], Hf[4], Hc[4], GU[4], GV[4], U[4], V[4], dxe[4], dye[4], Un[4], Vn[4];
  lds4(e, B.s_eta + p);
  if (iq > 0) eW = B.s_eta[p - 1];
  else if (!XP) eW = B.s_eta[p + Nx - 1];
  else eW = ll_load1(a.inbox + inbox_W(B.Ny) + (j - 1), seq, B.G.s_abort, B.G.err);      // west tile: eta(0, j) of this substep
  cst4<NCS, C_DXFC>(B, p, q2, dxf);
#pragma unroll
  for (int c = 0; c < 4; c++) dxe[c] = div_nr(__fsub_rn(e[c], c > 0 ? e[c > 0 ? c - 1 : 0] : eW), dxf[c]);
  if (j == 1 && (B.bflags & 2)) {
#pragma unroll
    for (int c = 0; c < 4; c++) dye[c] = 0.f;
  } else {
    float eS[4];
    if (jr > 0) lds4(eS, B.s_eta + p - Nx);
    else {
      const u64* src = !B.bot ? a.ll_eta + (size_t)(B.b - 1) * Nx + ir : a.inbox + inbox_S(B.Ny) + ir;   // the band below / the south tile
      f4_to(eS, ll_load4(src, seq, B.G.s_abort, B.G.err));
    }
    cst4<NCS, C_DYCF>(B, p, q2, dyc);
#pragma unroll
    for (int c = 0; c < 4; c++) dye[c] = div_nr(__fsub_rn(e[c], eS[c]), dyc[c]);
  }
  cst4<NCS, C_HFC>(B, p, q2, Hf); cst4<NCS, C_HCF>(B, p, q2, Hc);
  cst4<NCS, C_GU>(B, p, q2, GU); cst4<NCS, C_GV>(B, p, q2, GV);
  lds4(U, B.s_U + p);
  lds4(V, B.s_V + p);
#pragma unroll
  for (int c = 0; c < 4; c++) {
    Un[c] = __fmaf_rn(__fmaf_rn(__fmul_rn(Hf[c], B.grav), -dxe[c], GU[c]), B.dtau, U[c]);
    Vn[c] = __fmaf_rn(__fmaf_rn(__fmul_rn(Hc[c], B.grav), -dye[c], GV[c]), B.dtau, V[c]);
  }
  *reinterpret_cast<float4*>(B.s_U + p) = to_f4(Un);
  *reinterpret_cast<float4*>(B.s_V + p) = to_f4(Vn);
  if (jr == 0 && !B.bot) ll_store4(a.ll_v + (size_t)B.b * Nx + ir, Vn[0], Vn[1], Vn[2], Vn[3], seq);            // -> the band below
  if (XP && iq == 0) ll_store1(a.in_W + inbox_E(B.Ny) + (j - 1), Un[0], seq);                                   // -> west tile: U(Nx+1, j)
  if (a.in_S && j == 1) ll_store4(a.in_S + inbox_N(Nx, B.Ny) + ir, Vn[0], Vn[1], Vn[2], Vn[3], seq);            // -> south tile: V(i, Ny+1)
  if (XP && a.in_F && j == B.Ny)                                                                                // fold: V(i', Ny+1) = -V(Nx-i'+1, Ny)
    ll_store4(a.in_F + inbox_F(Nx, B.Ny, seq & 3) + (Nx - ir - 4), -Vn[3], -Vn[2], -Vn[1], -Vn[0], seq);
#pragma unroll
  for (int c = 0; c < 4; c++) { ae[c] = __fmaf_rn(e[c], wgt, ae[c]); au[c] = __fmaf_rn(Un[c], wgt, au[c]); av[c] = __fmaf_rn(Vn[c], wgt, av[c]); }
}

template <int NCS, bool XP>
__global__ void __launch_bounds__(BP_MAXT, 1) k_baro_persistent(DevGrid g, DevFields f, const __grid_constant__ BaroArgs a) {
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_abort;
  const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  BandCtx B;
  B.Nx = g.Nx; B.NQ = g.Nx >> 2; B.Ny = g.Ny; B.PX = g.PX; B.b = b;
  B.R = a.rows_base + (b < a.rows_extra ? 1 : 0);
  B.j0 = 1 + b * a.rows_base + min(b, a.rows_extra);
  B.cap = a.cap; B.bflags = a.bflags;
  B.top = b == a.nb - 1; B.bot = b == 0;
  B.dtau = a.dtau; B.grav = g.g;
  B.s_eta = sm; B.s_U = sm + a.cap; B.s_V = sm + 2 * a.cap; B.s_c = sm + 3 * a.cap;
  B.G.s_abort = &s_abort; B.G.err = a.err;
  {
    const float* gcs[BP_NCONST] = {g.dyfc, g.dxcf, g.azcc, g.dxfc, g.dycf, g.Hfc, g.Hcf, f.gU, f.gV};
#pragma unroll
    for (int c = 0; c < BP_NCONST; c++) B.gc[c] = gcs[c];
  }
  const int nq = B.R * B.NQ;
  if (tid == 0) s_abort = 0;
  const bool foldpush = XP && B.top && a.in_F != nullptr;
  // ---- prologue: band -> shared memory; the rows / columns the neighbours read in their first eta phase are
  // published as pairs numbered seq0 (the values the halo cells hold at this point)
  for (int q = tid; q < nq; q += T) {
    const int jr = __umulhi((unsigned)q, a.magic), iq = q - jr * B.NQ, p = 4 * q, ir = 4 * iq, j = B.j0 + jr;
    const int q2 = id2(g, ir + 1, j);
    const float4 u4 = *reinterpret_cast<const float4*>(f.bu + q2), v4 = *reinterpret_cast<const float4*>(f.bv + q2);
    *reinterpret_cast<float4*>(B.s_eta + p) = *reinterpret_cast<const float4*>(f.eta + q2);
    *reinterpret_cast<float4*>(B.s_U + p) = u4;
    *reinterpret_cast<float4*>(B.s_V + p) = v4;
#pragma unroll
    for (int c = 0; c < NCS; c++) *reinterpret_cast<float4*>(sm + (3 + c) * a.cap + p) = *reinterpret_cast<const float4*>(B.gc[c] + q2);
    if (jr == 0 && !B.bot) ll_store4(a.ll_v + (size_t)b * B.Nx + ir, v4.x, v4.y, v4.z, v4.w, a.seq0);
    if (XP && iq == 0) ll_store1(a.in_W + inbox_E(B.Ny) + (j - 1), u4.x, a.seq0);
    if (a.in_S && j == 1) ll_store4(a.in_S + inbox_N(B.Nx, B.Ny) + ir, v4.x, v4.y, v4.z, v4.w, a.seq0);
    if (foldpush && j == B.Ny) ll_store4(a.in_F + inbox_F(B.Nx, B.Ny, a.seq0 & 3) + (B.Nx - ir - 4), -v4.w, -v4.z, -v4.y, -v4.x, a.seq0);
  }
  // slot descriptors: (row << 16) | quad column, -1: no quad
  int desc[BP_MAXQ];
#pragma unroll
  for (int n = 0; n < BP_MAXQ; n++) {
    const int q = tid + n * T;
    const int jr = __umulhi((unsigned)q, a.magic), iq = q - jr * B.NQ;
    desc[n] = q < nq ? ((jr << 16) | iq) : -1;
  }
  float ae[BP_MAXQ][4], au[BP_MAXQ][4], av[BP_MAXQ][4];
#pragma unroll
  for (int n = 0; n < BP_MAXQ; n++)
#pragma unroll
    for (int c = 0; c < 4; c++) { ae[n][c] = 0.f; au[n][c] = 0.f; av[n][c] = 0.f; }
  __syncthreads();
#pragma unroll 1
  for (int m = 0; m < a.nsub; m++) {
    const int seq = a.seq0 + m + 1;
    // ================= eta phase.  Pass 0: the quads a neighbour waits for (last row; last column on a partitioned grid)
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
#pragma unroll
      for (int n = 0; n < BP_MAXQ; n++) {
        if (desc[n] >= 0) {
          const int jr = desc[n] >> 16, iq = desc[n] & 0xffff;
          const bool edge = jr == B.R - 1 || (XP && iq == B.NQ - 1);
          if (edge == (pass == 0)) eta_quad<NCS, XP>(B, g, a, jr, iq, seq);
        }
      }
    }
    __syncthreads();
    // ================= U,V phase.  Pass 0: first row; first column on a partitioned grid; the row pushed through the fold
    const float wgt = a.wgt[m];
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
#pragma unroll
      for (int n = 0; n < BP_MAXQ; n++) {
        if (desc[n] >= 0) {
          const int jr = desc[n] >> 16, iq = desc[n] & 0xffff;
          const bool edge = jr == 0 || (XP && iq == 0) || (foldpush && jr == B.R - 1);
          if (edge == (pass == 0)) uv_quad<NCS, XP>(B, g, a, jr, iq, seq, wgt, ae[n], au[n], av[n]);
        }
      }
    }
    __syncthreads();
  }
  // ---- epilogue: eta, U, V <- the weighted averages (k_baro_finish)
#pragma unroll
  for (int n = 0; n < BP_MAXQ; n++) {
    if (desc[n] >= 0) {
      const int jr = desc[n] >> 16, iq = desc[n] & 0xffff;
      const int q2 = id2(g, 4 * iq + 1, B.j0 + jr);
      *reinterpret_cast<float4*>(f.feta + q2) = to_f4(ae[n]); *reinterpret_cast<float4*>(f.fu + q2) = to_f4(au[n]);
      *reinterpret_cast<float4*>(f.fv + q2) = to_f4(av[n]);
      *reinterpret_cast<float4*>(f.eta + q2) = to_f4(ae[n]); *reinterpret_cast<float4*>(f.bu + q2) = to_f4(au[n]);
      *reinterpret_cast<float4*>(f.bv + q2) = to_f4(av[n]);
    }
  }
}

// ---------------------------------------------------------------------------------------- host side
struct BaroPlan {
  bool ok = false;
  int nb = 0, rows_base = 0, rows_extra = 0, cap = 0, ncs = 0, threads = 0;
  size_t smem = 0;
  u64* ll_local = nullptr;      // [2][nb][Nx] pairs: band-to-band rows
  int* local_err = nullptr;     // single-tile handles (a partitioned handle reports through the exchange's flag buffer)
  int seq = 1;                  // never 0: the pair buffers start zeroed, and 0 must not look like a published value
};

typedef void (*BaroKernel)(DevGrid, DevFields, const BaroArgs);
// shared-memory resident constants: 0, 2, 5 or all 9 of them (the largest count that fits)
static int baro_ncs_round(int ncs) { return ncs >= 9 ? 9 : ncs >= 5 ? 5 : ncs >= 2 ? 2 : 0; }
static BaroKernel baro_kernel(int ncs, bool xp) {
  switch (ncs) {
    case 0: return xp ? k_baro_persistent<0, true> : k_baro_persistent<0, false>;
    case 2: return xp ? k_baro_persistent<2, true> : k_baro_persistent<2, false>;
    case 5: return xp ? k_baro_persistent<5, true> : k_baro_persistent<5, false>;
    default: return xp ? k_baro_persistent<9, true> : k_baro_persistent<9, false>;
  }
}

static BaroPlan* baro_plan(Handle* h) {
  if (h->baro_plan) return (BaroPlan*)h->baro_plan;
  BaroPlan* P = new BaroPlan();
  h->baro_plan = P;
  if (!h->use_baro_persistent) return P;
  const DevGrid& g = h->g;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess || !prop.cooperativeLaunch) { cudaGetLastError(); return P; }
  int nb = prop.multiProcessorCount < g.Ny ? prop.multiProcessorCount : g.Ny;
  if (const char* nbe = getenv("GB25_BARO_BANDS")) { const int v = atoi(nbe); if (v >= 1 && v <= nb) nb = v; }
  P->nb = nb; P->rows_base = g.Ny / nb; P->rows_extra = g.Ny % nb;
  const int rmax = P->rows_base + (P->rows_extra ? 1 : 0);
  P->cap = rmax * g.Nx;
  // threads per CTA: the smallest warp multiple that covers the largest band in BP_MAXQ slots (the kernel is bound by
  // per-warp latency: measured at 1440 x 600, 480 threads x 4 slots 0.236 ms, 360 x 5 (one thread per quad column)
  // 0.306 ms, 384 x 5 0.252 ms, 608 x 3 with spills 0.268 ms)
  {
    const int NQ = g.Nx / 4, nqmax = rmax * NQ;
    int T = ((nqmax + BP_MAXQ - 1) / BP_MAXQ + 31) / 32 * 32;
    if (const char* te = getenv("GB25_BARO_THREADS")) { const int v = atoi(te); if (v >= 32 && v <= BP_MAXT && v * BP_MAXQ >= nqmax) T = v; }
    P->threads = T;
  }
  if (P->threads > BP_MAXT || P->cap > 4 * BP_MAXQ * P->threads || (g.Nx % 4) || (g.Hx % 4) || g.Nx > 32768 || h->cfg.nsubsteps > 64) return P;
  if ((size_t)BP_INBOX_OFFSET + (size_t)inbox_pairs(g.Nx, g.Ny) * sizeof(u64) > (size_t)(2 << 20)) return P;   // inbox lives in the 2 MiB flag buffer
  const size_t per = (size_t)P->cap * sizeof(float);
  const size_t maxsm = (size_t)prop.sharedMemPerBlockOptin - 1024;   // static shared + reserve
  if (3 * per > maxsm) return P;
  int ncs = (int)((maxsm - 3 * per) / per);
  if (ncs > BP_NCONST) ncs = BP_NCONST;
  if (const char* ce = getenv("GB25_BARO_NCS")) { const int v = atoi(ce); if (v >= 0 && v <= ncs) ncs = v; }
  ncs = baro_ncs_round(ncs);
  P->ncs = ncs; P->smem = (3 + ncs) * per;
  for (int xp = 0; xp < 2; xp++) {
    BaroKernel k = baro_kernel(ncs, xp != 0);
    if (cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P->smem) != cudaSuccess) { cudaGetLastError(); return P; }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, P->threads, P->smem) != cudaSuccess || per_sm < 1 ||
        per_sm * prop.multiProcessorCount < nb) { cudaGetLastError(); return P; }
  }
  const size_t llb = (size_t)2 * nb * g.Nx * sizeof(u64);
  if (cudaMalloc(&P->ll_local, llb) != cudaSuccess) { cudaGetLastError(); return P; }
  if (cudaMalloc(&P->local_err, 256) != cudaSuccess) { cudaGetLastError(); return P; }
  cudaMemset(P->ll_local, 0, llb);
  cudaMemset(P->local_err, 0, 256);
  P->ok = true;
  return P;
}
void baro_plan_free(Handle* h) {
  BaroPlan* P = (BaroPlan*)h->baro_plan;
  if (!P) return;
  if (P->ll_local) cudaFree(P->ll_local);
  if (P->local_err) cudaFree(P->local_err);
  delete P;
  h->baro_plan = nullptr;
}
int baro_check_timeout(Handle* h) {
  BaroPlan* P = (BaroPlan*)h->baro_plan;
  if (!P || !P->local_err || h->ex.on) return 0;   // (a partitioned handle reports through exchange_check_timeout)
  int flag = 0;
  cudaMemcpy(&flag, P->local_err, sizeof(int), cudaMemcpyDeviceToHost);
  return flag;
}

// returns false when the persistent kernel cannot run this configuration (the caller falls back to the substep kernels)
bool launch_barotropic_persistent(Handle* h, float dt) {
  BaroPlan* P = baro_plan(h);
  if (!P->ok) return false;
  const DevGrid& g = h->g;
  const gb25_config& c = h->cfg;
  if ((c.Rx > 1 || c.Ry > 1) && !h->ex.on) return false;   // a tile of a partitioned grid that is not connected yet
  BaroArgs a;
  a.nb = P->nb; a.rows_base = P->rows_base; a.rows_extra = P->rows_extra; a.cap = P->cap;
  a.magic = (unsigned)((0x100000000ULL + (unsigned long long)(g.Nx / 4) - 1) / (unsigned long long)(g.Nx / 4));
  a.nsub = c.nsubsteps;
  a.dtau = c.dtau_frac * dt;
  a.bflags = (c.Rx == 1 ? 1 : 0) | (c.ry == 0 ? 2 : 0) | ((c.ry == c.Ry - 1 && c.topo_y == GB25_TOPO_BOUNDED) ? 4 : 0) |
             ((c.ry == c.Ry - 1 && c.topo_y == GB25_TOPO_FOLD && c.Rx == 1) ? 8 : 0);
  a.seq0 = P->seq;
  a.err = h->ex.on ? h->ex.flags + BPF_ERR : P->local_err;
  a.ll_eta = P->ll_local; a.ll_v = P->ll_local + (size_t)P->nb * g.Nx;
  a.inbox = nullptr;
  a.in_E = a.in_W = a.in_N = a.in_S = a.in_F = nullptr;
  if (h->ex.on) {
    const Exchange& X = h->ex;
    auto inbox_of = [](int* flags) { return (u64*)((char*)flags + BP_INBOX_OFFSET); };
    a.inbox = inbox_of(X.flags);
    if (c.Rx > 1) { a.in_E = inbox_of(X.to[SLOT_E].flags); a.in_W = inbox_of(X.to[SLOT_W].flags); }
    if (c.ry < c.Ry - 1) a.in_N = inbox_of(X.to[SLOT_N].flags);
    if (c.ry > 0) a.in_S = inbox_of(X.to[SLOT_S].flags);
    if (c.topo_y == GB25_TOPO_FOLD && c.ry == c.Ry - 1 && c.Rx > 1) a.in_F = inbox_of(X.to[SLOT_FOLD].flags);
  }
  for (int m = 0; m < 64; m++) a.wgt[m] = m < c.nsubsteps ? h->weights[m] : 0.f;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(P->nb); cfg.blockDim = dim3(P->threads); cfg.dynamicSmemBytes = P->smem; cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  StageScope ts(h, "kernel:k_baro_persistent");
  cudaError_t ce = cudaLaunchKernelEx(&cfg, baro_kernel(P->ncs, a.in_E != nullptr), g, h->f, a);
  if (ce != cudaSuccess) { cudaGetLastError(); P->ok = false; return false; }
  h->count_launch();
  P->seq += c.nsubsteps + 1;   // seq0 itself numbers the initial values published in the prologue
  return true;
}

void baro_plan_prepare(Handle* h) { (void)baro_plan(h); }

// ---------------------------------------------------------------- kernel table (preload_kernels, gb25_api.cu)
KernelTable kernel_table_baro() {
  static const void* const k[] = {
    (const void*)k_baro_persistent<0, false>, (const void*)k_baro_persistent<0, true>, (const void*)k_baro_persistent<2, false>,
    (const void*)k_baro_persistent<2, true>, (const void*)k_baro_persistent<5, false>, (const void*)k_baro_persistent<5, true>,
    (const void*)k_baro_persistent<9, false>, (const void*)k_baro_persistent<9, true>,
  };
  return {k, (int)(sizeof k / sizeof k[0])};
}
