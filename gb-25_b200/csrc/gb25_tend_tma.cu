// gb25_tend_tma.cu — third-generation momentum tendency kernels: TMA-staged halo'd tiles, k-marching.
//
// ncu of the second generation (profiles/): half the instructions of the first, but only ~42 % issue-slot
// utilisation at 16-20 % occupancy — every level starts with ~40 dependent global loads per thread and there are
// too few warps to hide them.  Here the loads leave the instruction stream altogether:
//   * a CTA owns a 32 x 8 tile of columns and marches k = 1..Nz, one cell per thread;
//   * a dedicated producer warp issues, level by level, 3-D TMA box loads (cp.async.bulk.tensor) of the halo'd
//     tiles of u, v, zeta, dxU, dyV, w(k+1) and p into a 4-stage shared-memory ring; a "full" mbarrier per stage
//     tracks the bytes (expect_tx / complete_tx), an "empty" mbarrier per stage is armed by the consumer warps, so
//     there is no CTA-wide barrier in the k loop and warps drift up to three levels apart;
//   * the arithmetic reads its stencils from shared memory with immediate offsets (no address arithmetic, no
//     dependence on L1 hit rates), the vertical stencil of the own velocity lives in a register window, the
//     vertical momentum flux is carried from face to face;
//   * cells whose stencil touches bathymetry or a wall use the generic per-cell function (global memory), and
//     cells buried in rock are written as zero, exactly as in gb25_tend_v2.cu.
// The arithmetic (expression by expression) is that of k_gu_v2 / k_gv_v2, so the results are bit-identical.
#include <cuda.h>   // CUtensorMap types only; the encoder is fetched with cudaGetDriverEntryPoint (no libcuda link)

#include "gb25_internal.h"
#include "gb25_tend_generic.cuh"
#include "gb25_packed.cuh"

#define TMA_TX 32
#define TMA_TY 8
#ifndef TMA_MINB
#define TMA_MINB 2
#endif
#ifndef TMA_NST
#define TMA_NST 4
#endif
#define TMA_CW (TMA_TY)          // consumer warps (one per tile row); warp TMA_TY is the TMA producer

// ----------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  } while (!ok);
}
// Producer without a producer warp.  A dedicated producer warp costs a fifth of the CTA's registers (128 x 32) for seven
// instructions per level.  Instead every consumer warp counts itself out of a stage with one shared-memory atomic when it has
// finished a level, and the warp that arrives LAST issues the boxes of level k + NST into the stage it has just freed: the
// refill starts the moment the stage is free, nobody polls, and the CTA is 4 warps (four CTAs per SM at 128 registers).
// (A first version let lane 0 of warp 0 probe the empty barriers at the top of its own levels; whenever another warp was a few
// cycles behind, the probe failed and the refill slipped by a whole level: 1.6 - 2x slower than the producer warp, measured.)
__device__ __forceinline__ bool ring_release(int* cnt, int nwarps) {   // lane 0 of a warp, after __syncwarp(): true = last one out
  const int old = atomicAdd_block(cnt, 1);
  if (old != nwarps - 1) return false;
  *cnt = 0;             // re-armed before the refill is issued: nobody counts on this stage again until its data has landed
  return true;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

template <int DIR>
static __device__ __noinline__ float momentum_G_call3(const DevGrid* __restrict__ gp, const float* __restrict__ own,
                                                      const float* __restrict__ oth, const float* __restrict__ w,
                                                      const float* __restrict__ p, int i, int j, int k, float* wtop) {
  return momentum_G<DIR>(*gp, own, oth, w, p, i, j, k, wtop);
}
__device__ __forceinline__ float weno_sel_B3(float q0, float q1, float q2, float q3, float q4, float q5, int B, bool left, float eps) {
  if (B == 3) {
    const float v0 = left ? q0 : q5, v1 = left ? q1 : q4, v2 = left ? q2 : q3, v3 = left ? q3 : q2, v4 = left ? q4 : q1;
    return weno5(v0, v1, v2, v3, v4, eps);
  }
  if (B == 2) return left ? weno3(q1, q2, q3, eps) : weno3(q4, q3, q2, eps);
  return left ? q2 : q3;
}
__device__ __forceinline__ float weno5_fs_sel3(const float (&q)[6], const float (&s)[6], bool left, float eps) {
  return weno5_fs_selq(q, s, left, eps);
}

struct TmaMaps7 { CUtensorMap m[7]; };

// shared-memory tile geometry (floats).  Gu: own direction x.
//   U  (TX+8) x (TY+8) origin (-4,-4) | V (TX+4) x (TY+8) origin (-4,-4) | Z TX x (TY+8) origin (0,-4)
//   DX, DY, W (TX+8) x TY origin (-4,0) | P (TX+4) x TY origin (-4,0)
#define GU_PU (TMA_TX + 8)
#define GU_PV (TMA_TX + 4)
#define GU_PZ (TMA_TX)
#define GU_PD (TMA_TX + 8)
#define GU_PP (TMA_TX + 4)
#define GU_OFF_U 0
#define GU_OFF_V (GU_OFF_U + GU_PU * (TMA_TY + 8))
#define GU_OFF_Z (GU_OFF_V + GU_PV * (TMA_TY + 8))
#define GU_OFF_DX (GU_OFF_Z + GU_PZ * (TMA_TY + 8))
#define GU_OFF_DY (GU_OFF_DX + GU_PD * TMA_TY)
#define GU_OFF_W (GU_OFF_DY + GU_PD * TMA_TY)
#define GU_OFF_P (GU_OFF_W + GU_PD * TMA_TY)
#define GU_STAGE (GU_OFF_P + GU_PP * TMA_TY)

__global__ void __launch_bounds__(TMA_TX * (TMA_TY + 1), TMA_MINB)
k_gu_tma(DevGrid g, const DevGrid* __restrict__ gp, const __grid_constant__ TmaMaps7 tm, const float* __restrict__ u,
         const float* __restrict__ v, const float* __restrict__ w, const float* __restrict__ p, float* __restrict__ G,
         const float* __restrict__ carry) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar[TMA_NST], ebar[TMA_NST];   // full (TMA landed) / empty (all consumer warps done) per stage
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TMA_TX + tx;
  const int i0 = blockIdx.x * TMA_TX + 1, j0 = blockIdx.y * TMA_TY + 1;
  const int I0 = i0 + g.Hx - 1, J0 = j0 + g.Hy - 1;
  const int i = i0 + tx, j = j0 + ty;
  const bool valid = i <= g.Nx && j <= g.Ny;
  const int PX = g.PX, n2 = g.n2, Nz = g.Nz;
  const int q2 = id2(g, min(i, g.Nx), min(j, g.Ny));
  const float eps = g.eps;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TMA_NST; s++) { mbar_init(&bar[s], 1); mbar_init(&ebar[s], TMA_CW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int k) {   // level k -> stage (k-1) % NST   (called by thread 0 only)
    const int s = (k - 1) % TMA_NST;
    float* sm = smem + s * GU_STAGE;
    const int K = k + g.Hz - 1;
    mbar_expect_tx(&bar[s], GU_STAGE * sizeof(float));
    tma_load_3d(sm + GU_OFF_U, &tm.m[0], &bar[s], I0 - 4, J0 - 4, K);
    tma_load_3d(sm + GU_OFF_V, &tm.m[1], &bar[s], I0 - 4, J0 - 4, K);
    tma_load_3d(sm + GU_OFF_Z, &tm.m[2], &bar[s], I0, J0 - 4, K);
    tma_load_3d(sm + GU_OFF_DX, &tm.m[3], &bar[s], I0 - 4, J0, K);
    tma_load_3d(sm + GU_OFF_DY, &tm.m[4], &bar[s], I0 - 4, J0, K);
    tma_load_3d(sm + GU_OFF_W, &tm.m[5], &bar[s], I0 - 4, J0, K + 1);
    tma_load_3d(sm + GU_OFF_P, &tm.m[6], &bar[s], I0 - 4, J0, K);
  };
  if (ty == TMA_TY) {   // ===== producer warp: keeps the ring full, one elected lane issues the TMA loads
    if (tx == 0)
      for (int k = 1; k <= Nz; k++) {
        if (k > TMA_NST) mbar_wait(&ebar[(k - 1) % TMA_NST], (((k - 1) / TMA_NST) - 1) & 1);
        issue(k);
      }
    return;
  }
  // ---- hoisted 2-D data
  const float m1 = g.dxfc[q2], rV0 = g.azfc[q2];
  const float fbar = (g.fff[q2] + g.fff[q2 + PX]) * 0.5f;
  const float mv00 = g.dxcf[q2 - 1], mv01 = g.dxcf[q2], mv10 = g.dxcf[q2 - 1 + PX], mv11 = g.dxcf[q2 + PX];
  const float az0 = g.azcc[q2 - 2], az1 = g.azcc[q2 - 1], az2 = g.azcc[q2], az3 = g.azcc[q2 + 1];
  const int kbc = g.kb[q2];
  const int kgen = g.kgen2[q2], kzero = g.kzero2[q2];   // generic cells (kzero < k <= kgen) belong to k_generic_list
  // ---- vertical register window of u
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  float WU[7];
#pragma unroll
  for (int m = 0; m < 7; m++) WU[m] = __ldg(u + q3 + (ptrdiff_t)(m - 3) * n2);
  float Wb = 0.f;
  const int ou = (ty + 4) * GU_PU + (tx + 4), ov = (ty + 4) * GU_PV + (tx + 4), oz = (ty + 4) * GU_PZ + tx;
  const int od = ty * GU_PD + (tx + 4), op = ty * GU_PP + (tx + 4);
  for (int k = 1; k <= Nz; k++, q3 += n2) {
    const int s = (k - 1) % TMA_NST;
    mbar_wait(&bar[s], ((k - 1) / TMA_NST) & 1);
    const float* sm = smem + s * GU_STAGE;
    const float* U = sm + GU_OFF_U + ou; const float* V = sm + GU_OFF_V + ov; const float* Z = sm + GU_OFF_Z + oz;
    const float* DX = sm + GU_OFF_DX + od; const float* DY = sm + GU_OFF_DY + od; const float* W = sm + GU_OFF_W + od;
    const float* P = sm + GU_OFF_P + op;
    float out = 0.f;
    bool skip = false;
    if (valid) {
      if (k <= kzero) {
        out = 0.f; Wb = 0.f;
      } else if (k <= kgen) {
        skip = true;
      } else {
        if (k == kgen + 1 && kgen > 0) Wb = carry[q2];
        const float dz = g.dzc[k + g.Hz - 1];
        const float own0 = WU[3];
        const bool lown = own0 > 0.f;
        const float xm0 = mv00 * V[-1], xm1 = mv10 * V[GU_PV - 1];
        const float x00 = mv01 * V[0], x01 = mv11 * V[GU_PV];
        const float oavg = ((xm0 + xm1) * 0.5f + (x00 + x01) * 0.5f) * 0.5f;
        const float ohat = oavg / m1;
        float zq[6], zs[6], zr[6];
#pragma unroll
        for (int m = 0; m < 6; m++) {
          const int b = m - 2;
          zq[m] = Z[b * GU_PZ];
          zs[m] = (U[(b - 1) * GU_PU] + U[b * GU_PU]) * 0.5f;
          zr[m] = (V[b * GU_PV - 1] + V[b * GU_PV]) * 0.5f;
        }
        const float zR = weno5_vs_selq(zq, zs, zr, ohat > 0.f, eps);
        const float Hterm = -ohat * zR;
        float dOw[6], dv[6], dK[6], sK[6];
#pragma unroll
        for (int m = 0; m < 6; m++) {
          const int a = m - 3;
          dOw[m] = DX[a];
          dv[m] = DX[a] + DY[a];
          const float o0 = U[a], o1 = U[a + 1];
          dK[m] = o1 * o1 * 0.5f - o0 * o0 * 0.5f;
          sK[m] = (o0 + o1) * 0.5f;
        }
        const float dvs = sym4(DY[-2], DY[-1], DY[0], DY[1], 2);
        const float duR = weno5_fs_sel3(dOw, dv, lown, eps);
        const float Phi = own0 * (dvs + duR);
        const float dKo = weno5_fs_sel3(dK, sK, lown, eps);
        float kc[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
          const int b = m - 1;
          const float t0 = V[b * GU_PV], tm_ = V[b * GU_PV - 1];
          kc[m] = t0 * t0 * 0.5f - tm_ * tm_ * 0.5f;
        }
        const float Bterm = (dKo + sym4(kc[0], kc[1], kc[2], kc[3], 2)) / m1;
        const int Bw = (g.immersed && k + 1 > Nz) ? 1 : 2;
        const float wt = sym4(az0 * W[-2], az1 * W[-1], az2 * W[0], az3 * W[1], Bw);
        const int Bz = zbuf(g, kbc, k + 1, 3);
        const float Wt = wt * weno_sel_B3(WU[1], WU[2], WU[3], WU[4], WU[5], WU[6], Bz, wt > 0.f, eps);
        const float Vterm = (1.f / (rV0 * dz)) * (Phi + (Wt - Wb));
        Wb = Wt;
        const float cor = -(fbar * oavg / m1);
        const float dp = (P[0] - P[-1]) / m1;
        out = -(Hterm + Vterm + Bterm) - cor - dp;
      }
      if (!skip) G[q3] = out;
#pragma unroll
      for (int m = 0; m < 6; m++) WU[m] = WU[m + 1];
      WU[6] = __ldg(u + q3 + (size_t)4 * n2);
    }
    __syncwarp();
    if (tx == 0) mbar_arrive(&ebar[s]);   // this warp is done with stage s
  }
}

// Gv: own direction y.  V (TX+8) x (TY+8) origin (-4,-4) | U (TX+8) x (TY+8) origin (-4,-4) | Z (TX+8) x TY origin (-4,0)
//     DY, DX, W, P: TX x (TY+8) origin (0,-4)
#define GV_PV (TMA_TX + 8)
#define GV_PU (TMA_TX + 8)
#define GV_PZ (TMA_TX + 8)
#define GV_PD (TMA_TX)
#define GV_OFF_V 0
#define GV_OFF_U (GV_OFF_V + GV_PV * (TMA_TY + 8))
#define GV_OFF_Z (GV_OFF_U + GV_PU * (TMA_TY + 8))
#define GV_OFF_DY (GV_OFF_Z + GV_PZ * TMA_TY)
#define GV_OFF_DX (GV_OFF_DY + GV_PD * (TMA_TY + 8))
#define GV_OFF_W (GV_OFF_DX + GV_PD * (TMA_TY + 8))
#define GV_OFF_P (GV_OFF_W + GV_PD * (TMA_TY + 8))
#define GV_STAGE (GV_OFF_P + GV_PD * (TMA_TY + 8))

__global__ void __launch_bounds__(TMA_TX * (TMA_TY + 1), TMA_MINB)
k_gv_tma(DevGrid g, const DevGrid* __restrict__ gp, const __grid_constant__ TmaMaps7 tm, const float* __restrict__ u,
         const float* __restrict__ v, const float* __restrict__ w, const float* __restrict__ p, float* __restrict__ G,
         const float* __restrict__ carry) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar[TMA_NST], ebar[TMA_NST];   // full (TMA landed) / empty (all consumer warps done) per stage
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TMA_TX + tx;
  const int i0 = blockIdx.x * TMA_TX + 1, j0 = blockIdx.y * TMA_TY + 1;
  const int I0 = i0 + g.Hx - 1, J0 = j0 + g.Hy - 1;
  const int i = i0 + tx, j = j0 + ty;
  const bool valid = i <= g.Nx && j <= g.Ny;
  const int PX = g.PX, n2 = g.n2, Nz = g.Nz;
  const int q2 = id2(g, min(i, g.Nx), min(j, g.Ny));
  const float eps = g.eps;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TMA_NST; s++) { mbar_init(&bar[s], 1); mbar_init(&ebar[s], TMA_CW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int k) {
    const int s = (k - 1) % TMA_NST;
    float* sm = smem + s * GV_STAGE;
    const int K = k + g.Hz - 1;
    mbar_expect_tx(&bar[s], GV_STAGE * sizeof(float));
    tma_load_3d(sm + GV_OFF_V, &tm.m[0], &bar[s], I0 - 4, J0 - 4, K);
    tma_load_3d(sm + GV_OFF_U, &tm.m[1], &bar[s], I0 - 4, J0 - 4, K);
    tma_load_3d(sm + GV_OFF_Z, &tm.m[2], &bar[s], I0 - 4, J0, K);
    tma_load_3d(sm + GV_OFF_DY, &tm.m[3], &bar[s], I0, J0 - 4, K);
    tma_load_3d(sm + GV_OFF_DX, &tm.m[4], &bar[s], I0, J0 - 4, K);
    tma_load_3d(sm + GV_OFF_W, &tm.m[5], &bar[s], I0, J0 - 4, K + 1);
    tma_load_3d(sm + GV_OFF_P, &tm.m[6], &bar[s], I0, J0 - 4, K);
  };
  if (ty == TMA_TY) {
    if (tx == 0)
      for (int k = 1; k <= Nz; k++) {
        if (k > TMA_NST) mbar_wait(&ebar[(k - 1) % TMA_NST], (((k - 1) / TMA_NST) - 1) & 1);
        issue(k);
      }
    return;
  }
  const float m1 = g.dycf[q2], rV0 = g.azcf[q2];
  const float fbar = (g.fff[q2] + g.fff[q2 + 1]) * 0.5f;
  const float mu00 = g.dyfc[q2 - PX], mu01 = g.dyfc[q2 - PX + 1], mu10 = g.dyfc[q2], mu11 = g.dyfc[q2 + 1];
  const float az0 = g.azcc[q2 - 2 * PX], az1 = g.azcc[q2 - PX], az2 = g.azcc[q2], az3 = g.azcc[q2 + PX];
  const int kbc = g.kb[q2];
  const int kgen = g.kgen2[q2], kzero = g.kzero2[q2];   // generic cells (kzero < k <= kgen) belong to k_generic_list
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  float WV[7];
#pragma unroll
  for (int m = 0; m < 7; m++) WV[m] = __ldg(v + q3 + (ptrdiff_t)(m - 3) * n2);
  float Wb = 0.f;
  const int ov = (ty + 4) * GV_PV + (tx + 4), ou = (ty + 4) * GV_PU + (tx + 4), oz = ty * GV_PZ + (tx + 4);
  const int od = (ty + 4) * GV_PD + tx;
  for (int k = 1; k <= Nz; k++, q3 += n2) {
    const int s = (k - 1) % TMA_NST;
    mbar_wait(&bar[s], ((k - 1) / TMA_NST) & 1);
    const float* sm = smem + s * GV_STAGE;
    const float* V = sm + GV_OFF_V + ov; const float* U = sm + GV_OFF_U + ou; const float* Z = sm + GV_OFF_Z + oz;
    const float* DY = sm + GV_OFF_DY + od; const float* DX = sm + GV_OFF_DX + od; const float* W = sm + GV_OFF_W + od;
    const float* P = sm + GV_OFF_P + od;
    float out = 0.f;
    bool skip = false;
    if (valid) {
      if (k <= kzero) {
        out = 0.f; Wb = 0.f;
      } else if (k <= kgen) {
        skip = true;
      } else {
        if (k == kgen + 1 && kgen > 0) Wb = carry[q2];
        const float dz = g.dzc[k + g.Hz - 1];
        const float own0 = WV[3];
        const bool lown = own0 > 0.f;
        const float xm0 = mu00 * U[-GV_PU], xm1 = mu01 * U[-GV_PU + 1];
        const float x00 = mu10 * U[0], x01 = mu11 * U[1];
        const float oavg = ((xm0 + xm1) * 0.5f + (x00 + x01) * 0.5f) * 0.5f;
        const float ohat = oavg / m1;
        float zq[6], zs[6], zr[6];
#pragma unroll
        for (int m = 0; m < 6; m++) {           // window along x: cols i-2 .. i+3
          const int b = m - 2;
          zq[m] = Z[b];
          zs[m] = (V[b - 1] + V[b]) * 0.5f;
          zr[m] = (U[b - GV_PU] + U[b]) * 0.5f;
        }
        const float zR = weno5_vs_selq(zq, zs, zr, ohat > 0.f, eps);
        const float Hterm = ohat * zR;
        float dOw[6], dv[6], dK[6], sK[6];
#pragma unroll
        for (int m = 0; m < 6; m++) {           // window along y: rows j-3 .. j+2
          const int a = m - 3;
          dOw[m] = DY[a * GV_PD];
          dv[m] = DX[a * GV_PD] + DY[a * GV_PD];
          const float o0 = V[a * GV_PV], o1 = V[(a + 1) * GV_PV];
          dK[m] = o1 * o1 * 0.5f - o0 * o0 * 0.5f;
          sK[m] = (o0 + o1) * 0.5f;
        }
        const float dus = sym4(DX[-2 * GV_PD], DX[-GV_PD], DX[0], DX[GV_PD], 2);
        const float dvR = weno5_fs_sel3(dOw, dv, lown, eps);
        const float Phi = own0 * (dus + dvR);
        const float dKo = weno5_fs_sel3(dK, sK, lown, eps);
        float kc[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {           // cols i-1 .. i+2
          const int b = m - 1;
          const float t0 = U[b], tm_ = U[b - GV_PU];
          kc[m] = t0 * t0 * 0.5f - tm_ * tm_ * 0.5f;
        }
        const float Bterm = (dKo + sym4(kc[0], kc[1], kc[2], kc[3], 2)) / m1;
        const int Bw = (g.immersed && k + 1 > Nz) ? 1 : 2;
        const float wt = sym4(az0 * W[-2 * GV_PD], az1 * W[-GV_PD], az2 * W[0], az3 * W[GV_PD], Bw);
        const int Bz = zbuf(g, kbc, k + 1, 3);
        const float Wt = wt * weno_sel_B3(WV[1], WV[2], WV[3], WV[4], WV[5], WV[6], Bz, wt > 0.f, eps);
        const float Vterm = (1.f / (rV0 * dz)) * (Phi + (Wt - Wb));
        Wb = Wt;
        const float cor = fbar * oavg / m1;
        const float dp = (P[0] - P[-GV_PD]) / m1;
        out = -(Hterm + Vterm + Bterm) - cor - dp;
      }
      if (!skip) G[q3] = out;
#pragma unroll
      for (int m = 0; m < 6; m++) WV[m] = WV[m + 1];
      WV[6] = __ldg(v + q3 + (size_t)4 * n2);
    }
    __syncwarp();
    if (tx == 0) mbar_arrive(&ebar[s]);
  }
}

// =====================================================================================
// Packed (FP32x2) variants of the momentum kernels: the same tiles, ring and barriers, but 128 consumer threads, each
// owning TWO x-adjacent cells whose arithmetic runs two lanes wide in FFMA2 / FADD2 / FMUL2 (gb25_packed.cuh).
// Windows across y and z are natural float2 loads (the pair is contiguous in shared memory); windows along x are
// assembled from two 32-bit loads when the offset is odd.  Divisions by metrics use reciprocals hoisted out of the
// k loop (one IEEE division per column instead of five per cell and level).
// =====================================================================================
__device__ __forceinline__ float2 ldp(const float* p) { return make_float2(p[0], p[1]); }                       // any alignment
__device__ __forceinline__ float2 ldpa(const float* p) { return *reinterpret_cast<const float2*>(p); }           // 8-byte aligned
__device__ __forceinline__ float2 psym4(float2 q0, float2 q1, float2 q2, float2 q3, int B) {
  return B >= 2 ? pfmas(q0, -1.f / 12.f, pfmas(q1, 7.f / 12.f, pfmas(q2, 7.f / 12.f, pmuls(q3, -1.f / 12.f))))
                : pmuls(padd(q1, q2), 0.5f);
}
__device__ __forceinline__ float2 pneg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 pweno5_fs_sel(const float2* q, const float2* s, bool lx, bool ly, float eps) {
  return pweno5_fs(psel(q[0], q[5], lx, ly), psel(q[1], q[4], lx, ly), psel(q[2], q[3], lx, ly), psel(q[3], q[2], lx, ly), psel(q[4], q[1], lx, ly),
                   psel(s[0], s[5], lx, ly), psel(s[1], s[4], lx, ly), psel(s[2], s[3], lx, ly), psel(s[3], s[2], lx, ly), psel(s[4], s[1], lx, ly), eps);
}
__device__ __forceinline__ float2 pweno5_vs_sel(const float2* q, const float2* s, const float2* r, bool lx, bool ly, float eps) {
  return pweno5_vs(psel(q[0], q[5], lx, ly), psel(q[1], q[4], lx, ly), psel(q[2], q[3], lx, ly), psel(q[3], q[2], lx, ly), psel(q[4], q[1], lx, ly),
                   psel(s[0], s[5], lx, ly), psel(s[1], s[4], lx, ly), psel(s[2], s[3], lx, ly), psel(s[3], s[2], lx, ly), psel(s[4], s[1], lx, ly),
                   psel(r[0], r[5], lx, ly), psel(r[1], r[4], lx, ly), psel(r[2], r[3], lx, ly), psel(r[3], r[2], lx, ly), psel(r[4], r[1], lx, ly), eps);
}
// vertical reconstruction of a pair from the two register windows; order reduction falls back to the scalar code
__device__ __forceinline__ float2 pvert(const float (&A)[7], const float (&B)[7], int BA, int BB, float2 wt, float eps) {
  if (BA == 3 && BB == 3) {
    float2 q[6];
#pragma unroll
    for (int m = 0; m < 6; m++) q[m] = make_float2(A[m + 1], B[m + 1]);
    return pweno5_sel(q, wt.x > 0.f, wt.y > 0.f, eps);
  }
  return make_float2(weno_sel_B3(A[1], A[2], A[3], A[4], A[5], A[6], BA, wt.x > 0.f, eps),
                     weno_sel_B3(B[1], B[2], B[3], B[4], B[5], B[6], BB, wt.y > 0.f, eps));
}

// Packed kernels: the Gv stage holds exactly the rows that are read — P rows j-1 .. j+7 (TY + 1), W rows j-2 .. j+8 (TY + 3) —
// and, with the AB2 epilogue, one more box: the TX x TY tile of G- (a global load in the consumer's instruction stream is
// needed the moment it is issued: +0.10 ms per kernel, measured; as an eighth TMA box its latency is hidden by the ring).
#define GVP_OFF_W (GV_OFF_DX + GV_PD * (TMA_TY + 8))
#define GVP_OFF_P (GVP_OFF_W + GV_PD * (TMA_TY + 3))
#define GVP_STAGE (GVP_OFF_P + GV_PD * (TMA_TY + 1))
#define MOM_GM_TILE (TMA_TX * TMA_TY)
// ring depth: measured at 1440 x 600 x 50 (ms per launch pair): 3 stages 1.64, 4 1.63, 5 1.55, 6 1.66 (occupancy drops to 2 CTAs)
// Producer choice per kernel, measured at 1440 x 600 x 50 (ms per launch; Gu / Gv / tracers):
//   producer warp, 3 CTAs per SM, 5 / 6 stages   0.662 / 0.662 / 0.845
//   last-arriver refill, 3 CTAs, 5 / 6 stages    0.703 / 0.691 / 0.804
//   last-arriver refill, 4 CTAs, 4 / 5 stages    0.681 / 0.789 / 0.807   (the Gv stages are 15.9 KB: only 3 CTAs fit)
// 16 instead of 12 consumer warps per SM buy nothing: the kernels are bound by the FP32 pipes, not by latency.
#ifndef MOM_OPP
#define MOM_OPP 0       // momentum: dedicated producer warp (5-warp CTAs)
#endif
#ifndef TR_OPP
#define TR_OPP 1        // tracers: last-arriver refill (4-warp CTAs)
#endif
#ifndef MOM_NST
#define MOM_NST (MOM_OPP ? 4 : 5)
#endif
#ifndef MOM_MINB
#define MOM_MINB (MOM_OPP ? 4 : 3)
#endif
#define MOM_THREADS (MOM_OPP ? 128 : 160)
#define TR_THREADS (TR_OPP ? 128 : 160)
// AB2 epilogue (template flag AB2): the thread that owns a column pair also applies the NEXT step's QuasiAdamsBashforth2
// update to its own velocity component — u* = mask(u + dt (c1 Gn - c2 G-)) into the other state buffer — and accumulates,
// level by level as it marches, the barotropic forcing sum dz mask(c1 Gn - c2 G-) and the transport sum dz u* that the
// split-explicit solve and the corrector of the next step need (rows A8, A9, A1 of SURVEY 8a: the arithmetic and the
// summation order of k_ab2_uv, bit for bit).  Generic cells (bathymetry in the stencil) were written by k_generic_list
// earlier in the stream and are read back here.
struct MomAb2 {
  float* own_next;      // the other state buffer of this component
  float* gsum;          // 2-D: barotropic forcing  (-> GU / GV)
  float* usum;          // 2-D: sum dz u*           (-> the corrector's column sums)
  float dt, c1, c2;
};
template <int DIR, bool AB2>
__global__ void __launch_bounds__(MOM_THREADS, MOM_MINB)
k_mom_tma_p2(DevGrid g, const __grid_constant__ TmaMaps7 tm, const __grid_constant__ CUtensorMap tmg, const float* __restrict__ own_g,
             float* __restrict__ G, const float* __restrict__ carry, const MomAb2 ab) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar[MOM_NST], ebar[MOM_NST];
  __shared__ int rcnt[MOM_NST];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
  const int i0 = blockIdx.x * TMA_TX + 1, j0 = blockIdx.y * TMA_TY + 1;
  const int I0 = i0 + g.Hx - 1, J0 = j0 + g.Hy - 1;
  const int PX = g.PX, n2 = g.n2, Nz = g.Nz;
  const float eps = g.eps;
  constexpr int OFF_GM = DIR == 0 ? GU_STAGE : GVP_STAGE;
  constexpr int STAGE = OFF_GM + (AB2 ? MOM_GM_TILE : 0);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < MOM_NST; s++) { mbar_init(&bar[s], 1); mbar_init(&ebar[s], 4); rcnt[s] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int k, int s) {
        float* sm = smem + s * STAGE;
        const int K = k + g.Hz - 1;
        mbar_expect_tx(&bar[s], STAGE * sizeof(float));
        if (DIR == 0) {
          tma_load_3d(sm + GU_OFF_U, &tm.m[0], &bar[s], I0 - 4, J0 - 4, K);
          tma_load_3d(sm + GU_OFF_V, &tm.m[1], &bar[s], I0 - 4, J0 - 4, K);
          tma_load_3d(sm + GU_OFF_Z, &tm.m[2], &bar[s], I0, J0 - 4, K);
          tma_load_3d(sm + GU_OFF_DX, &tm.m[3], &bar[s], I0 - 4, J0, K);
          tma_load_3d(sm + GU_OFF_DY, &tm.m[4], &bar[s], I0 - 4, J0, K);
          tma_load_3d(sm + GU_OFF_W, &tm.m[5], &bar[s], I0 - 4, J0, K + 1);
          tma_load_3d(sm + GU_OFF_P, &tm.m[6], &bar[s], I0 - 4, J0, K);
        } else {
          tma_load_3d(sm + GV_OFF_V, &tm.m[0], &bar[s], I0 - 4, J0 - 4, K);
          tma_load_3d(sm + GV_OFF_U, &tm.m[1], &bar[s], I0 - 4, J0 - 4, K);
          tma_load_3d(sm + GV_OFF_Z, &tm.m[2], &bar[s], I0 - 4, J0, K);
          tma_load_3d(sm + GV_OFF_DY, &tm.m[3], &bar[s], I0, J0 - 4, K);
          tma_load_3d(sm + GV_OFF_DX, &tm.m[4], &bar[s], I0, J0 - 4, K);
          tma_load_3d(sm + GVP_OFF_W, &tm.m[5], &bar[s], I0, J0 - 2, K + 1);
          tma_load_3d(sm + GVP_OFF_P, &tm.m[6], &bar[s], I0, J0 - 1, K);
        }
        if (AB2) tma_load_3d(sm + OFF_GM, &tmg, &bar[s], I0, J0, K);
  };
#if MOM_OPP
  if (tid == 0)
    for (int k = 1; k <= min(Nz, MOM_NST); k++) issue(k, k - 1);
#else
  if (ty >= TMA_TY) {   // ===== producer warp
    if (tid == 128)
      for (int k = 1; k <= Nz; k++) {
        const int s = (k - 1) % MOM_NST;
        if (k > MOM_NST) mbar_wait(&ebar[s], (((k - 1) / MOM_NST) - 1) & 1);
        issue(k, s);
      }
    return;
  }
#endif
  const int lx = 2 * tx;
  const int i = min(i0 + lx, g.Nx - 1), j = min(j0 + ty, g.Ny);
  const bool valid = (i0 + lx + 1) <= g.Nx && (j0 + ty) <= g.Ny;
  const int q2 = id2(g, i, j);
  // ---- hoisted 2-D data (pairs over the two cells c = 0, 1)
  float2 rm1, rAz, fbar, mA0, mA1, mB0, mB1, azp[4];
  int kbA, kbB;
  if (DIR == 0) {
    rm1 = make_float2(1.f / g.dxfc[q2], 1.f / g.dxfc[q2 + 1]);
    rAz = make_float2(1.f / g.azfc[q2], 1.f / g.azfc[q2 + 1]);
    fbar = make_float2((g.fff[q2] + g.fff[q2 + PX]) * 0.5f, (g.fff[q2 + 1] + g.fff[q2 + 1 + PX]) * 0.5f);
    mA0 = make_float2(g.dxcf[q2 - 1], g.dxcf[q2]); mA1 = make_float2(g.dxcf[q2 - 1 + PX], g.dxcf[q2 + PX]);        // v at i-1: rows j, j+1
    mB0 = make_float2(g.dxcf[q2], g.dxcf[q2 + 1]); mB1 = make_float2(g.dxcf[q2 + PX], g.dxcf[q2 + 1 + PX]);        // v at i
#pragma unroll
    for (int n = 0; n < 4; n++) azp[n] = make_float2(g.azcc[q2 - 2 + n], g.azcc[q2 - 1 + n]);
  } else {
    rm1 = make_float2(1.f / g.dycf[q2], 1.f / g.dycf[q2 + 1]);
    rAz = make_float2(1.f / g.azcf[q2], 1.f / g.azcf[q2 + 1]);
    fbar = make_float2((g.fff[q2] + g.fff[q2 + 1]) * 0.5f, (g.fff[q2 + 1] + g.fff[q2 + 2]) * 0.5f);
    mA0 = make_float2(g.dyfc[q2 - PX], g.dyfc[q2 - PX + 1]); mA1 = make_float2(g.dyfc[q2 - PX + 1], g.dyfc[q2 - PX + 2]);  // u at j-1: cols i, i+1
    mB0 = make_float2(g.dyfc[q2], g.dyfc[q2 + 1]); mB1 = make_float2(g.dyfc[q2 + 1], g.dyfc[q2 + 2]);                      // u at j
#pragma unroll
    for (int n = 0; n < 4; n++) azp[n] = make_float2(g.azcc[q2 + (n - 2) * PX], g.azcc[q2 + (n - 2) * PX + 1]);
  }
  kbA = g.kb[q2]; kbB = g.kb[q2 + 1];
  const int kgen = g.kgen2[q2], kzero = g.kzero2[q2];   // pair-based (identical for both cells)
  // AB2 epilogue: highest solid level next to each of the two velocity nodes (the node is peripheral for k <= km)
  int kmA = 0, kmB = 0;
  float2 gs = make_float2(0.f, 0.f), us = make_float2(0.f, 0.f);
  if (AB2) {
    if (DIR == 0) { kmA = max(kbA, (int)g.kb[q2 - 1]); kmB = max(kbB, kbA); }
    else {
      const bool ywall = y_outside(g, j) || y_outside(g, j - 1);
      kmA = ywall ? GB25_BIG : max(kbA, (int)g.kb[q2 - PX]);
      kmB = ywall ? GB25_BIG : max(kbB, (int)g.kb[q2 - PX + 1]);
    }
  }
  // ---- vertical register windows of the own velocity
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  float WA[7], WB[7];
#pragma unroll
  for (int m = 0; m < 7; m++) {
    const float2 a = __ldg(reinterpret_cast<const float2*>(own_g + q3 + (ptrdiff_t)(m - 3) * n2));
    WA[m] = a.x; WB[m] = a.y;
  }
  float2 Wb = make_float2(0.f, 0.f);
  int s = 0; uint32_t fph = 0;     // stage and full-barrier parity of the level being consumed
  for (int k = 1; k <= Nz; k++, q3 += n2) {
    mbar_wait(&bar[s], fph);
    const float* sm = smem + s * STAGE;
    float2 out = make_float2(0.f, 0.f);
    bool store = false;
    if (valid) {
      const float dz = g.dzc[k + g.Hz - 1];
      if (k <= kzero) { Wb = make_float2(0.f, 0.f); store = true; }
      else if (k > kgen) {
        store = true;
        if (k == kgen + 1 && kgen > 0) Wb = make_float2(carry[q2], carry[q2 + 1]);
        const float2 own0 = make_float2(WA[3], WB[3]);
        const bool lA = own0.x > 0.f, lB = own0.y > 0.f;
        const int Bw = (g.immersed && k + 1 > Nz) ? 1 : 2;
        float2 Hterm, Phi, Bsum, wt, oavg, dpp;
        if (DIR == 0) {
          const float* U = sm + GU_OFF_U + (ty + 4) * GU_PU + (lx + 4); const float* V = sm + GU_OFF_V + (ty + 4) * GU_PV + (lx + 4);
          const float* Z = sm + GU_OFF_Z + (ty + 4) * GU_PZ + lx; const float* DX = sm + GU_OFF_DX + ty * GU_PD + (lx + 4);
          const float* DY = sm + GU_OFF_DY + ty * GU_PD + (lx + 4); const float* W = sm + GU_OFF_W + ty * GU_PD + (lx + 4);
          const float* P = sm + GU_OFF_P + ty * GU_PP + (lx + 4);
          // rows j-3 .. j+3 of u at the pair, rows j-2 .. j+3 of v at (i-1, i) and (i, i+1), of zeta at the pair
          float2 uy[7], vw[6], ve[6], zq[6], zs[6], zr[6];
#pragma unroll
          for (int m = 0; m < 7; m++) uy[m] = ldpa(U + (m - 3) * GU_PU);
#pragma unroll
          for (int m = 0; m < 6; m++) {
            vw[m] = ldp(V + (m - 2) * GU_PV - 1); ve[m] = ldpa(V + (m - 2) * GU_PV);
            zq[m] = ldpa(Z + (m - 2) * GU_PZ);
            zs[m] = pmuls(padd(uy[m], uy[m + 1]), 0.5f);
            zr[m] = pmuls(padd(vw[m], ve[m]), 0.5f);
          }
          const float2 xm0 = pmul(mA0, vw[2]), xm1 = pmul(mA1, vw[3]), x00 = pmul(mB0, ve[2]), x01 = pmul(mB1, ve[3]);
          oavg = pmuls(padd(pmuls(padd(xm0, xm1), 0.5f), pmuls(padd(x00, x01), 0.5f)), 0.5f);
          const float2 ohat = pmul(oavg, rm1);
          Hterm = pneg(pmul(ohat, pweno5_vs_sel(zq, zs, zr, ohat.x > 0.f, ohat.y > 0.f, eps)));
          // x windows: cells i-3 .. i+3
          float2 ux[7], dOw[6], dv[6], dK[6], sK[6], hs[7], dyx[6];
#pragma unroll
          for (int n = 0; n < 7; n++) { ux[n] = (n & 1) ? ldpa(U + n - 3 + 0) : ldp(U + n - 3); }
#pragma unroll
          for (int n = 0; n < 7; n++) hs[n] = pmuls(pmul(ux[n], ux[n]), 0.5f);
#pragma unroll
          for (int m = 0; m < 6; m++) {
            dOw[m] = (m & 1) ? ldpa(DX + m - 3) : ldp(DX + m - 3);
            dyx[m] = (m & 1) ? ldpa(DY + m - 3) : ldp(DY + m - 3);
            dv[m] = padd(dOw[m], dyx[m]);
            dK[m] = psub(hs[m + 1], hs[m]);
            sK[m] = pmuls(padd(ux[m], ux[m + 1]), 0.5f);
          }
          const float2 dvs = psym4(dyx[1], dyx[2], dyx[3], dyx[4], 2);
          Phi = pmul(own0, padd(dvs, pweno5_fs_sel(dOw, dv, lA, lB, eps)));
          const float2 dKo = pweno5_fs_sel(dK, sK, lA, lB, eps);
          float2 kc[4];
#pragma unroll
          for (int m = 0; m < 4; m++) kc[m] = psub(pmuls(pmul(ve[m + 1], ve[m + 1]), 0.5f), pmuls(pmul(vw[m + 1], vw[m + 1]), 0.5f));
          Bsum = padd(dKo, psym4(kc[0], kc[1], kc[2], kc[3], 2));
          wt = psym4(pmul(azp[0], ldpa(W - 2)), pmul(azp[1], ldp(W - 1)), pmul(azp[2], ldpa(W)), pmul(azp[3], ldp(W + 1)), Bw);
          dpp = psub(ldpa(P), ldp(P - 1));
        } else {
          const float* V = sm + GV_OFF_V + (ty + 4) * GV_PV + (lx + 4); const float* U = sm + GV_OFF_U + (ty + 4) * GV_PU + (lx + 4);
          const float* Z = sm + GV_OFF_Z + ty * GV_PZ + (lx + 4); const float* DY = sm + GV_OFF_DY + (ty + 4) * GV_PD + lx;
          const float* DX = sm + GV_OFF_DX + (ty + 4) * GV_PD + lx; const float* W = sm + GVP_OFF_W + (ty + 2) * GV_PD + lx;
          const float* P = sm + GVP_OFF_P + (ty + 1) * GV_PD + lx;
          // x windows (cols i-2 .. i+3): zeta, v row j (cols i-3 .. i+3), u rows j-1 and j
          float2 vx[7], us[6], un[6], zq[6], zs[6], zr[6];
#pragma unroll
          for (int n = 0; n < 7; n++) vx[n] = (n & 1) ? ldpa(V + n - 3) : ldp(V + n - 3);
#pragma unroll
          for (int m = 0; m < 6; m++) {
            us[m] = (m & 1) ? ldp(U - GV_PU + m - 2) : ldpa(U - GV_PU + m - 2);
            un[m] = (m & 1) ? ldp(U + m - 2) : ldpa(U + m - 2);
            zq[m] = (m & 1) ? ldp(Z + m - 2) : ldpa(Z + m - 2);
            zs[m] = pmuls(padd(vx[m], vx[m + 1]), 0.5f);
            zr[m] = pmuls(padd(us[m], un[m]), 0.5f);
          }
          const float2 xm0 = pmul(mA0, us[2]), xm1 = pmul(mA1, us[3]), x00 = pmul(mB0, un[2]), x01 = pmul(mB1, un[3]);
          oavg = pmuls(padd(pmuls(padd(xm0, xm1), 0.5f), pmuls(padd(x00, x01), 0.5f)), 0.5f);
          const float2 ohat = pmul(oavg, rm1);
          Hterm = pmul(ohat, pweno5_vs_sel(zq, zs, zr, ohat.x > 0.f, ohat.y > 0.f, eps));
          // y windows: rows j-3 .. j+3
          float2 vy[7], dOw[6], dv[6], dK[6], sK[6], hs[7], dxy[6];
#pragma unroll
          for (int n = 0; n < 7; n++) { vy[n] = ldpa(V + (n - 3) * GV_PV); hs[n] = pmuls(pmul(vy[n], vy[n]), 0.5f); }
#pragma unroll
          for (int m = 0; m < 6; m++) {
            dOw[m] = ldpa(DY + (m - 3) * GV_PD);
            dxy[m] = ldpa(DX + (m - 3) * GV_PD);
            dv[m] = padd(dxy[m], dOw[m]);
            dK[m] = psub(hs[m + 1], hs[m]);
            sK[m] = pmuls(padd(vy[m], vy[m + 1]), 0.5f);
          }
          const float2 dus = psym4(dxy[1], dxy[2], dxy[3], dxy[4], 2);
          Phi = pmul(own0, padd(dus, pweno5_fs_sel(dOw, dv, lA, lB, eps)));
          const float2 dKo = pweno5_fs_sel(dK, sK, lA, lB, eps);
          float2 kc[4];
#pragma unroll
          for (int m = 0; m < 4; m++) kc[m] = psub(pmuls(pmul(un[m + 1], un[m + 1]), 0.5f), pmuls(pmul(us[m + 1], us[m + 1]), 0.5f));
          Bsum = padd(dKo, psym4(kc[0], kc[1], kc[2], kc[3], 2));
          wt = psym4(pmul(azp[0], ldpa(W - 2 * GV_PD)), pmul(azp[1], ldpa(W - GV_PD)), pmul(azp[2], ldpa(W)), pmul(azp[3], ldpa(W + GV_PD)), Bw);
          dpp = psub(ldpa(P), ldpa(P - GV_PD));
        }
        const float2 Wt = pmul(wt, pvert(WA, WB, zbuf(g, kbA, k + 1, 3), zbuf(g, kbB, k + 1, 3), wt, eps));
        const float2 Vterm = pmul(pmuls(rAz, rcp_refined(dz)), padd(Phi, psub(Wt, Wb)));   // (1/dz without the range-check branch)
        Wb = Wt;
        const float2 ct = pmul(pmul(fbar, oavg), rm1);          // Gu: -(-ct) ; Gv: -(+ct)
        const float2 sum = padd(padd(Hterm, Vterm), pmul(Bsum, rm1));
        const float2 rest = padd(DIR == 0 ? pneg(ct) : ct, pmul(dpp, rm1));
        out = pneg(padd(sum, rest));
      }
      if (store) *reinterpret_cast<float2*>(G + q3) = out;
      if (AB2) {
        const float2 gn = store ? out : *reinterpret_cast<const float2*>(G + q3);      // generic cell: left by k_generic_list
        const float2 gmv = ldpa(sm + OFF_GM + ty * TMA_TX + lx);
        const float gA = ab2_g(ab.c1, ab.c2, gn.x, gmv.x), gB = ab2_g(ab.c1, ab.c2, gn.y, gmv.y);
        const bool pA = k <= kmA, pB = k <= kmB;
        const float tA = __fmul_rn(dz, pA ? 0.f : gA), tB = __fmul_rn(dz, pB ? 0.f : gB);
        gs.x = (k == 1) ? tA : __fadd_rn(gs.x, tA); gs.y = (k == 1) ? tB : __fadd_rn(gs.y, tB);
        float uA = ab2_upd(WA[3], ab.dt, gA), uB = ab2_upd(WB[3], ab.dt, gB);
        if (g.immersed) { if (pA) uA = 0.f; if (pB) uB = 0.f; }
        *reinterpret_cast<float2*>(ab.own_next + q3) = make_float2(uA, uB);
        const float wA = __fmul_rn(dz, uA), wB = __fmul_rn(dz, uB);
        us.x = (k == 1) ? wA : __fadd_rn(us.x, wA); us.y = (k == 1) ? wB : __fadd_rn(us.y, wB);
      }
      const float2 a = __ldg(reinterpret_cast<const float2*>(own_g + q3 + (size_t)4 * n2));
#pragma unroll
      for (int m = 0; m < 6; m++) { WA[m] = WA[m + 1]; WB[m] = WB[m + 1]; }
      WA[6] = a.x; WB[6] = a.y;
    }
    __syncwarp();
#if MOM_OPP
    if ((tid & 31) == 0 && k + MOM_NST <= Nz && ring_release(&rcnt[s], 4)) issue(k + MOM_NST, s);
#else
    if ((tid & 31) == 0) mbar_arrive(&ebar[s]);
#endif
    if (++s == MOM_NST) { s = 0; fph ^= 1u; }
  }
  if (AB2 && valid) {
    *reinterpret_cast<float2*>(ab.gsum + q2) = gs;
    *reinterpret_cast<float2*>(ab.usum + q2) = us;
  }
}

// =====================================================================================
// Tracer tendencies (row A6), TMA-staged.  CTA = 32 x 16 columns; 128 consumer threads, each owning a 2 x 2 patch
// of columns so that the x- and y-face fluxes inside the patch are computed once (16 reconstructions for 4 cells:
// 4.0 per cell instead of 4.5 in the register-blocked kernel and 6 in the per-cell one); one tracer per blockIdx.z.
// Same ring / barrier scheme as the momentum kernels; generic cells belong to k_generic_list.
// =====================================================================================
#define TR_TX 32
#define TR_TY 16
#define TR_PT (TR_TX + 8)
#define TR_PU (TR_TX + 4)
#define TR_PV (TR_TX)
#define TR_OFF_T 0
#define TR_OFF_U (TR_OFF_T + TR_PT * (TR_TY + 8))
#define TR_OFF_V (TR_OFF_U + TR_PU * TR_TY)
#define TR_OFF_W (TR_OFF_V + TR_PV * (TR_TY + 1))
#define TR_STAGE (TR_OFF_W + TR_PV * TR_TY)
#define TR_GM_TILE (TR_TX * TR_TY)
#define TR_CW 4   // consumer warps

struct TmaMaps5 { CUtensorMap m[5]; };   // T, S, u, v, w
// AB2 epilogue of the tracer kernel (template flag AB2): c' = mask(c + dt (c1 Gn - c2 G-)) into the other state buffer
// (rows A9, A1: the arithmetic of k_ab2_ts_3d, bit for bit)
struct TrAb2 { float* next[2]; float dt, c1, c2; int zhalo; };
struct TmaMaps2 { CUtensorMap m[2]; };   // G- of T and of S

__device__ __forceinline__ float weno5_selp(const float* q, bool left, float eps) {   // q[0..5], face between q[2], q[3]
  const float v0 = left ? q[0] : q[5], v1 = left ? q[1] : q[4], v2 = left ? q[2] : q[3], v3 = left ? q[3] : q[2], v4 = left ? q[4] : q[1];
  return weno5(v0, v1, v2, v3, v4, eps);
}

// ring depth (ms per launch): 4 stages 1.08, 5 1.00, 6 1.00
#ifndef TR_NST
#define TR_NST (TR_OPP ? 4 : 6)
#endif
#ifndef TR_MINB
#define TR_MINB (TR_OPP ? 4 : 3)
#endif
template <bool AB2>
__global__ void __launch_bounds__(TR_THREADS, TR_MINB)
k_tracer_tma(DevGrid g, const __grid_constant__ TmaMaps5 tm, const __grid_constant__ TmaMaps2 tmg, const float* __restrict__ T0, const float* __restrict__ T1,
             float* __restrict__ G0, float* __restrict__ G1, const float* __restrict__ carry0, const float* __restrict__ carry1,
             const TrAb2 ab) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bar[TR_NST], ebar[TR_NST];
  __shared__ int rcnt[TR_NST];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 16 + tx;
  const int i0 = blockIdx.x * TR_TX + 1, j0 = blockIdx.y * TR_TY + 1;
  const int I0 = i0 + g.Hx - 1, J0 = j0 + g.Hy - 1;
  const int tr = blockIdx.z;
  const int PX = g.PX, n2 = g.n2, Nz = g.Nz;
  const float eps = g.eps;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TR_NST; s++) { mbar_init(&bar[s], 1); mbar_init(&ebar[s], TR_CW); rcnt[s] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr int STAGE = TR_STAGE + (AB2 ? TR_GM_TILE : 0);
  auto issue = [&](int k, int s) {
        float* sm = smem + s * STAGE;
        const int K = k + g.Hz - 1;
        mbar_expect_tx(&bar[s], STAGE * sizeof(float));
        if (AB2) tma_load_3d(sm + TR_STAGE, &tmg.m[tr], &bar[s], I0, J0, K);
        tma_load_3d(sm + TR_OFF_T, &tm.m[tr], &bar[s], I0 - 4, J0 - 4, K);
        tma_load_3d(sm + TR_OFF_U, &tm.m[2], &bar[s], I0, J0, K);
        tma_load_3d(sm + TR_OFF_V, &tm.m[3], &bar[s], I0, J0, K);
        tma_load_3d(sm + TR_OFF_W, &tm.m[4], &bar[s], I0, J0, K + 1);
  };
#if TR_OPP
  if (tid == 0)
    for (int k = 1; k <= min(Nz, TR_NST); k++) issue(k, k - 1);
#else
  if (ty >= 8) {   // ===== producer warp
    if (tid == 128)
      for (int k = 1; k <= Nz; k++) {
        const int s = (k - 1) % TR_NST;
        if (k > TR_NST) mbar_wait(&ebar[s], (((k - 1) / TR_NST) - 1) & 1);
        issue(k, s);
      }
    return;
  }
#endif
  const float* __restrict__ T = tr == 0 ? T0 : T1;
  float* __restrict__ G = tr == 0 ? G0 : G1;
  const float* __restrict__ carry = tr == 0 ? carry0 : carry1;
  float* __restrict__ Tn = tr == 0 ? ab.next[0] : ab.next[1];
  const int lx = 2 * tx, ly = 2 * ty;                       // tile-local column / row of the patch
  // clamped patch origin (for the hoisted loads).  A patch may start on the last row when Ny is odd: its second row is then
  // the halo row Ny + 1, which exists in memory and is never stored (vrow); clamping the origin to Ny - 1 instead would make
  // the patch read rows (Ny, Ny + 1) from shared memory but address rows (Ny - 1, Ny) in global memory.
  const int ic = min(i0 + lx, g.Nx - 1), jc = min(j0 + ly, g.Ny);
  const bool vx = (i0 + lx + 1) <= g.Nx;
  bool vrow[2] = {vx && (j0 + ly) <= g.Ny, vx && (j0 + ly + 1) <= g.Ny};
  const int q2 = id2(g, ic, jc);
  // ---- hoisted 2-D data
  float dyf[2][3], dxf[3][2], az[2][2];
  int kbc[2][2], kg[2], kz[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
#pragma unroll
    for (int e = 0; e < 3; e++) dyf[r][e] = g.dyfc[q2 + r * PX + e];
#pragma unroll
    for (int c = 0; c < 2; c++) { az[r][c] = g.azcc[q2 + r * PX + c]; kbc[r][c] = g.kb[q2 + r * PX + c]; }
    kg[r] = g.kgen2[q2 + r * PX]; kz[r] = g.kzero2[q2 + r * PX];
  }
#pragma unroll
  for (int e = 0; e < 3; e++) { dxf[e][0] = g.dxcf[q2 + e * PX]; dxf[e][1] = g.dxcf[q2 + e * PX + 1]; }
  // ---- vertical register windows W[r][c][m] = T(i, j, k-3+m)
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  float W[2][2][7];
#pragma unroll
  for (int m = 0; m < 7; m++)
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const float2 a = __ldg(reinterpret_cast<const float2*>(T + q3 + (ptrdiff_t)(m - 3) * n2 + r * PX));
      W[r][0][m] = a.x; W[r][1][m] = a.y;
    }
  float Fz[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  const int oT = (ly + 4) * TR_PT + (lx + 4), oU = ly * TR_PU + lx, oV = ly * TR_PV + lx;
  int s = 0; uint32_t fph = 0;     // stage and full-barrier parity of the level being consumed
  for (int k = 1; k <= Nz; k++, q3 += n2) {
    mbar_wait(&bar[s], fph);
    const float* sm = smem + s * STAGE;
    const bool fast0 = k > kg[0], fast1 = k > kg[1];
    float out[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    if (fast0 || fast1) {
#pragma unroll
      for (int r = 0; r < 2; r++)
        if (k == kg[r] + 1 && kg[r] > 0) { Fz[r][0] = carry[q2 + r * PX]; Fz[r][1] = carry[q2 + r * PX + 1]; }
      const float dz = g.dzc[k + g.Hz - 1];
      const float* St = sm + TR_OFF_T + oT;
      // All reconstructions are evaluated two at a time with packed FP32x2 arithmetic (gb25_packed.cuh):
      // x faces pair the two rows of the patch, y and z faces pair its two columns.
      // ---- x faces: 3 per row.  rT[n] = (row 0, row 1) at local column lx-4+n
      float2 fx[3];
      {
        float2 rT[10];
#pragma unroll
        for (int n = 0; n < 10; n++) rT[n] = make_float2(St[n - 4], St[TR_PT + n - 4]);
#pragma unroll
        for (int e = 0; e < 3; e++) {
          const float2 uu = make_float2(sm[TR_OFF_U + oU + e], sm[TR_OFF_U + oU + TR_PU + e]);
          const float2 av = pmul(pmuls(make_float2(dyf[0][e], dyf[1][e]), dz), uu);
          fx[e] = pmul(av, pweno5_sel(&rT[e + 1], uu.x > 0.f, uu.y > 0.f, eps));
        }
      }
      // ---- y faces: 3 per column.  cT[m] = (col 0, col 1) at local row ly-4+m
      float2 fy[3];
      {
        float2 cT[10];
#pragma unroll
        for (int m = 0; m < 10; m++) cT[m] = *reinterpret_cast<const float2*>(St + (m - 4) * TR_PT);
#pragma unroll
        for (int e = 0; e < 3; e++) {
          const float2 vv = *reinterpret_cast<const float2*>(sm + TR_OFF_V + oV + e * TR_PV);
          const float2 av = pmul(pmuls(make_float2(dxf[e][0], dxf[e][1]), dz), vv);
          fy[e] = pmul(av, pweno5_sel(&cT[e + 1], vv.x > 0.f, vv.y > 0.f, eps));
        }
      }
      // ---- z: top faces from the register windows (pairs of columns); bottom fluxes are carried
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const float2 ww = *reinterpret_cast<const float2*>(sm + TR_OFF_W + oV + r * TR_PV);
        const int B0 = zbuf(g, kbc[r][0], k + 1, 3), B1 = zbuf(g, kbc[r][1], k + 1, 3);
        float2 rec;
        if (B0 == 3 && B1 == 3) {
          float2 q[6];
#pragma unroll
          for (int m = 0; m < 6; m++) q[m] = make_float2(W[r][0][m + 1], W[r][1][m + 1]);
          rec = pweno5_sel(q, ww.x > 0.f, ww.y > 0.f, eps);
        } else {
          rec.x = weno_sel_B3(W[r][0][1], W[r][0][2], W[r][0][3], W[r][0][4], W[r][0][5], W[r][0][6], B0, ww.x > 0.f, eps);
          rec.y = weno_sel_B3(W[r][1][1], W[r][1][2], W[r][1][3], W[r][1][4], W[r][1][5], W[r][1][6], B1, ww.y > 0.f, eps);
        }
        const float2 azp = make_float2(az[r][0], az[r][1]);
        const float2 ft = pmul(pmul(azp, ww), rec);
        const float2 rV = prcp_nr(pmuls(azp, dz));
        const float fxr0 = r == 0 ? fx[0].x : fx[0].y, fxr1 = r == 0 ? fx[1].x : fx[1].y, fxr2 = r == 0 ? fx[2].x : fx[2].y;
        const float2 dfx = make_float2(fxr1 - fxr0, fxr2 - fxr1);
        const float2 dfy = psub(fy[r + 1], fy[r]);
        const float2 fzo = make_float2(Fz[r][0], Fz[r][1]);
        const float2 o = pmul(rV, padd(padd(dfx, dfy), psub(ft, fzo)));
        out[r][0] = -o.x; out[r][1] = -o.y;
        Fz[r][0] = ft.x; Fz[r][1] = ft.y;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const bool fast = r == 0 ? fast0 : fast1;
      if (k <= kz[r]) { Fz[r][0] = 0.f; Fz[r][1] = 0.f; out[r][0] = 0.f; out[r][1] = 0.f; }
      const bool st = fast || k <= kz[r];
      if (vrow[r] && st) *reinterpret_cast<float2*>(G + q3 + r * PX) = make_float2(out[r][0], out[r][1]);
      if (AB2 && vrow[r]) {
        const float2 gn = st ? make_float2(out[r][0], out[r][1]) : *reinterpret_cast<const float2*>(G + q3 + r * PX);   // generic: k_generic_list
        const float2 gmv = ldpa(sm + TR_STAGE + (ly + r) * TR_TX + lx);
        float c0 = ab2_upd(W[r][0][3], ab.dt, ab2_g(ab.c1, ab.c2, gn.x, gmv.x)), c1n = ab2_upd(W[r][1][3], ab.dt, ab2_g(ab.c1, ab.c2, gn.y, gmv.y));
        if (g.immersed) { if (k <= kbc[r][0]) c0 = 0.f; if (k <= kbc[r][1]) c1n = 0.f; }
        *reinterpret_cast<float2*>(Tn + q3 + r * PX) = make_float2(c0, c1n);
        if (ab.zhalo) {   // no-flux z halos of the updated tracer (the mirror of k_halo_bottom_top)
          if (k <= g.Hz) *reinterpret_cast<float2*>(Tn + q3 + r * PX - (size_t)(2 * k - 1) * n2) = make_float2(c0, c1n);
          if (k > Nz - g.Hz) *reinterpret_cast<float2*>(Tn + q3 + r * PX + (size_t)(2 * (Nz - k) + 1) * n2) = make_float2(c0, c1n);
        }
      }
      const float2 a = __ldg(reinterpret_cast<const float2*>(T + q3 + (size_t)4 * n2 + r * PX));
#pragma unroll
      for (int m = 0; m < 6; m++) { W[r][0][m] = W[r][0][m + 1]; W[r][1][m] = W[r][1][m + 1]; }
      W[r][0][6] = a.x; W[r][1][6] = a.y;
    }
    __syncwarp();
#if TR_OPP
    if ((tid & 31) == 0 && k + TR_NST <= Nz && ring_release(&rcnt[s], TR_CW)) issue(k + TR_NST, s);
#else
    if ((tid & 31) == 0) mbar_arrive(&ebar[s]);
#endif
    if (++s == TR_NST) { s = 0; fph ^= 1u; }
  }
}

// ----------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encoder() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}
static bool make_map(const DevGrid& g, const float* base, int bx, int by, CUtensorMap* out) {
  PFN_encodeTiled enc = get_encoder();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)g.PX, (cuuint64_t)g.PY, (cuuint64_t)g.PZ};
  cuuint64_t strides[2] = {(cuuint64_t)g.PX * 4, (cuuint64_t)g.PX * g.PY * 4};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [parity of the state buffers]; gm: G- tiles of u, v, T, S for the AB2 epilogues ([parity of the Gn / G- pointer swap])
struct TmaState { bool ready = false, ok = false; TmaMaps7 gu[2], gv[2]; TmaMaps5 tr[2]; CUtensorMap gm[2][4]; const float* gm_base[2][4]; };
static TmaState* tma_state(Handle* h) {
  if (!h->tma) h->tma = new TmaState();
  TmaState* t = (TmaState*)h->tma;
  if (t->ready) return t;
  t->ready = true;
  const DevGrid& g = h->g;
  if (g.PX % 4) return t;   // TMA needs 16-byte row pitch
  const int TX = TMA_TX, TY = TMA_TY;
  bool ok = true;
  for (int par = 0; par < 2; par++) {
    const float *u = h->state_buf[par][0], *v = h->state_buf[par][1], *T = h->state_buf[par][2], *S = h->state_buf[par][3];
    ok &= make_map(g, u, TX + 8, TY + 8, &t->gu[par].m[0]);
    ok &= make_map(g, v, TX + 4, TY + 8, &t->gu[par].m[1]);
    ok &= make_map(g, h->zeta, TX, TY + 8, &t->gu[par].m[2]);
    ok &= make_map(g, h->dxU, TX + 8, TY, &t->gu[par].m[3]);
    ok &= make_map(g, h->dyV, TX + 8, TY, &t->gu[par].m[4]);
    ok &= make_map(g, h->f.w, TX + 8, TY, &t->gu[par].m[5]);
    ok &= make_map(g, h->f.p, TX + 4, TY, &t->gu[par].m[6]);
    ok &= make_map(g, v, TX + 8, TY + 8, &t->gv[par].m[0]);
    ok &= make_map(g, u, TX + 8, TY + 8, &t->gv[par].m[1]);
    ok &= make_map(g, h->zeta, TX + 8, TY, &t->gv[par].m[2]);
    ok &= make_map(g, h->dyV, TX, TY + 8, &t->gv[par].m[3]);
    ok &= make_map(g, h->dxU, TX, TY + 8, &t->gv[par].m[4]);
    ok &= make_map(g, h->f.w, TX, TY + 3, &t->gv[par].m[5]);
    ok &= make_map(g, h->f.p, TX, TY + 1, &t->gv[par].m[6]);
    ok &= make_map(g, T, TR_TX + 8, TR_TY + 8, &t->tr[par].m[0]);
    ok &= make_map(g, S, TR_TX + 8, TR_TY + 8, &t->tr[par].m[1]);
    ok &= make_map(g, u, TR_TX + 4, TR_TY, &t->tr[par].m[2]);
    ok &= make_map(g, v, TR_TX, TR_TY + 1, &t->tr[par].m[3]);
    ok &= make_map(g, h->f.w, TR_TX, TR_TY, &t->tr[par].m[4]);
  }
  // the Gn / G- arrays swap roles every step: one G- map per array, looked up by base pointer at launch
  for (int q = 0; q < 4; q++) {
    t->gm_base[0][q] = h->f.gm[q]; t->gm_base[1][q] = h->f.gn[q];
    for (int r = 0; r < 2; r++) ok &= make_map(g, t->gm_base[r][q], q < 2 ? TX : TR_TX, q < 2 ? TY : TR_TY, &t->gm[r][q]);
  }
  if (ok) {
    const int f4 = (int)sizeof(float);
    ok &= cudaFuncSetAttribute(k_tracer_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_NST * TR_STAGE * f4) == cudaSuccess;
    ok &= cudaFuncSetAttribute(k_tracer_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_NST * (TR_STAGE + TR_GM_TILE) * f4) == cudaSuccess;
    ok &= cudaFuncSetAttribute(k_mom_tma_p2<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MOM_NST * GU_STAGE * f4) == cudaSuccess;
    ok &= cudaFuncSetAttribute(k_mom_tma_p2<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MOM_NST * (GU_STAGE + MOM_GM_TILE) * f4) == cudaSuccess;
    ok &= cudaFuncSetAttribute(k_mom_tma_p2<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MOM_NST * GVP_STAGE * f4) == cudaSuccess;
    ok &= cudaFuncSetAttribute(k_mom_tma_p2<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MOM_NST * (GVP_STAGE + MOM_GM_TILE) * f4) == cudaSuccess;
    ok &= cudaFuncSetAttribute(k_gu_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_NST * GU_STAGE * (int)sizeof(float)) == cudaSuccess;
    ok &= cudaFuncSetAttribute(k_gv_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_NST * GV_STAGE * (int)sizeof(float)) == cudaSuccess;
  }
  t->ok = ok;
  return t;
}
bool tma_available(Handle* h) { return tma_state(h)->ok; }
void tma_free(Handle* h) { delete (TmaState*)h->tma; h->tma = nullptr; }

static const CUtensorMap& gm_map(TmaState* t, const float* base, int q) { return t->gm[t->gm_base[0][q] == base ? 0 : 1][q]; }
void launch_momentum_tendency_tma(Handle* h, const Ab2Spec* spec) {
  TmaState* t = tma_state(h);
  const DevGrid& g = h->g;
  const int par = h->parity;
  if (h->use_packed) {
    dim3 b(16, MOM_THREADS / 16), gr((g.Nx + TMA_TX - 1) / TMA_TX, (g.Ny + TMA_TY - 1) / TMA_TY);
    const int gmt = spec ? MOM_GM_TILE : 0;
    const size_t smu = MOM_NST * (GU_STAGE + gmt) * sizeof(float), smv = MOM_NST * (GVP_STAGE + gmt) * sizeof(float);
    MomAb2 au = {}, av = {};
    if (spec) {
      au = MomAb2{h->state_buf[1 - par][0], h->spec2d[0], h->spec2d[2], spec->dt, spec->c1, spec->c2};
      av = MomAb2{h->state_buf[1 - par][1], h->spec2d[1], h->spec2d[3], spec->dt, spec->c1, spec->c2};
    }
    { StageScope ts(h, "kernel:k_gu_tma");
      if (spec) k_mom_tma_p2<0, true><<<gr, b, smu, h->stream>>>(g, t->gu[par], gm_map(t, h->f.gm[0], 0), h->f.u, h->f.gn[0], h->carry[0], au);
      else k_mom_tma_p2<0, false><<<gr, b, smu, h->stream>>>(g, t->gu[par], gm_map(t, h->f.gm[0], 0), h->f.u, h->f.gn[0], h->carry[0], au); }
    h->count_launch();
    { StageScope ts(h, "kernel:k_gv_tma");
      if (spec) k_mom_tma_p2<1, true><<<gr, b, smv, h->stream>>>(g, t->gv[par], gm_map(t, h->f.gm[1], 1), h->f.v, h->f.gn[1], h->carry[1], av);
      else k_mom_tma_p2<1, false><<<gr, b, smv, h->stream>>>(g, t->gv[par], gm_map(t, h->f.gm[1], 1), h->f.v, h->f.gn[1], h->carry[1], av); }
    h->count_launch();
    return;
  }
  // scalar TMA kernels (GB25_PACKED=0): their Gv stage keeps full-height W and P boxes
  TmaMaps7 gvs = t->gv[par];
  make_map(g, h->f.w, TMA_TX, TMA_TY + 8, &gvs.m[5]);
  make_map(g, h->f.p, TMA_TX, TMA_TY + 8, &gvs.m[6]);
  dim3 b(TMA_TX, TMA_TY + 1), gr((g.Nx + TMA_TX - 1) / TMA_TX, (g.Ny + TMA_TY - 1) / TMA_TY);
  { StageScope ts(h, "kernel:k_gu_tma");
  k_gu_tma<<<gr, b, TMA_NST * GU_STAGE * sizeof(float), h->stream>>>(g, h->g_dev, t->gu[par], h->f.u, h->f.v, h->f.w, h->f.p, h->f.gn[0], h->carry[0]); }
  h->count_launch();
  { StageScope ts(h, "kernel:k_gv_tma");
  k_gv_tma<<<gr, b, TMA_NST * GV_STAGE * sizeof(float), h->stream>>>(g, h->g_dev, gvs, h->f.u, h->f.v, h->f.w, h->f.p, h->f.gn[1], h->carry[1]); }
  h->count_launch();
}

void launch_tracer_tendency_tma(Handle* h, const Ab2Spec* spec) {
  TmaState* t = tma_state(h);
  const DevGrid& g = h->g;
  const int par = h->parity;
  dim3 b(16, TR_THREADS / 16), gr((g.Nx + TR_TX - 1) / TR_TX, (g.Ny + TR_TY - 1) / TR_TY, 2);
  const size_t sm = TR_NST * (TR_STAGE + (spec ? TR_GM_TILE : 0)) * sizeof(float);
  TrAb2 ab = {};
  TmaMaps2 tmg;
  tmg.m[0] = gm_map(t, h->f.gm[2], 2); tmg.m[1] = gm_map(t, h->f.gm[3], 3);
  if (spec) ab = TrAb2{{h->state_buf[1 - par][2], h->state_buf[1 - par][3]}, spec->dt, spec->c1, spec->c2, spec->zhalo};
  StageScope ts(h, "kernel:k_tracer_tma");
  if (spec) k_tracer_tma<true><<<gr, b, sm, h->stream>>>(g, t->tr[par], tmg, h->f.T, h->f.S, h->f.gn[2], h->f.gn[3], h->carry[2], h->carry[3], ab);
  else k_tracer_tma<false><<<gr, b, sm, h->stream>>>(g, t->tr[par], tmg, h->f.T, h->f.S, h->f.gn[2], h->f.gn[3], h->carry[2], h->carry[3], ab);
  h->count_launch();
}

// ---------------------------------------------------------------- kernel table (preload_kernels, gb25_api.cu)
KernelTable kernel_table_tend_tma() {
  static const void* const k[] = {
    (const void*)k_gu_tma, (const void*)k_gv_tma, (const void*)k_mom_tma_p2<0, false>, (const void*)k_mom_tma_p2<0, true>,
    (const void*)k_mom_tma_p2<1, false>, (const void*)k_mom_tma_p2<1, true>, (const void*)k_tracer_tma<false>, (const void*)k_tracer_tma<true>,
  };
  return {k, (int)(sizeof k / sizeof k[0])};
}
