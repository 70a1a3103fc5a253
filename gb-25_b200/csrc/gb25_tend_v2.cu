// gb25_tend_v2.cu — second-generation tendency kernels: x-blocked, k-marching.
//
// The first-generation kernels (thread per cell) are FP32-issue bound, and more than half of the issued
// instructions are not WENO arithmetic but scalar loads with 64-bit address arithmetic (ncu, profiles/).
// Here one thread owns FOUR consecutive cells in x (i0..i0+3, 16-byte aligned since Hx = 8 and Nx % 4 == 0)
// and marches k = 1..Nz:
//   * every stencil row is fetched with 128-bit loads (one LDG.128 per row per field for 4 cells),
//   * the vertical stencil lives in a 7-deep register window that shifts as the thread marches,
//   * the vertical face flux is computed once per face and carried to the next level,
//   * 2-D metrics are hoisted out of the k loop,
//   * upwinding is branch-free (select the mirrored window, then one WENO evaluation),
//   * cells whose stencil is not clear of bathymetry / walls (k-1 <= knear) fall back to the generic
//     per-cell function, so the fast path carries no masks and no order reduction in x and y.
#include <cstdlib>

#include "gb25_internal.h"
#include "gb25_tend_generic.cuh"

#ifndef GB25_TRACER_NC
#define GB25_TRACER_NC 2
#endif

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float comp(const float4& a, int c) { return c == 0 ? a.x : (c == 1 ? a.y : (c == 2 ? a.z : a.w)); }

// branch-free upwind-biased WENO5 at the face between q2 and q3 of (q0..q5)
__device__ __forceinline__ float weno5_sel(float q0, float q1, float q2, float q3, float q4, float q5, bool left, float eps) {
  const float v0 = left ? q0 : q5, v1 = left ? q1 : q4, v2 = left ? q2 : q3, v3 = left ? q3 : q2, v4 = left ? q4 : q1;
  return weno5(v0, v1, v2, v3, v4, eps);
}
// variable-order version for the vertical direction (B is warp-uniform there)
__device__ __forceinline__ float weno_sel_B(float q0, float q1, float q2, float q3, float q4, float q5, int B, bool left, float eps) {
  if (B == 3) return weno5_sel(q0, q1, q2, q3, q4, q5, left, eps);
  if (B == 2) return left ? weno3(q1, q2, q3, eps) : weno3(q4, q3, q2, eps);
  return left ? q2 : q3;
}

// vector load of NC consecutive floats (NC = 2: LDG.64, NC = 4: LDG.128)
template <int NC> __device__ __forceinline__ void ldv(const float* p, float (&o)[NC]);
template <> __device__ __forceinline__ void ldv<4>(const float* p, float (&o)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)); o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
}
template <> __device__ __forceinline__ void ldv<2>(const float* p, float (&o)[2]) {
  const float2 a = __ldg(reinterpret_cast<const float2*>(p)); o[0] = a.x; o[1] = a.y;
}
template <int NC> __device__ __forceinline__ void stv(float* p, const float (&o)[NC]);
template <> __device__ __forceinline__ void stv<4>(float* p, const float (&o)[4]) { *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]); }
template <> __device__ __forceinline__ void stv<2>(float* p, const float (&o)[2]) { *reinterpret_cast<float2*>(p) = make_float2(o[0], o[1]); }

// One tracer per blockIdx.z (T or S); NC consecutive cells in x per thread.
template <int NC>
__global__ void __launch_bounds__(128) k_tracer_tendency_v2(DevGrid g, const DevGrid* __restrict__ gp, const float* __restrict__ u,
                                                             const float* __restrict__ v, const float* __restrict__ w,
                                                             const float* __restrict__ T0, const float* __restrict__ T1,
                                                             float* __restrict__ G0, float* __restrict__ G1,
                                                             const float* __restrict__ carry0, const float* __restrict__ carry1) {
  const int i0 = NC * (blockIdx.x * blockDim.x + threadIdx.x) + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i0 > g.Nx || j > g.Ny) return;
  const float* __restrict__ T = blockIdx.z == 0 ? T0 : T1;
  float* __restrict__ GT = blockIdx.z == 0 ? G0 : G1;
  const float* __restrict__ carry = blockIdx.z == 0 ? carry0 : carry1;
  const int PX = g.PX, n2 = g.n2, Nz = g.Nz;
  const int q2 = id2(g, i0, j);
  const float eps = g.eps;
  // ---- hoisted 2-D data
  float dyf[NC + 1], dxs[NC], dxn[NC], az[NC];
  int kgen = 0;   // levels k <= kgen (bathymetry / walls somewhere in the stencil) take the generic path
  int kzero = GB25_BIG;   // levels k <= kzero: the whole stencil is solid rock, every flux is masked, G = 0
  int kbc[NC];
#pragma unroll
  for (int e = 0; e <= NC; e++) dyf[e] = g.dyfc[q2 + e];
#pragma unroll
  for (int c = 0; c < NC; c++) {
    dxs[c] = g.dxcf[q2 + c]; dxn[c] = g.dxcf[q2 + PX + c]; az[c] = g.azcc[q2 + c];
    kbc[c] = g.kb[q2 + c];
  }
  kgen = g.kgen2[q2]; kzero = g.kzero2[q2];   // pair-based: identical for all columns of this thread
  // ---- vertical register window: WT[c][m] = T(i0+c, j, k-3+m)
  size_t q3 = q2 + (size_t)n2 * g.Hz;  // level k = 1
  float WT[NC][7];
#pragma unroll
  for (int m = 0; m < 7; m++) {
    float a[NC];
    ldv<NC>(T + q3 + (ptrdiff_t)(m - 3) * n2, a);
#pragma unroll
    for (int c = 0; c < NC; c++) WT[c][m] = a[c];
  }
  float FzT[NC];   // carried flux through the bottom face of level k (set by the generic path at k = 1)
#pragma unroll
  for (int c = 0; c < NC; c++) FzT[c] = 0.f;
  for (int k = 1; k <= Nz; k++, q3 += n2) {
    float oT[NC];
    bool skip = false;
    if (k <= kzero) {
#pragma unroll
      for (int c = 0; c < NC; c++) { oT[c] = 0.f; FzT[c] = 0.f; }
    } else if (k <= kgen) {
      // generic cells (bathymetry / a wall in the stencil) are computed by k_generic_list: nothing to do here
      skip = true;
    } else {
      if (k == kgen + 1 && kgen > 0) {   // first fast level: the flux through its bottom face was left by k_generic_list
#pragma unroll
        for (int c = 0; c < NC; c++) FzT[c] = carry[q2 + c];
      }
      const float dz = g.dzc[k + g.Hz - 1];
      // ------------- x: faces i0 .. i0+NC from the row (i0-4 .. i0+NC+3)
      float rT[NC + 8];
#pragma unroll
      for (int b = 0; b < 4 / NC; b++) {
        float a[NC], d[NC];
        ldv<NC>(T + q3 - 4 + b * NC, a);
        ldv<NC>(T + q3 + NC + b * NC, d);
#pragma unroll
        for (int c = 0; c < NC; c++) { rT[b * NC + c] = a[c]; rT[4 + NC + b * NC + c] = d[c]; }
      }
#pragma unroll
      for (int c = 0; c < NC; c++) rT[4 + c] = WT[c][3];
      float uu[NC + 1];
      {
        float a[NC];
        ldv<NC>(u + q3, a);
#pragma unroll
        for (int c = 0; c < NC; c++) uu[c] = a[c];
        uu[NC] = __ldg(u + q3 + NC);
      }
      float fxT[NC + 1];
#pragma unroll
      for (int e = 0; e <= NC; e++)
        fxT[e] = dyf[e] * dz * uu[e] * weno5_sel(rT[e + 1], rT[e + 2], rT[e + 3], rT[e + 4], rT[e + 5], rT[e + 6], uu[e] > 0.f, eps);
      // ------------- y: faces j and j+1 from rows j-3 .. j+3
      float vs[NC], vn[NC], wt[NC];
      ldv<NC>(v + q3, vs); ldv<NC>(v + q3 + PX, vn); ldv<NC>(w + q3 + n2, wt);
      float RT[7][NC];
#pragma unroll
      for (int m = 0; m < 7; m++) {
        if (m == 3) continue;
        ldv<NC>(T + q3 + (m - 3) * PX, RT[m]);
      }
#pragma unroll
      for (int c = 0; c < NC; c++) {
        const float t3 = WT[c][3];
        const float fsT = dxs[c] * dz * vs[c] * weno5_sel(RT[0][c], RT[1][c], RT[2][c], t3, RT[4][c], RT[5][c], vs[c] > 0.f, eps);
        const float fnT = dxn[c] * dz * vn[c] * weno5_sel(RT[1][c], RT[2][c], t3, RT[4][c], RT[5][c], RT[6][c], vn[c] > 0.f, eps);
        // ------------- z: top face k+1 (order reduces near the bottom and the surface); the bottom face flux is carried
        const int Bt = zbuf(g, kbc[c], k + 1, 3);
        const float ftT = az[c] * wt[c] * weno_sel_B(WT[c][1], WT[c][2], WT[c][3], WT[c][4], WT[c][5], WT[c][6], Bt, wt[c] > 0.f, eps);
        const float rV = 1.f / (az[c] * dz);
        oT[c] = -(rV * (((fxT[c + 1] - fxT[c]) + (fnT - fsT)) + (ftT - FzT[c])));
        FzT[c] = ftT;
      }
    }
    if (!skip) stv<NC>(GT + q3, oT);
    // ---- shift the vertical window and fetch level k+4
    {
      float a[NC];
      ldv<NC>(T + q3 + (size_t)4 * n2, a);
#pragma unroll
      for (int c = 0; c < NC; c++) {
#pragma unroll
        for (int m = 0; m < 6; m++) WT[c][m] = WT[c][m + 1];
        WT[c][6] = a[c];
      }
    }
  }
}

void launch_tracer_tendency_v2(Handle* h) {
  const DevGrid& g = h->g;
  constexpr int NC = GB25_TRACER_NC;
  dim3 b(32 / NC, 128 * NC / 32), gr((g.Nx / NC + b.x - 1) / b.x, (g.Ny + b.y - 1) / b.y, 2);
  StageScope ts(h, "kernel:k_tracer_tendency_v2");
  k_tracer_tendency_v2<NC><<<gr, b, 0, h->stream>>>(g, h->g_dev, h->f.u, h->f.v, h->f.w, h->f.T, h->f.S, h->f.gn[2], h->f.gn[3], h->carry[2], h->carry[3]);
  h->count_launch();
}

// =====================================================================================
// Fused auxiliary column pass: compute_w_from_continuity! (row A3) plus the by-products the momentum kernel
// would otherwise recompute 12 times per value: the horizontal flux differences dxU = delta_x(Ax u),
// dyV = delta_y(Ay v) at (C,C,C) and the vertical vorticity zeta at (F,F,C), with the immersed-aware
// (conditional) differences already applied.  One thread per column of the extended range, marching k.
// =====================================================================================
#ifndef AUX_MINB
#define AUX_MINB 8
#endif
__global__ void __launch_bounds__(128, AUX_MINB) k_aux_columns(DevGrid g, const float* __restrict__ u, const float* __restrict__ v,
                                                     float* __restrict__ w, float* __restrict__ zeta, float* __restrict__ dxU,
                                                     float* __restrict__ dyV) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + (-g.Hx + 2);
  const int j = blockIdx.y + (-g.Hy + 2);
  if (i > g.Nx + g.Hx - 1) return;
  const int q2 = id2(g, i, j), PX = g.PX, n2 = g.n2;
  const float dyE = g.dyfc[q2 + 1], dyW = g.dyfc[q2], dxN = g.dxcf[q2 + PX], dxS = g.dxcf[q2];
  const float az = g.azcc[q2];
  // vorticity metrics and conditional-difference thresholds
  const float zyE = g.dycf[q2], zyW = g.dycf[q2 - 1], zxN = g.dxfc[q2], zxS = g.dxfc[q2 - PX], azff = g.azff[q2];
  int t1 = -1, t2 = -1;
  if (g.immersed && g.cond_diff) {
    const int c00 = y_outside(g, j) ? GB25_BIG : (int)g.kb[q2], c0m = y_outside(g, j - 1) ? GB25_BIG : (int)g.kb[q2 - PX];
    const int cm0 = y_outside(g, j) ? GB25_BIG : (int)g.kb[q2 - 1], cmm = y_outside(g, j - 1) ? GB25_BIG : (int)g.kb[q2 - PX - 1];
    t1 = max(min(c00, c0m), min(cm0, cmm));   // delta_x(dy v) vanishes for k <= t1 (an inactive v node on either side)
    t2 = max(min(c00, cm0), min(c0m, cmm));   // delta_y(dx u) vanishes for k <= t2
  }
  size_t q3 = q2 + (size_t)n2 * g.Hz;  // k = 1
  float wk = 0.f;
  w[q3] = 0.f;
  // (divisions stay IEEE: w and zeta are compared element-wise.)  With both areas inside the normal range the
  // divisions are issued without the range-check branch (div_by, gb25_device.cuh) and with the reciprocals hoisted:
  // the loads of four levels are then in flight together instead of one level's (0.33 -> 0.30 ms at 64 registers; 96 registers 0.35 ms, 40 registers 0.48 ms)
  const bool safe = az > 1e-10f && az < 1e30f && azff > 1e-10f && azff < 1e30f;
  if (safe) {
    const float raz = rcp_refined(az), razff = rcp_refined(azff);
#pragma unroll 4
    for (int k = 1; k <= g.Nz; k++, q3 += n2) {
      const float dz = g.dzc[k + g.Hz - 1];
      const float u0 = u[q3], v0 = v[q3];
      const float dU = dyE * dz * u[q3 + 1] - dyW * dz * u0;
      const float dV = dxN * dz * v[q3 + PX] - dxS * dz * v0;
      dxU[q3] = dU; dyV[q3] = dV;
      wk = wk - div_by(dU + dV, az, raz);
      w[q3 + n2] = wk;
      float d1 = zyE * v0 - zyW * v[q3 - 1];
      float d2 = zxN * u0 - zxS * u[q3 - PX];
      if (k <= t1) d1 = 0.f;
      if (k <= t2) d2 = 0.f;
      zeta[q3] = div_by(d1 - d2, azff, razff);
    }
    return;
  }
#pragma unroll 4
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {
    const float dz = g.dzc[k + g.Hz - 1];
    const float u0 = u[q3], v0 = v[q3];
    const float dU = dyE * dz * u[q3 + 1] - dyW * dz * u0;
    const float dV = dxN * dz * v[q3 + PX] - dxS * dz * v0;
    dxU[q3] = dU; dyV[q3] = dV;
    wk = wk - (dU + dV) / az;
    w[q3 + n2] = wk;
    float d1 = zyE * v0 - zyW * v[q3 - 1];
    float d2 = zxN * u0 - zxS * u[q3 - PX];
    if (k <= t1) d1 = 0.f;
    if (k <= t2) d2 = 0.f;
    zeta[q3] = (d1 - d2) / azff;
  }
}
// The same pass with NC = 2 or 4 x-adjacent columns per thread and 64- / 128-bit accesses: one thread per aligned group of
// storage columns (the whole parent width; the outermost ring is computed with clamped neighbours and NOT stored, as the
// scalar kernel leaves it untouched).  Per level and group: 4 vector loads + 2 scalar loads and 4 vector stores instead of
// 6 NC scalar loads and 4 NC scalar stores — the pass is bound by memory latency (ncu: long-scoreboard 18 stalls per issue,
// 3.6 TB/s), and a thread now keeps NC times the bytes in flight.  Same expressions per column, bit-identical results.
template <int NC> struct VecT;
template <> struct VecT<2> { typedef float2 T; };
template <> struct VecT<4> { typedef float4 T; };
template <int NC> __device__ __forceinline__ void ldvec(const float* p, float (&o)[NC]) {
  const typename VecT<NC>::T a = *reinterpret_cast<const typename VecT<NC>::T*>(p);
  const float* q = reinterpret_cast<const float*>(&a);
#pragma unroll
  for (int c = 0; c < NC; c++) o[c] = q[c];
}
template <int NC> __device__ __forceinline__ void stvec(float* p, const float (&o)[NC], int lo, int hi) {   // components lo..hi-1 are stored
  if (lo == 0 && hi == NC) {
    typename VecT<NC>::T a;
    float* q = reinterpret_cast<float*>(&a);
#pragma unroll
    for (int c = 0; c < NC; c++) q[c] = o[c];
    *reinterpret_cast<typename VecT<NC>::T*>(p) = a;
  } else {
#pragma unroll
    for (int c = 0; c < NC; c++) if (c >= lo && c < hi) p[c] = o[c];
  }
}
#ifndef AUXV_MINB
#define AUXV_MINB 4
#endif
#ifndef AUXV_UNROLL
#define AUXV_UNROLL 2
#endif
template <int NC>
__global__ void __launch_bounds__(128, AUXV_MINB) k_aux_columns_vec(DevGrid g, const float* __restrict__ u, const float* __restrict__ v,
                                                                   float* __restrict__ w, float* __restrict__ zeta, float* __restrict__ dxU,
                                                                   float* __restrict__ dyV) {
  const int I0 = NC * (blockIdx.x * blockDim.x + threadIdx.x);    // first storage column of the group
  const int J = blockIdx.y + 1;                                    // storage row 1 .. PY-3
  const int PX = g.PX, n2 = g.n2;
  if (I0 >= PX) return;
  const int lo = I0 == 0 ? 1 : 0, hi = (I0 + NC == PX) ? NC - 1 : NC;      // the outermost ring is not stored
  const int j = J - g.Hy + 1;
  const int q2 = I0 + PX * J;
  const int qE = min(I0 + NC, PX - 1) + PX * J, qW = max(I0 - 1, 0) + PX * J;   // neighbours of the group (clamped at the ring)
  float dyf[NC + 1], dxN[NC], dxS[NC], az[NC], zy[NC + 1], zxN[NC], zxS[NC], azff[NC];
  int t1[NC], t2[NC];
  bool safe = true;
#pragma unroll
  for (int c = 0; c < NC; c++) {
    dyf[c] = g.dyfc[q2 + c]; dxN[c] = g.dxcf[q2 + PX + c]; dxS[c] = g.dxcf[q2 + c]; az[c] = g.azcc[q2 + c];
    zy[c + 1] = g.dycf[q2 + c]; zxN[c] = g.dxfc[q2 + c]; zxS[c] = g.dxfc[q2 - PX + c]; azff[c] = g.azff[q2 + c];
    safe = safe && az[c] > 1e-10f && az[c] < 1e30f && azff[c] > 1e-10f && azff[c] < 1e30f;
    t1[c] = -1; t2[c] = -1;
  }
  dyf[NC] = g.dyfc[qE]; zy[0] = g.dycf[qW];
  if (g.immersed && g.cond_diff) {
    const bool o0 = y_outside(g, j), om = y_outside(g, j - 1);
    int kc[NC + 1], km[NC + 1];      // kb of columns I0-1 .. I0+NC-1 on rows j and j-1
    kc[0] = o0 ? GB25_BIG : (int)g.kb[qW]; km[0] = om ? GB25_BIG : (int)g.kb[qW - PX];
#pragma unroll
    for (int c = 0; c < NC; c++) { kc[c + 1] = o0 ? GB25_BIG : (int)g.kb[q2 + c]; km[c + 1] = om ? GB25_BIG : (int)g.kb[q2 - PX + c]; }
#pragma unroll
    for (int c = 0; c < NC; c++) {
      t1[c] = max(min(kc[c + 1], km[c + 1]), min(kc[c], km[c]));
      t2[c] = max(min(kc[c + 1], kc[c]), min(km[c + 1], km[c]));
    }
  }
  float raz[NC], razff[NC], wk[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) { raz[c] = safe ? rcp_refined(az[c]) : 0.f; razff[c] = safe ? rcp_refined(azff[c]) : 0.f; wk[c] = 0.f; }
  size_t q3 = q2 + (size_t)n2 * g.Hz;  // k = 1
  stvec<NC>(w + q3, wk, lo, hi);
  const int oE = qE - q2, oW = qW - q2;
  constexpr int kUnroll = AUXV_UNROLL;
#pragma unroll kUnroll
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {
    const float dz = g.dzc[k + g.Hz - 1];
    float uu[NC + 1], vv[NC + 1], vn[NC], us[NC];
    ldvec<NC>(u + q3, *reinterpret_cast<float(*)[NC]>(uu));
    ldvec<NC>(v + q3, *reinterpret_cast<float(*)[NC]>(vv + 1));
    ldvec<NC>(v + q3 + PX, vn);
    ldvec<NC>(u + q3 - PX, us);
    uu[NC] = u[q3 + oE]; vv[0] = v[q3 + oW];
    float dU[NC], dV[NC], zt[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
      dU[c] = dyf[c + 1] * dz * uu[c + 1] - dyf[c] * dz * uu[c];
      dV[c] = dxN[c] * dz * vn[c] - dxS[c] * dz * vv[c + 1];
      float d1 = zy[c + 1] * vv[c + 1] - zy[c] * vv[c];
      float d2 = zxN[c] * uu[c] - zxS[c] * us[c];
      if (k <= t1[c]) d1 = 0.f;
      if (k <= t2[c]) d2 = 0.f;
      if (safe) { wk[c] = wk[c] - div_by(dU[c] + dV[c], az[c], raz[c]); zt[c] = div_by(d1 - d2, azff[c], razff[c]); }
      else { wk[c] = wk[c] - (dU[c] + dV[c]) / az[c]; zt[c] = (d1 - d2) / azff[c]; }
    }
    stvec<NC>(dxU + q3, dU, lo, hi); stvec<NC>(dyV + q3, dV, lo, hi);
    stvec<NC>(w + q3 + n2, wk, lo, hi); stvec<NC>(zeta + q3, zt, lo, hi);
  }
}
void launch_aux_columns(Handle* h) {
  const DevGrid& g = h->g;
  const int nx = g.Nx + 2 * g.Hx - 2, ny = g.Ny + 2 * g.Hy - 2;
  StageScope ts(h, "kernel:k_aux_columns");
  static const int nc = []() { const char* e = getenv("GB25_AUX_NC"); return e ? atoi(e) : 4; }();
  if ((nc == 4 || nc == 2) && g.PX % 4 == 0) {
    dim3 b(128), gr((g.PX / nc + 127) / 128, ny);
    if (nc == 4) k_aux_columns_vec<4><<<gr, b, 0, h->stream>>>(g, h->f.u, h->f.v, h->f.w, h->zeta, h->dxU, h->dyV);
    else k_aux_columns_vec<2><<<gr, b, 0, h->stream>>>(g, h->f.u, h->f.v, h->f.w, h->zeta, h->dxU, h->dyV);
    h->count_launch();
    return;
  }
  dim3 b(128), gr((nx + 127) / 128, ny);
  k_aux_columns<<<gr, b, 0, h->stream>>>(g, h->f.u, h->f.v, h->f.w, h->zeta, h->dxU, h->dyV);
  h->count_launch();
}

// =====================================================================================
// Momentum tendencies, second generation (row A5).  Same arithmetic as momentum_G<DIR> (gb25_tend_generic.cuh)
// for cells whose stencil is clear of bathymetry and walls; those cells take no masks, no horizontal order
// reduction and read zeta / dxU / dyV from the scratch arrays of k_aux_columns instead of rebuilding them from
// u, v and six metric arrays.  NC = 2 consecutive cells in x per thread, k-marching with the vertical momentum
// flux carried from face to face and the own-velocity column held in a register window.  Everything else falls
// back to the generic function (not inlined).
// =====================================================================================
template <int DIR>
static __device__ __noinline__ float momentum_G_call(const DevGrid* __restrict__ gp, const float* __restrict__ own,
                                                     const float* __restrict__ oth, const float* __restrict__ w,
                                                     const float* __restrict__ p, int i, int j, int k, float* wtop) {
  return momentum_G<DIR>(*gp, own, oth, w, p, i, j, k, wtop);
}
__device__ __forceinline__ float2 ld2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
// load 2*NP consecutive floats starting at an 8-byte aligned address
template <int NP>
__device__ __forceinline__ void ldrow(const float* p, float (&o)[2 * NP]) {
#pragma unroll
  for (int b = 0; b < NP; b++) { const float2 a = ld2(p + 2 * b); o[2 * b] = a.x; o[2 * b + 1] = a.y; }
}
__device__ __forceinline__ float weno5_fs_sel(const float (&q)[6], const float (&s)[6], bool left, float eps) {
  return weno5_fs_selq(q, s, left, eps);
}

// ---- Gu: own direction = x (contiguous), cross = y
__global__ void __launch_bounds__(128) k_gu_v2(DevGrid g, const DevGrid* __restrict__ gp, const float* __restrict__ u,
                                               const float* __restrict__ v, const float* __restrict__ w,
                                               const float* __restrict__ p, const float* __restrict__ zeta,
                                               const float* __restrict__ dxU, const float* __restrict__ dyV,
                                               float* __restrict__ G, const float* __restrict__ carry) {
  constexpr int NC = 2;
  const int i0 = NC * (blockIdx.x * blockDim.x + threadIdx.x) + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i0 > g.Nx || j > g.Ny) return;
  const int PX = g.PX, n2 = g.n2, Nz = g.Nz;
  const int q2 = id2(g, i0, j);
  const float eps = g.eps;
  // ---- hoisted 2-D data
  float m1[NC], rV0[NC], fbar[NC], mv[2][NC + 1], azw[NC + 3];
  int kbc[NC], kgen = 0, kzero = GB25_BIG;
#pragma unroll
  for (int c = 0; c < NC; c++) {
    m1[c] = g.dxfc[q2 + c];
    rV0[c] = g.azfc[q2 + c];
    fbar[c] = (g.fff[q2 + c] + g.fff[q2 + c + PX]) * 0.5f;
    kbc[c] = g.kb[q2 + c];
  }
  kgen = g.kgen2[q2]; kzero = g.kzero2[q2];
#pragma unroll
  for (int c = 0; c <= NC; c++) { mv[0][c] = g.dxcf[q2 + c - 1]; mv[1][c] = g.dxcf[q2 + c - 1 + PX]; }
#pragma unroll
  for (int c = 0; c < NC + 3; c++) azw[c] = g.azcc[q2 + c - 2];
  // ---- vertical register window of u: WU[c][m] = u(i0+c, j, k-3+m)
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  float WU[NC][7];
#pragma unroll
  for (int m = 0; m < 7; m++) {
    const float2 a = ld2(u + q3 + (ptrdiff_t)(m - 3) * n2);
    WU[0][m] = a.x; WU[1][m] = a.y;
  }
  float Wb[NC] = {0.f, 0.f};   // carried vertical flux through the bottom face (set by the generic path at k = 1)
  for (int k = 1; k <= Nz; k++, q3 += n2) {
    float out[NC];
    bool skip = false;
    if (k <= kzero) {   // solid rock all around: u = v = 0, every flux masked, pressure difference conditional => G = 0
      out[0] = out[1] = 0.f; Wb[0] = Wb[1] = 0.f;
    } else if (k <= kgen) {
      skip = true;   // generic cells are computed by k_generic_list
    } else {
      if (k == kgen + 1 && kgen > 0) { Wb[0] = carry[q2]; Wb[1] = carry[q2 + 1]; }
      const float dz = g.dzc[k + g.Hz - 1];
      // ---- rows of u (row j: cols i0-4 .. i0+5; rows j-3..j+3: cols i0, i0+1)
      float ur[10];
      {
        float a[4], b[4];
        ldrow<2>(u + q3 - 4, a); ldrow<2>(u + q3 + 2, b);
#pragma unroll
        for (int n = 0; n < 4; n++) { ur[n] = a[n]; ur[6 + n] = b[n]; }
        ur[4] = WU[0][3]; ur[5] = WU[1][3];
      }
      float uy[7][NC];
#pragma unroll
      for (int m = 0; m < 7; m++) {
        if (m == 3) { uy[3][0] = ur[4]; uy[3][1] = ur[5]; continue; }
        const float2 a = ld2(u + q3 + (m - 3) * PX); uy[m][0] = a.x; uy[m][1] = a.y;
      }
      // ---- rows j-2..j+3 of v: cols i0-2 .. i0+1 (col i0-2 unused), and of zeta: cols i0, i0+1
      float vy[6][4], zy[6][NC];
#pragma unroll
      for (int m = 0; m < 6; m++) {
        ldrow<2>(v + q3 - 2 + (m - 2) * PX, vy[m]);
        const float2 a = ld2(zeta + q3 + (m - 2) * PX); zy[m][0] = a.x; zy[m][1] = a.y;
      }
      // ---- row j of dxU, dyV: cols i0-4 .. i0+3
      float dxr[8], dyr[8];
      ldrow<4>(dxU + q3 - 4, dxr); ldrow<4>(dyV + q3 - 4, dyr);
      // ---- w at the top face, cols i0-2 .. i0+3 (i0+3 unused); p at cols i0-2(unused), i0-1, i0, i0+1
      float wr[6], pr[4];
      ldrow<3>(w + q3 + n2 - 2, wr); ldrow<2>(p + q3 - 2, pr);
      const int Bw = (g.immersed && k + 1 > Nz) ? 1 : 2;
#pragma unroll
      for (int c = 0; c < NC; c++) {
        const float own0 = ur[4 + c];
        const bool lown = own0 > 0.f;
        // vorticity flux: -v^ zeta^R (VelocityStencil smoothness)
        const float xm0 = mv[0][c] * vy[2][1 + c], xm1 = mv[1][c] * vy[3][1 + c];
        const float x00 = mv[0][c + 1] * vy[2][2 + c], x01 = mv[1][c + 1] * vy[3][2 + c];
        const float oavg = ((xm0 + xm1) * 0.5f + (x00 + x01) * 0.5f) * 0.5f;
        const float ohat = oavg / m1[c];
        float zq[6], zs[6], zr[6];
#pragma unroll
        for (int m = 0; m < 6; m++) {
          zq[m] = zy[m][c];
          zs[m] = (uy[m][c] + uy[m + 1][c]) * 0.5f;
          zr[m] = (vy[m][1 + c] + vy[m][2 + c]) * 0.5f;
        }
        const float zR = weno5_vs_selq(zq, zs, zr, ohat > 0.f, eps);
        const float Hterm = -ohat * zR;
        // divergence flux and kinetic-energy gradient along x (cells i-3 .. i+2  <->  dxr[1+c .. 6+c])
        float dOw[6], dv[6], dK[6], sK[6];
#pragma unroll
        for (int m = 0; m < 6; m++) {
          dOw[m] = dxr[1 + c + m];
          dv[m] = dxr[1 + c + m] + dyr[1 + c + m];
          const float o0 = ur[1 + c + m], o1 = ur[2 + c + m];
          dK[m] = o1 * o1 * 0.5f - o0 * o0 * 0.5f;
          sK[m] = (o0 + o1) * 0.5f;
        }
        const float dvs = sym4(dyr[2 + c], dyr[3 + c], dyr[4 + c], dyr[5 + c], 2);
        const float duR = weno5_fs_sel(dOw, dv, lown, eps);
        const float Phi = own0 * (dvs + duR);
        const float dKo = weno5_fs_sel(dK, sK, lown, eps);
        float kc[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
          const float t0 = vy[m + 1][2 + c], tm = vy[m + 1][1 + c];
          kc[m] = t0 * t0 * 0.5f - tm * tm * 0.5f;
        }
        const float Bterm = (dKo + sym4(kc[0], kc[1], kc[2], kc[3], 2)) / m1[c];
        // vertical advection: top face k+1, bottom face carried
        const float wt = sym4(azw[c] * wr[c], azw[c + 1] * wr[c + 1], azw[c + 2] * wr[c + 2], azw[c + 3] * wr[c + 3], Bw);
        const int Bz = zbuf(g, kbc[c], k + 1, 3);
        const float Wt = wt * weno_sel_B(WU[c][1], WU[c][2], WU[c][3], WU[c][4], WU[c][5], WU[c][6], Bz, wt > 0.f, eps);
        const float Vterm = (1.f / (rV0[c] * dz)) * (Phi + (Wt - Wb[c]));
        Wb[c] = Wt;
        const float cor = -(fbar[c] * oavg / m1[c]);
        const float dp = (pr[2 + c] - pr[1 + c]) / m1[c];
        out[c] = -(Hterm + Vterm + Bterm) - cor - dp;
      }
    }
    if (!skip) *reinterpret_cast<float2*>(G + q3) = make_float2(out[0], out[1]);
    {
      const float2 a = ld2(u + q3 + (size_t)4 * n2);
#pragma unroll
      for (int c = 0; c < NC; c++) {
#pragma unroll
        for (int m = 0; m < 6; m++) WU[c][m] = WU[c][m + 1];
      }
      WU[0][6] = a.x; WU[1][6] = a.y;
    }
  }
}

// ---- Gv: own direction = y, cross = x.  The swapped vorticity of momentum_G<1> is -zeta and the
// reconstruction is odd, so +u^ zeta^R with zeta from the scratch array is the identical value.
__global__ void __launch_bounds__(128) k_gv_v2(DevGrid g, const DevGrid* __restrict__ gp, const float* __restrict__ u,
                                               const float* __restrict__ v, const float* __restrict__ w,
                                               const float* __restrict__ p, const float* __restrict__ zeta,
                                               const float* __restrict__ dxU, const float* __restrict__ dyV,
                                               float* __restrict__ G, const float* __restrict__ carry) {
  constexpr int NC = 2;
  const int i0 = NC * (blockIdx.x * blockDim.x + threadIdx.x) + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i0 > g.Nx || j > g.Ny) return;
  const int PX = g.PX, n2 = g.n2, Nz = g.Nz;
  const int q2 = id2(g, i0, j);
  const float eps = g.eps;
  // ---- hoisted 2-D data
  float m1[NC], rV0[NC], fbar[NC], mu[2][NC + 1], azw[4][NC];
  int kbc[NC], kgen = 0, kzero = GB25_BIG;
#pragma unroll
  for (int c = 0; c < NC; c++) {
    m1[c] = g.dycf[q2 + c];
    rV0[c] = g.azcf[q2 + c];
    fbar[c] = (g.fff[q2 + c] + g.fff[q2 + c + 1]) * 0.5f;
    kbc[c] = g.kb[q2 + c];
#pragma unroll
    for (int m = 0; m < 4; m++) azw[m][c] = g.azcc[q2 + c + (m - 2) * PX];
  }
#pragma unroll
  for (int c = 0; c <= NC; c++) { mu[0][c] = g.dyfc[q2 + c - PX]; mu[1][c] = g.dyfc[q2 + c]; }
  kgen = g.kgen2[q2]; kzero = g.kzero2[q2];
  // ---- vertical register window of v
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  float WV[NC][7];
#pragma unroll
  for (int m = 0; m < 7; m++) {
    const float2 a = ld2(v + q3 + (ptrdiff_t)(m - 3) * n2);
    WV[0][m] = a.x; WV[1][m] = a.y;
  }
  float Wb[NC] = {0.f, 0.f};
  for (int k = 1; k <= Nz; k++, q3 += n2) {
    float out[NC];
    bool skip = false;
    if (k <= kzero) {
      out[0] = out[1] = 0.f; Wb[0] = Wb[1] = 0.f;
    } else if (k <= kgen) {
      skip = true;   // generic cells are computed by k_generic_list
    } else {
      if (k == kgen + 1 && kgen > 0) { Wb[0] = carry[q2]; Wb[1] = carry[q2 + 1]; }
      const float dz = g.dzc[k + g.Hz - 1];
      // ---- v: row j cols i0-4 .. i0+5; rows j-3..j+3 cols i0, i0+1
      float vr[10];
      {
        float a[4], b[4];
        ldrow<2>(v + q3 - 4, a); ldrow<2>(v + q3 + 2, b);
#pragma unroll
        for (int n = 0; n < 4; n++) { vr[n] = a[n]; vr[6 + n] = b[n]; }
        vr[4] = WV[0][3]; vr[5] = WV[1][3];
      }
      float vy[7][NC];
#pragma unroll
      for (int m = 0; m < 7; m++) {
        if (m == 3) { vy[3][0] = vr[4]; vy[3][1] = vr[5]; continue; }
        const float2 a = ld2(v + q3 + (m - 3) * PX); vy[m][0] = a.x; vy[m][1] = a.y;
      }
      // ---- u rows j-1, j and zeta row j: cols i0-2 .. i0+5
      float us[8], un[8], zr8[8];
      ldrow<4>(u + q3 - 2 - PX, us); ldrow<4>(u + q3 - 2, un); ldrow<4>(zeta + q3 - 2, zr8);
      // ---- dyV, dxU rows j-3 .. j+2, w (top face) rows j-2 .. j+1, p rows j-1, j: cols i0, i0+1
      float dyc[6][NC], dxc[6][NC], wc[4][NC], ps[NC], pn[NC];
#pragma unroll
      for (int m = 0; m < 6; m++) {
        const float2 a = ld2(dyV + q3 + (m - 3) * PX), b = ld2(dxU + q3 + (m - 3) * PX);
        dyc[m][0] = a.x; dyc[m][1] = a.y; dxc[m][0] = b.x; dxc[m][1] = b.y;
      }
#pragma unroll
      for (int m = 0; m < 4; m++) { const float2 a = ld2(w + q3 + n2 + (m - 2) * PX); wc[m][0] = a.x; wc[m][1] = a.y; }
      { const float2 a = ld2(p + q3 - PX), b = ld2(p + q3); ps[0] = a.x; ps[1] = a.y; pn[0] = b.x; pn[1] = b.y; }
      const int Bw = (g.immersed && k + 1 > Nz) ? 1 : 2;
#pragma unroll
      for (int c = 0; c < NC; c++) {
        const float own0 = vr[4 + c];
        const bool lown = own0 > 0.f;
        // us/un index n <-> col i0-2+n ; cell i = i0+c  ->  u(i) at n = c+2
        const float xm0 = mu[0][c] * us[c + 2], xm1 = mu[0][c + 1] * us[c + 3];
        const float x00 = mu[1][c] * un[c + 2], x01 = mu[1][c + 1] * un[c + 3];
        const float oavg = ((xm0 + xm1) * 0.5f + (x00 + x01) * 0.5f) * 0.5f;
        const float ohat = oavg / m1[c];
        float zq[6], zs[6], zr[6];
#pragma unroll
        for (int m = 0; m < 6; m++) {           // window along x: cols i-2 .. i+3
          zq[m] = zr8[c + m];
          zs[m] = (vr[1 + c + m] + vr[2 + c + m]) * 0.5f;   // (v(i+b-1, j) + v(i+b, j)) / 2
          zr[m] = (us[c + m] + un[c + m]) * 0.5f;            // (u(i+b, j-1) + u(i+b, j)) / 2
        }
        const float zR = weno5_vs_selq(zq, zs, zr, ohat > 0.f, eps);
        const float Hterm = ohat * zR;
        float dOw[6], dv[6], dK[6], sK[6];
#pragma unroll
        for (int m = 0; m < 6; m++) {           // window along y: rows j-3 .. j+2
          dOw[m] = dyc[m][c];
          dv[m] = dxc[m][c] + dyc[m][c];
          const float o0 = vy[m][c], o1 = vy[m + 1][c];
          dK[m] = o1 * o1 * 0.5f - o0 * o0 * 0.5f;
          sK[m] = (o0 + o1) * 0.5f;
        }
        const float dus = sym4(dxc[1][c], dxc[2][c], dxc[3][c], dxc[4][c], 2);
        const float dvR = weno5_fs_sel(dOw, dv, lown, eps);
        const float Phi = own0 * (dus + dvR);
        const float dKo = weno5_fs_sel(dK, sK, lown, eps);
        float kc[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {           // cols i-1 .. i+2
          const float t0 = un[c + 1 + m], tm = us[c + 1 + m];
          kc[m] = t0 * t0 * 0.5f - tm * tm * 0.5f;
        }
        const float Bterm = (dKo + sym4(kc[0], kc[1], kc[2], kc[3], 2)) / m1[c];
        const float wt = sym4(azw[0][c] * wc[0][c], azw[1][c] * wc[1][c], azw[2][c] * wc[2][c], azw[3][c] * wc[3][c], Bw);
        const int Bz = zbuf(g, kbc[c], k + 1, 3);
        const float Wt = wt * weno_sel_B(WV[c][1], WV[c][2], WV[c][3], WV[c][4], WV[c][5], WV[c][6], Bz, wt > 0.f, eps);
        const float Vterm = (1.f / (rV0[c] * dz)) * (Phi + (Wt - Wb[c]));
        Wb[c] = Wt;
        const float cor = fbar[c] * oavg / m1[c];
        const float dp = (pn[c] - ps[c]) / m1[c];
        out[c] = -(Hterm + Vterm + Bterm) - cor - dp;
      }
    }
    if (!skip) *reinterpret_cast<float2*>(G + q3) = make_float2(out[0], out[1]);
    {
      const float2 a = ld2(v + q3 + (size_t)4 * n2);
#pragma unroll
      for (int c = 0; c < NC; c++) {
#pragma unroll
        for (int m = 0; m < 6; m++) WV[c][m] = WV[c][m + 1];
      }
      WV[0][6] = a.x; WV[1][6] = a.y;
    }
  }
}

void launch_momentum_tendency_v2(Handle* h) {
  const DevGrid& g = h->g;
  dim3 b(16, 8), gr((g.Nx / 2 + b.x - 1) / b.x, (g.Ny + b.y - 1) / b.y);
  k_gu_v2<<<gr, b, 0, h->stream>>>(g, h->g_dev, h->f.u, h->f.v, h->f.w, h->f.p, h->zeta, h->dxU, h->dyV, h->f.gn[0], h->carry[0]); h->count_launch();
  k_gv_v2<<<gr, b, 0, h->stream>>>(g, h->g_dev, h->f.u, h->f.v, h->f.w, h->f.p, h->zeta, h->dxU, h->dyV, h->f.gn[1], h->carry[1]); h->count_launch();
}

// =====================================================================================
// Generic cells (bathymetry or a wall somewhere in the stencil: ~1.5 % of the cells of the 1/4-degree tripolar
// workload) handled as a flat list, one thread per cell, all four tendencies at once.  Inside the blocked kernels
// these cells stalled whole warps / CTAs (a generic cell costs ~3x a fast one and reads global memory); here they
// run at full occupancy.  The topmost generic cell of a column leaves the vertical fluxes through its top face in
// the 2-D carry arrays, where the fast kernels pick them up.
// =====================================================================================
#ifndef GENERIC_MINB
#define GENERIC_MINB 12   // CTAs per SM: 1 (80 registers) 0.154 ms per launch, 8: 0.124, 12: 0.116, 16: 0.114 (latency-bound, divergent)
#endif
__global__ void __launch_bounds__(128, GENERIC_MINB) k_generic_list(DevGrid g, const DevGrid* __restrict__ gp, const float* __restrict__ u,
                                                      const float* __restrict__ v, const float* __restrict__ w,
                                                      const float* __restrict__ p, const float* __restrict__ T,
                                                      const float* __restrict__ S, float* __restrict__ Gu, float* __restrict__ Gv,
                                                      float* __restrict__ GT, float* __restrict__ GS, float* __restrict__ cu,
                                                      float* __restrict__ cv, float* __restrict__ cT, float* __restrict__ cS,
                                                      int do_mom, int do_trc, const float* __restrict__ zeta,
                                                      const float* __restrict__ dxU, const float* __restrict__ dyV) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.nglist) return;
  const int q3 = g.glist[idx];
  const int K = q3 / g.n2, r = q3 - K * g.n2, J = r / g.PX, I = r - J * g.PX;
  const int i = I - g.Hx + 1, j = J - g.Hy + 1, k = K - g.Hz + 1;
  const bool top = (k == (int)g.kgen2[r]) && k < g.Nz;
  if (do_mom) {
    float wu, wv;
    // (zeta, dx(Ax u), dy(Ay v) come from the scratch arrays of the column pass that precedes the tendency kernels)
    Gu[q3] = momentum_G<0, true>(g, u, v, w, p, i, j, k, &wu, zeta, dxU, dyV);
    Gv[q3] = momentum_G<1, true>(g, v, u, w, p, i, j, k, &wv, zeta, dxU, dyV);
    if (top) { cu[r] = wu; cv[r] = wv; }
  }
  if (do_trc) {
    float fT, fS, gT, gS;
    tracer_cell_generic(g, u, v, w, T, S, i, j, k, gT, gS, &fT, &fS);   // both tracers share velocities, areas, orders and masks
    GT[q3] = gT; GS[q3] = gS;
    if (top) { cT[r] = fT; cS[r] = fS; }
  }
}
void launch_generic_list(Handle* h, bool momentum, bool tracers) {
  const DevGrid& g = h->g;
  if (g.nglist == 0) return;
  StageScope ts(h, "kernel:k_generic_list");
  k_generic_list<<<(g.nglist + 127) / 128, 128, 0, h->stream>>>(g, h->g_dev, h->f.u, h->f.v, h->f.w, h->f.p, h->f.T, h->f.S,
                                                              h->f.gn[0], h->f.gn[1], h->f.gn[2], h->f.gn[3], h->carry[0],
                                                              h->carry[1], h->carry[2], h->carry[3], momentum ? 1 : 0, tracers ? 1 : 0,
                                                              h->zeta, h->dxU, h->dyV);
  h->count_launch();
}

// ---------------------------------------------------------------- kernel table (preload_kernels, gb25_api.cu)
KernelTable kernel_table_tend_v2() {
  static const void* const k[] = {
    (const void*)k_tracer_tendency_v2<2>, (const void*)k_aux_columns, (const void*)k_aux_columns_vec<2>, (const void*)k_aux_columns_vec<4>,
    (const void*)k_gu_v2, (const void*)k_gv_v2, (const void*)k_generic_list,
  };
  return {k, (int)(sizeof k / sizeof k[0])};
}
