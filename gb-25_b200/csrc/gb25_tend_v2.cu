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
                                                             float* __restrict__ G0, float* __restrict__ G1) {
  const int i0 = NC * (blockIdx.x * blockDim.x + threadIdx.x) + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i0 > g.Nx || j > g.Ny) return;
  const float* __restrict__ T = blockIdx.z == 0 ? T0 : T1;
  float* __restrict__ GT = blockIdx.z == 0 ? G0 : G1;
  const int PX = g.PX, n2 = g.n2, Nz = g.Nz;
  const int q2 = id2(g, i0, j);
  const float eps = g.eps;
  // ---- hoisted 2-D data
  float dyf[NC + 1], dxs[NC], dxn[NC], az[NC];
  int kgen = 0;   // levels k <= kgen (bathymetry / walls somewhere in the stencil) take the generic path
  int kbc[NC];
#pragma unroll
  for (int e = 0; e <= NC; e++) dyf[e] = g.dyfc[q2 + e];
#pragma unroll
  for (int c = 0; c < NC; c++) {
    dxs[c] = g.dxcf[q2 + c]; dxn[c] = g.dxcf[q2 + PX + c]; az[c] = g.azcc[q2 + c];
    kbc[c] = g.kb[q2 + c];
    kgen = max(kgen, (int)g.knear[q2 + c] + 1);
  }
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
    if (k <= kgen) {   // always taken at k = 1, so the carried flux is valid from the first fast level on
      for (int c = 0; c < NC; c++) {
        float ft;
        const float o = tracer_cell_generic1(gp, u, v, w, T, i0 + c, j, k, &ft);
#pragma unroll
        for (int cc = 0; cc < NC; cc++) if (cc == c) { oT[cc] = o; FzT[cc] = ft; }
      }
    } else {
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
    stv<NC>(GT + q3, oT);
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
  k_tracer_tendency_v2<NC><<<gr, b, 0, h->stream>>>(g, h->g_dev, h->f.u, h->f.v, h->f.w, h->f.T, h->f.S, h->f.gn[2], h->f.gn[3]);
  h->count_launch();
}
