// gb25_packed.cuh — packed FP32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2) for the WENO kernels.
//
// sm_100 adds `add/sub/mul/fma.rn.f32x2`: one instruction operates on an aligned register pair.  The tendency
// kernels are bound by the instruction-issue rate, not by the FMA pipe (ncu: issue slots 72-75 % active, FMA pipe
// 43-50 %), and a microbenchmark on this B200 (scripts/micro/ffma2_bench.cu) measures 42.4 TFLOP/s for a stream of
// independent scalar 3-register FFMAs against 64.0 TFLOP/s for the same FMAs issued as FFMA2.  So every thread
// evaluates TWO reconstructions at a time with its operands held as float2; selects, min/max and MUFU.RCP have no
// packed form and stay per component.  Constants appear as immediates (the assembler broadcasts them).
#pragma once
#include <cuda_runtime.h>

#include "gb25_device.cuh"

typedef unsigned long long gb25_u64;
#define GB25_P(x) (*reinterpret_cast<const gb25_u64*>(&(x)))
__device__ __forceinline__ float2 padd(float2 a, float2 b) { gb25_u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(GB25_P(a)), "l"(GB25_P(b))); return *reinterpret_cast<float2*>(&r); }
__device__ __forceinline__ float2 psub(float2 a, float2 b) { gb25_u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(GB25_P(a)), "l"(GB25_P(b))); return *reinterpret_cast<float2*>(&r); }
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { gb25_u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(GB25_P(a)), "l"(GB25_P(b))); return *reinterpret_cast<float2*>(&r); }
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { gb25_u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(GB25_P(a)), "l"(GB25_P(b)), "l"(GB25_P(c))); return *reinterpret_cast<float2*>(&r); }
__device__ __forceinline__ float2 pbc(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 pmuls(float2 a, float s) { return pmul(a, pbc(s)); }
__device__ __forceinline__ float2 pfmas(float2 a, float s, float2 c) { return pfma(a, pbc(s), c); }   // a*s + c
__device__ __forceinline__ float2 prcp(float2 a) { return make_float2(frcp(a.x), frcp(a.y)); }
__device__ __forceinline__ float2 prcp_exact(float2 a) { return make_float2(1.f / a.x, 1.f / a.y); }   // IEEE: feeds a model field
// the same correctly rounded reciprocal without the range-check branch of `1.f / x` (cell volumes are normal numbers)
__device__ __forceinline__ float2 prcp_nr(float2 a) { return make_float2(rcp_refined(a.x), rcp_refined(a.y)); }

// smoothness indicators / 3.25 (see beta5_* in gb25_device.cuh)
__device__ __forceinline__ float2 pbeta0(float2 a, float2 b, float2 c) {
  const float2 d2 = padd(pfmas(b, -2.f, a), c), d1 = pfmas(a, 3.f, pfmas(b, -4.f, c));
  return pfma(d2, d2, pmul(pmuls(d1, 3.f / 13.f), d1));
}
__device__ __forceinline__ float2 pbeta1(float2 a, float2 b, float2 c) {
  const float2 d2 = padd(pfmas(b, -2.f, a), c), d1 = psub(a, c);
  return pfma(d2, d2, pmul(pmuls(d1, 3.f / 13.f), d1));
}
__device__ __forceinline__ float2 pbeta2(float2 a, float2 b, float2 c) {
  const float2 d2 = padd(pfmas(b, -2.f, a), c), d1 = pfmas(c, 3.f, pfmas(b, -4.f, a));
  return pfma(d2, d2, pmul(pmuls(d1, 3.f / 13.f), d1));
}
// two WENO5-Z reconstructions at once; arguments far-upwind -> downwind (already mirrored per component),
// b0..b2 = smoothness indicators / 3.25.  Same algebra as weno5_combine; tau enters squared, so no |.| is needed.
__device__ __forceinline__ float2 pweno5_combine(float2 v0, float2 v1, float2 v2, float2 v3, float2 v4,
                                                 float2 b0, float2 b1, float2 b2, float eps) {
  const float2 es = pbc(eps * GB25_BETA_SCALE);
  const float2 d = psub(b0, b2);
  float2 t0 = pmul(d, prcp(padd(b0, es))), t1 = pmul(d, prcp(padd(b1, es))), t2 = pmul(d, prcp(padd(b2, es)));
  t0 = pmul(t0, t0); t1 = pmul(t1, t1); t2 = pmul(t2, t2);
  t0 = make_float2(fminf(t0.x, 1e36f), fminf(t0.y, 1e36f));
  t1 = make_float2(fminf(t1.x, 1e36f), fminf(t1.y, 1e36f));
  t2 = make_float2(fminf(t2.x, 1e36f), fminf(t2.y, 1e36f));
  const float2 a0 = pfmas(t0, 0.3f, pbc(0.3f)), a1 = pfmas(t1, 0.6f, pbc(0.6f)), a2 = pfmas(t2, 0.1f, pbc(0.1f));
  const float2 p0 = pfmas(v2, 1.f / 3.f, pfmas(v3, 5.f / 6.f, pmuls(v4, -1.f / 6.f)));
  const float2 p1 = pfmas(v1, -1.f / 6.f, pfmas(v2, 5.f / 6.f, pmuls(v3, 1.f / 3.f)));
  const float2 p2 = pfmas(v0, 1.f / 3.f, pfmas(v1, -7.f / 6.f, pmuls(v2, 11.f / 6.f)));
  const float2 rs = prcp(padd(padd(a0, a1), a2));
  return pfma(pmul(a2, rs), p2, pfma(pmul(a1, rs), p1, pmul(pmul(a0, rs), p0)));
}
__device__ __forceinline__ float2 pweno5(float2 v0, float2 v1, float2 v2, float2 v3, float2 v4, float eps) {
  return pweno5_combine(v0, v1, v2, v3, v4, pbeta0(v2, v3, v4), pbeta1(v1, v2, v3), pbeta2(v0, v1, v2), eps);
}
__device__ __forceinline__ float2 pweno5_fs(float2 v0, float2 v1, float2 v2, float2 v3, float2 v4,
                                            float2 s0, float2 s1, float2 s2, float2 s3, float2 s4, float eps) {
  return pweno5_combine(v0, v1, v2, v3, v4, pbeta0(s2, s3, s4), pbeta1(s1, s2, s3), pbeta2(s0, s1, s2), eps);
}
__device__ __forceinline__ float2 pweno5_vs(float2 v0, float2 v1, float2 v2, float2 v3, float2 v4,
                                            float2 s0, float2 s1, float2 s2, float2 s3, float2 s4,
                                            float2 r0, float2 r1, float2 r2, float2 r3, float2 r4, float eps) {
  const float2 b0 = pmuls(padd(pbeta0(s2, s3, s4), pbeta0(r2, r3, r4)), 0.5f);
  const float2 b1 = pmuls(padd(pbeta1(s1, s2, s3), pbeta1(r1, r2, r3)), 0.5f);
  const float2 b2 = pmuls(padd(pbeta2(s0, s1, s2), pbeta2(r0, r1, r2)), 0.5f);
  return pweno5_combine(v0, v1, v2, v3, v4, b0, b1, b2, eps);
}
// per-component upwind mirror of a six-point window q[0..5] (face between q[2] and q[3]) + packed reconstruction
__device__ __forceinline__ float2 psel(float2 a, float2 b, bool lx, bool ly) { return make_float2(lx ? a.x : b.x, ly ? a.y : b.y); }
__device__ __forceinline__ float2 pweno5_sel(const float2* q, bool lx, bool ly, float eps) {
  return pweno5(psel(q[0], q[5], lx, ly), psel(q[1], q[4], lx, ly), psel(q[2], q[3], lx, ly), psel(q[3], q[2], lx, ly),
                psel(q[4], q[1], lx, ly), eps);
}
