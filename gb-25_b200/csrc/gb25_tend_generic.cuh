// gb25_tend_generic.cuh — generic (any cell: walls, bathymetry, reduced order) per-cell tendency functions.
// Used directly by the first-generation thread-per-cell kernels and as the fallback of the blocked
// k-marching kernels (gb25_tend_v2.cu) for cells whose stencil is not clear of bathymetry / walls.
#pragma once
#include "gb25_device.cuh"

// =====================================================================================
// Tracer tendencies (row A6): Gc = -div(U c), WENO5-Z upwind flux form, T and S together
// =====================================================================================
__device__ __forceinline__ void tracer_cell_generic(const DevGrid& g, const real* __restrict__ u, const real* __restrict__ v,
                                                    const real* __restrict__ w, const real* __restrict__ T,
                                                    const real* __restrict__ S, int i, int j, int k, real& outT, real& outS,
                                                    real* ftopT = nullptr, real* ftopS = nullptr) {
  const int q2 = id2(g, i, j), PX = g.PX, n2 = g.n2;
  const size_t q3 = q2 + (size_t)n2 * (k + g.Hz - 1);
  const real eps = g.eps;
  const bool clear = (k - 1) > (int)g.knear[q2];
  const bool imm = g.immersed && !clear;
  const real dz = g.dzc[k + g.Hz - 1];
  const real* Tc = T + q3; const real* Sc = S + q3;
  real dT = R(0.), dS = R(0.);
  // ---- x faces i (e=0) and i+1 (e=1)
  real fT[2], fS[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const real vel = u[q3 + e];
    const real area = g.dyfc[q2 + e] * dz;
    const int B = clear ? 3 : buf_from(g.fx3, g.fx2, q2 + e, k);
    const bool masked = imm && (k <= (int)g.kb[q2 + e] || k <= (int)g.kb[q2 + e - 1]);
    const bool left = vel > R(0.);
    fT[e] = masked ? R(0.) : area * vel * recon_mem(Tc + e, 1, B, left, eps);
    fS[e] = masked ? R(0.) : area * vel * recon_mem(Sc + e, 1, B, left, eps);
  }
  dT = fT[1] - fT[0]; dS = fS[1] - fS[0];
  // ---- y faces j and j+1
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const real vel = v[q3 + e * PX];
    const real area = g.dxcf[q2 + e * PX] * dz;
    const int B = clear ? 3 : buf_from(g.fy3, g.fy2, q2 + e * PX, k);
    const bool wall = y_outside(g, j + e) || y_outside(g, j + e - 1);
    const bool masked = imm && !wall && (k <= (int)g.kb[q2 + e * PX] || k <= (int)g.kb[q2 + (e - 1) * PX]);
    const bool left = vel > R(0.);
    fT[e] = masked ? R(0.) : area * vel * recon_mem(Tc + e * PX, PX, B, left, eps);
    fS[e] = masked ? R(0.) : area * vel * recon_mem(Sc + e * PX, PX, B, left, eps);
  }
  dT += fT[1] - fT[0]; dS += fS[1] - fS[0];
  // ---- z faces k and k+1
  const int kbc = g.kb[q2];
  const real az = g.azcc[q2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int kk = k + e;
    const real vel = w[q3 + (size_t)e * n2];
    const int B = zbuf(g, kbc, kk, 3);
    const bool masked = g.immersed && kk != 1 && kk != g.Nz + 1 && (kk - 1 <= kbc);
    const bool left = vel > R(0.);
    fT[e] = masked ? R(0.) : az * vel * recon_mem(Tc + e * n2, n2, B, left, eps);
    fS[e] = masked ? R(0.) : az * vel * recon_mem(Sc + e * n2, n2, B, left, eps);
  }
  dT += fT[1] - fT[0]; dS += fS[1] - fS[0];
  const real rV = R(1.) / (az * dz);
  outT = -(rV * dT);
  outS = -(rV * dS);
  if (ftopT) { *ftopT = fT[1]; *ftopS = fS[1]; }   // (masked) fluxes through the top face, for k-marching callers
}

// single-tracer variant (same arithmetic) used by the blocked kernel's fallback
// Returns G; *ftop receives the (masked) flux through the top face so that a k-marching caller can carry it.
// (takes the grid descriptor through a global-memory pointer: a by-reference kernel parameter would be
// copied to the local stack at every call of a non-inlined function)
static __device__ __noinline__ real tracer_cell_generic1(const DevGrid* __restrict__ gp, const real* __restrict__ u,
                                                          const real* __restrict__ v, const real* __restrict__ w,
                                                          const real* __restrict__ T, int i, int j, int k, real* ftop) {
  const DevGrid& g = *gp;
  const int q2 = id2(g, i, j), PX = g.PX, n2 = g.n2;
  const size_t q3 = q2 + (size_t)n2 * (k + g.Hz - 1);
  const real eps = g.eps;
  const bool clear = (k - 1) > (int)g.knear[q2];
  const bool imm = g.immersed && !clear;
  const real dz = g.dzc[k + g.Hz - 1];
  const real* Tc = T + q3;
  real fT[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const real vel = u[q3 + e];
    const int B = clear ? 3 : buf_from(g.fx3, g.fx2, q2 + e, k);
    const bool masked = imm && (k <= (int)g.kb[q2 + e] || k <= (int)g.kb[q2 + e - 1]);
    fT[e] = masked ? R(0.) : g.dyfc[q2 + e] * dz * vel * recon_mem(Tc + e, 1, B, vel > R(0.), eps);
  }
  real dT = fT[1] - fT[0];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const real vel = v[q3 + e * PX];
    const int B = clear ? 3 : buf_from(g.fy3, g.fy2, q2 + e * PX, k);
    const bool wall = y_outside(g, j + e) || y_outside(g, j + e - 1);
    const bool masked = imm && !wall && (k <= (int)g.kb[q2 + e * PX] || k <= (int)g.kb[q2 + (e - 1) * PX]);
    fT[e] = masked ? R(0.) : g.dxcf[q2 + e * PX] * dz * vel * recon_mem(Tc + e * PX, PX, B, vel > R(0.), eps);
  }
  dT += fT[1] - fT[0];
  const int kbc = g.kb[q2];
  const real az = g.azcc[q2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int kk = k + e;
    const real vel = w[q3 + (size_t)e * n2];
    const int B = zbuf(g, kbc, kk, 3);
    const bool masked = g.immersed && kk != 1 && kk != g.Nz + 1 && (kk - 1 <= kbc);
    fT[e] = masked ? R(0.) : az * vel * recon_mem(Tc + e * n2, n2, B, vel > R(0.), eps);
  }
  dT += fT[1] - fT[0];
  *ftop = fT[1];
  return -((R(1.) / (az * dz)) * dT);
}

// =====================================================================================
// Momentum tendencies (row A5): WENO5 vector-invariant, self-upwinding,
// enstrophy-conserving Coriolis, hydrostatic pressure gradient.  One device function for both
// components: DIR = 0 computes Gu, DIR = 1 computes Gv with the roles of (x,u) and (y,v) swapped;
// the swapped vorticity is -zeta and WENO reconstruction is odd, so the result is identical.
// (a, b) below are offsets along the component's own / cross horizontal direction.
// =====================================================================================
// SCR: the vorticity and the two flux differences are read from the scratch arrays of k_aux_columns (same expressions,
// conditional differences included; the swapped vorticity of DIR = 1 is -zeta) instead of being rebuilt from u, v and six
// metric arrays at each of the 18 stencil points.
template <int DIR, bool SCR = false>
__device__ __forceinline__ real momentum_G(const DevGrid& g, const real* __restrict__ own, const real* __restrict__ oth,
                                            const real* __restrict__ w, const real* __restrict__ p, int i, int j, int k,
                                            real* wtop = nullptr, const real* __restrict__ zeta = nullptr,
                                            const real* __restrict__ dxU = nullptr, const real* __restrict__ dyV = nullptr) {
  const int PX = g.PX, n2 = g.n2;
  const int sO = DIR == 0 ? 1 : PX, sC = DIR == 0 ? PX : 1;
  const real* __restrict__ M1 = DIR == 0 ? g.dxfc : g.dycf;  // own-direction spacing at the own-velocity point
  const real* __restrict__ M2 = DIR == 0 ? g.dxcf : g.dyfc;  // own-direction spacing at the other-velocity point
  const real* __restrict__ M3 = DIR == 0 ? g.dyfc : g.dxcf;  // cross spacing at the own-velocity point
  const real* __restrict__ M4 = DIR == 0 ? g.dycf : g.dxfc;  // cross spacing at the other-velocity point
  const real* __restrict__ AZo = DIR == 0 ? g.azfc : g.azcf;
  const short* fO3 = DIR == 0 ? g.fx3 : g.fy3; const short* fO2 = DIR == 0 ? g.fx2 : g.fy2;
  const short* cC3 = DIR == 0 ? g.cy3 : g.cx3; const short* cC2 = DIR == 0 ? g.cy2 : g.cx2;
  const int q2 = id2(g, i, j);
  const size_t q3 = q2 + (size_t)n2 * (k + g.Hz - 1);
  const real* O = own + q3; const real* X = oth + q3;
  const real eps = g.eps;
  const real dz = g.dzc[k + g.Hz - 1];
  const bool clear = (k - 1) > (int)g.knear[q2];
  const bool imm = g.immersed && !clear;
#define OF(a, b) ((a) * sO + (b) * sC)
  // inactive cell at offset (a,b) and level kk
  auto inact = [&](int a, int b, int kk) -> bool {
    const int jj = DIR == 0 ? j + b : j + a;
    return kk < 1 || kk > g.Nz || y_outside(g, jj) || kk <= (int)g.kb[q2 + OF(a, b)];
  };
  const real own0 = O[0];
  const real m1 = M1[q2];

  // ---------------- horizontal: -(other-hat) * zeta^R  (vorticity flux)
  const real x00 = M2[q2 + OF(0, 0)] * X[OF(0, 0)], x01 = M2[q2 + OF(0, 1)] * X[OF(0, 1)];
  const real xm0 = M2[q2 + OF(-1, 0)] * X[OF(-1, 0)], xm1 = M2[q2 + OF(-1, 1)] * X[OF(-1, 1)];
  const real oavg = ((xm0 + xm1) * R(0.5) + (x00 + x01) * R(0.5)) * R(0.5);
  const real ohat = oavg / m1;
  real zq[6], zs[6], zr[6];
#pragma unroll
  for (int m = 0; m < 6; m++) {
    const int b = m - 2;
    if (SCR) {
      const real z = zeta[q3 + OF(0, b)];
      zq[m] = DIR == 0 ? z : -z;
    } else {
      real d1 = M4[q2 + OF(0, b)] * X[OF(0, b)] - M4[q2 + OF(-1, b)] * X[OF(-1, b)];
      real d2 = M1[q2 + OF(0, b)] * O[OF(0, b)] - M1[q2 + OF(0, b - 1)] * O[OF(0, b - 1)];
      if (imm && g.cond_diff) {
        const bool c00 = inact(0, b, k), c0m = inact(0, b - 1, k), cm0 = inact(-1, b, k), cmm = inact(-1, b - 1, k);
        if ((c00 && c0m) || (cm0 && cmm)) d1 = R(0.);   // inactive other-velocity nodes
        if ((c00 && cm0) || (c0m && cmm)) d2 = R(0.);   // inactive own-velocity nodes
      }
      zq[m] = (d1 - d2) / g.azff[q2 + OF(0, b)];
    }
    zs[m] = (O[OF(0, b - 1)] + O[OF(0, b)]) * R(0.5);
    zr[m] = (X[OF(-1, b)] + X[OF(0, b)]) * R(0.5);
  }
  const int Bc = clear ? 3 : buf_from(cC3, cC2, q2, k);
  const real zR = recon_w_vs(zq, zs, zr, Bc, ohat > R(0.), eps);
  const real Hterm = -ohat * zR;

  // ---------------- divergence flux (self-upwinding) and kinetic-energy gradient along the own direction
  real dOw[6], dv[6], dK[6], sK[6], dOt[6];
#pragma unroll
  for (int m = 0; m < 6; m++) {
    const int a = m - 3;
    const real o0 = O[OF(a, 0)], o1 = O[OF(a + 1, 0)];
    if (SCR) {
      const real fx = dxU[q3 + OF(a, 0)], fy = dyV[q3 + OF(a, 0)];
      dOw[m] = DIR == 0 ? fx : fy;
      dOt[m] = DIR == 0 ? fy : fx;
    } else {
      dOw[m] = M3[q2 + OF(a + 1, 0)] * dz * o1 - M3[q2 + OF(a, 0)] * dz * o0;
      dOt[m] = M2[q2 + OF(a, 1)] * dz * X[OF(a, 1)] - M2[q2 + OF(a, 0)] * dz * X[OF(a, 0)];
    }
    dv[m] = DIR == 0 ? dOw[m] + dOt[m] : dOt[m] + dOw[m];
    dK[m] = o1 * o1 * R(0.5) - o0 * o0 * R(0.5);
    sK[m] = (o0 + o1) * R(0.5);
  }
  const int Bf = clear ? 3 : buf_from(fO3, fO2, q2, k);
  const int Bs = clear ? 2 : (k > (int)fO2[q2] ? 2 : 1);
  const bool lown = own0 > R(0.);
  const real dvs = sym4(dOt[1], dOt[2], dOt[3], dOt[4], Bs);
  const real duR = recon_w_fs(dOw, dv, Bf, lown, eps);
  const real Phi = own0 * (dvs + duR);
  const real dKo = recon_w_fs(dK, sK, Bf, lown, eps);
  // cross kinetic-energy gradient, centred along the cross direction
  const int Bsc = clear ? 2 : (k > (int)cC2[q2] ? 2 : 1);
  real kc[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const int b = m - 1;
    const real t0 = X[OF(0, b)], t1 = X[OF(-1, b)];
    kc[m] = t0 * t0 * R(0.5) - t1 * t1 * R(0.5);
  }
  const real dKc = sym4(kc[0], kc[1], kc[2], kc[3], Bsc);
  const real Bterm = (dKo + dKc) / m1;

  // ---------------- vertical advection: delta_z ( w~ * own^R )
  const int kb0 = g.kb[q2], kbm = g.kb[q2 + OF(-1, 0)];
  real Wf[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int kk = k + e;
    bool masked = false;
    if (g.immersed && kk != 1 && kk != g.Nz + 1) {
      bool wall = false;
      if (DIR == 1) wall = y_outside(g, j) || y_outside(g, j - 1);
      masked = !wall && (kk - 1 <= kb0 || kk - 1 <= kbm);
    }
    const real* wl = w + q3 + (size_t)e * n2;
    const int Bw = (g.immersed && kk > g.Nz) ? 1 : (kk > (int)fO2[q2] ? 2 : 1);
    const real wt = sym4(g.azcc[q2 + OF(-2, 0)] * wl[OF(-2, 0)], g.azcc[q2 + OF(-1, 0)] * wl[OF(-1, 0)],
                          g.azcc[q2 + OF(0, 0)] * wl[OF(0, 0)], g.azcc[q2 + OF(1, 0)] * wl[OF(1, 0)], Bw);
    const int Bz = zbuf(g, kb0, kk, 3);
    const real oR = recon_mem(O + (size_t)e * n2, n2, Bz, wt > R(0.), eps);
    Wf[e] = masked ? R(0.) : wt * oR;
  }
  if (wtop) *wtop = Wf[1];   // (masked) vertical momentum flux through the top face, for k-marching callers
  const real Vterm = (R(1.) / (AZo[q2] * dz)) * (Phi + (Wf[1] - Wf[0]));

  // ---------------- Coriolis (enstrophy conserving, optionally active-cell weighted)
  const real fbar = (g.fff[q2] + g.fff[q2 + OF(0, 1)]) * R(0.5);
  real avg = oavg;
  if (g.coriolis_scheme == 1 && !clear) {
    // other-velocity nodes are Faces along the cross direction: peripheral = cell (a,b) or (a,b-1) inactive
    int nact = 0;
#pragma unroll
    for (int a = -1; a <= 0; a++)
#pragma unroll
      for (int b = 0; b <= 1; b++) nact += !(inact(a, b, k) || inact(a, b - 1, k));
    avg = nact == 0 ? R(0.) : oavg / ((real)nact * R(0.25));
  }
  const real ct = fbar * avg / m1;
  const real cor = DIR == 0 ? -ct : ct;
  // ---------------- hydrostatic pressure gradient
  const real* pc = p + q3;
  real dp = (pc[0] - pc[OF(-1, 0)]) / m1;
  if (imm && g.cond_diff && (inact(0, 0, k) || inact(-1, 0, k))) dp = R(0.);
#undef OF
  return -(Hterm + Vterm + Bterm) - cor - dp;
}

