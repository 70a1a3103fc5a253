// gb25_kernels.cu — first-generation (operator-per-kernel) sm_100a kernels of libgb25cuda.
// One kernel per stage of the Oceananigans HydrostaticFreeSurfaceModel step, in the order of
// /root/reference/src/precompile.jl:31-42.  All kernels are HBM/FP32-issue bound stencil or scan
// work: x is the coalesced thread dimension everywhere, column scans run one thread per column.
#include "gb25_internal.h"

// =====================================================================================
// Halo fills (row A2).  Bit-exact contract: copies and sign flips only.
// =====================================================================================
struct HaloField { float* a; int lx, ly, lz; float sign; };
struct HaloBatch { HaloField f[4]; int n; };

// south/north of 3-D or 2-D fields: threads over (i, k, field); loop over the halo depth
__global__ void k_halo_south_north(DevGrid g, HaloBatch hb, int three_d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int fidx = blockIdx.z;
  if (i > g.Nx) return;
  const HaloField hf = hb.f[fidx];
  const int nk = three_d ? g.Nz + hf.lz : 1;
  const int k = blockIdx.y + 1;
  if (k > nk) return;
  float* a = hf.a + (three_d ? (size_t)g.n2 * (k + g.Hz - 1) : 0);
  const int PX = g.PX, Hx = g.Hx, Hy = g.Hy, Ny = g.Ny, Nx = g.Nx;
  const int I = i + Hx - 1;
#define A2(ii, jj) a[(ii) + PX * ((jj) + Hy - 1)]
  if (hf.ly == 0) { for (int m = 1; m <= Hy; m++) A2(I, 1 - m) = A2(I, m); }
  else A2(I, 1) = 0.f;
  if (g.topo_y == 0) {
    if (hf.ly == 0) { for (int m = 1; m <= Hy; m++) A2(I, Ny + m) = A2(I, Ny + 1 - m); }
    else A2(I, Ny + 1) = 0.f;
  } else {
    int ip; float sg = hf.sign;
    if (hf.lx == 0) ip = Nx - i + 1;
    else { ip = Nx - i + 2; if (ip > Nx) { ip -= Nx; sg = fabsf(sg); } }
    const int IP = ip + Hx - 1;
    for (int m = 1; m <= Hy; m++) {
      const int js = hf.ly == 0 ? Ny - m : Ny - m + 1;
      A2(I, Ny + m) = sg * A2(IP, js);
    }
  }
#undef A2
}
// fold variant 1: overwrite the redundant half of row Ny (Centre-y fields); separate launch because it
// reads and writes the same row (i > Nx/2 reads i' <= Nx/2: disjoint halves, race-free)
__global__ void k_halo_fold_row(DevGrid g, HaloBatch hb, int three_d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1 + g.Nx / 2;
  if (i > g.Nx) return;
  const HaloField hf = hb.f[blockIdx.z];
  if (hf.ly != 0) return;
  const int nk = three_d ? g.Nz + hf.lz : 1;
  const int k = blockIdx.y + 1;
  if (k > nk) return;
  float* a = hf.a + (three_d ? (size_t)g.n2 * (k + g.Hz - 1) : 0);
  int ip; float sg = hf.sign;
  if (hf.lx == 0) ip = g.Nx - i + 1;
  else { ip = g.Nx - i + 2; if (ip > g.Nx) { ip -= g.Nx; sg = fabsf(sg); } }
  const int row = g.PX * (g.Ny + g.Hy - 1);
  a[row + i + g.Hx - 1] = sg * a[row + ip + g.Hx - 1];
}
// bottom/top: threads over (i, j, field)
__global__ void k_halo_bottom_top(DevGrid g, HaloBatch hb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y + 1;
  const HaloField hf = hb.f[blockIdx.z];
  const int jt = g.Ny + ((hf.ly && g.topo_y == 0) ? 1 : 0);
  if (i > g.Nx || j > jt) return;
  float* a = hf.a + id2(g, i, j);
  const size_t n2 = g.n2; const int Hz = g.Hz, Nz = g.Nz;
#define AK(kk) a[n2 * (size_t)((kk) + Hz - 1)]
  if (hf.lz == 0) {
    for (int m = 1; m <= Hz; m++) { AK(1 - m) = AK(m); AK(Nz + m) = AK(Nz + 1 - m); }
  } else { AK(1) = 0.f; AK(Nz + 1) = 0.f; }
#undef AK
}
// periodic x over the full parent extent in (j,k): threads x = 2*Hx halo cells
__global__ void k_halo_periodic_x(DevGrid g, HaloBatch hb, int three_d) {
  const int t = threadIdx.x;                      // 0 .. 2Hx-1
  const int J = blockIdx.x * blockDim.y + threadIdx.y;  // storage row
  const int K = blockIdx.y;                       // storage plane
  if (t >= 2 * g.Hx || J >= g.PY) return;
  float* a = hb.f[blockIdx.z].a + (three_d ? (size_t)g.n2 * K : 0) + (size_t)g.PX * J;
  // west halo cell I = t (t < Hx)  <- I + Nx ; east halo cell I = Nx + t (t >= Hx) <- I - Nx
  if (t < g.Hx) a[t] = a[t + g.Nx];
  else a[g.Nx + t] = a[t];
}

void launch_fill_halo(Handle* h, const HaloSpec* specs, int n, bool three_d) {
  const DevGrid& g = h->g;
  HaloBatch hb; hb.n = n;
  int maxlz = 0;
  for (int q = 0; q < n; q++) { hb.f[q] = HaloField{specs[q].a, specs[q].lx, specs[q].ly, specs[q].lz, specs[q].sign}; maxlz = max(maxlz, specs[q].lz); }
  const int nk = three_d ? g.Nz + maxlz : 1;
  dim3 b1(128), g1((g.Nx + 127) / 128, nk, n);
  k_halo_south_north<<<g1, b1, 0, h->stream>>>(g, hb, three_d); h->count_launch();
  if (g.topo_y == 1 && g.fold_variant == 1) {
    dim3 gf((g.Nx / 2 + 127) / 128, nk, n);
    k_halo_fold_row<<<gf, b1, 0, h->stream>>>(g, hb, three_d); h->count_launch();
  }
  if (three_d) {
    dim3 g2((g.Nx + 127) / 128, g.Ny + 1, n);
    k_halo_bottom_top<<<g2, b1, 0, h->stream>>>(g, hb); h->count_launch();
  }
  dim3 b3(2 * g.Hx, 16), g3((g.PY + 15) / 16, three_d ? g.PZ : 1, n);
  k_halo_periodic_x<<<g3, b3, 0, h->stream>>>(g, hb, three_d); h->count_launch();
}

// =====================================================================================
// mask_immersed_field! (row A1)
// =====================================================================================
__global__ void k_mask_fields(DevGrid g, float* u, float* v, float* T, float* S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j);
  const size_t q3 = q2 + (size_t)g.n2 * (k + g.Hz - 1);
  const bool c0 = inactive_cell(g, i, j, k);
  if (u && (c0 || inactive_cell(g, i - 1, j, k))) u[q3] = 0.f;
  if (v && (c0 || inactive_cell(g, i, j - 1, k))) v[q3] = 0.f;
  if (T && c0) T[q3] = 0.f;
  if (S && c0) S[q3] = 0.f;
}
// barotropic transports (decision U11): masked where the surface-level velocity node is peripheral
__global__ void k_mask_barotropic(DevGrid g, float* U, float* V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j);
  const bool c0 = inactive_cell(g, i, j, g.Nz);
  if (c0 || inactive_cell(g, i - 1, j, g.Nz)) U[q2] = 0.f;
  if (c0 || inactive_cell(g, i, j - 1, g.Nz)) V[q2] = 0.f;
}
void launch_mask(Handle* h, bool uv_only) {
  if (!h->g.immersed) return;
  const DevGrid& g = h->g;
  dim3 b(128), gr((g.Nx + 127) / 128, g.Ny, g.Nz);
  k_mask_fields<<<gr, b, 0, h->stream>>>(g, h->f.u, h->f.v, uv_only ? nullptr : h->f.T, uv_only ? nullptr : h->f.S);
  h->count_launch();
  if (!uv_only) {
    dim3 g2((g.Nx + 127) / 128, g.Ny);
    k_mask_barotropic<<<g2, b, 0, h->stream>>>(g, h->f.bu, h->f.bv);
    h->count_launch();
  }
}

// =====================================================================================
// compute_w_from_continuity! (row A3): one thread per column of the extended range, upward scan
// =====================================================================================
__global__ void k_compute_w(DevGrid g, const float* __restrict__ u, const float* __restrict__ v, float* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + (-g.Hx + 2);
  const int j = blockIdx.y + (-g.Hy + 2);
  if (i > g.Nx + g.Hx - 1) return;
  const int q2 = id2(g, i, j), PX = g.PX;
  const float dyE = g.dyfc[q2 + 1], dyW = g.dyfc[q2], dxN = g.dxcf[q2 + PX], dxS = g.dxcf[q2];
  const float az = g.azcc[q2];
  size_t q3 = q2 + (size_t)g.n2 * g.Hz;  // k = 1
  float wk = 0.f;
  w[q3] = 0.f;
  for (int k = 2; k <= g.Nz + 1; k++) {
    const float dz = g.dzc[k - 1 + g.Hz - 1];
    const float dU = dyE * dz * u[q3 + 1] - dyW * dz * u[q3];
    const float dV = dxN * dz * v[q3 + PX] - dxS * dz * v[q3];
    wk = wk - (dU + dV) / az;
    q3 += g.n2;
    w[q3] = wk;
  }
}
// =====================================================================================
// update_hydrostatic_pressure! (row A4): one thread per column, downward scan, one EOS call per cell
// =====================================================================================
__global__ void k_compute_p(DevGrid g, const float* __restrict__ T, const float* __restrict__ S, float* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. Nx+1
  const int j = blockIdx.y;                             // 0 .. Ny+1
  if (i > g.Nx + 1) return;
  const int q2 = id2(g, i, j);
  size_t q3 = q2 + (size_t)g.n2 * (g.Nz + 1 + g.Hz - 1);
  const float gr = g.g, r0 = g.rho0;
  float bup = -(gr * teos10_rho_prime(T[q3], S[q3], g.zc[g.Nz + 1 + g.Hz - 1], r0, g.eos_r0) / r0);
  float pk = 0.f;
  for (int k = g.Nz; k >= 1; k--) {
    q3 -= g.n2;
    const float b = -(gr * teos10_rho_prime(T[q3], S[q3], g.zc[k + g.Hz - 1], r0, g.eos_r0) / r0);
    const float bbar = (b + bup) * 0.5f;
    const float dzf = g.dzf[k + 1 + g.Hz - 1];
    pk = (k == g.Nz) ? -bbar * dzf : pk - bbar * dzf;
    p[q3] = pk;
    bup = b;
  }
}
void launch_compute_w(Handle* h) {
  const DevGrid& g = h->g;
  const int nx = g.Nx + 2 * g.Hx - 2, ny = g.Ny + 2 * g.Hy - 2;
  dim3 b(128), gr((nx + 127) / 128, ny);
  k_compute_w<<<gr, b, 0, h->stream>>>(g, h->f.u, h->f.v, h->f.w); h->count_launch();
}
void launch_compute_p(Handle* h) {
  const DevGrid& g = h->g;
  dim3 b(128), gr((g.Nx + 2 + 127) / 128, g.Ny + 2);
  k_compute_p<<<gr, b, 0, h->stream>>>(g, h->f.T, h->f.S, h->f.p); h->count_launch();
}

// =====================================================================================
// Tracer tendencies (row A6): Gc = -div(U c), WENO5-Z upwind flux form, T and S together
// =====================================================================================
__global__ void __launch_bounds__(256) k_tracer_tendency(DevGrid g, const float* __restrict__ u, const float* __restrict__ v,
                                                          const float* __restrict__ w, const float* __restrict__ T,
                                                          const float* __restrict__ S, float* __restrict__ GT, float* __restrict__ GS) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.Nx || j > g.Ny) return;
  const int q2 = id2(g, i, j), PX = g.PX, n2 = g.n2;
  const size_t q3 = q2 + (size_t)n2 * (k + g.Hz - 1);
  const float eps = g.eps;
  const bool clear = (k - 1) > (int)g.knear[q2];
  const bool imm = g.immersed && !clear;
  const float dz = g.dzc[k + g.Hz - 1];
  const float* Tc = T + q3; const float* Sc = S + q3;
  float dT = 0.f, dS = 0.f;
  // ---- x faces i (e=0) and i+1 (e=1)
  float fT[2], fS[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const float vel = u[q3 + e];
    const float area = g.dyfc[q2 + e] * dz;
    const int B = clear ? 3 : buf_from(g.fx3, g.fx2, q2 + e, k);
    const bool masked = imm && (k <= (int)g.kb[q2 + e] || k <= (int)g.kb[q2 + e - 1]);
    const bool left = vel > 0.f;
    fT[e] = masked ? 0.f : area * vel * recon_mem(Tc + e, 1, B, left, eps);
    fS[e] = masked ? 0.f : area * vel * recon_mem(Sc + e, 1, B, left, eps);
  }
  dT = fT[1] - fT[0]; dS = fS[1] - fS[0];
  // ---- y faces j and j+1
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const float vel = v[q3 + e * PX];
    const float area = g.dxcf[q2 + e * PX] * dz;
    const int B = clear ? 3 : buf_from(g.fy3, g.fy2, q2 + e * PX, k);
    const bool wall = y_outside(g, j + e) || y_outside(g, j + e - 1);
    const bool masked = imm && !wall && (k <= (int)g.kb[q2 + e * PX] || k <= (int)g.kb[q2 + (e - 1) * PX]);
    const bool left = vel > 0.f;
    fT[e] = masked ? 0.f : area * vel * recon_mem(Tc + e * PX, PX, B, left, eps);
    fS[e] = masked ? 0.f : area * vel * recon_mem(Sc + e * PX, PX, B, left, eps);
  }
  dT += fT[1] - fT[0]; dS += fS[1] - fS[0];
  // ---- z faces k and k+1
  const int kbc = g.kb[q2];
  const float az = g.azcc[q2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int kk = k + e;
    const float vel = w[q3 + (size_t)e * n2];
    const int B = zbuf(g, kbc, kk, 3);
    const bool masked = g.immersed && kk != 1 && kk != g.Nz + 1 && (kk - 1 <= kbc);
    const bool left = vel > 0.f;
    fT[e] = masked ? 0.f : az * vel * recon_mem(Tc + e * n2, n2, B, left, eps);
    fS[e] = masked ? 0.f : az * vel * recon_mem(Sc + e * n2, n2, B, left, eps);
  }
  dT += fT[1] - fT[0]; dS += fS[1] - fS[0];
  const float rV = 1.f / (az * dz);
  GT[q3] = -(rV * dT);
  GS[q3] = -(rV * dS);
}
void launch_tracer_tendency(Handle* h) {
  const DevGrid& g = h->g;
  dim3 b(64, 4), gr((g.Nx + 63) / 64, (g.Ny + 3) / 4, g.Nz);
  k_tracer_tendency<<<gr, b, 0, h->stream>>>(g, h->f.u, h->f.v, h->f.w, h->f.T, h->f.S, h->f.gn[2], h->f.gn[3]);
  h->count_launch();
}

// =====================================================================================
// Momentum tendencies (row A5): WENO5 vector-invariant, self-upwinding,
// enstrophy-conserving Coriolis, hydrostatic pressure gradient.  One device function for both
// components: DIR = 0 computes Gu, DIR = 1 computes Gv with the roles of (x,u) and (y,v) swapped;
// the swapped vorticity is -zeta and WENO reconstruction is odd, so the result is identical.
// (a, b) below are offsets along the component's own / cross horizontal direction.
// =====================================================================================
template <int DIR>
__device__ __forceinline__ float momentum_G(const DevGrid& g, const float* __restrict__ own, const float* __restrict__ oth,
                                            const float* __restrict__ w, const float* __restrict__ p, int i, int j, int k) {
  const int PX = g.PX, n2 = g.n2;
  const int sO = DIR == 0 ? 1 : PX, sC = DIR == 0 ? PX : 1;
  const float* __restrict__ M1 = DIR == 0 ? g.dxfc : g.dycf;  // own-direction spacing at the own-velocity point
  const float* __restrict__ M2 = DIR == 0 ? g.dxcf : g.dyfc;  // own-direction spacing at the other-velocity point
  const float* __restrict__ M3 = DIR == 0 ? g.dyfc : g.dxcf;  // cross spacing at the own-velocity point
  const float* __restrict__ M4 = DIR == 0 ? g.dycf : g.dxfc;  // cross spacing at the other-velocity point
  const float* __restrict__ AZo = DIR == 0 ? g.azfc : g.azcf;
  const short* fO3 = DIR == 0 ? g.fx3 : g.fy3; const short* fO2 = DIR == 0 ? g.fx2 : g.fy2;
  const short* cC3 = DIR == 0 ? g.cy3 : g.cx3; const short* cC2 = DIR == 0 ? g.cy2 : g.cx2;
  const int q2 = id2(g, i, j);
  const size_t q3 = q2 + (size_t)n2 * (k + g.Hz - 1);
  const float* O = own + q3; const float* X = oth + q3;
  const float eps = g.eps;
  const float dz = g.dzc[k + g.Hz - 1];
  const bool clear = (k - 1) > (int)g.knear[q2];
  const bool imm = g.immersed && !clear;
#define OF(a, b) ((a) * sO + (b) * sC)
  // inactive cell at offset (a,b) and level kk
  auto inact = [&](int a, int b, int kk) -> bool {
    const int jj = DIR == 0 ? j + b : j + a;
    return kk < 1 || kk > g.Nz || y_outside(g, jj) || kk <= (int)g.kb[q2 + OF(a, b)];
  };
  const float own0 = O[0];
  const float m1 = M1[q2];

  // ---------------- horizontal: -(other-hat) * zeta^R  (vorticity flux)
  const float x00 = M2[q2 + OF(0, 0)] * X[OF(0, 0)], x01 = M2[q2 + OF(0, 1)] * X[OF(0, 1)];
  const float xm0 = M2[q2 + OF(-1, 0)] * X[OF(-1, 0)], xm1 = M2[q2 + OF(-1, 1)] * X[OF(-1, 1)];
  const float oavg = ((xm0 + xm1) * 0.5f + (x00 + x01) * 0.5f) * 0.5f;
  const float ohat = oavg / m1;
  float zq[6], zs[6], zr[6];
#pragma unroll
  for (int m = 0; m < 6; m++) {
    const int b = m - 2;
    float d1 = M4[q2 + OF(0, b)] * X[OF(0, b)] - M4[q2 + OF(-1, b)] * X[OF(-1, b)];
    float d2 = M1[q2 + OF(0, b)] * O[OF(0, b)] - M1[q2 + OF(0, b - 1)] * O[OF(0, b - 1)];
    if (imm && g.cond_diff) {
      const bool c00 = inact(0, b, k), c0m = inact(0, b - 1, k), cm0 = inact(-1, b, k), cmm = inact(-1, b - 1, k);
      if ((c00 && c0m) || (cm0 && cmm)) d1 = 0.f;   // inactive other-velocity nodes
      if ((c00 && cm0) || (c0m && cmm)) d2 = 0.f;   // inactive own-velocity nodes
    }
    zq[m] = (d1 - d2) / g.azff[q2 + OF(0, b)];
    zs[m] = (O[OF(0, b - 1)] + O[OF(0, b)]) * 0.5f;
    zr[m] = (X[OF(-1, b)] + X[OF(0, b)]) * 0.5f;
  }
  const int Bc = clear ? 3 : buf_from(cC3, cC2, q2, k);
  const float zR = recon_w_vs(zq, zs, zr, Bc, ohat > 0.f, eps);
  const float Hterm = -ohat * zR;

  // ---------------- divergence flux (self-upwinding) and kinetic-energy gradient along the own direction
  float dOw[6], dv[6], dK[6], sK[6], dOt[6];
#pragma unroll
  for (int m = 0; m < 6; m++) {
    const int a = m - 3;
    const float o0 = O[OF(a, 0)], o1 = O[OF(a + 1, 0)];
    dOw[m] = M3[q2 + OF(a + 1, 0)] * dz * o1 - M3[q2 + OF(a, 0)] * dz * o0;
    dOt[m] = M2[q2 + OF(a, 1)] * dz * X[OF(a, 1)] - M2[q2 + OF(a, 0)] * dz * X[OF(a, 0)];
    dv[m] = DIR == 0 ? dOw[m] + dOt[m] : dOt[m] + dOw[m];
    dK[m] = o1 * o1 * 0.5f - o0 * o0 * 0.5f;
    sK[m] = (o0 + o1) * 0.5f;
  }
  const int Bf = clear ? 3 : buf_from(fO3, fO2, q2, k);
  const int Bs = clear ? 2 : (k > (int)fO2[q2] ? 2 : 1);
  const bool lown = own0 > 0.f;
  const float dvs = sym4(dOt[1], dOt[2], dOt[3], dOt[4], Bs);
  const float duR = recon_w_fs(dOw, dv, Bf, lown, eps);
  const float Phi = own0 * (dvs + duR);
  const float dKo = recon_w_fs(dK, sK, Bf, lown, eps);
  // cross kinetic-energy gradient, centred along the cross direction
  const int Bsc = clear ? 2 : (k > (int)cC2[q2] ? 2 : 1);
  float kc[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const int b = m - 1;
    const float t0 = X[OF(0, b)], t1 = X[OF(-1, b)];
    kc[m] = t0 * t0 * 0.5f - t1 * t1 * 0.5f;
  }
  const float dKc = sym4(kc[0], kc[1], kc[2], kc[3], Bsc);
  const float Bterm = (dKo + dKc) / m1;

  // ---------------- vertical advection: delta_z ( w~ * own^R )
  const int kb0 = g.kb[q2], kbm = g.kb[q2 + OF(-1, 0)];
  float Wf[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int kk = k + e;
    bool masked = false;
    if (g.immersed && kk != 1 && kk != g.Nz + 1) {
      bool wall = false;
      if (DIR == 1) wall = y_outside(g, j) || y_outside(g, j - 1);
      masked = !wall && (kk - 1 <= kb0 || kk - 1 <= kbm);
    }
    const float* wl = w + q3 + (size_t)e * n2;
    const int Bw = (g.immersed && kk > g.Nz) ? 1 : (kk > (int)fO2[q2] ? 2 : 1);
    const float wt = sym4(g.azcc[q2 + OF(-2, 0)] * wl[OF(-2, 0)], g.azcc[q2 + OF(-1, 0)] * wl[OF(-1, 0)],
                          g.azcc[q2 + OF(0, 0)] * wl[OF(0, 0)], g.azcc[q2 + OF(1, 0)] * wl[OF(1, 0)], Bw);
    const int Bz = zbuf(g, kb0, kk, 3);
    const float oR = recon_mem(O + (size_t)e * n2, n2, Bz, wt > 0.f, eps);
    Wf[e] = masked ? 0.f : wt * oR;
  }
  const float Vterm = (1.f / (AZo[q2] * dz)) * (Phi + (Wf[1] - Wf[0]));

  // ---------------- Coriolis (enstrophy conserving, optionally active-cell weighted)
  const float fbar = (g.fff[q2] + g.fff[q2 + OF(0, 1)]) * 0.5f;
  float avg = oavg;
  if (g.coriolis_scheme == 1 && !clear) {
    // other-velocity nodes are Faces along the cross direction: peripheral = cell (a,b) or (a,b-1) inactive
    int nact = 0;
#pragma unroll
    for (int a = -1; a <= 0; a++)
#pragma unroll
      for (int b = 0; b <= 1; b++) nact += !(inact(a, b, k) || inact(a, b - 1, k));
    avg = nact == 0 ? 0.f : oavg / ((float)nact * 0.25f);
  }
  const float ct = fbar * avg / m1;
  const float cor = DIR == 0 ? -ct : ct;
  // ---------------- hydrostatic pressure gradient
  const float* pc = p + q3;
  float dp = (pc[0] - pc[OF(-1, 0)]) / m1;
  if (imm && g.cond_diff && (inact(0, 0, k) || inact(-1, 0, k))) dp = 0.f;
#undef OF
  return -(Hterm + Vterm + Bterm) - cor - dp;
}

template <int DIR>
__global__ void __launch_bounds__(256) k_momentum_tendency(DevGrid g, const float* __restrict__ u, const float* __restrict__ v,
                                                            const float* __restrict__ w, const float* __restrict__ p,
                                                            float* __restrict__ G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.Nx || j > g.Ny) return;
  const size_t q3 = id2(g, i, j) + (size_t)g.n2 * (k + g.Hz - 1);
  G[q3] = DIR == 0 ? momentum_G<0>(g, u, v, w, p, i, j, k) : momentum_G<1>(g, v, u, w, p, i, j, k);
}
void launch_momentum_tendency(Handle* h) {
  const DevGrid& g = h->g;
  dim3 b(64, 4), gr((g.Nx + 63) / 64, (g.Ny + 3) / 4, g.Nz);
  k_momentum_tendency<0><<<gr, b, 0, h->stream>>>(g, h->f.u, h->f.v, h->f.w, h->f.p, h->f.gn[0]); h->count_launch();
  k_momentum_tendency<1><<<gr, b, 0, h->stream>>>(g, h->f.u, h->f.v, h->f.w, h->f.p, h->f.gn[1]); h->count_launch();
}

// =====================================================================================
// ab2_step! part 1 (rows A8 + A9): barotropic forcing column integral fused with the AB2 update
// =====================================================================================
__global__ void k_ab2_columns(DevGrid g, DevFields f, float dt, float chi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j), n2 = g.n2;
  const float c1 = 1.5f + chi, c2 = 0.5f + chi;
  const float ne = (chi != -0.5f) ? 1.f : 0.f;
  const int kb0 = g.kb[q2], kbw = g.kb[q2 - 1], kbs = g.kb[q2 - g.PX];
  const bool ywall = y_outside(g, j) || y_outside(g, j - 1);
  float su = 0.f, sv = 0.f;
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {
    const float dz = g.dzc[k + g.Hz - 1];
    const float gu = c1 * f.gn[0][q3] - c2 * f.gm[0][q3] * ne;
    const float gv = c1 * f.gn[1][q3] - c2 * f.gm[1][q3] * ne;
    const bool pu = k <= kb0 || k <= kbw;
    const bool pv = ywall || k <= kb0 || k <= kbs;
    const float tu = dz * (pu ? 0.f : gu), tv = dz * (pv ? 0.f : gv);
    su = (k == 1) ? tu : su + tu;
    sv = (k == 1) ? tv : sv + tv;
    f.u[q3] += dt * gu;
    f.v[q3] += dt * gv;
    f.T[q3] = f.T[q3] + dt * (c1 * f.gn[2][q3] - c2 * f.gm[2][q3]);
    f.S[q3] = f.S[q3] + dt * (c1 * f.gn[3][q3] - c2 * f.gm[3][q3]);
  }
  f.gU[q2] = su; f.gV[q2] = sv;
}
void launch_ab2_columns(Handle* h, float dt, float chi) {
  const DevGrid& g = h->g;
  dim3 b(128), gr((g.Nx + 127) / 128, g.Ny);
  k_ab2_columns<<<gr, b, 0, h->stream>>>(g, h->f, dt, chi); h->count_launch();
}

// =====================================================================================
// Split-explicit substeps (row A10): forward-backward, topology-aware differences
// =====================================================================================
__global__ void k_baro_eta(DevGrid g, float* __restrict__ eta, const float* __restrict__ U, const float* __restrict__ V, float dtau) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j), PX = g.PX;
  const int qe = (i == g.Nx) ? id2(g, 1, j) : q2 + 1;
  const float dU = g.dyfc[qe] * U[qe] - g.dyfc[q2] * U[q2];
  float dV;
  if (j == 1) dV = g.dxcf[q2 + PX] * V[q2 + PX];
  else if (j == g.Ny) {
    if (g.topo_y == 0) dV = -(g.dxcf[q2] * V[q2]);
    else {  // folded row Ny+1: V[i,Ny+1] = -V[Nx-i+1,Ny]
      const float vn = -V[id2(g, g.Nx - i + 1, g.Ny)];
      dV = g.dxcf[q2 + PX] * vn - g.dxcf[q2] * V[q2];
    }
  } else dV = g.dxcf[q2 + PX] * V[q2 + PX] - g.dxcf[q2] * V[q2];
  eta[q2] -= dtau * (dU + dV) / g.azcc[q2];
}
__global__ void k_baro_uv(DevGrid g, DevFields f, float dtau, float wgt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j);
  const int qw = (i == 1) ? id2(g, g.Nx, j) : q2 - 1;
  const float e0 = f.eta[q2];
  const float dxe = (e0 - f.eta[qw]) / g.dxfc[q2];
  const float dye = (j == 1) ? 0.f : (e0 - f.eta[q2 - g.PX]) / g.dycf[q2];
  const float Un = f.bu[q2] + dtau * (-g.g * g.Hfc[q2] * dxe + f.gU[q2]);
  const float Vn = f.bv[q2] + dtau * (-g.g * g.Hcf[q2] * dye + f.gV[q2]);
  f.bu[q2] = Un; f.bv[q2] = Vn;
  f.feta[q2] += wgt * e0;
  f.fu[q2] += wgt * Un;
  f.fv[q2] += wgt * Vn;
}
__global__ void k_baro_finish(DevGrid g, DevFields f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j);
  f.eta[q2] = f.feta[q2]; f.bu[q2] = f.fu[q2]; f.bv[q2] = f.fv[q2];
}
void launch_barotropic(Handle* h, float dt) {
  const DevGrid& g = h->g;
  const size_t b2 = (size_t)g.n2 * sizeof(float);
  cudaMemsetAsync(h->f.feta, 0, b2, h->stream);
  cudaMemsetAsync(h->f.fu, 0, b2, h->stream);
  cudaMemsetAsync(h->f.fv, 0, b2, h->stream);
  const float dtau = h->cfg.dtau_frac * dt;
  dim3 b(128), gr((g.Nx + 127) / 128, g.Ny);
  for (int m = 0; m < h->cfg.nsubsteps; m++) {
    k_baro_eta<<<gr, b, 0, h->stream>>>(g, h->f.eta, h->f.bu, h->f.bv, dtau); h->count_launch();
    k_baro_uv<<<gr, b, 0, h->stream>>>(g, h->f, dtau, h->weights[m]); h->count_launch();
  }
  k_baro_finish<<<gr, b, 0, h->stream>>>(g, h->f); h->count_launch();
}

// =====================================================================================
// Corrector + cache (rows A11, A12)
// =====================================================================================
__global__ void k_correct_cache(DevGrid g, DevFields f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j), n2 = g.n2;
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  float su = 0.f, sv = 0.f;
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {
    const float dz = g.dzc[k + g.Hz - 1];
    const float tu = dz * f.u[q3], tv = dz * f.v[q3];
    su = (k == 1) ? tu : su + tu;
    sv = (k == 1) ? tv : sv + tv;
  }
  f.fu[q2] = su; f.fv[q2] = sv;
  const float cu = (f.bu[q2] - su) / g.Hfc[q2], cv = (f.bv[q2] - sv) / g.Hcf[q2];
  q3 = q2 + (size_t)n2 * g.Hz;
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {
    f.u[q3] = f.u[q3] + cu;
    f.v[q3] = f.v[q3] + cv;
#pragma unroll
    for (int q = 0; q < 4; q++) f.gm[q][q3] = f.gn[q][q3];
  }
  f.gmU[q2] = f.gU[q2]; f.gmV[q2] = f.gV[q2];
}
void launch_correct_cache(Handle* h) {
  const DevGrid& g = h->g;
  dim3 b(128), gr((g.Nx + 127) / 128, g.Ny);
  k_correct_cache<<<gr, b, 0, h->stream>>>(g, h->f); h->count_launch();
}
// barotropic mode only (initialize!)
__global__ void k_barotropic_mode(DevGrid g, DevFields f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j), n2 = g.n2;
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  float su = 0.f, sv = 0.f;
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {
    const float dz = g.dzc[k + g.Hz - 1];
    const float tu = dz * f.u[q3], tv = dz * f.v[q3];
    su = (k == 1) ? tu : su + tu;
    sv = (k == 1) ? tv : sv + tv;
  }
  f.bu[q2] = su; f.bv[q2] = sv;
}
void launch_barotropic_mode(Handle* h) {
  const DevGrid& g = h->g;
  dim3 b(128), gr((g.Nx + 127) / 128, g.Ny);
  k_barotropic_mode<<<gr, b, 0, h->stream>>>(g, h->f); h->count_launch();
}
