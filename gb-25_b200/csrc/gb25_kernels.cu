// gb25_kernels.cu — first-generation (operator-per-kernel) sm_100a kernels of libgb25cuda.
// One kernel per stage of the Oceananigans HydrostaticFreeSurfaceModel step, in the order of
// /root/reference/src/precompile.jl:31-42.  All kernels are HBM/FP32-issue bound stencil or scan
// work: x is the coalesced thread dimension everywhere, column scans run one thread per column.
#include <cstdlib>

#include "gb25_internal.h"
#ifndef GB25_F64
#include "gb25_packed.cuh"
#endif

// =====================================================================================
// Halo fills (row A2).  Bit-exact contract: copies and sign flips only.
// =====================================================================================
struct HaloField { real* a; int lx, ly, lz; real sign; int flat; };
struct HaloBatch { HaloField f[9]; int n; };

// south/north of 3-D or 2-D fields: threads over (i, k, field); loop over the halo depth
// mode_s: 0 = nothing (a neighbour tile fills it), 1 = local wall BC.  mode_n: 0 = nothing, 1 = wall BC, 2 = local fold.
__global__ void k_halo_south_north(DevGrid g, HaloBatch hb, int three_d, int mode_s, int mode_n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int fidx = blockIdx.z;
  if (i > g.Nx) return;
  const HaloField hf = hb.f[fidx];
  if (hf.flat) three_d = 0;
  const int nk = three_d ? g.Nz + hf.lz : 1;
  const int k = blockIdx.y + 1;
  if (k > nk) return;
  real* a = hf.a + (three_d ? (size_t)g.n2 * (k + g.Hz - 1) : 0);
  const int PX = g.PX, Hx = g.Hx, Hy = g.Hy, Ny = g.Ny, Nx = g.Nx;
  const int I = i + Hx - 1;
#define A2(ii, jj) a[(ii) + PX * ((jj) + Hy - 1)]
  // loads first, stores afterwards (the rows never overlap, but the compiler cannot know: a load-store chain would
  // serialise Hy round trips per thread)
  const bool h8 = Hy == 8 && Ny >= 8;
  if (mode_s == 1) {
    if (hf.ly == 0) {
      if (h8) {
        real t[8];
#pragma unroll
        for (int m = 1; m <= 8; m++) t[m - 1] = A2(I, m);
#pragma unroll
        for (int m = 1; m <= 8; m++) A2(I, 1 - m) = t[m - 1];
      } else { for (int m = 1; m <= Hy; m++) A2(I, 1 - m) = A2(I, m); }
    }
    else A2(I, 1) = R(0.);
  }
  if (mode_n == 1) {
    if (hf.ly == 0) {
      if (h8) {
        real t[8];
#pragma unroll
        for (int m = 1; m <= 8; m++) t[m - 1] = A2(I, Ny + 1 - m);
#pragma unroll
        for (int m = 1; m <= 8; m++) A2(I, Ny + m) = t[m - 1];
      } else { for (int m = 1; m <= Hy; m++) A2(I, Ny + m) = A2(I, Ny + 1 - m); }
    }
    else A2(I, Ny + 1) = R(0.);
  } else if (mode_n == 2) {
    int ip; real sg = hf.sign;
    if (hf.lx == 0) ip = Nx - i + 1;
    else { ip = Nx - i + 2; if (ip > Nx) { ip -= Nx; sg = rabs(sg); } }
    const int IP = ip + Hx - 1;
    const int jo = hf.ly == 0 ? 0 : 1;
    if (h8) {
      real t[8];
#pragma unroll
      for (int m = 1; m <= 8; m++) t[m - 1] = A2(IP, Ny - m + jo);
#pragma unroll
      for (int m = 1; m <= 8; m++) A2(I, Ny + m) = sg * t[m - 1];
    } else {
      for (int m = 1; m <= Hy; m++) A2(I, Ny + m) = sg * A2(IP, Ny - m + jo);
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
  if (hf.flat) three_d = 0;
  const int nk = three_d ? g.Nz + hf.lz : 1;
  const int k = blockIdx.y + 1;
  if (k > nk) return;
  real* a = hf.a + (three_d ? (size_t)g.n2 * (k + g.Hz - 1) : 0);
  int ip; real sg = hf.sign;
  if (hf.lx == 0) ip = g.Nx - i + 1;
  else { ip = g.Nx - i + 2; if (ip > g.Nx) { ip -= g.Nx; sg = rabs(sg); } }
  const int row = g.PX * (g.Ny + g.Hy - 1);
  a[row + i + g.Hx - 1] = sg * a[row + ip + g.Hx - 1];
}
// bottom/top: threads over (i, j, field)
__global__ void k_halo_bottom_top(DevGrid g, HaloBatch hb, int row0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y + row0;
  const HaloField hf = hb.f[blockIdx.z];
  const int jt = g.Ny + ((hf.ly && g.wall_n) ? 1 : 0);
  if (i > g.Nx || j > jt || hf.flat) return;
  real* a = hf.a + id2(g, i, j);
  const size_t n2 = g.n2; const int Hz = g.Hz, Nz = g.Nz;
#define AK(kk) a[n2 * (size_t)((kk) + Hz - 1)]
  if (hf.lz == 0) {
    // all loads first, then all stores (source and destination alias as far as the compiler knows: a
    // load-store-load-store chain is 2 Hz dependent round trips to HBM per thread)
    if (Hz == 8 && Nz >= 8) {
      real lo[8], hi[8];
#pragma unroll
      for (int m = 1; m <= 8; m++) { lo[m - 1] = AK(m); hi[m - 1] = AK(Nz + 1 - m); }
#pragma unroll
      for (int m = 1; m <= 8; m++) { AK(1 - m) = lo[m - 1]; AK(Nz + m) = hi[m - 1]; }
    } else {
      for (int m = 1; m <= Hz; m++) { AK(1 - m) = AK(m); AK(Nz + m) = AK(Nz + 1 - m); }
    }
  } else { AK(1) = R(0.); AK(Nz + 1) = R(0.); }
#undef AK
}
// periodic x over the full parent extent in (j,k): threads x = 2*Hx halo cells
__global__ void k_halo_periodic_x(DevGrid g, HaloBatch hb, int three_d) {
  const int t = threadIdx.x;                      // 0 .. 2Hx-1
  const int J = blockIdx.x * blockDim.y + threadIdx.y;  // storage row
  const int K = blockIdx.y;                       // storage plane
  if (t >= 2 * g.Hx || J >= g.PY) return;
  if (hb.f[blockIdx.z].flat) { if (K > 0) return; three_d = 0; }
  real* a = hb.f[blockIdx.z].a + (three_d ? (size_t)g.n2 * K : 0) + (size_t)g.PX * J;
  // west halo cell I = t (t < Hx)  <- I + Nx ; east halo cell I = Nx + t (t >= Hx) <- I - Nx
  if (t < g.Hx) a[t] = a[t + g.Nx];
  else a[g.Nx + t] = a[t];
}

static HaloBatch make_batch(const HaloSpec* specs, int n, int* maxlz) {
  HaloBatch hb; hb.n = n; *maxlz = 0;
  for (int q = 0; q < n; q++) {
    hb.f[q] = HaloField{specs[q].a, specs[q].lx, specs[q].ly, specs[q].lz, specs[q].sign, specs[q].flat};
    if (!specs[q].flat) *maxlz = max(*maxlz, specs[q].lz);
  }
  return hb;
}
void launch_halo_south_north(Handle* h, const HaloSpec* specs, int n, bool three_d, int mode_s, int mode_n) {
  const DevGrid& g = h->g;
  int maxlz; HaloBatch hb = make_batch(specs, n, &maxlz);
  const int nk = three_d ? g.Nz + maxlz : 1;
  dim3 b1(128), g1((g.Nx + 127) / 128, nk, n);
  if (mode_s || mode_n) { k_halo_south_north<<<g1, b1, 0, h->stream>>>(g, hb, three_d, mode_s, mode_n); h->count_launch(); }
  if (mode_n == 2 && g.fold_variant == 1) {
    dim3 gf((g.Nx / 2 + 127) / 128, nk, n);
    k_halo_fold_row<<<gf, b1, 0, h->stream>>>(g, hb, three_d); h->count_launch();
  }
}
void launch_halo_bottom_top(Handle* h, const HaloSpec* specs, int n) {
  const DevGrid& g = h->g;
  // fields whose producer already wrote the z halos drop out; of a Face-y field on a wall tile only the wall row is left
  HaloSpec todo[9], wall[9]; int nt = 0, nw = 0;
  for (int q = 0; q < n; q++) {
    if (specs[q].flat) continue;
    if (!specs[q].zdone) todo[nt++] = specs[q];
    else if (specs[q].zdone == 2 && specs[q].ly && g.wall_n) wall[nw++] = specs[q];
  }
  int maxlz;
  dim3 b1(128);
  if (nt) {
    HaloBatch hb = make_batch(todo, nt, &maxlz);
    dim3 g2((g.Nx + 127) / 128, g.Ny + 1, nt);
    k_halo_bottom_top<<<g2, b1, 0, h->stream>>>(g, hb, 1); h->count_launch();
  }
  if (nw) {
    HaloBatch hb = make_batch(wall, nw, &maxlz);
    dim3 g2((g.Nx + 127) / 128, 1, nw);
    k_halo_bottom_top<<<g2, b1, 0, h->stream>>>(g, hb, g.Ny + 1); h->count_launch();
  }
}
void launch_halo_periodic_x(Handle* h, const HaloSpec* specs, int n, bool three_d) {
  const DevGrid& g = h->g;
  int maxlz; HaloBatch hb = make_batch(specs, n, &maxlz);
  dim3 b3(2 * g.Hx, 16), g3((g.PY + 15) / 16, three_d ? g.PZ : 1, n);
  k_halo_periodic_x<<<g3, b3, 0, h->stream>>>(g, hb, three_d); h->count_launch();
}
void launch_fill_halo(Handle* h, const HaloSpec* specs, int n, bool three_d) {
  if (h->ex.on) { launch_fill_halo_dist(h, specs, n, three_d); return; }
  const DevGrid& g = h->g;
  launch_halo_south_north(h, specs, n, three_d, 1, g.topo_y == 0 ? 1 : 2);
  if (three_d) launch_halo_bottom_top(h, specs, n);
  launch_halo_periodic_x(h, specs, n, three_d);
}

// =====================================================================================
// mask_immersed_field! (row A1)
// =====================================================================================
__global__ void k_mask_fields(DevGrid g, real* u, real* v, real* T, real* S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j);
  const size_t q3 = q2 + (size_t)g.n2 * (k + g.Hz - 1);
  const bool c0 = inactive_cell(g, i, j, k);
  if (u && (c0 || inactive_cell(g, i - 1, j, k))) u[q3] = R(0.);
  if (v && (c0 || inactive_cell(g, i, j - 1, k))) v[q3] = R(0.);
  if (T && c0) T[q3] = R(0.);
  if (S && c0) S[q3] = R(0.);
}
// barotropic transports (decision U11): masked where the surface-level velocity node is peripheral
__global__ void k_mask_barotropic(DevGrid g, real* U, real* V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j);
  const bool c0 = inactive_cell(g, i, j, g.Nz);
  if (c0 || inactive_cell(g, i - 1, j, g.Nz)) U[q2] = R(0.);
  if (c0 || inactive_cell(g, i, j - 1, g.Nz)) V[q2] = R(0.);
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
__global__ void k_compute_w(DevGrid g, const real* __restrict__ u, const real* __restrict__ v, real* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + (-g.Hx + 2);
  const int j = blockIdx.y + (-g.Hy + 2);
  if (i > g.Nx + g.Hx - 1) return;
  const int q2 = id2(g, i, j), PX = g.PX;
  const real dyE = g.dyfc[q2 + 1], dyW = g.dyfc[q2], dxN = g.dxcf[q2 + PX], dxS = g.dxcf[q2];
  const real az = g.azcc[q2];
  size_t q3 = q2 + (size_t)g.n2 * g.Hz;  // k = 1
  real wk = R(0.);
  w[q3] = R(0.);
  for (int k = 2; k <= g.Nz + 1; k++) {
    const real dz = g.dzc[k - 1 + g.Hz - 1];
    const real dU = dyE * dz * u[q3 + 1] - dyW * dz * u[q3];
    const real dV = dxN * dz * v[q3 + PX] - dxS * dz * v[q3];
    wk = wk - (dU + dV) / az;
    q3 += g.n2;
    w[q3] = wk;
  }
}
// =====================================================================================
// update_hydrostatic_pressure! (row A4): one thread per column, downward scan, one EOS call per cell
// =====================================================================================
__global__ void k_compute_p(DevGrid g, const real* __restrict__ T, const real* __restrict__ S, real* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. Nx+1
  const int j = blockIdx.y;                             // 0 .. Ny+1
  if (i > g.Nx + 1) return;
  const int q2 = id2(g, i, j);
  size_t q3 = q2 + (size_t)g.n2 * (g.Nz + 1 + g.Hz - 1);
  const real gr = g.g, r0 = g.rho0;
  real bup = -(gr * teos10_rho_prime(T[q3], S[q3], g.zc[g.Nz + 1 + g.Hz - 1], r0, g.eos_r0) / r0);
  real pk = R(0.);
#pragma unroll 4
  for (int k = g.Nz; k >= 1; k--) {
    q3 -= g.n2;
    const real b = -(gr * teos10_rho_prime(T[q3], S[q3], g.zc[k + g.Hz - 1], r0, g.eos_r0) / r0);
    const real bbar = (b + bup) * R(0.5);
    const real dzf = g.dzf[k + 1 + g.Hz - 1];
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
#ifndef GB25_F64
// Two x-adjacent columns per thread with the 55-term polynomial in packed FP32x2 (FFMA2): k_compute_p is bound by the
// issue rate (ncu: issue slots 79 % active, 118 instructions per cell), and FFMA2 rounds each lane exactly like FFMA,
// so the result is bit-identical.  Pairs start at i = -1 (8-byte aligned storage); columns outside 0..Nx+1 are not stored.
__device__ __forceinline__ float2 pteos10_rho_prime(float2 Theta, float2 SA, real Z, real rho0, int with_r0) {
  const float2 t = pmuls(Theta, R(0.025));
  const float2 sa = pmuls(padd(SA, pbc(R(32.))), R(1.) / R(40.18861714285714));
  const float2 s = make_float2(sqrt_nr(sa.x), sqrt_nr(sa.y));   // (S_A + 32) / 40.19 is of order one: no range check needed
  const real z = Z * -R(1e-4);
#define PF(a, b, c) pfma(a, b, c)
#define PC(x) pbc(x)
  float2 r3 = PF(PC(R(3.7969820455e-01)), t, PF(PC(-R(1.8507636718e-02)), s, PC(-R(2.3342758797e-02))));
  float2 r2 = PF(t, PF(t, PC(-R(1.2419983026)), PF(s, PC(-R(2.1311365518e-01)), PC(R(2.0564311499)))),
                 PF(s, PF(s, PC(R(2.5019633244)), PC(-R(4.9527603989))), PC(R(2.0660924175))));
  float2 r1 = PF(t,
                 PF(t,
                    PF(t, PF(t, PC(R(5.5927935970e-01)), PF(s, PC(-R(5.5077101279e-01)), PC(-R(2.4649669534)))),
                       PF(s, PF(s, PC(-R(1.8795372996)), PC(R(3.5063081279))), PC(R(6.7080479603)))),
                    PF(s, PF(s, PF(s, PC(-R(6.5399043664e-01)), PC(R(5.0042598061))), PC(-R(4.4870114575))), PC(-R(1.3336301113e+01)))),
                 PF(s, PF(s, PF(s, PF(s, PC(R(6.6051753097)), PC(-R(3.0938076334e+01))), PC(R(5.0774768218e+01))), PC(-R(4.2549998214e+01))), PC(R(1.9681925209e+01))));
  float2 q5 = PF(t, PC(-R(1.9083568888e-01)), PF(s, PC(R(4.8169980163e-01)), PC(R(5.4048723791e-01))));
  float2 q4 = PF(t, q5, PF(s, PF(s, PC(-R(5.3563304045)), PC(R(1.1311538584e+01))), PC(-R(8.3627885467))));
  float2 q3 = PF(t, q4, PF(s, PF(s, PF(s, PC(-R(3.1742946532)), PC(R(1.9717078466e+01))), PC(-R(3.3449108469e+01))), PC(R(2.1661789529e+01))));
  float2 q2 = PF(t, q3, PF(s, PF(s, PF(s, PF(s, PC(-R(5.4723692739)), PC(R(2.9130021253e+01))), PC(-R(6.0362551501e+01))), PC(R(6.1548258127e+01))), PC(-R(3.7074170417e+01))));
  float2 q1 = PF(t, q2, PF(s, PF(s, PF(s, PF(s, PF(s, PC(-R(1.9193502195)), PC(R(1.7681814114e+01))), PC(-R(5.6888046321e+01))), PC(R(8.1770425108e+01))), PC(-R(6.5281885265e+01))), PC(R(2.6010145068e+01))));
  float2 r0 = PF(t, q1,
                 PF(s, PF(s, PF(s, PF(s, PF(s, PF(s, PC(-R(6.0579916612e+01)), PC(R(4.3227585684e+02))), PC(-R(1.2849161071e+03))), PC(R(2.0375295546e+03))), PC(-R(1.7864682637e+03))), PC(R(8.6672408165e+02))), PC(R(8.0189615746e+02))));
  float2 r = PF(PF(PF(r3, PC(z), r2), PC(z), r1), PC(z), r0);
#undef PF
#undef PC
  if (with_r0) {
    const real rz = rfma(rfma(rfma(rfma(rfma(rfma(-R(1.7243708991e-03), z, R(1.5616995503e-02)), z, R(6.4326772569e-02)), z, R(2.2601900708e-01)), z, -R(5.2099962525)), z, R(4.6494977072e+01)), z, R(0.));
    r = padd(r, pbc(rz));
  }
  return psub(r, pbc(rho0));
}
__global__ void __launch_bounds__(128) k_compute_p2(DevGrid g, const real* __restrict__ T, const real* __restrict__ S, real* __restrict__ p) {
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x) - 1;   // pair (i, i+1), i = -1, 1, ..., Nx+1
  const int j = blockIdx.y;                                        // 0 .. Ny+1
  if (i > g.Nx + 1) return;
  const bool st0 = i >= 0, st1 = i + 1 <= g.Nx + 1;
  const int q2 = id2(g, i, j);
  size_t q3 = q2 + (size_t)g.n2 * (g.Nz + 1 + g.Hz - 1);
  const real gr = g.g, r0 = g.rho0;
  const real rr0 = rcp_refined(r0);   // rho0 ~ 1e3: the division by it needs no range check (div_by, gb25_device.cuh)
  auto buoy = [&](size_t q, real z) {
    const float2 rp = pteos10_rho_prime(*reinterpret_cast<const float2*>(T + q), *reinterpret_cast<const float2*>(S + q), z, r0, g.eos_r0);
    return make_float2(-div_by(gr * rp.x, r0, rr0), -div_by(gr * rp.y, r0, rr0));
  };
  float2 bup = buoy(q3, g.zc[g.Nz + 1 + g.Hz - 1]);
  float2 pk = make_float2(R(0.), R(0.));
#pragma unroll 2
  for (int k = g.Nz; k >= 1; k--) {
    q3 -= g.n2;
    const float2 b = buoy(q3, g.zc[k + g.Hz - 1]);
    const float2 bbar = pmuls(padd(b, bup), R(0.5));
    const real dzf = g.dzf[k + 1 + g.Hz - 1];
    pk.x = (k == g.Nz) ? -bbar.x * dzf : pk.x - bbar.x * dzf;
    pk.y = (k == g.Nz) ? -bbar.y * dzf : pk.y - bbar.y * dzf;
    if (st0 && st1) *reinterpret_cast<float2*>(p + q3) = pk;
    else if (st0) p[q3] = pk.x;
    else if (st1) p[q3 + 1] = pk.y;
    bup = b;
  }
}
#endif
void launch_compute_p(Handle* h) {
  const DevGrid& g = h->g;
#ifndef GB25_F64
  if (h->use_packed && (g.Hx % 2) == 0 && (g.PX % 2) == 0) {
    const int npairs = (g.Nx + 2) / 2 + 1;
    dim3 b(128), gr((npairs + 127) / 128, g.Ny + 2);
    k_compute_p2<<<gr, b, 0, h->stream>>>(g, h->f.T, h->f.S, h->f.p); h->count_launch();
    return;
  }
#endif
  dim3 b(128), gr((g.Nx + 2 + 127) / 128, g.Ny + 2);
  k_compute_p<<<gr, b, 0, h->stream>>>(g, h->f.T, h->f.S, h->f.p); h->count_launch();
}

#include "gb25_tend_generic.cuh"

__global__ void __launch_bounds__(256) k_tracer_tendency(DevGrid g, const real* __restrict__ u, const real* __restrict__ v,
                                                          const real* __restrict__ w, const real* __restrict__ T,
                                                          const real* __restrict__ S, real* __restrict__ GT, real* __restrict__ GS) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.Nx || j > g.Ny) return;
  const size_t q3 = id2(g, i, j) + (size_t)g.n2 * (k + g.Hz - 1);
  real a, b;
  tracer_cell_generic(g, u, v, w, T, S, i, j, k, a, b);
  GT[q3] = a; GS[q3] = b;
}
void launch_tracer_tendency_v1(Handle* h) {
  const DevGrid& g = h->g;
  dim3 b(64, 4), gr((g.Nx + 63) / 64, (g.Ny + 3) / 4, g.Nz);
  k_tracer_tendency<<<gr, b, 0, h->stream>>>(g, h->f.u, h->f.v, h->f.w, h->f.T, h->f.S, h->f.gn[2], h->f.gn[3]);
  h->count_launch();
}
bool spec_possible(Handle* h) {
  return h->use_spec && h->use_fused && h->use_tma && h->use_tma_tracer && h->use_packed && h->g.Nx % 4 == 0 && h->cfg.closure != 1 &&
         !h->has_bflux &&     // (flux boundary conditions are added to Gn after the tendency kernels)
         tma_available(h);
}
void launch_tracer_tendency(Handle* h, const Ab2Spec* spec) {
  if (h->use_fused && h->use_tma && h->use_tma_tracer && h->g.Nx % 4 == 0 && tma_available(h)) { launch_generic_list(h, false, true); launch_tracer_tendency_tma(h, spec); }
  else if (h->use_fused && h->g.Nx % 4 == 0) { launch_generic_list(h, false, true); launch_tracer_tendency_v2(h); }
  else launch_tracer_tendency_v1(h);
}

template <int DIR>
__global__ void __launch_bounds__(256) k_momentum_tendency(DevGrid g, const real* __restrict__ u, const real* __restrict__ v,
                                                            const real* __restrict__ w, const real* __restrict__ p,
                                                            real* __restrict__ G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.Nx || j > g.Ny) return;
  const size_t q3 = id2(g, i, j) + (size_t)g.n2 * (k + g.Hz - 1);
  G[q3] = DIR == 0 ? momentum_G<0>(g, u, v, w, p, i, j, k) : momentum_G<1>(g, v, u, w, p, i, j, k);
}
void launch_momentum_tendency_v1(Handle* h) {
  const DevGrid& g = h->g;
  dim3 b(64, 4), gr((g.Nx + 63) / 64, (g.Ny + 3) / 4, g.Nz);
  k_momentum_tendency<0><<<gr, b, 0, h->stream>>>(g, h->f.u, h->f.v, h->f.w, h->f.p, h->f.gn[0]); h->count_launch();
  k_momentum_tendency<1><<<gr, b, 0, h->stream>>>(g, h->f.u, h->f.v, h->f.w, h->f.p, h->f.gn[1]); h->count_launch();
}

void launch_momentum_tendency(Handle* h, const Ab2Spec* spec) {
  if (h->use_fused && h->use_tma && h->g.Nx % 4 == 0 && tma_available(h)) { launch_generic_list(h, true, false); launch_momentum_tendency_tma(h, spec); }
  else if (h->use_fused && h->g.Nx % 2 == 0) { launch_generic_list(h, true, false); launch_momentum_tendency_v2(h); }
  else launch_momentum_tendency_v1(h);
}

// =====================================================================================
// ab2_step! part 1 (rows A8 + A9): barotropic forcing column integral fused with the AB2 update
// =====================================================================================
__global__ void k_ab2_columns(DevGrid g, DevFields f, real dt, real chi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j), n2 = g.n2;
  const real c1 = R(1.5) + chi, c2 = R(0.5) + chi;
  const int kb0 = g.kb[q2], kbw = g.kb[q2 - 1], kbs = g.kb[q2 - g.PX];
  const bool ywall = y_outside(g, j) || y_outside(g, j - 1);
  real su = R(0.), sv = R(0.);
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {
    const real dz = g.dzc[k + g.Hz - 1];
    const real gu = ab2_g(c1, c2, f.gn[0][q3], f.gm[0][q3]);
    const real gv = ab2_g(c1, c2, f.gn[1][q3], f.gm[1][q3]);
    const bool pu = k <= kb0 || k <= kbw;
    const bool pv = ywall || k <= kb0 || k <= kbs;
    const real tu = rmul_rn(dz, pu ? R(0.) : gu), tv = rmul_rn(dz, pv ? R(0.) : gv);
    su = (k == 1) ? tu : radd_rn(su, tu);
    sv = (k == 1) ? tv : radd_rn(sv, tv);
    f.u[q3] = ab2_upd(f.u[q3], dt, gu);
    f.v[q3] = ab2_upd(f.v[q3], dt, gv);
    f.T[q3] = ab2_upd(f.T[q3], dt, ab2_g(c1, c2, f.gn[2][q3], f.gm[2][q3]));
    f.S[q3] = ab2_upd(f.S[q3], dt, ab2_g(c1, c2, f.gn[3][q3], f.gm[3][q3]));
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
// Edge handling: a tile that spans the whole direction applies the topology itself (periodic wrap, wall, fold);
// on a partitioned grid the one-cell halos of eta / U / V are refreshed by the neighbours after every kernel.
//   bflags bit 0: wrap in x locally (Rx == 1); bit 1: south wall; bit 2: north wall; bit 3: local north fold
// On a partitioned grid the kernels also PUSH their edge values straight into the neighbours' one-cell halos
// (peer stores over NVLink from the threads that own the edge): no separate copy kernels between the substeps.
struct BaroPeers { real *eta_E, *eta_N, *bu_W, *bv_S, *bv_F; };
__global__ void k_baro_eta(DevGrid g, real* __restrict__ eta, const real* __restrict__ U, const real* __restrict__ V, real dtau, int bflags,
                           BaroPeers pe) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j), PX = g.PX;
  const int qe = (i == g.Nx && (bflags & 1)) ? id2(g, 1, j) : q2 + 1;
  const real dU = g.dyfc[qe] * U[qe] - g.dyfc[q2] * U[q2];
  real dV;
  if (j == 1 && (bflags & 2)) dV = g.dxcf[q2 + PX] * V[q2 + PX];
  else if (j == g.Ny && (bflags & 12)) {
    if (bflags & 4) dV = -(g.dxcf[q2] * V[q2]);
    else {  // folded row Ny+1: V[i,Ny+1] = -V[Nx-i+1,Ny]
      const real vn = -V[id2(g, g.Nx - i + 1, g.Ny)];
      dV = g.dxcf[q2 + PX] * vn - g.dxcf[q2] * V[q2];
    }
  } else dV = g.dxcf[q2 + PX] * V[q2 + PX] - g.dxcf[q2] * V[q2];
  const real en = eta[q2] - dtau * (dU + dV) / g.azcc[q2];
  eta[q2] = en;
  if (pe.eta_E && i == g.Nx) pe.eta_E[id2(g, 0, j)] = en;          // my last column  -> the east tile's west halo
  if (pe.eta_N && j == g.Ny) pe.eta_N[id2(g, i, 0)] = en;          // my last row     -> the north tile's south halo
}
__global__ void k_baro_uv(DevGrid g, DevFields f, real dtau, real wgt, int bflags, BaroPeers pe) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j);
  const int qw = (i == 1 && (bflags & 1)) ? id2(g, g.Nx, j) : q2 - 1;
  const real e0 = f.eta[q2];
  const real dxe = (e0 - f.eta[qw]) / g.dxfc[q2];
  const real dye = (j == 1 && (bflags & 2)) ? R(0.) : (e0 - f.eta[q2 - g.PX]) / g.dycf[q2];
  const real Un = f.bu[q2] + dtau * (-g.g * g.Hfc[q2] * dxe + f.gU[q2]);
  const real Vn = f.bv[q2] + dtau * (-g.g * g.Hcf[q2] * dye + f.gV[q2]);
  f.bu[q2] = Un; f.bv[q2] = Vn;
  if (pe.bu_W && i == 1) pe.bu_W[id2(g, g.Nx + 1, j)] = Un;                       // my first column -> the west tile's east halo
  if (pe.bv_S && j == 1) pe.bv_S[id2(g, i, g.Ny + 1)] = Vn;                       // my first row    -> the south tile's north halo
  if (pe.bv_F && j == g.Ny) pe.bv_F[id2(g, g.Nx - i + 1, g.Ny + 1)] = -Vn;        // fold: V(i', Ny+1) = -V(Nx-i'+1, Ny) on the partner
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
  if (launch_barotropic_persistent(h, dt)) return;   // gb25_baro.cu: all substeps in one persistent kernel
  const DevGrid& g = h->g;
  const size_t b2 = (size_t)g.n2 * sizeof(real);
  cudaMemsetAsync(h->f.feta, 0, b2, h->stream);
  cudaMemsetAsync(h->f.fu, 0, b2, h->stream);
  cudaMemsetAsync(h->f.fv, 0, b2, h->stream);
  const real dtau = (real)h->cfg.dtau_frac * (real)dt;   // (the product in the scalar type of the build)
  dim3 b(128), gr((g.Nx + 127) / 128, g.Ny);
  const gb25_config& c = h->cfg;
  const int bflags = (c.Rx == 1 ? 1 : 0) | (c.ry == 0 ? 2 : 0) | ((c.ry == c.Ry - 1 && c.topo_y == GB25_TOPO_BOUNDED) ? 4 : 0) |
                     ((c.ry == c.Ry - 1 && c.topo_y == GB25_TOPO_FOLD && c.Rx == 1) ? 8 : 0);
  BaroPeers pe = {nullptr, nullptr, nullptr, nullptr, nullptr};
  if (h->ex.on) {
    const Exchange& X = h->ex;
    if (c.Rx > 1) { pe.eta_E = X.to[SLOT_E].fld[EX_ETA]; pe.bu_W = X.to[SLOT_W].fld[EX_BU]; }
    if (c.ry < c.Ry - 1) pe.eta_N = X.to[SLOT_N].fld[EX_ETA];
    if (c.ry > 0) pe.bv_S = X.to[SLOT_S].fld[EX_BV];
    if (c.topo_y == GB25_TOPO_FOLD && c.ry == c.Ry - 1 && c.Rx > 1) pe.bv_F = X.to[SLOT_FOLD].fld[EX_BV];
  }
  for (int m = 0; m < h->cfg.nsubsteps; m++) {
    k_baro_eta<<<gr, b, 0, h->stream>>>(g, h->f.eta, h->f.bu, h->f.bv, dtau, bflags, pe); h->count_launch();
    if (h->ex.on) exchange_baro_eta(h);
    k_baro_uv<<<gr, b, 0, h->stream>>>(g, h->f, dtau, h->weights[m], bflags, pe); h->count_launch();
    if (h->ex.on) exchange_baro_uv(h);
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
  real su = R(0.), sv = R(0.);
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {
    const real dz = g.dzc[k + g.Hz - 1];
    const real tu = dz * f.u[q3], tv = dz * f.v[q3];
    su = (k == 1) ? tu : su + tu;
    sv = (k == 1) ? tv : sv + tv;
  }
  f.fu[q2] = su; f.fv[q2] = sv;
  const real cu = (f.bu[q2] - su) / g.Hfc[q2], cv = (f.bv[q2] - sv) / g.Hcf[q2];
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
  real su = R(0.), sv = R(0.);
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {
    const real dz = g.dzc[k + g.Hz - 1];
    const real tu = dz * f.u[q3], tv = dz * f.v[q3];
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

// =====================================================================================
// Fused step path (gb25_time_step / gb25_loop).  Same arithmetic as the operator-level kernels above, with the
// pointwise stages folded into the two column passes that have to touch the state anyway:
//   k_ab2_fused     = compute_free_surface_tendency! + ab2_step_velocities!/tracers! + mask_immersed_field!(u,v)
//                     (done by step_free_surface! after the substeps; it does not interact with them) + the T,S
//                     mask of the next update_state! + the column sums the corrector needs, and
//   k_correct_fused = barotropic corrector + mask_immersed_model_fields!(u,v,U,V) of update_state!.
// G- <- Gn becomes a pointer swap on the host.
// =====================================================================================
__global__ void k_ab2_fused(DevGrid g, DevFields f, real* __restrict__ us2, real* __restrict__ vs2, real dt, real chi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j), n2 = g.n2;
  const real c1 = R(1.5) + chi, c2 = R(0.5) + chi;
  const int kb0 = g.kb[q2], kbw = g.kb[q2 - 1], kbs = g.kb[q2 - g.PX];
  const bool ywall = y_outside(g, j) || y_outside(g, j - 1);
  const bool imm = g.immersed;
  real su = R(0.), sv = R(0.), bu = R(0.), bv = R(0.);
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {   // (unrolling this loop costs occupancy: 0.54 -> 0.68 ms, measured)
    const real dz = g.dzc[k + g.Hz - 1];
    const real gu = ab2_g(c1, c2, f.gn[0][q3], f.gm[0][q3]);
    const real gv = ab2_g(c1, c2, f.gn[1][q3], f.gm[1][q3]);
    const bool pu = k <= kb0 || k <= kbw;
    const bool pv = ywall || k <= kb0 || k <= kbs;
    const real tu = rmul_rn(dz, pu ? R(0.) : gu), tv = rmul_rn(dz, pv ? R(0.) : gv);
    su = (k == 1) ? tu : radd_rn(su, tu);
    sv = (k == 1) ? tv : radd_rn(sv, tv);
    real un = ab2_upd(f.u[q3], dt, gu), vn = ab2_upd(f.v[q3], dt, gv);
    real Tn = ab2_upd(f.T[q3], dt, ab2_g(c1, c2, f.gn[2][q3], f.gm[2][q3]));
    real Sn = ab2_upd(f.S[q3], dt, ab2_g(c1, c2, f.gn[3][q3], f.gm[3][q3]));
    if (imm) {
      if (pu) un = R(0.);
      if (pv) vn = R(0.);
      if (k <= kb0) { Tn = R(0.); Sn = R(0.); }
    }
    f.u[q3] = un; f.v[q3] = vn; f.T[q3] = Tn; f.S[q3] = Sn;
    const real wu = rmul_rn(dz, un), wv = rmul_rn(dz, vn);
    bu = (k == 1) ? wu : radd_rn(bu, wu);
    bv = (k == 1) ? wv : radd_rn(bv, wv);
  }
  f.gU[q2] = su; f.gV[q2] = sv;
  us2[q2] = bu; vs2[q2] = bv;
}
__global__ void k_correct_fused(DevGrid g, DevFields f, const real* __restrict__ us2, const real* __restrict__ vs2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j), n2 = g.n2;
  const real su = us2[q2], sv = vs2[q2];
  f.fu[q2] = su; f.fv[q2] = sv;     // filtered_state.U/V are reused as scratch by the reference
  const real cu = (f.bu[q2] - su) / g.Hfc[q2], cv = (f.bv[q2] - sv) / g.Hcf[q2];
  const int kb0 = g.kb[q2], kbw = g.kb[q2 - 1], kbs = g.kb[q2 - g.PX];
  const bool ywall = y_outside(g, j) || y_outside(g, j - 1);
  const bool imm = g.immersed;
  size_t q3 = q2 + (size_t)n2 * g.Hz;
#pragma unroll 4
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {
    real un = f.u[q3] + cu, vn = f.v[q3] + cv;
    if (imm) {
      if (k <= kb0 || k <= kbw) un = R(0.);
      if (ywall || k <= kb0 || k <= kbs) vn = R(0.);
    }
    f.u[q3] = un; f.v[q3] = vn;
  }
  if (imm) {   // barotropic transports: masked where the surface node is peripheral (decision U11)
    if (g.Nz <= kb0 || g.Nz <= kbw) f.bu[q2] = R(0.);
    if (ywall || g.Nz <= kb0 || g.Nz <= kbs) f.bv[q2] = R(0.);
  }
  f.gmU[q2] = f.gU[q2]; f.gmV[q2] = f.gV[q2];
}
// The same stage in two kernels: the u, v half needs the column sums (thread per column, as above); the T, S half has no
// dependence along k and streams like k_correct_3d (one thread per four x-adjacent cells x AB2_KCH levels, 128-bit accesses).
#define AB2_KCH 10
__global__ void __launch_bounds__(128) k_ab2_uv(DevGrid g, DevFields f, real* __restrict__ us2, real* __restrict__ vs2, real dt, real chi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j), n2 = g.n2;
  const real c1 = R(1.5) + chi, c2 = R(0.5) + chi;
  const int kb0 = g.kb[q2], kbw = g.kb[q2 - 1], kbs = g.kb[q2 - g.PX];
  const bool ywall = y_outside(g, j) || y_outside(g, j - 1);
  const bool imm = g.immersed;
  real su = R(0.), sv = R(0.), bu = R(0.), bv = R(0.);
  size_t q3 = q2 + (size_t)n2 * g.Hz;
  for (int k = 1; k <= g.Nz; k++, q3 += n2) {
    const real dz = g.dzc[k + g.Hz - 1];
    const real gu = ab2_g(c1, c2, f.gn[0][q3], f.gm[0][q3]);
    const real gv = ab2_g(c1, c2, f.gn[1][q3], f.gm[1][q3]);
    const bool pu = k <= kb0 || k <= kbw;
    const bool pv = ywall || k <= kb0 || k <= kbs;
    const real tu = rmul_rn(dz, pu ? R(0.) : gu), tv = rmul_rn(dz, pv ? R(0.) : gv);
    su = (k == 1) ? tu : radd_rn(su, tu);
    sv = (k == 1) ? tv : radd_rn(sv, tv);
    real un = ab2_upd(f.u[q3], dt, gu), vn = ab2_upd(f.v[q3], dt, gv);
    if (imm) {
      if (pu) un = R(0.);
      if (pv) vn = R(0.);
    }
    f.u[q3] = un; f.v[q3] = vn;
    const real wu = rmul_rn(dz, un), wv = rmul_rn(dz, vn);
    bu = (k == 1) ? wu : radd_rn(bu, wu);
    bv = (k == 1) ? wv : radd_rn(bv, wv);
  }
  f.gU[q2] = su; f.gV[q2] = sv;
  us2[q2] = bu; vs2[q2] = bv;
}
#ifndef GB25_F64
__global__ void __launch_bounds__(128) k_ab2_ts_3d(DevGrid g, DevFields f, real dt, real chi) {
  const int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x) + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int k0 = blockIdx.z * AB2_KCH + 1, k1 = min(k0 + AB2_KCH - 1, g.Nz);
  const int q2 = id2(g, i, j), n2 = g.n2;
  const real c1 = R(1.5) + chi, c2 = R(0.5) + chi;
  const int kb[4] = {g.kb[q2], g.kb[q2 + 1], g.kb[q2 + 2], g.kb[q2 + 3]};
  const bool imm = g.immersed;
  size_t q3 = q2 + (size_t)n2 * (k0 + g.Hz - 1);
#pragma unroll 2
  for (int k = k0; k <= k1; k++, q3 += n2) {
#pragma unroll
    for (int q = 2; q < 4; q++) {
      real* __restrict__ x = q == 2 ? f.T : f.S;
      const float4 x4 = *reinterpret_cast<const float4*>(x + q3);
      const float4 n4 = *reinterpret_cast<const float4*>(f.gn[q] + q3), m4 = *reinterpret_cast<const float4*>(f.gm[q] + q3);
      real xn[4] = {ab2_upd(x4.x, dt, ab2_g(c1, c2, n4.x, m4.x)), ab2_upd(x4.y, dt, ab2_g(c1, c2, n4.y, m4.y)),
                     ab2_upd(x4.z, dt, ab2_g(c1, c2, n4.z, m4.z)), ab2_upd(x4.w, dt, ab2_g(c1, c2, n4.w, m4.w))};
      if (imm) {
#pragma unroll
        for (int c = 0; c < 4; c++) if (k <= kb[c]) xn[c] = R(0.);
      }
      *reinterpret_cast<float4*>(x + q3) = make_float4(xn[0], xn[1], xn[2], xn[3]);
    }
  }
}
#endif
#ifndef GB25_F64
// the 2-D by-products of an AB2 epilogue (barotropic forcing, transport sums) move into their model arrays when the
// speculation is consumed: one launch, four arrays
__global__ void k_commit_spec(int n4, const float4* __restrict__ a0, const float4* __restrict__ a1, const float4* __restrict__ a2,
                              const float4* __restrict__ a3, float4* __restrict__ b0, float4* __restrict__ b1, float4* __restrict__ b2,
                              float4* __restrict__ b3) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n4) return;
  b0[q] = a0[q]; b1[q] = a1[q]; b2[q] = a2[q]; b3[q] = a3[q];
}
void launch_commit_spec(Handle* h) {
  const int n4 = h->g.n2 / 4;   // PX % 4 == 0 on this path
  k_commit_spec<<<(n4 + 255) / 256, 256, 0, h->stream>>>(n4, (const float4*)h->spec2d[0], (const float4*)h->spec2d[1], (const float4*)h->spec2d[2],
                                                        (const float4*)h->spec2d[3], (float4*)h->f.gU, (float4*)h->f.gV, (float4*)h->us2, (float4*)h->vs2);
  h->count_launch();
}
#else
void launch_commit_spec(Handle*) {}
#endif
void launch_ab2_fused(Handle* h, float dt, float chi) {
  const DevGrid& g = h->g;
  StageScope ts(h, "kernel:k_ab2_fused");
  static const bool split = []() { const char* e = getenv("GB25_AB2_SPLIT_TS"); return !(e && e[0] == '0'); }();
#ifndef GB25_F64
  if (split && g.Nx % 4 == 0 && g.Hx % 4 == 0 && g.PX % 4 == 0) {
    dim3 b(128), gr((g.Nx + 127) / 128, g.Ny);
    k_ab2_uv<<<gr, b, 0, h->stream>>>(g, h->f, h->us2, h->vs2, dt, chi); h->count_launch();
    dim3 g3((g.Nx / 4 + 127) / 128, g.Ny, (g.Nz + AB2_KCH - 1) / AB2_KCH);
    k_ab2_ts_3d<<<g3, b, 0, h->stream>>>(g, h->f, dt, chi); h->count_launch();
    return;
  }
#else
  (void)split;
#endif
  dim3 b(128), gr((g.Nx + 127) / 128, g.Ny);
  k_ab2_fused<<<gr, b, 0, h->stream>>>(g, h->f, h->us2, h->vs2, dt, chi); h->count_launch();
}
// The corrector has no dependence along k (the correction is a 2-D field), so the 3-D part streams: one thread per four
// x-adjacent cells and CORR_KCH levels, 128-bit accesses in memory order, instead of one thread marching a whole column.
// Same per-cell expressions as k_correct_fused.  The 2-D part (filtered-state scratch, U,V masks, G- <- Gn of the
// barotropic tendencies) runs first in its own small kernel and leaves the unmasked transports in two scratch arrays,
// from which the 3-D kernel recomputes the correction.
#define CORR_KCH 10
#ifndef GB25_F64
template <bool ZH>
__global__ void __launch_bounds__(128) k_correct_3d(DevGrid g, DevFields f, const real* __restrict__ us2, const real* __restrict__ vs2,
                                                    const real* __restrict__ ubar, const real* __restrict__ vbar, int bottom_tile) {
  const int i = 4 * (blockIdx.x * blockDim.x + threadIdx.x) + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  // the impenetrable condition of the halo fill sets v(i, 1, k) = 0 on the bottom row of tiles right after this kernel; the
  // z halos written here (ZH) must mirror that value, so it is applied now
  const bool vwall = ZH && bottom_tile && j == 1;
  const int k0 = blockIdx.z * CORR_KCH + 1, k1 = min(k0 + CORR_KCH - 1, g.Nz);
  const int q2 = id2(g, i, j), n2 = g.n2;
  const float4 su = *reinterpret_cast<const float4*>(us2 + q2), sv = *reinterpret_cast<const float4*>(vs2 + q2);
  const float4 bu = *reinterpret_cast<const float4*>(ubar + q2), bv = *reinterpret_cast<const float4*>(vbar + q2);
  const float4 hf = *reinterpret_cast<const float4*>(g.Hfc + q2), hc = *reinterpret_cast<const float4*>(g.Hcf + q2);
  const real cu[4] = {(bu.x - su.x) / hf.x, (bu.y - su.y) / hf.y, (bu.z - su.z) / hf.z, (bu.w - su.w) / hf.w};
  const real cv[4] = {(bv.x - sv.x) / hc.x, (bv.y - sv.y) / hc.y, (bv.z - sv.z) / hc.z, (bv.w - sv.w) / hc.w};
  int kbu[4], kbv[4];   // highest solid level next to the u / v node
  {
    const int kw = g.kb[q2 - 1];
    const int kc[4] = {g.kb[q2], g.kb[q2 + 1], g.kb[q2 + 2], g.kb[q2 + 3]};
    const bool ywall = y_outside(g, j) || y_outside(g, j - 1);
#pragma unroll
    for (int c = 0; c < 4; c++) {
      kbu[c] = max(kc[c], c == 0 ? kw : kc[c > 0 ? c - 1 : 0]);
      kbv[c] = ywall ? GB25_BIG : max(kc[c], (int)g.kb[q2 - g.PX + c]);
    }
  }
  const bool imm = g.immersed;
  size_t q3 = q2 + (size_t)n2 * (k0 + g.Hz - 1);
#pragma unroll 2
  for (int k = k0; k <= k1; k++, q3 += n2) {
    float4 u4 = *reinterpret_cast<const float4*>(f.u + q3), v4 = *reinterpret_cast<const float4*>(f.v + q3);
    real un[4] = {u4.x + cu[0], u4.y + cu[1], u4.z + cu[2], u4.w + cu[3]};
    real vn[4] = {v4.x + cv[0], v4.y + cv[1], v4.z + cv[2], v4.w + cv[3]};
    if (imm) {
#pragma unroll
      for (int c = 0; c < 4; c++) {
        if (k <= kbu[c]) un[c] = R(0.);
        if (k <= kbv[c]) vn[c] = R(0.);
      }
    }
    if (vwall) { vn[0] = R(0.); vn[1] = R(0.); vn[2] = R(0.); vn[3] = R(0.); }
    const float4 uo = make_float4(un[0], un[1], un[2], un[3]), vo = make_float4(vn[0], vn[1], vn[2], vn[3]);
    *reinterpret_cast<float4*>(f.u + q3) = uo;
    *reinterpret_cast<float4*>(f.v + q3) = vo;
    if (ZH) {   // no-flux z halos (mirror): psi(1-m) = psi(m), psi(Nz+m) = psi(Nz+1-m), m = 1..Hz  (k_halo_bottom_top)
      if (k <= g.Hz) {
        const size_t qm = q3 - (size_t)(2 * k - 1) * n2;
        *reinterpret_cast<float4*>(f.u + qm) = uo; *reinterpret_cast<float4*>(f.v + qm) = vo;
      }
      if (k > g.Nz - g.Hz) {
        const size_t qm = q3 + (size_t)(2 * (g.Nz - k) + 1) * n2;
        *reinterpret_cast<float4*>(f.u + qm) = uo; *reinterpret_cast<float4*>(f.v + qm) = vo;
      }
    }
  }
}
#endif
// 2-D part of the corrector; leaves the unmasked transports in ubar / vbar for k_correct_3d
__global__ void k_correct_2d(DevGrid g, DevFields f, const real* __restrict__ us2, const real* __restrict__ vs2,
                             real* __restrict__ ubar, real* __restrict__ vbar) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j);
  f.fu[q2] = us2[q2]; f.fv[q2] = vs2[q2];     // filtered_state.U/V are reused as scratch by the reference
  ubar[q2] = f.bu[q2]; vbar[q2] = f.bv[q2];
  if (g.immersed) {   // barotropic transports: masked where the surface node is peripheral (decision U11)
    const int kb0 = g.kb[q2], kbw = g.kb[q2 - 1], kbs = g.kb[q2 - g.PX];
    const bool ywall = y_outside(g, j) || y_outside(g, j - 1);
    if (g.Nz <= kb0 || g.Nz <= kbw) f.bu[q2] = R(0.);
    if (ywall || g.Nz <= kb0 || g.Nz <= kbs) f.bv[q2] = R(0.);
  }
  f.gmU[q2] = f.gU[q2]; f.gmV[q2] = f.gV[q2];
}
bool launch_correct_fused(Handle* h) {
  const DevGrid& g = h->g;
  static const bool streamed = []() { const char* e = getenv("GB25_CORRECT_3D"); return !(e && e[0] == '0'); }();
#ifndef GB25_F64
  if (streamed && g.Nx % 4 == 0 && g.Hx % 4 == 0 && g.PX % 4 == 0) {
    dim3 b2(128), g2((g.Nx + 127) / 128, g.Ny);
    k_correct_2d<<<g2, b2, 0, h->stream>>>(g, h->f, h->us2, h->vs2, h->corr_u, h->corr_v); h->count_launch();
    dim3 b(128), gr((g.Nx / 4 + 127) / 128, g.Ny, (g.Nz + CORR_KCH - 1) / CORR_KCH);
    const bool zh = h->use_zfold && g.Nz >= g.Hz;
    const int bottom = h->cfg.ry == 0 ? 1 : 0;
    if (zh) k_correct_3d<true><<<gr, b, 0, h->stream>>>(g, h->f, h->us2, h->vs2, h->corr_u, h->corr_v, bottom);
    else k_correct_3d<false><<<gr, b, 0, h->stream>>>(g, h->f, h->us2, h->vs2, h->corr_u, h->corr_v, bottom);
    h->count_launch();
    return zh;
  }
#else
  (void)streamed;
#endif
  dim3 b(128), gr((g.Nx + 127) / 128, g.Ny);
  k_correct_fused<<<gr, b, 0, h->stream>>>(g, h->f, h->us2, h->vs2); h->count_launch();
  return false;
}

// =====================================================================================
// Vertical diffusion (row A13): VerticalScalarDiffusivity(kappa, nu), constant coefficients.
// A face is "open" when it is not a peripheral node: 2 <= k <= Nz and every adjacent cell is fluid.
// =====================================================================================
__device__ __forceinline__ int vdiff_kbm(const DevGrid& g, int q2, int j, int fld) {
  // highest solid level among the columns adjacent to the node (u: i-1,i ; v: j-1,j ; T,S: own); walls -> BIG
  const int kb0 = g.kb[q2];
  if (fld == 0) return max(kb0, (int)g.kb[q2 - 1]);
  if (fld == 1) return (y_outside(g, j) || y_outside(g, j - 1)) ? GB25_BIG : max(kb0, (int)g.kb[q2 - g.PX]);
  return kb0;
}
__global__ void k_vdiff_explicit(DevGrid g, DevFields f, real kappa, real nu) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j), n2 = g.n2;
  const size_t q3 = q2 + (size_t)n2 * (k + g.Hz - 1);
  const real dzc = g.dzc[k + g.Hz - 1], dzt = g.dzf[k + 1 + g.Hz - 1], dzb = g.dzf[k + g.Hz - 1];
  real* fld[4] = {f.u, f.v, f.T, f.S};
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int kbm = vdiff_kbm(g, q2, j, q);
    const real K = q < 2 ? nu : kappa;
    const real* c = fld[q] + q3;
    const bool open_t = (k + 1 <= g.Nz) && (k > kbm), open_b = (k >= 2) && (k - 1 > kbm);
    const real qt = open_t ? K * (c[n2] - c[0]) / dzt : R(0.);
    const real qb = open_b ? K * (c[0] - c[-n2]) / dzb : R(0.);
    f.gn[q][q3] = f.gn[q][q3] + (qt - qb) / dzc;
  }
}
void launch_vdiff_explicit(Handle* h) {
  const DevGrid& g = h->g;
  dim3 b(128), gr((g.Nx + 127) / 128, g.Ny, g.Nz);
  k_vdiff_explicit<<<gr, b, 0, h->stream>>>(g, h->f, h->cfg.kappa, h->cfg.nu); h->count_launch();
}
// =====================================================================================
// Boundary tendency contributions (row A7): compute_hydrostatic_boundary_tendency_contributions! -> apply_z_bcs!
// (/root/reference/src/precompile.jl:52-61).  Gc[i,j,1] += J_bottom Az / V, Gc[i,j,Nz] -= J_top Az / V, Az and V at the
// location of the field; only fields with a flux array set take part.
// =====================================================================================
struct FluxSet { const real* a[4][2]; };
__global__ void k_boundary_tendencies(DevGrid g, DevFields f, FluxSet fs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j);
  const real* az[4] = {g.azfc, g.azcf, g.azcc, g.azcc};
#pragma unroll
  for (int q = 0; q < 4; q++)
#pragma unroll
    for (int side = 0; side < 2; side++) {
      if (!fs.a[q][side]) continue;
      const int k = side ? g.Nz : 1;
      const size_t q3 = q2 + (size_t)g.n2 * (k + g.Hz - 1);
      const real A = az[q][q2], V = A * g.dzc[k + g.Hz - 1];
      const real d = fs.a[q][side][q2] * A / V;
      f.gn[q][q3] = side ? f.gn[q][q3] - d : f.gn[q][q3] + d;
    }
}
void launch_boundary_tendencies(Handle* h) {
  if (!h->has_bflux) return;
  const DevGrid& g = h->g;
  FluxSet fs;
  for (int q = 0; q < 4; q++) for (int s = 0; s < 2; s++) fs.a[q][s] = h->bflux[q][s];
  dim3 b(128), gr((g.Nx + 127) / 128, g.Ny);
  k_boundary_tendencies<<<gr, b, 0, h->stream>>>(g, h->f, fs); h->count_launch();
}
// implicit: (I - dt d_z K d_z) c = c*, Thomas algorithm per column (Oceananigans' batched tridiagonal solver);
// the elimination factors t[k] go through a 3-D scratch array (the zeta scratch, rebuilt later in the step)
__global__ void k_implicit_columns(DevGrid g, DevFields f, real* __restrict__ scratch, real dt, real kappa, real nu,
                                   real* __restrict__ us2, real* __restrict__ vs2, int with_sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.Nx) return;
  const int q2 = id2(g, i, j), n2 = g.n2, Nz = g.Nz;
  const size_t q1 = q2 + (size_t)n2 * g.Hz;   // level 1
  real* fld[4] = {f.u, f.v, f.T, f.S};
#pragma unroll 1
  for (int q = 0; q < 4; q++) {
    const int kbm = vdiff_kbm(g, q2, j, q);
    const real K = q < 2 ? nu : kappa;
    real* c = fld[q] + q1;     // c[(k-1)*n2]
    real* t = scratch + q1;
    auto upper = [&](int k) -> real {   // face k+1
      if (k > Nz - 1) return R(0.);
      return (k > kbm) ? -dt * K / (g.dzc[k + g.Hz - 1] * g.dzf[k + 1 + g.Hz - 1]) : R(0.);
    };
    auto lower = [&](int kp) -> real {  // face k = kp+1
      if (kp < 1) return R(0.);
      const int k = kp + 1;
      return (k - 1 > kbm) ? -dt * K / (g.dzc[k + g.Hz - 1] * g.dzf[k + g.Hz - 1]) : R(0.);
    };
    real beta = R(1.) - upper(1) - lower(0);
    real prev = c[0] / beta;
    c[0] = prev;
    for (int k = 2; k <= Nz; k++) {
      const real lo = lower(k - 1);
      const real tk = upper(k - 1) / beta;
      t[(size_t)(k - 1) * n2] = tk;
      beta = (R(1.) - upper(k) - lo) - lo * tk;
      prev = (c[(size_t)(k - 1) * n2] - lo * prev) / beta;
      c[(size_t)(k - 1) * n2] = prev;
    }
    for (int k = Nz - 1; k >= 1; k--) {
      prev = c[(size_t)(k - 1) * n2] - t[(size_t)k * n2] * prev;
      c[(size_t)(k - 1) * n2] = prev;
    }
  }
  if (with_sums) {   // the corrector's column sums must see the diffused velocities
    real bu = R(0.), bv = R(0.);
    size_t q3 = q1;
    for (int k = 1; k <= Nz; k++, q3 += n2) {
      const real dz = g.dzc[k + g.Hz - 1];
      const real wu = dz * f.u[q3], wv = dz * f.v[q3];
      bu = (k == 1) ? wu : bu + wu;
      bv = (k == 1) ? wv : bv + wv;
    }
    us2[q2] = bu; vs2[q2] = bv;
  }
}
void launch_implicit_columns(Handle* h, float dt, bool with_sums) {
  const DevGrid& g = h->g;
  dim3 b(128), gr((g.Nx + 127) / 128, g.Ny);
  k_implicit_columns<<<gr, b, 0, h->stream>>>(g, h->f, h->zeta, dt, h->cfg.kappa, h->cfg.nu, h->us2, h->vs2, with_sums ? 1 : 0);
  h->count_launch();
}

// ---------------------------------------------------------------- kernel table (preload_kernels, gb25_api.cu)
KernelTable kernel_table_core() {
  static const void* const k[] = {
    (const void*)k_halo_south_north, (const void*)k_halo_fold_row, (const void*)k_halo_bottom_top, (const void*)k_halo_periodic_x,
    (const void*)k_mask_fields, (const void*)k_mask_barotropic, (const void*)k_compute_w, (const void*)k_compute_p,
    (const void*)k_tracer_tendency, (const void*)k_momentum_tendency<0>, (const void*)k_momentum_tendency<1>,
    (const void*)k_ab2_columns, (const void*)k_baro_eta, (const void*)k_baro_uv, (const void*)k_baro_finish,
    (const void*)k_correct_cache, (const void*)k_barotropic_mode, (const void*)k_ab2_fused, (const void*)k_correct_fused,
    (const void*)k_ab2_uv, (const void*)k_correct_2d, (const void*)k_vdiff_explicit, (const void*)k_boundary_tendencies,
    (const void*)k_implicit_columns,
#ifndef GB25_F64
    (const void*)k_compute_p2, (const void*)k_ab2_ts_3d, (const void*)k_commit_spec, (const void*)k_correct_3d<false>,
    (const void*)k_correct_3d<true>,
#endif
  };
  return {k, (int)(sizeof k / sizeof k[0])};
}
