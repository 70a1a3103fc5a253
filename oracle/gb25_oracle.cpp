// oracle/gb25_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of the Oceananigans
// HydrostaticFreeSurfaceModel time step that GB-25 drives
// (/root/reference/src/timestepping_utils.jl:21-45, stage order
// /root/reference/src/precompile.jl:31-42, physics configuration
// /root/reference/src/baroclinic_instability_model.jl:17-40).
//
// *** PARITY UNPINNED ***  The arithmetic of this path lives in un-vendored
// third-party Julia packages (Oceananigans =0.96.26, SeawaterPolynomials 0.3.9;
// /root/reference/Project.toml:37,42) that are not on disk here; Julia is not
// installed; the reference ships no golden vectors.  This file restates the
// published algorithm and is pinned only by
// self-contained known answers (tests/test_oracle_known_answers.py): TEOS-10
// check value, WENO exactness / ideal weights, split-explicit weights,
// conservation and fold identities.  Every fidelity decision is a named flag
// in OConfig (DESIGN.md "uncertainty register").
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product (libgb25cuda)
// never links, loads or calls it.
//
// Build: oracle/Makefile  ->  oracle/libgb25oracle.so
//
// Layout: every 3-D array is (PX,PY,PZ) = (Nx+2Hx, Ny+2Hy+1, Nz+2Hz+1),
// x fastest; interior index (i,j,k) (1-based, Julia convention) lives at
// storage (i+Hx-1, j+Hy-1, k+Hz-1).  2-D arrays are (PX,PY).  The extra
// row/plane holds the Ny+1 / Nz+1 faces of Bounded directions.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {
struct OConfig {
  int Nx, Ny, Nz, Hx, Hy, Hz;
  int topo_y;           // 0 = Bounded (lat-lon), 1 = RightConnected (tripolar north fold)
  int immersed;         // 0 = plain grid, 1 = ImmersedBoundaryGrid(GridFittedBottom)
  int nsub;             // number of barotropic substeps actually taken (21 for substeps=30)
  int coriolis_scheme;  // U2: 0 = EnstrophyConserving, 1 = ActiveCellEnstrophyConserving
  int fold_variant;     // U1: 0 = plain zipper, 1 = also overwrite redundant half of row Ny
  int south_inactive;   // U4: RightConnected: cells j<1 count as outside the domain (1) or not (0)
  int cond_diff;        // U15: immersed-aware (conditional) differences in zeta and grad p
  int eos_r0;           // U8: include r0(z) in rho'
  int beta_form;        // D1: 0 = expanded smoothness indicators (recalled reference form), 1 = sum of squares
  int closure;          // A13: 0 = nothing, 1 = VerticalScalarDiffusivity explicit, 2 = vertically implicit
  double g, rho0, chi, dtau_frac, weno_eps, kappa, nu;
};
}

enum { TOPO_BOUNDED = 0, TOPO_FOLD = 1 };
enum FieldId {
  F_U = 0, F_V, F_W, F_T, F_S, F_P,
  F_GNU, F_GNV, F_GNT, F_GNS, F_GMU, F_GMV, F_GMT, F_GMS,
  F_ETA, F_BU, F_BV, F_FETA, F_FU, F_FV, F_GU, F_GV, F_GMBU, F_GMBV,
  F_COUNT
};
// location codes: 0 = Center, 1 = Face
struct Loc { int x, y, z; };

template <class FT>
struct Oracle {
  OConfig c;
  int Nx, Ny, Nz, Hx, Hy, Hz, PX, PY, PZ;
  size_t n2, n3;
  std::vector<FT> dxcc, dxfc, dxcf, dxff, dycc, dyfc, dycf, dyff, azcc, azfc, azcf, azff, fff;
  std::vector<FT> zf, zc, dzc, dzf;
  std::vector<int> kbot;
  std::vector<FT> Hcc, Hfc, Hcf;
  std::vector<FT> wts;
  std::vector<FT> fld[F_COUNT];
  std::vector<FT> bflux[4][2];   // A7: flux boundary conditions of u, v, T, S at the bottom (0) / top (1); empty = no-flux
  double time = 0;
  long iteration = 0;
  FT last_dt = 0;
  FT eps;

  // ---------------------------------------------------------------- indexing
  inline size_t id2(int i, int j) const { return (size_t)(i + Hx - 1) + (size_t)PX * (size_t)(j + Hy - 1); }
  inline size_t id3(int i, int j, int k) const { return id2(i, j) + n2 * (size_t)(k + Hz - 1); }
  inline FT& at(int f, int i, int j, int k) { return fld[f][id3(i, j, k)]; }
  inline FT at(int f, int i, int j, int k) const { return fld[f][id3(i, j, k)]; }
  inline FT& at2(int f, int i, int j) { return fld[f][id2(i, j)]; }
  inline FT at2(int f, int i, int j) const { return fld[f][id2(i, j)]; }
  inline FT Dzc(int k) const { return dzc[k + Hz - 1]; }
  inline FT Dzf(int k) const { return dzf[k + Hz - 1]; }
  inline FT Zc(int k) const { return zc[k + Hz - 1]; }
  static bool is3d(int f) { return f < F_ETA; }

  Oracle(const OConfig& cfg, const FT* const* g2, const FT* const* gz, const FT* bottom, const FT* weights)
      : c(cfg) {
    Nx = c.Nx; Ny = c.Ny; Nz = c.Nz; Hx = c.Hx; Hy = c.Hy; Hz = c.Hz;
    PX = Nx + 2 * Hx; PY = Ny + 2 * Hy + 1; PZ = Nz + 2 * Hz + 1;
    n2 = (size_t)PX * PY; n3 = n2 * PZ;
    std::vector<FT>* m2[13] = {&dxcc, &dxfc, &dxcf, &dxff, &dycc, &dyfc, &dycf, &dyff,
                               &azcc, &azfc, &azcf, &azff, &fff};
    for (int a = 0; a < 13; a++) m2[a]->assign(g2[a], g2[a] + n2);
    std::vector<FT>* mz[4] = {&zf, &zc, &dzc, &dzf};
    for (int a = 0; a < 4; a++) mz[a]->assign(gz[a], gz[a] + PZ);
    wts.assign(weights, weights + c.nsub);
    eps = (FT)c.weno_eps;
    for (int f = 0; f < F_COUNT; f++) fld[f].assign(is3d(f) ? n3 : n2, (FT)0);
    // --- immersed boundary products (ImmersedBoundaries/grid_fitted_bottom.jl [OCN-recall])
    kbot.assign(n2, 0);
    Hcc.assign(n2, 0); Hfc.assign(n2, 0); Hcf.assign(n2, 0);
    FT ztop = zf[Nz + 1 + Hz - 1], zbot = zf[1 + Hz - 1];
    for (size_t q = 0; q < n2; q++) {
      if (c.immersed && bottom) {
        FT bh = std::min(std::max(bottom[q], zbot), ztop);
        int kb = 0;
        for (int k = 1; k <= Nz; k++) if (Zc(k) <= bh) kb = k;  // z_c is increasing: prefix of solid cells
        kbot[q] = kb;
        // snapped bottom height = top face of the highest solid cell (CenterImmersedCondition)
        FT snapped = zf[kb + 1 + Hz - 1];
        Hcc[q] = ztop - snapped;
      } else {
        Hcc[q] = ztop - zbot;
      }
    }
    for (int j = 1 - Hy; j <= Ny + Hy + 1; j++)
      for (int i = 1 - Hx; i <= Nx + Hx; i++) {
        int im = std::max(i - 1, 1 - Hx), jm = std::max(j - 1, 1 - Hy);
        Hfc[id2(i, j)] = std::min(Hcc[id2(im, j)], Hcc[id2(i, j)]);
        Hcf[id2(i, j)] = std::min(Hcc[id2(i, jm)], Hcc[id2(i, j)]);
      }
  }

  // ------------------------------------------------- immersed predicates
  inline bool outside(int, int j, int k) const {
    if (k < 1 || k > Nz) return true;
    if (c.topo_y == TOPO_BOUNDED) return (j < 1 || j > Ny);
    return c.south_inactive ? (j < 1) : false;
  }
  inline bool immersed_cell(int i, int j, int k) const { return c.immersed && k <= kbot[id2(i, j)]; }
  inline bool inactive(int i, int j, int k) const { return outside(i, j, k) || immersed_cell(i, j, k); }
  // peripheral_node: any adjacent cell inactive; inactive_node: all adjacent cells inactive
  inline bool peripheral(Loc l, int i, int j, int k, bool underlying = false) const {
    bool r = false;
    for (int di = 0; di <= l.x; di++)
      for (int dj = 0; dj <= l.y; dj++)
        for (int dk = 0; dk <= l.z; dk++)
          r = r || (underlying ? outside(i - di, j - dj, k - dk) : inactive(i - di, j - dj, k - dk));
    return r;
  }
  inline bool inactive_node(Loc l, int i, int j, int k) const {
    bool r = true;
    for (int di = 0; di <= l.x; di++)
      for (int dj = 0; dj <= l.y; dj++)
        for (int dk = 0; dk <= l.z; dk++) r = r && inactive(i - di, j - dj, k - dk);
    return r;
  }
  inline bool immersed_peripheral(Loc l, int i, int j, int k) const {
    if (!c.immersed) return false;
    return peripheral(l, i, j, k, false) && !peripheral(l, i, j, k, true);
  }

  // ------------------------------------------------- reconstructions
  // order reduction: the largest buffer B <= Bmax such that the window along `dir` is clear.
  // Face reconstruction at index n: cells n-B .. n+B-1 must be active.
  // plain (non-immersed) grids: topological rule, only the index along `dir` matters
  inline int topo_buffer(int dir, int n, int Bmax, bool center) const {
    if (dir == 0) return Bmax;  // periodic x
    int N = dir == 1 ? Ny : Nz;
    bool lower_only = false;
    if (dir == 1 && c.topo_y == TOPO_FOLD) { if (!c.south_inactive) return Bmax; lower_only = true; }
    for (int B = Bmax; B >= 2; B--) {
      bool lo = center ? (n >= B) : (n >= B + 1);
      bool hi = lower_only ? true : (n <= N + 1 - B);
      if (lo && hi) return B;
    }
    return 1;
  }
  inline int face_buffer(int dir, int i, int j, int k, int Bmax) const {
    if (!c.immersed) return topo_buffer(dir, dir == 0 ? i : (dir == 1 ? j : k), Bmax, false);
    for (int B = Bmax; B >= 2; B--) {
      bool near = false;
      for (int m = -B; m <= B - 1; m++)
        near = near || inactive(i + (dir == 0) * m, j + (dir == 1) * m, k + (dir == 2) * m);
      if (!near) return B;
    }
    return 1;
  }
  // Centre reconstruction at index n from Face data: faces n-B+1 .. n+B must not be inactive nodes.
  inline int center_buffer(int dir, int i, int j, int k, int Bmax) const {
    if (!c.immersed) return topo_buffer(dir, dir == 0 ? i : (dir == 1 ? j : k), Bmax, true);
    Loc l{dir == 0, dir == 1, dir == 2};
    for (int B = Bmax; B >= 2; B--) {
      bool near = false;
      for (int m = -B + 1; m <= B; m++)
        near = near || inactive_node(l, i + (dir == 0) * m, j + (dir == 1) * m, k + (dir == 2) * m);
      if (!near) return B;
    }
    return 1;
  }

  inline FT beta3(int r, FT a, FT b, FT cc) const {
    if (c.beta_form == 1) {
      FT d2 = (a - (FT)2 * b) + cc;
      FT d1 = r == 0 ? ((FT)3 * a - (FT)4 * b) + cc : (r == 1 ? a - cc : (a - (FT)4 * b) + (FT)3 * cc);
      return (FT)3.25 * d2 * d2 + (FT)0.75 * d1 * d1;
    }
    switch (r) {
      case 0: return a * ((FT)10 * a - (FT)31 * b + (FT)11 * cc) + b * ((FT)25 * b - (FT)19 * cc) + (FT)4 * cc * cc;
      case 1: return a * ((FT)4 * a - (FT)13 * b + (FT)5 * cc) + b * ((FT)13 * b - (FT)13 * cc) + (FT)4 * cc * cc;
      default: return a * ((FT)4 * a - (FT)19 * b + (FT)11 * cc) + b * ((FT)25 * b - (FT)31 * cc) + (FT)10 * cc * cc;
    }
  }
  static inline FT cand3(int r, FT a, FT b, FT cc) {
    switch (r) {
      case 0: return (FT)(1.0 / 3) * a + (FT)(5.0 / 6) * b - (FT)(1.0 / 6) * cc;
      case 1: return -(FT)(1.0 / 6) * a + (FT)(5.0 / 6) * b + (FT)(1.0 / 3) * cc;
      default: return (FT)(1.0 / 3) * a - (FT)(7.0 / 6) * b + (FT)(11.0 / 6) * cc;
    }
  }
  // q: window of 2B values centred on the face (cells n-B .. n+B-1); s1,s2: smoothness sources
  // (s1 == nullptr -> q itself; s2 != nullptr -> VelocityStencil average of the two).
  inline FT weno_window(int B, bool left, const FT* q, const FT* s1, const FT* s2) const {
    FT qq[6], a1[6], a2[6];
    int W = 2 * B;
    for (int m = 0; m < W; m++) {
      int src = left ? m : W - 1 - m;
      qq[m] = q[src];
      a1[m] = s1 ? s1[src] : q[src];
      a2[m] = s2 ? s2[src] : (FT)0;
    }
    if (B == 1) return qq[0];
    if (B == 2) {
      // WENO3-Z: S0 = (q1,q2), S1 = (q0,q1)
      auto b2f = [&](FT x, FT y) { return c.beta_form == 1 ? (x - y) * (x - y) : x * (x - (FT)2 * y) + y * y; };
      FT b0 = b2f(a1[1], a1[2]);
      FT b1 = b2f(a1[0], a1[1]);
      if (s2) {
        FT c0 = b2f(a2[1], a2[2]);
        FT c1 = b2f(a2[0], a2[1]);
        b0 = (b0 + c0) / (FT)2; b1 = (b1 + c1) / (FT)2;
      }
      FT tau = std::fabs(b0 - b1);
      FT r0 = tau / (b0 + eps), r1 = tau / (b1 + eps);
      FT al0 = (FT)(2.0 / 3) * ((FT)1 + r0 * r0), al1 = (FT)(1.0 / 3) * ((FT)1 + r1 * r1);
      FT p0 = (FT)0.5 * qq[1] + (FT)0.5 * qq[2];
      FT p1 = -(FT)0.5 * qq[0] + (FT)1.5 * qq[1];
      FT s = al0 + al1;
      return (al0 / s) * p0 + (al1 / s) * p1;
    }
    // WENO5-Z: S0 = (q2,q3,q4), S1 = (q1,q2,q3), S2 = (q0,q1,q2)
    FT be[3] = {0, 0, 0}, al[3], p[3];
    for (int r = 0; r < 3; r++) {
      be[r] = beta3(r, a1[2 - r], a1[3 - r], a1[4 - r]);
      if (s2) be[r] = (be[r] + beta3(r, a2[2 - r], a2[3 - r], a2[4 - r])) / (FT)2;
      p[r] = cand3(r, qq[2 - r], qq[3 - r], qq[4 - r]);
    }
    FT tau = std::fabs(be[0] - be[2]);
    const FT C[3] = {(FT)(3.0 / 10), (FT)(3.0 / 5), (FT)(1.0 / 10)};
    FT s = 0;
    for (int r = 0; r < 3; r++) {
      FT t = tau / (be[r] + eps);
      al[r] = C[r] * ((FT)1 + t * t);
      s += al[r];
    }
    return (al[0] / s) * p[0] + (al[1] / s) * p[1] + (al[2] / s) * p[2];
  }
  // generic biased reconstruction at "face" index n from a sampler q(m), m = index along dir
  template <class Q, class S1, class S2>
  inline FT biased(int B, bool left, int n, Q q, S1 s1, S2 s2, int mode) const {
    FT qw[6], a1[6], a2[6];
    for (int m = 0; m < 2 * B; m++) {
      qw[m] = q(n - B + m);
      if (mode >= 1) a1[m] = s1(n - B + m);
      if (mode == 2) a2[m] = s2(n - B + m);
    }
    return weno_window(B, left, qw, mode >= 1 ? a1 : nullptr, mode == 2 ? a2 : nullptr);
  }
  template <class Q>
  inline FT biased(int B, bool left, int n, Q q) const {
    return biased(B, left, n, q, q, q, 0);
  }
  // symmetric (centred) reconstruction at face n: B=2 -> 4th order, B=1 -> 2nd order
  template <class Q>
  inline FT symmetric(int B, int n, Q q) const {
    if (B >= 2)
      return -(FT)(1.0 / 12) * q(n - 2) + (FT)(7.0 / 12) * q(n - 1) + (FT)(7.0 / 12) * q(n) - (FT)(1.0 / 12) * q(n + 1);
    return (FT)0.5 * q(n - 1) + (FT)0.5 * q(n);
  }

  // ------------------------------------------------- TEOS-10 (Roquet et al. 2015, 55-term)
  inline FT rho_prime(FT Theta, FT SA, FT Z) const {
    const FT tau = Theta / (FT)40;
    const FT s = std::sqrt((SA + (FT)32) / (FT)(40.0 * 35.16504 / 35.0));
    const FT z = -Z / (FT)1e4;
    FT r3 = (FT)-2.3342758797e-02 + (FT)-1.8507636718e-02 * s + (FT)3.7969820455e-01 * tau;
    FT r2 = (FT)2.0660924175 + s * ((FT)-4.9527603989 + s * (FT)2.5019633244) +
            tau * ((FT)2.0564311499 + s * (FT)-2.1311365518e-01 + tau * (FT)-1.2419983026);
    FT r1 = (FT)1.9681925209e+01 +
            s * ((FT)-4.2549998214e+01 + s * ((FT)5.0774768218e+01 + s * ((FT)-3.0938076334e+01 + s * (FT)6.6051753097))) +
            tau * ((FT)-1.3336301113e+01 + s * ((FT)-4.4870114575 + s * ((FT)5.0042598061 + s * (FT)-6.5399043664e-01)) +
                   tau * ((FT)6.7080479603 + s * ((FT)3.5063081279 + s * (FT)-1.8795372996) +
                          tau * ((FT)-2.4649669534 + s * (FT)-5.5077101279e-01 + tau * (FT)5.5927935970e-01)));
    FT r0 = (FT)8.0189615746e+02 +
            s * ((FT)8.6672408165e+02 + s * ((FT)-1.7864682637e+03 + s * ((FT)2.0375295546e+03 + s * ((FT)-1.2849161071e+03 + s * ((FT)4.3227585684e+02 + s * (FT)-6.0579916612e+01))))) +
            tau * ((FT)2.6010145068e+01 + s * ((FT)-6.5281885265e+01 + s * ((FT)8.1770425108e+01 + s * ((FT)-5.6888046321e+01 + s * ((FT)1.7681814114e+01 + s * (FT)-1.9193502195)))) +
                   tau * ((FT)-3.7074170417e+01 + s * ((FT)6.1548258127e+01 + s * ((FT)-6.0362551501e+01 + s * ((FT)2.9130021253e+01 + s * (FT)-5.4723692739))) +
                          tau * ((FT)2.1661789529e+01 + s * ((FT)-3.3449108469e+01 + s * ((FT)1.9717078466e+01 + s * (FT)-3.1742946532)) +
                                 tau * ((FT)-8.3627885467 + s * ((FT)1.1311538584e+01 + s * (FT)-5.3563304045) +
                                        tau * ((FT)5.4048723791e-01 + s * (FT)4.8169980163e-01 + tau * (FT)-1.9083568888e-01)))));
    FT r = ((r3 * z + r2) * z + r1) * z + r0;
    if (c.eos_r0) {
      FT rz = ((((((FT)-1.7243708991e-03 * z + (FT)1.5616995503e-02) * z + (FT)6.4326772569e-02) * z + (FT)2.2601900708e-01) * z - (FT)5.2099962525) * z + (FT)4.6494977072e+01) * z;
      r += rz;
    }
    return r - (FT)c.rho0;
  }
  inline FT buoyancy(int i, int j, int k) const {
    return -((FT)c.g * rho_prime(at(F_T, i, j, k), at(F_S, i, j, k), Zc(k)) / (FT)c.rho0);
  }

  // ------------------------------------------------- halo fills, bit-exact contract
  // loc: staggering; sign: -1 for vector components under the fold; three_d: has z halos
  void fill_halo(int f, Loc l, FT sign, bool three_d) {
    std::vector<FT>& a = fld[f];
    int k0 = three_d ? 1 : 1, k1 = three_d ? Nz + (l.z ? 1 : 0) : 1;
    auto A = [&](int i, int j, int k) -> FT& { return three_d ? a[id3(i, j, k)] : a[id2(i, j)]; };
    // --- south / north over interior i, k
    for (int k = k0; k <= k1; k++)
      for (int i = 1; i <= Nx; i++) {
        // south (both topologies behave as Bounded there)
        if (l.y == 0) for (int m = 1; m <= Hy; m++) A(i, 1 - m, k) = A(i, m, k);
        else A(i, 1, k) = 0;
        if (c.topo_y == TOPO_BOUNDED) {
          if (l.y == 0) for (int m = 1; m <= Hy; m++) A(i, Ny + m, k) = A(i, Ny + 1 - m, k);
          else A(i, Ny + 1, k) = 0;
        }
      }
    if (c.topo_y == TOPO_FOLD) {
      // zipper: read only interior rows <= Ny, write rows > Ny (and optionally half of row Ny)
      for (int k = k0; k <= k1; k++)
        for (int i = 1; i <= Nx; i++) {
          int ip; FT sg = sign;
          if (l.x == 0) ip = Nx - i + 1;
          else { ip = Nx - i + 2; if (ip > Nx) { ip -= Nx; sg = std::fabs(sign); } }
          for (int m = 1; m <= Hy; m++) {
            int js = (l.y == 0) ? Ny - m : Ny - m + 1;
            A(i, Ny + m, k) = sg * A(ip, js, k);
          }
        }
      if (c.fold_variant == 1 && l.y == 0)
        for (int k = k0; k <= k1; k++)
          for (int i = Nx / 2 + 1; i <= Nx; i++) {
            int ip; FT sg = sign;
            if (l.x == 0) ip = Nx - i + 1;
            else { ip = Nx - i + 2; if (ip > Nx) { ip -= Nx; sg = std::fabs(sign); } }
            A(i, Ny, k) = sg * A(ip, Ny, k);
          }
    }
    // --- bottom / top over interior i, j
    if (three_d) {
      int jt = Ny + ((l.y && c.topo_y == TOPO_BOUNDED) ? 1 : 0);
      for (int j = 1; j <= jt; j++)
        for (int i = 1; i <= Nx; i++) {
          if (l.z == 0) {
            for (int m = 1; m <= Hz; m++) { A(i, j, 1 - m) = A(i, j, m); A(i, j, Nz + m) = A(i, j, Nz + 1 - m); }
          } else { A(i, j, 1) = 0; A(i, j, Nz + 1) = 0; }
        }
    }
    // --- periodic x over the full parent extent in j, k
    int K0 = three_d ? 1 - Hz : 1, K1 = three_d ? Nz + Hz + 1 : 1;
    for (int k = K0; k <= K1; k++)
      for (int j = 1 - Hy; j <= Ny + Hy + 1; j++)
        for (int m = 1; m <= Hx; m++) {
          A(1 - m, j, k) = A(Nx + 1 - m, j, k);
          A(Nx + m, j, k) = A(m, j, k);
        }
  }
  void fill_halo_prognostic() {
    fill_halo(F_U, {1, 0, 0}, -1, true);
    fill_halo(F_V, {0, 1, 0}, -1, true);
    fill_halo(F_T, {0, 0, 0}, 1, true);
    fill_halo(F_S, {0, 0, 0}, 1, true);
    fill_halo(F_ETA, {0, 0, 1}, 1, false);
    fill_halo(F_BU, {1, 0, 0}, -1, false);
    fill_halo(F_BV, {0, 1, 0}, -1, false);
  }

  // ------------------------------------------------- mask_immersed_field! (row A1)
  void mask_field(int f, Loc l) {
    if (!c.immersed) return;
    for (int k = 1; k <= Nz; k++)
      for (int j = 1; j <= Ny; j++)
        for (int i = 1; i <= Nx; i++)
          if (peripheral(l, i, j, k)) at(f, i, j, k) = 0;
  }
  // barotropic transports are prognostic fields too (decision U11): masked where the surface-level
  // velocity node is peripheral, i.e. where the column depth at the node is zero
  void mask_barotropic() {
    if (!c.immersed) return;
    for (int j = 1; j <= Ny; j++)
      for (int i = 1; i <= Nx; i++) {
        if (peripheral({1, 0, 0}, i, j, Nz)) at2(F_BU, i, j) = 0;
        if (peripheral({0, 1, 0}, i, j, Nz)) at2(F_BV, i, j) = 0;
      }
  }
  void mask_immersed_fields() {
    mask_field(F_U, {1, 0, 0}); mask_field(F_V, {0, 1, 0});
    mask_field(F_T, {0, 0, 0}); mask_field(F_S, {0, 0, 0});
    mask_barotropic();
  }

  // ------------------------------------------------- compute_w_from_continuity! (row A3)
  inline FT flux_div_xy(int i, int j, int k) const {
    FT dz = Dzc(k);
    FT ax1 = dyfc[id2(i + 1, j)] * dz * at(F_U, i + 1, j, k), ax0 = dyfc[id2(i, j)] * dz * at(F_U, i, j, k);
    FT ay1 = dxcf[id2(i, j + 1)] * dz * at(F_V, i, j + 1, k), ay0 = dxcf[id2(i, j)] * dz * at(F_V, i, j, k);
    return (ax1 - ax0) + (ay1 - ay0);
  }
  void compute_w() {
#pragma omp parallel for
    for (int j = -Hy + 2; j <= Ny + Hy - 1; j++)
      for (int i = -Hx + 2; i <= Nx + Hx - 1; i++) {
        FT w = 0;
        at(F_W, i, j, 1) = 0;
        for (int k = 2; k <= Nz + 1; k++) {
          w = w - flux_div_xy(i, j, k - 1) / azcc[id2(i, j)];
          at(F_W, i, j, k) = w;
        }
      }
  }
  // ------------------------------------------------- update_hydrostatic_pressure! (row A4)
  void compute_p() {
#pragma omp parallel for
    for (int j = 0; j <= Ny + 1; j++)
      for (int i = 0; i <= Nx + 1; i++) {
        FT bup = buoyancy(i, j, Nz + 1);
        FT p = 0;
        for (int k = Nz; k >= 1; k--) {
          FT b = buoyancy(i, j, k);
          FT bbar = (b + bup) / (FT)2;  // ℑz b at face k+1
          p = (k == Nz ? -bbar * Dzf(k + 1) : p - bbar * Dzf(k + 1));
          at(F_P, i, j, k) = p;
          bup = b;
        }
      }
  }

  // ------------------------------------------------- tracer tendency (row A6)
  inline FT tracer_flux(int dir, int f, int i, int j, int k) const {
    Loc l{dir == 0, dir == 1, dir == 2};
    if (immersed_peripheral(l, i, j, k)) return 0;
    FT vel, area;
    if (dir == 0) { vel = at(F_U, i, j, k); area = dyfc[id2(i, j)] * Dzc(k); }
    else if (dir == 1) { vel = at(F_V, i, j, k); area = dxcf[id2(i, j)] * Dzc(k); }
    else { vel = at(F_W, i, j, k); area = azcc[id2(i, j)]; }
    int n = dir == 0 ? i : (dir == 1 ? j : k);
    int B = face_buffer(dir, i, j, k, 3);
    auto q = [&](int m) { return dir == 0 ? at(f, m, j, k) : (dir == 1 ? at(f, i, m, k) : at(f, i, j, m)); };
    FT cr = biased(B, vel > 0, n, q);
    return area * vel * cr;
  }
  void tracer_tendency(int fc, int fg) {
#pragma omp parallel for collapse(2)
    for (int k = 1; k <= Nz; k++)
      for (int j = 1; j <= Ny; j++)
        for (int i = 1; i <= Nx; i++) {
          FT dfx = tracer_flux(0, fc, i + 1, j, k) - tracer_flux(0, fc, i, j, k);
          FT dfy = tracer_flux(1, fc, i, j + 1, k) - tracer_flux(1, fc, i, j, k);
          FT dfz = tracer_flux(2, fc, i, j, k + 1) - tracer_flux(2, fc, i, j, k);
          FT V = azcc[id2(i, j)] * Dzc(k);
          at(fg, i, j, k) = -(((FT)1 / V) * (dfx + dfy + dfz));
        }
  }

  // ------------------------------------------------- momentum tendency (row A5)
  inline FT U(int i, int j, int k) const { return at(F_U, i, j, k); }
  inline FT Vv(int i, int j, int k) const { return at(F_V, i, j, k); }
  inline FT zeta(int i, int j, int k) const {
    const Loc cfc{0, 1, 0}, fcc{1, 0, 0};
    FT dxv = dycf[id2(i, j)] * Vv(i, j, k) - dycf[id2(i - 1, j)] * Vv(i - 1, j, k);
    FT dyu = dxfc[id2(i, j)] * U(i, j, k) - dxfc[id2(i, j - 1)] * U(i, j - 1, k);
    if (c.immersed && c.cond_diff) {
      if (inactive_node(cfc, i, j, k) || inactive_node(cfc, i - 1, j, k)) dxv = 0;
      if (inactive_node(fcc, i, j, k) || inactive_node(fcc, i, j - 1, k)) dyu = 0;
    }
    return (dxv - dyu) / azff[id2(i, j)];
  }
  inline FT div_xy_cc(int i, int j, int k) const { return flux_div_xy(i, j, k); }  // δx(Ax u) + δy(Ay v) at (C,C,C)
  inline FT dxU(int i, int j, int k) const {
    FT dz = Dzc(k);
    return dyfc[id2(i + 1, j)] * dz * U(i + 1, j, k) - dyfc[id2(i, j)] * dz * U(i, j, k);
  }
  inline FT dyV(int i, int j, int k) const {
    FT dz = Dzc(k);
    return dxcf[id2(i, j + 1)] * dz * Vv(i, j + 1, k) - dxcf[id2(i, j)] * dz * Vv(i, j, k);
  }
  inline FT half_sq(FT a) const { return a * a / (FT)2; }

  FT gu(int i, int j, int k) const {
    const FT uc = U(i, j, k);
    // horizontal: -v̂ ζᴿ
    FT vavg = ((dxcf[id2(i - 1, j)] * Vv(i - 1, j, k) + dxcf[id2(i - 1, j + 1)] * Vv(i - 1, j + 1, k)) / (FT)2 +
               (dxcf[id2(i, j)] * Vv(i, j, k) + dxcf[id2(i, j + 1)] * Vv(i, j + 1, k)) / (FT)2) / (FT)2;  // ℑx ℑy (Δx v)
    FT vhat = vavg / dxfc[id2(i, j)];
    int Bc = center_buffer(1, i, j, k, 3);
    FT zR = biased(Bc, vhat > 0, j + 1,
                   [&](int m) { return zeta(i, m, k); },
                   [&](int m) { return (U(i, m - 1, k) + U(i, m, k)) / (FT)2; },
                   [&](int m) { return (Vv(i - 1, m, k) + Vv(i, m, k)) / (FT)2; }, 2);
    FT HU = -vhat * zR;
    // vertical + divergence (self-upwinding)
    int Bs = face_buffer(0, i, j, k, 2);
    FT dvs = symmetric(Bs, i, [&](int m) { return dyV(m, j, k); });
    int Bf = face_buffer(0, i, j, k, 3);
    FT duR = biased(Bf, uc > 0, i, [&](int m) { return dxU(m, j, k); },
                    [&](int m) { return div_xy_cc(m, j, k); }, [&](int) { return (FT)0; }, 1);
    FT Phi = uc * (dvs + duR);
    auto Wu = [&](int kk) -> FT {
      if (immersed_peripheral({1, 0, 1}, i, j, kk)) return 0;
      int Bw = face_buffer(0, i, j, kk, 2);
      FT wt = symmetric(Bw, i, [&](int m) { return azcc[id2(m, j)] * at(F_W, m, j, kk); });
      int Bz = face_buffer(2, i, j, kk, 3);
      FT uR = biased(Bz, wt > 0, kk, [&](int m) { return U(i, j, m); });
      return wt * uR;
    };
    FT Az_ = Wu(k + 1) - Wu(k);
    FT Vfcc = azfc[id2(i, j)] * Dzc(k);
    FT VU = ((FT)1 / Vfcc) * (Phi + Az_);
    // Bernoulli head
    FT dKu = biased(Bf, uc > 0, i, [&](int m) { return half_sq(U(m + 1, j, k)) - half_sq(U(m, j, k)); },
                    [&](int m) { return (U(m, j, k) + U(m + 1, j, k)) / (FT)2; }, [&](int) { return (FT)0; }, 1);
    int Bsy = center_buffer(1, i, j, k, 2);
    FT dKv = symmetric(Bsy, j + 1, [&](int m) { return half_sq(Vv(i, m, k)) - half_sq(Vv(i - 1, m, k)); });
    FT BU = (dKu + dKv) / dxfc[id2(i, j)];
    // Coriolis
    FT fbar = (fff[id2(i, j)] + fff[id2(i, j + 1)]) / (FT)2;
    FT vsum4 = vavg;
    FT cor;
    if (c.coriolis_scheme == 1) {
      const Loc cfc{0, 1, 0};
      int nact = (!peripheral(cfc, i - 1, j, k)) + (!peripheral(cfc, i - 1, j + 1, k)) +
                 (!peripheral(cfc, i, j, k)) + (!peripheral(cfc, i, j + 1, k));
      FT wgt = (FT)nact / (FT)4;
      FT avg = nact == 0 ? (FT)0 : vsum4 / wgt;
      cor = -fbar * avg / dxfc[id2(i, j)];
    } else cor = -fbar * vsum4 / dxfc[id2(i, j)];
    // pressure gradient
    FT dpx = (at(F_P, i, j, k) - at(F_P, i - 1, j, k)) / dxfc[id2(i, j)];
    if (c.immersed && c.cond_diff && (inactive(i, j, k) || inactive(i - 1, j, k))) dpx = 0;
    return -(HU + VU + BU) - cor - dpx;
  }

  FT gv(int i, int j, int k) const {
    const FT vc = Vv(i, j, k);
    FT uavg = ((dyfc[id2(i, j - 1)] * U(i, j - 1, k) + dyfc[id2(i + 1, j - 1)] * U(i + 1, j - 1, k)) / (FT)2 +
               (dyfc[id2(i, j)] * U(i, j, k) + dyfc[id2(i + 1, j)] * U(i + 1, j, k)) / (FT)2) / (FT)2;  // ℑy ℑx (Δy u)
    FT uhat = uavg / dycf[id2(i, j)];
    int Bc = center_buffer(0, i, j, k, 3);
    FT zR = biased(Bc, uhat > 0, i + 1,
                   [&](int m) { return zeta(m, j, k); },
                   [&](int m) { return (U(m, j - 1, k) + U(m, j, k)) / (FT)2; },
                   [&](int m) { return (Vv(m - 1, j, k) + Vv(m, j, k)) / (FT)2; }, 2);
    FT HV = uhat * zR;
    int Bs = face_buffer(1, i, j, k, 2);
    FT dus = symmetric(Bs, j, [&](int m) { return dxU(i, m, k); });
    int Bf = face_buffer(1, i, j, k, 3);
    FT dvR = biased(Bf, vc > 0, j, [&](int m) { return dyV(i, m, k); },
                    [&](int m) { return div_xy_cc(i, m, k); }, [&](int) { return (FT)0; }, 1);
    FT Phi = vc * (dus + dvR);
    auto Wv = [&](int kk) -> FT {
      if (immersed_peripheral({0, 1, 1}, i, j, kk)) return 0;
      int Bw = face_buffer(1, i, j, kk, 2);
      FT wt = symmetric(Bw, j, [&](int m) { return azcc[id2(i, m)] * at(F_W, i, m, kk); });
      int Bz = face_buffer(2, i, j, kk, 3);
      FT vR = biased(Bz, wt > 0, kk, [&](int m) { return Vv(i, j, m); });
      return wt * vR;
    };
    FT Az_ = Wv(k + 1) - Wv(k);
    FT Vcfc = azcf[id2(i, j)] * Dzc(k);
    FT VV = ((FT)1 / Vcfc) * (Phi + Az_);
    FT dKv = biased(Bf, vc > 0, j, [&](int m) { return half_sq(Vv(i, m + 1, k)) - half_sq(Vv(i, m, k)); },
                    [&](int m) { return (Vv(i, m, k) + Vv(i, m + 1, k)) / (FT)2; }, [&](int) { return (FT)0; }, 1);
    int Bsx = center_buffer(0, i, j, k, 2);
    FT dKu = symmetric(Bsx, i + 1, [&](int m) { return half_sq(U(m, j, k)) - half_sq(U(m, j - 1, k)); });
    FT BV = (dKv + dKu) / dycf[id2(i, j)];
    FT fbar = (fff[id2(i, j)] + fff[id2(i + 1, j)]) / (FT)2;
    FT usum4 = uavg;
    FT cor;
    if (c.coriolis_scheme == 1) {
      const Loc fcc{1, 0, 0};
      int nact = (!peripheral(fcc, i, j - 1, k)) + (!peripheral(fcc, i + 1, j - 1, k)) +
                 (!peripheral(fcc, i, j, k)) + (!peripheral(fcc, i + 1, j, k));
      FT wgt = (FT)nact / (FT)4;
      FT avg = nact == 0 ? (FT)0 : usum4 / wgt;
      cor = fbar * avg / dycf[id2(i, j)];
    } else cor = fbar * usum4 / dycf[id2(i, j)];
    FT dpy = (at(F_P, i, j, k) - at(F_P, i, j - 1, k)) / dycf[id2(i, j)];
    if (c.immersed && c.cond_diff && (inactive(i, j, k) || inactive(i, j - 1, k))) dpy = 0;
    return -(HV + VV + BV) - cor - dpy;
  }
  void momentum_tendency() {
#pragma omp parallel for collapse(2)
    for (int k = 1; k <= Nz; k++)
      for (int j = 1; j <= Ny; j++)
        for (int i = 1; i <= Nx; i++) {
          at(F_GNU, i, j, k) = gu(i, j, k);
          at(F_GNV, i, j, k) = gv(i, j, k);
        }
  }
  // ------------------------------------------------- vertical diffusion (row A13)
  // VerticalScalarDiffusivity(kappa, nu), constant coefficients.  Explicit: G += (1/V) delta_z(Az K d_z c) with zero
  // flux through peripheral faces (walls, bathymetry).  Implicit: after the AB2 update solve, column by column,
  // (I - dt d_z K d_z) c = c* with the Thomas algorithm of Oceananigans' batched tridiagonal solver.
  void vertical_diffusion_explicit(int fc, int fg, Loc l, FT K) {
    const Loc lf{l.x, l.y, 1};
#pragma omp parallel for collapse(2)
    for (int k = 1; k <= Nz; k++)
      for (int j = 1; j <= Ny; j++)
        for (int i = 1; i <= Nx; i++) {
          FT qt = peripheral(lf, i, j, k + 1) ? (FT)0 : K * (at(fc, i, j, k + 1) - at(fc, i, j, k)) / Dzf(k + 1);
          FT qb = peripheral(lf, i, j, k) ? (FT)0 : K * (at(fc, i, j, k) - at(fc, i, j, k - 1)) / Dzf(k);
          at(fg, i, j, k) = at(fg, i, j, k) + (qt - qb) / Dzc(k);
        }
  }
  void implicit_vertical_diffusion(int fc, Loc l, FT K, FT dt) {
    const Loc lf{l.x, l.y, 1};
#pragma omp parallel for
    for (int j = 1; j <= Ny; j++) {
      std::vector<FT> t(Nz + 2);
      for (int i = 1; i <= Nx; i++) {
        auto upper = [&](int k) -> FT {   // couples k and k+1 through face k+1
          if (k > Nz - 1) return 0;
          return peripheral(lf, i, j, k + 1) ? (FT)0 : -dt * K / (Dzc(k) * Dzf(k + 1));
        };
        auto lower = [&](int kp) -> FT {  // kp = k-1: couples k and k-1 through face k
          if (kp < 1) return 0;
          int k = kp + 1;
          return peripheral(lf, i, j, k) ? (FT)0 : -dt * K / (Dzc(k) * Dzf(k));
        };
        auto diag = [&](int k) -> FT { return (FT)1 - upper(k) - lower(k - 1); };
        FT beta = diag(1);
        at(fc, i, j, 1) = at(fc, i, j, 1) / beta;
        for (int k = 2; k <= Nz; k++) {
          t[k] = upper(k - 1) / beta;
          beta = diag(k) - lower(k - 1) * t[k];
          at(fc, i, j, k) = (at(fc, i, j, k) - lower(k - 1) * at(fc, i, j, k - 1)) / beta;
        }
        for (int k = Nz - 1; k >= 1; k--) at(fc, i, j, k) -= t[k + 1] * at(fc, i, j, k + 1);
      }
    }
  }
  // ------------------------------------------------- boundary tendency contributions (row A7)
  // compute_hydrostatic_boundary_tendency_contributions! -> apply_z_bcs! [OCN-recall: BoundaryConditions/apply_flux_bcs.jl]:
  //   Gc[i,j,1]  += J_bottom * Az / V(i,j,1),   Gc[i,j,Nz] -= J_top * Az / V(i,j,Nz)      (a positive top flux leaves the domain)
  // with Az and V at the location of the field.  Reference call site: /root/reference/src/precompile.jl:52-61.
  void boundary_tendencies() {
    const int fg[4] = {F_GNU, F_GNV, F_GNT, F_GNS};
    const std::vector<FT>* az[4] = {&azfc, &azcf, &azcc, &azcc};
    for (int q = 0; q < 4; q++)
      for (int side = 0; side < 2; side++) {
        if (bflux[q][side].empty()) continue;
        const int k = side ? Nz : 1;
#pragma omp parallel for
        for (int j = 1; j <= Ny; j++)
          for (int i = 1; i <= Nx; i++) {
            const FT A = (*az[q])[id2(i, j)], V = A * Dzc(k);
            const FT d = bflux[q][side][id2(i, j)] * A / V;
            if (side) at(fg[q], i, j, k) -= d; else at(fg[q], i, j, k) += d;
          }
      }
  }
  void compute_tendencies() {
    momentum_tendency();
    tracer_tendency(F_T, F_GNT);
    tracer_tendency(F_S, F_GNS);
    if (c.closure == 1) {
      vertical_diffusion_explicit(F_U, F_GNU, {1, 0, 0}, (FT)c.nu);
      vertical_diffusion_explicit(F_V, F_GNV, {0, 1, 0}, (FT)c.nu);
      vertical_diffusion_explicit(F_T, F_GNT, {0, 0, 0}, (FT)c.kappa);
      vertical_diffusion_explicit(F_S, F_GNS, {0, 0, 0}, (FT)c.kappa);
    }
    boundary_tendencies();
  }
  void compute_auxiliaries() { compute_w(); compute_p(); }
  void update_state() {
    mask_immersed_fields();
    fill_halo_prognostic();
    compute_auxiliaries();
    compute_tendencies();
  }

  // ------------------------------------------------- barotropic mode
  void barotropic_mode(int fU, int fV) {
#pragma omp parallel for
    for (int j = 1; j <= Ny; j++)
      for (int i = 1; i <= Nx; i++) {
        FT su = Dzc(1) * U(i, j, 1), sv = Dzc(1) * Vv(i, j, 1);
        for (int k = 2; k <= Nz; k++) { su += Dzc(k) * U(i, j, k); sv += Dzc(k) * Vv(i, j, k); }
        at2(fU, i, j) = su; at2(fV, i, j) = sv;
      }
  }
  void initialize() {
    barotropic_mode(F_BU, F_BV);
    fill_halo(F_BU, {1, 0, 0}, -1, false);
    fill_halo(F_BV, {0, 1, 0}, -1, false);
  }

  // ------------------------------------------------- ab2_step! (rows A8-A10)
  void free_surface_tendency(FT chi) {
    const Loc fcc{1, 0, 0}, cfc{0, 1, 0};
    FT c1 = (FT)1.5 + chi, c2 = (FT)0.5 + chi;
    FT ne = (chi != (FT)-0.5) ? (FT)1 : (FT)0;
#pragma omp parallel for
    for (int j = 1; j <= Ny; j++)
      for (int i = 1; i <= Nx; i++) {
        FT su = 0, sv = 0;
        for (int k = 1; k <= Nz; k++) {
          FT gu_ = peripheral(fcc, i, j, k) ? (FT)0 : c1 * at(F_GNU, i, j, k) - c2 * at(F_GMU, i, j, k) * ne;
          FT gv_ = peripheral(cfc, i, j, k) ? (FT)0 : c1 * at(F_GNV, i, j, k) - c2 * at(F_GMV, i, j, k) * ne;
          if (k == 1) { su = Dzc(k) * gu_; sv = Dzc(k) * gv_; }
          else { su += Dzc(k) * gu_; sv += Dzc(k) * gv_; }
        }
        at2(F_GU, i, j) = su; at2(F_GV, i, j) = sv;
      }
    fill_halo(F_GU, {1, 0, 0}, -1, false);
    fill_halo(F_GV, {0, 1, 0}, -1, false);
  }
  void ab2_fields(FT dt, FT chi) {
    FT c1 = (FT)1.5 + chi, c2 = (FT)0.5 + chi;
    FT ne = (chi != (FT)-0.5) ? (FT)1 : (FT)0;
#pragma omp parallel for collapse(2)
    for (int k = 1; k <= Nz; k++)
      for (int j = 1; j <= Ny; j++)
        for (int i = 1; i <= Nx; i++) {
          at(F_U, i, j, k) += dt * (c1 * at(F_GNU, i, j, k) - c2 * at(F_GMU, i, j, k) * ne);
          at(F_V, i, j, k) += dt * (c1 * at(F_GNV, i, j, k) - c2 * at(F_GMV, i, j, k) * ne);
          at(F_T, i, j, k) = at(F_T, i, j, k) + dt * (c1 * at(F_GNT, i, j, k) - c2 * at(F_GMT, i, j, k));
          at(F_S, i, j, k) = at(F_S, i, j, k) + dt * (c1 * at(F_GNS, i, j, k) - c2 * at(F_GMS, i, j, k));
        }
  }
  // one forward-backward substep over the interior; halo/topology handled by explicit BC application
  void barotropic_substep(FT dtau, FT wgt) {
    std::vector<FT>&eta = fld[F_ETA], &BU = fld[F_BU], &BV = fld[F_BV];
    const bool fold = c.topo_y == TOPO_FOLD;
    // (fold: the rows above Ny that the differences touch must hold the folded values)
    if (fold) for (int i = 1; i <= Nx; i++) BV[id2(i, Ny + 1)] = -BV[id2(Nx - i + 1, Ny)];
#pragma omp parallel for
    for (int j = 1; j <= Ny; j++)
      for (int i = 1; i <= Nx; i++) {
        int ie = (i == Nx) ? 1 : i + 1;
        FT dxU_ = dyfc[id2(ie, j)] * BU[id2(ie, j)] - dyfc[id2(i, j)] * BU[id2(i, j)];
        FT dyV_;
        if (!fold && j == Ny) dyV_ = -(dxcf[id2(i, j)] * BV[id2(i, j)]);
        else if (j == 1) dyV_ = dxcf[id2(i, 2)] * BV[id2(i, 2)];
        else dyV_ = dxcf[id2(i, j + 1)] * BV[id2(i, j + 1)] - dxcf[id2(i, j)] * BV[id2(i, j)];
        eta[id2(i, j)] -= dtau * (dxU_ + dyV_) / azcc[id2(i, j)];
      }
#pragma omp parallel for
    for (int j = 1; j <= Ny; j++)
      for (int i = 1; i <= Nx; i++) {
        int iw = (i == 1) ? Nx : i - 1;
        FT dxe = (eta[id2(i, j)] - eta[id2(iw, j)]) / dxfc[id2(i, j)];
        FT dye = (j == 1) ? (FT)0 : (eta[id2(i, j)] - eta[id2(i, j - 1)]) / dycf[id2(i, j)];
        BU[id2(i, j)] += dtau * (-(FT)c.g * Hfc[id2(i, j)] * dxe + at2(F_GU, i, j));
        BV[id2(i, j)] += dtau * (-(FT)c.g * Hcf[id2(i, j)] * dye + at2(F_GV, i, j));
        at2(F_FETA, i, j) += wgt * eta[id2(i, j)];
        at2(F_FU, i, j) += wgt * BU[id2(i, j)];
        at2(F_FV, i, j) += wgt * BV[id2(i, j)];
      }
  }
  void step_free_surface(FT dt) {
    std::fill(fld[F_FETA].begin(), fld[F_FETA].end(), (FT)0);
    std::fill(fld[F_FU].begin(), fld[F_FU].end(), (FT)0);
    std::fill(fld[F_FV].begin(), fld[F_FV].end(), (FT)0);
    FT dtau = (FT)c.dtau_frac * dt;
    for (int m = 0; m < c.nsub; m++) barotropic_substep(dtau, wts[m]);
    for (int j = 1; j <= Ny; j++)
      for (int i = 1; i <= Nx; i++) {
        at2(F_ETA, i, j) = at2(F_FETA, i, j);
        at2(F_BU, i, j) = at2(F_FU, i, j);
        at2(F_BV, i, j) = at2(F_FV, i, j);
      }
    mask_field(F_U, {1, 0, 0});
    mask_field(F_V, {0, 1, 0});
    fill_halo(F_BU, {1, 0, 0}, -1, false);
    fill_halo(F_BV, {0, 1, 0}, -1, false);
  }
  void ab2_step(FT dt, FT chi) {
    free_surface_tendency(chi);
    ab2_fields(dt, chi);
    if (c.closure == 2) {
      implicit_vertical_diffusion(F_U, {1, 0, 0}, (FT)c.nu, dt);
      implicit_vertical_diffusion(F_V, {0, 1, 0}, (FT)c.nu, dt);
      implicit_vertical_diffusion(F_T, {0, 0, 0}, (FT)c.kappa, dt);
      implicit_vertical_diffusion(F_S, {0, 0, 0}, (FT)c.kappa, dt);
    }
    step_free_surface(dt);
  }
  // ------------------------------------------------- corrector + cache (rows A11, A12)
  void correct_and_cache() {
    barotropic_mode(F_FU, F_FV);
#pragma omp parallel for collapse(2)
    for (int k = 1; k <= Nz; k++)
      for (int j = 1; j <= Ny; j++)
        for (int i = 1; i <= Nx; i++) {
          at(F_U, i, j, k) = at(F_U, i, j, k) + (at2(F_BU, i, j) - at2(F_FU, i, j)) / Hfc[id2(i, j)];
          at(F_V, i, j, k) = at(F_V, i, j, k) + (at2(F_BV, i, j) - at2(F_FV, i, j)) / Hcf[id2(i, j)];
        }
    for (int q = 0; q < 4; q++)
      for (int k = 1; k <= Nz; k++)
        for (int j = 1; j <= Ny; j++)
          for (int i = 1; i <= Nx; i++) at(F_GMU + q, i, j, k) = at(F_GNU + q, i, j, k);
    for (int j = 1; j <= Ny; j++)
      for (int i = 1; i <= Nx; i++) { at2(F_GMBU, i, j) = at2(F_GU, i, j); at2(F_GMBV, i, j) = at2(F_GV, i, j); }
  }
  // ------------------------------------------------- whole step (row A0)
  void time_step(FT dt, bool euler) {
    euler = euler || (dt != last_dt);
    FT chi = euler ? (FT)-0.5 : (FT)c.chi;
    ab2_step(dt, chi);
    time += (double)dt; iteration += 1; last_dt = dt;
    correct_and_cache();
    update_state();
  }
  void first_time_step(FT dt) { initialize(); update_state(); time_step(dt, true); }
};

// ---------------------------------------------------------------- C API (ctypes)
#define ORACLE_API(SUF, FT)                                                                                  \
  extern "C" void* gb25o_create_##SUF(const OConfig* cfg, const FT* const* g2, const FT* const* gz,        \
                                      const FT* bottom, const FT* weights) {                                \
    return new Oracle<FT>(*cfg, g2, gz, bottom, weights);                                                    \
  }                                                                                                          \
  extern "C" void gb25o_destroy_##SUF(void* h) { delete (Oracle<FT>*)h; }                                    \
  extern "C" FT* gb25o_field_##SUF(void* h, int f) { return ((Oracle<FT>*)h)->fld[f].data(); }               \
  extern "C" int* gb25o_kbot_##SUF(void* h) { return ((Oracle<FT>*)h)->kbot.data(); }                        \
  extern "C" FT* gb25o_depth_##SUF(void* h, int which) {                                                     \
    Oracle<FT>* o = (Oracle<FT>*)h;                                                                          \
    return which == 0 ? o->Hcc.data() : (which == 1 ? o->Hfc.data() : o->Hcf.data());                        \
  }                                                                                                          \
  extern "C" void gb25o_set_clock_##SUF(void* h, double t, long it, double last_dt) {                        \
    Oracle<FT>* o = (Oracle<FT>*)h; o->time = t; o->iteration = it; o->last_dt = (FT)last_dt;                \
  }                                                                                                          \
  extern "C" void gb25o_op_##SUF(void* h, int op, double dt, double chi) {                                   \
    Oracle<FT>* o = (Oracle<FT>*)h;                                                                          \
    switch (op) {                                                                                            \
      case 0: o->initialize(); break;                                                                        \
      case 1: o->update_state(); break;                                                                      \
      case 2: o->first_time_step((FT)dt); break;                                                             \
      case 3: o->time_step((FT)dt, false); break;                                                            \
      case 4: o->mask_immersed_fields(); break;                                                              \
      case 5: o->fill_halo_prognostic(); break;                                                              \
      case 6: o->compute_auxiliaries(); break;                                                               \
      case 7: o->compute_tendencies(); break;                                                                \
      case 8: o->ab2_step((FT)dt, (FT)chi); break;                                                           \
      case 9: o->correct_and_cache(); break;                                                                 \
      case 10: o->compute_w(); break;                                                                        \
      case 11: o->compute_p(); break;                                                                        \
      case 12: o->momentum_tendency(); break;                                                                \
      case 13: o->tracer_tendency(F_T, F_GNT); o->tracer_tendency(F_S, F_GNS); break;                        \
      case 14: o->boundary_tendencies(); break;                                                              \
      default: break;                                                                                        \
    }                                                                                                        \
  }                                                                                                          \
  extern "C" void gb25o_set_flux_bc_##SUF(void* h, int q, int side, const FT* a) {                          \
    Oracle<FT>* o = (Oracle<FT>*)h;                                                                          \
    if (q < 0 || q > 3 || side < 0 || side > 1) return;                                                      \
    if (a) o->bflux[q][side].assign(a, a + o->n2); else o->bflux[q][side].clear();                           \
  }                                                                                                          \
  extern "C" void gb25o_fill_halo_##SUF(void* h, int f, int lx, int ly, int lz, double sign, int three_d) {  \
    ((Oracle<FT>*)h)->fill_halo(f, Loc{lx, ly, lz}, (FT)sign, three_d != 0);                                 \
  }                                                                                                          \
  extern "C" FT gb25o_rho_prime_##SUF(void* h, double T, double S, double Z) {                               \
    return ((Oracle<FT>*)h)->rho_prime((FT)T, (FT)S, (FT)Z);                                                 \
  }                                                                                                          \
  extern "C" FT gb25o_weno_##SUF(void* h, int B, int left, const FT* q, const FT* s1, const FT* s2) {        \
    return ((Oracle<FT>*)h)->weno_window(B, left != 0, q, s1, s2);                                           \
  }

ORACLE_API(f32, float)
ORACLE_API(f64, double)

// ---- OpenMP thread control for the timed CPU baseline (torchrun exports OMP_NUM_THREADS=1 to its workers)
extern "C" void gb25o_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
extern "C" int gb25o_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
