/* libgb25cuda from plain C: the model of src/baroclinic_instability_model.jl on a latitude-longitude grid, built, stepped
 * and read back through include/gb25cuda.h alone (no Python, no Julia, no C++).
 *
 *   gcc -std=c99 -O1 -Iinclude examples/lat_lon_from_c.c -o lat_lon_from_c -Lgb-25_b200/csrc -lgb25cuda -lm
 *   LD_LIBRARY_PATH=gb-25_b200/csrc ./lat_lon_from_c             first_time_step! + loop!(model, 10), prints max|u|, max|eta|
 *   ./lat_lon_from_c --checksums                                  grid products only (no device needed)
 *
 * The grid products follow /root/reference/src/model_utils.jl:56-65 (LatitudeLongitudeGrid, latitude (-80, 80), halo 8,
 * exponential z faces) the way gb-25_b200/grids.py builds them; tests/test_abi_and_host.py compares the checksums.
 * Exit codes: 0 ok, 2 no CUDA device (the library has no CPU path), 1 any other failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gb25cuda.h"

#define NX 64
#define NY 32
#define NZ 8
#define H 8
#define PX (NX + 2 * H)
#define PY (NY + 2 * H + 1)
#define PZ (NZ + 2 * H + 1)
#define SUBSTEPS 30

static const double kPi = 3.14159265358979323846, kRadius = 6371.0e3, kOmega = 7.292115e-5;

static gb25_real* plane(void) { return (gb25_real*)calloc((size_t)PX * PY, sizeof(gb25_real)); }
static void fill_rows(gb25_real* a, const double* per_row) {
  for (int j = 0; j < PY; j++)
    for (int i = 0; i < PX; i++) a[(size_t)j * PX + i] = (gb25_real)per_row[j];
}
static double sum(const gb25_real* a, size_t n) { double s = 0; for (size_t q = 0; q < n; q++) s += fabs((double)a[q]); return s; }   /* sum |a| */

/* SplitExplicitFreeSurface(substeps = 30): the averaging weights of the shape function, cut at its last positive value */
static int averaging_weights(float* w) {
  const double p = 2, q = 4, r = 0.18927, tau0 = (p + 2) * (p + q + 2) / ((p + 1) * (p + q + 1));
  double a[SUBSTEPS], total = 0;
  int n = 0;
  for (int m = 1; m <= SUBSTEPS; m++) {
    const double t = 2.0 * m / SUBSTEPS / tau0;
    a[m - 1] = pow(t, p) * (1 - pow(t, q)) - r * t;
  }
  for (int m = SUBSTEPS; m >= 1; m--) if (a[m - 1] >= 0) { n = m; break; }
  for (int m = 0; m < n; m++) total += a[m];
  for (int m = 0; m < n; m++) w[m] = (float)(a[m] / total);
  return n;
}

int main(int argc, char** argv) {
  const int checksums_only = argc > 1 && strcmp(argv[1], "--checksums") == 0;
  /* ---- horizontal metrics: functions of the row only */
  const double dlam = 360.0 / NX * kPi / 180.0, dphi = 160.0 / NY * kPi / 180.0;
  double dx_c[PY], dx_f[PY], dy[PY], az_c[PY], az_f[PY], f_f[PY];
  for (int r = 0; r < PY; r++) {
    const double phif = (-80.0 + (r - H) * (160.0 / NY)) * kPi / 180.0, phic = phif + dphi / 2;
    dx_c[r] = kRadius * cos(phic) * dlam;
    dx_f[r] = kRadius * cos(phif) * dlam;
    dy[r] = kRadius * dphi;
    az_c[r] = kRadius * kRadius * dlam * (sin(phif + dphi) - sin(phif));
    az_f[r] = kRadius * kRadius * dlam * (sin(phic) - sin(phic - dphi));
    f_f[r] = 2 * kOmega * sin(phif);
  }
  gb25_real *dxc = plane(), *dxf = plane(), *dyy = plane(), *azc = plane(), *azf = plane(), *fff = plane();
  fill_rows(dxc, dx_c); fill_rows(dxf, dx_f); fill_rows(dyy, dy); fill_rows(azc, az_c); fill_rows(azf, az_f); fill_rows(fff, f_f);
  /* ---- vertical: exponential z faces (depth 4000 m, scale 30), halos by linear extrapolation of the end spacings */
  double zf[PZ + 1], e[NZ + 1];
  {
    const double L = NZ + 1, hs = 30.0;
    for (int k = 1; k <= NZ + 1; k++) e[k - 1] = (exp(k / hs) - exp(-L / hs)) / (1 - exp(-L / hs));
    const double e0 = e[0];
    for (int k = 0; k <= NZ; k++) e[k] -= e0;
    const double scale = -4000.0 / e[NZ];
    for (int k = 0; k <= NZ; k++) e[k] *= scale;
    e[0] = 0.0;
    for (int k = 0; k <= NZ; k++) zf[H + k] = e[NZ - k];           /* bottom (-4000) first */
    const double dlo = zf[H + 1] - zf[H], dhi = zf[H + NZ] - zf[H + NZ - 1];
    for (int m = 1; m <= H; m++) zf[H - m] = zf[H] - m * dlo;
    for (int m = 1; m <= H + 1; m++) zf[H + NZ + m] = zf[H + NZ] + m * dhi;
  }
  gb25_real z_f[PZ], z_c[PZ], dz_c[PZ], dz_f[PZ];
  for (int k = 0; k < PZ; k++) { z_f[k] = (gb25_real)zf[k]; z_c[k] = (gb25_real)(0.5 * (zf[k] + zf[k + 1])); dz_c[k] = (gb25_real)(zf[k + 1] - zf[k]); }
  for (int k = 1; k < PZ; k++) dz_f[k] = (gb25_real)(0.5 * (zf[k + 1] - zf[k - 1]));
  dz_f[0] = dz_f[1];
  float weights[SUBSTEPS];
  const int nw = averaging_weights(weights);
  if (checksums_only) {
    double ws = 0;
    for (int m = 0; m < nw; m++) ws += fabs((double)weights[m]);
    printf("dx_cc %.9e\ndx_cf %.9e\ndy_cc %.9e\naz_cc %.9e\naz_cf %.9e\nf_ff %.9e\nz_f %.9e\nz_c %.9e\ndz_c %.9e\ndz_f %.9e\nnweights %d\nweights %.9e\n",
           sum(dxc, (size_t)PX * PY), sum(dxf, (size_t)PX * PY), sum(dyy, (size_t)PX * PY), sum(azc, (size_t)PX * PY),
           sum(azf, (size_t)PX * PY), sum(fff, (size_t)PX * PY), sum(z_f, PZ), sum(z_c, PZ), sum(dz_c, PZ), sum(dz_f, PZ), nw, ws);
    return 0;
  }

  /* ---- HydrostaticFreeSurfaceModel(; grid, free_surface = SplitExplicitFreeSurface(substeps = 30), ...) */
  gb25_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.Nx = NX; cfg.Ny = NY; cfg.Nz = NZ; cfg.Hx = cfg.Hy = cfg.Hz = H;
  cfg.topo_y = GB25_TOPO_BOUNDED; cfg.immersed = 0; cfg.nsubsteps = nw;
  cfg.coriolis_scheme = 1; cfg.fold_variant = 0; cfg.south_inactive = 1; cfg.cond_diff = 1; cfg.eos_r0 = 0;
  cfg.g = 9.80665f; cfg.rho0 = 1020.0f; cfg.chi = 0.1f; cfg.dtau_frac = 2.0f / SUBSTEPS; cfg.weno_eps = 1e-8f;
  cfg.Rx = cfg.Ry = 1; cfg.rx = cfg.ry = 0; cfg.device = -1; cfg.closure = 0;
  gb25_grid grid;
  memset(&grid, 0, sizeof grid);
  grid.dx_cc = grid.dx_fc = dxc; grid.dx_cf = grid.dx_ff = dxf;
  grid.dy_cc = grid.dy_fc = grid.dy_cf = grid.dy_ff = dyy;
  grid.az_cc = grid.az_fc = azc; grid.az_cf = grid.az_ff = azf;
  grid.f_ff = fff; grid.z_f = z_f; grid.z_c = z_c; grid.dz_c = dz_c; grid.dz_f = dz_f;
  grid.bottom_height = NULL; grid.avg_weights = weights;

  gb25_handle* h = NULL;
  int rc = gb25_create(&cfg, &grid, &h);
  if (rc != GB25_OK) {
    fprintf(stderr, "gb25_create failed (%d): %s\n", rc, gb25_last_error(NULL));
    return rc == GB25_ERR_NO_DEVICE ? 2 : 1;
  }
#define CHECK(call) do { rc = (call); if (rc != GB25_OK) { fprintf(stderr, #call " failed (%d): %s\n", rc, gb25_last_error(h)); gb25_destroy(h); return 1; } } while (0)
  /* ---- set!(model, T = ..., S = ..., u = ...): interior-shaped uploads */
  int shp[3];
  CHECK(gb25_interior_shape(h, GB25_T, shp));
  const size_t n3 = (size_t)shp[0] * shp[1] * shp[2];
  gb25_real* buf = (gb25_real*)malloc(n3 * sizeof(gb25_real));
  for (int k = 0; k < shp[2]; k++)
    for (int j = 0; j < shp[1]; j++)
      for (int i = 0; i < shp[0]; i++) {
        const double lat = -80.0 + (j + 0.5) * 160.0 / NY;
        buf[((size_t)k * shp[1] + j) * shp[0] + i] = (gb25_real)(5.0 + 20.0 * cos(lat * kPi / 180.0) + 0.002 * (double)z_c[H + k]);   /* warm equator, stratified */
      }
  CHECK(gb25_set_interior(h, GB25_T, buf));
  for (size_t q = 0; q < n3; q++) buf[q] = (gb25_real)35.0;
  CHECK(gb25_set_interior(h, GB25_S, buf));
  CHECK(gb25_interior_shape(h, GB25_U, shp));
  for (size_t q = 0; q < (size_t)shp[0] * shp[1] * shp[2]; q++) buf[q] = (gb25_real)(1e-3 * ((double)rand() / RAND_MAX));
  CHECK(gb25_set_interior(h, GB25_U, buf));

  /* ---- first_time_step!(model); loop!(model, 10)  (src/timestepping_utils.jl:21-45) */
  const float dt = 60.0f;
  CHECK(gb25_first_time_step(h, dt));
  CHECK(gb25_loop(h, dt, 10));
  CHECK(gb25_synchronize(h));

  double umax = 0, emax = 0;
  int finite = 1;
  CHECK(gb25_get_interior(h, GB25_U, buf));
  for (size_t q = 0; q < (size_t)shp[0] * shp[1] * shp[2]; q++) { if (!isfinite((double)buf[q])) finite = 0; if (fabs((double)buf[q]) > umax) umax = fabs((double)buf[q]); }
  CHECK(gb25_interior_shape(h, GB25_ETA, shp));
  CHECK(gb25_get_interior(h, GB25_ETA, buf));
  for (size_t q = 0; q < (size_t)shp[0] * shp[1]; q++) { if (!isfinite((double)buf[q])) finite = 0; if (fabs((double)buf[q]) > emax) emax = fabs((double)buf[q]); }
  double t = 0; long it = 0; float last_dt = 0; long launches = 0;
  CHECK(gb25_get_clock(h, &t, &it, &last_dt));
  CHECK(gb25_kernel_launch_count(h, &launches));
  printf("iteration %ld, time %.0f s, %ld kernel launches, max|u| = %.4e m/s, max|eta| = %.4e m, %s\n", it, t, launches, umax, emax,
         finite ? "all finite" : "NOT FINITE");
  free(buf);
  gb25_destroy(h);
  return finite && it == 11 ? 0 : 1;
}
