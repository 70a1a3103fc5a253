/* gb25cuda.h — C ABI of libgb25cuda: the B200-native (sm_100a) implementation of the Oceananigans
 * HydrostaticFreeSurfaceModel time step that PRONTOLab/GB-25 drives.
 *
 * Every entry point cites the reference interface it replaces (paths under /root/reference).
 * The reference host stays Julia: the GordonBell25 constructors and first_time_step!/time_step!/
 * loop! are kept, and call these symbols through `ccall` (INTEGRATION.md shows the binding).
 *
 * Conventions
 *  - plain C types only; host arrays are borrowed for the duration of the call;
 *  - all arrays are gb25_real (Float32 in libgb25cuda.so, Float64 in libgb25cuda_f64.so), column-major with x fastest
 *    and halos included ("parent" arrays of Oceananigans fields): a (Tx,Ty,Tz) Julia array is passed as-is;
 *  - every function returns 0 on success and a negative gb25_status otherwise; the message of the
 *    last failure is available from gb25_last_error(); no exception crosses the boundary;
 *  - a handle is driven by one thread at a time; compute calls are stream-ordered and asynchronous,
 *    gb25_get_field / gb25_synchronize / timing queries block;
 *  - there is no CPU fallback: with no usable CUDA device gb25_create fails with GB25_ERR_NO_DEVICE.
 */
#ifndef GB25CUDA_H
#define GB25CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define GB25_ABI_VERSION 1

/* Scalar type of every field and grid array that crosses the boundary.  libgb25cuda.so is the Float32 build (BASELINE:
 * Float32 cell-steps/s); libgb25cuda_f64.so is the same source compiled with -DGB25_F64 for `--float-type Float64`
 * (/root/reference/src/arg_parsing.jl:28-31, the reference CLI's default): it runs the operator-per-kernel generation of the
 * kernels in double precision; the TMA / packed-FP32x2 / persistent kernels exist in Float32 only.  gb25_real_bytes() tells
 * the two apart at run time.  Model parameters (g, rho0, chi, dt, ...) stay `float` in both: the host model holds them in
 * Float32 when it was built with a Float32 grid, and a Float64 host passes values that are exact in Float32 or accepts the
 * rounding (the parity tests round them first). */
#ifdef GB25_F64
typedef double gb25_real;
#else
typedef float gb25_real;
#endif

typedef struct gb25_handle gb25_handle;

typedef enum {
  GB25_OK = 0,
  GB25_ERR_INVALID = -1,    /* bad argument / unsupported configuration */
  GB25_ERR_NO_DEVICE = -2,  /* no CUDA device: the product has no CPU path */
  GB25_ERR_CUDA = -3,       /* CUDA runtime error (sticky until gb25_clear_error) */
  GB25_ERR_ALLOC = -4,
  GB25_ERR_COMM = -5        /* multi-GPU exchange set-up / transport error */
} gb25_status;

/* y topology of the horizontal grid */
#define GB25_TOPO_BOUNDED 0 /* LatitudeLongitudeGrid: (Periodic, Bounded, Bounded)            */
#define GB25_TOPO_FOLD 1    /* TripolarGrid: (Periodic, RightConnected, Bounded), zipper fold */

/* Model configuration: what HydrostaticFreeSurfaceModel(; grid, free_surface, buoyancy, coriolis,
 * momentum_advection, tracer_advection) fixes in src/baroclinic_instability_model.jl:17-70.
 * Sizes are those of THIS handle's tile (the whole domain when Rx = Ry = 1). */
typedef struct {
  int Nx, Ny, Nz;          /* interior size of this tile                                           */
  int Hx, Hy, Hz;          /* halo (8,8,8) in the reference: src/baroclinic_instability_model.jl:18 */
  int topo_y;              /* GB25_TOPO_*                                                          */
  int immersed;            /* 1: ImmersedBoundaryGrid(GridFittedBottom) (src/model_utils.jl:143)   */
  int nsubsteps;           /* barotropic substeps actually taken = length(averaging_weights)       */
  int coriolis_scheme;     /* 0 EnstrophyConserving, 1 ActiveCellEnstrophyConserving   (decision U2)  */
  int fold_variant;        /* 0 plain zipper, 1 also overwrite redundant half of row Ny (decision U1) */
  int south_inactive;      /* tripolar: cells south of j=1 are outside the domain      (decision U4)  */
  int cond_diff;           /* immersed-aware differences in vorticity and grad p       (decision U15) */
  int eos_r0;              /* include r0(z) in TEOS-10 rho'                            (decision U8)  */
  float g;                 /* 9.80665                                                              */
  float rho0;              /* 1020 (TEOS-10 reference density)                                     */
  float chi;               /* 0.1 (QuasiAdamsBashforth2)                                           */
  float dtau_frac;         /* barotropic step as a fraction of dt = 2/substeps                     */
  float weno_eps;          /* 1e-8                                                                 */
  /* x/y partition: Distributed(arch; partition=Partition(Rx,Ry,1)),
   * sharding/sharded_baroclinic_instability_simulation_run.jl:65-72; rank = rx + Rx*ry */
  int Rx, Ry, rx, ry;
  int device;              /* CUDA device ordinal to use (-1: current device)                      */
  /* closure: 0 = nothing (reference default, src/baroclinic_instability_model.jl:29),
   * 1 = VerticalScalarDiffusivity(kappa, nu) explicit, 2 = VerticallyImplicitTimeDiscretization (the commented
   * alternative at src/baroclinic_instability_model.jl:31 with kappa = 1e-5, nu = 1e-4)                           */
  int closure;
  float kappa, nu;
} gb25_config;

/* Grid products, all host pointers.  2-D arrays have (Nx+2Hx) x (Ny+2Hy+1) elements, x fastest;
 * interior (i,j) (1-based) at [(i+Hx-1) + (Nx+2Hx)*(j+Hy-1)]; the last row is only read for
 * Face-in-y quantities on Bounded grids.  Vertical arrays have Nz+2Hz+1 elements, k at [k+Hz-1].
 * Replaces what the kernels read from `grid` (grid.Δxᶠᶜᵃ, grid.Azᶜᶜᵃ, grid.z.cᵃᵃᶠ, …,
 * ibg.immersed_boundary.bottom_height) — built on the host by src/model_utils.jl:56-65,134-146. */
typedef struct {
  const gb25_real *dx_cc, *dx_fc, *dx_cf, *dx_ff;
  const gb25_real *dy_cc, *dy_fc, *dy_cf, *dy_ff;
  const gb25_real *az_cc, *az_fc, *az_cf, *az_ff;
  const gb25_real *f_ff;            /* 2 Ω sin φ at (Face,Face): HydrostaticSphericalCoriolis           */
  const gb25_real *z_f, *z_c;       /* face / centre heights                                            */
  const gb25_real *dz_c, *dz_f;     /* Δz at centres (zf[k+1]-zf[k]) and at faces (zc[k]-zc[k-1])       */
  const gb25_real *bottom_height;   /* GridFittedBottom height at (C,C), or NULL when not immersed       */
  const float *avg_weights;         /* nsubsteps split-explicit averaging weights (always Float32)       */
} gb25_grid;

/* Field identifiers for gb25_set_field / gb25_get_field / gb25_field_shape.
 * = Oceananigans.fields(model), timestepper.Gⁿ / G⁻, free_surface.{η, barotropic_velocities,
 *   filtered_state}: exactly the arrays src/correctness.jl:28-90 compares. */
typedef enum {
  GB25_U = 0, GB25_V, GB25_W, GB25_T, GB25_S, GB25_P,
  GB25_GN_U, GB25_GN_V, GB25_GN_T, GB25_GN_S,
  GB25_GM_U, GB25_GM_V, GB25_GM_T, GB25_GM_S,
  GB25_ETA, GB25_BARO_U, GB25_BARO_V,
  GB25_FILT_ETA, GB25_FILT_U, GB25_FILT_V,
  GB25_GN_BARO_U, GB25_GN_BARO_V, GB25_GM_BARO_U, GB25_GM_BARO_V,
  GB25_FIELD_COUNT
} gb25_field;

int gb25_abi_version(void);
int gb25_real_bytes(void);                            /* 4: libgb25cuda.so, 8: libgb25cuda_f64.so */

/* Construction / destruction.  Replaces the device-side allocation done by
 * HydrostaticFreeSurfaceModel(...) (src/baroclinic_instability_model.jl:67-70). */
int gb25_create(const gb25_config* cfg, const gb25_grid* grid, gb25_handle** out);
int gb25_destroy(gb25_handle* h);
const char* gb25_last_error(const gb25_handle* h); /* h may be NULL: error of a failed gb25_create */
int gb25_clear_error(gb25_handle* h);

/* State transfer in Oceananigans parent shape.  Replaces sync_states! (src/correctness.jl:92-103)
 * and the Array(parent(ψ)) reads of compare_parent (src/correctness.jl:4-15). */
int gb25_field_shape(const gb25_handle* h, int field, int shape[3]);
int gb25_set_field(gb25_handle* h, int field, const gb25_real* host_parent);
int gb25_get_field(gb25_handle* h, int field, gb25_real* host_parent);
/* Interior-shaped transfers: what Oceananigans' set!(model, u=..., v=...) writes and Array(interior(ψ)) reads
 * (correctness/correctness_baroclinic_instability_simulation_run.jl:40-42; src/model_utils.jl:99-131 for T, S).
 * The host array has the interior shape of the field — (Nx, Ny, Nz) for a (C,C,C) field, Ny+1 rows for a Face-y field
 * on a tile that owns the north wall of a Bounded grid, Nz+1 levels for w, one level for 2-D fields — and halos are
 * left untouched.  gb25_interior_shape returns it. */
int gb25_interior_shape(const gb25_handle* h, int field, int shape[3]);
int gb25_set_interior(gb25_handle* h, int field, const gb25_real* host_interior);
int gb25_get_interior(gb25_handle* h, int field, gb25_real* host_interior);
/* Batched transfers: n fields, all copies enqueued back to back on the handle's stream (one cudaMemcpy3DAsync per
 * field) and ONE synchronisation at the end.  interior = 0: parent shape, 1: interior shape.
 * Interior-shaped downloads of u, v, T, S, eta, U, V issued right after gb25_time_step / gb25_loop do not wait for the
 * tail of the step: the library knows since which kernel each of these interiors is final (T, S: the AB2 stage; the
 * others: the corrector) and copies them on a separate stream behind that event, under the tendency kernels. */
int gb25_set_fields(gb25_handle* h, int n, const int* fields, const gb25_real* const* host, int interior);
int gb25_get_fields(gb25_handle* h, int n, const int* fields, gb25_real* const* host, int interior);
/* model.clock: time, iteration, last_Δt (src/baroclinic_instability_model.jl:82) */
int gb25_set_clock(gb25_handle* h, double time, long iteration, float last_dt);
int gb25_get_clock(const gb25_handle* h, double* time, long* iteration, float* last_dt);

/* Whole-step API.  Replaces Oceananigans.initialize!, TimeSteppers.update_state!
 * (correctness/correctness_baroclinic_instability_simulation_run.jl:50-54) and
 * GordonBell25.first_time_step! / time_step! / loop! (src/timestepping_utils.jl:21-45).
 * dt <= 0 means "use clock.last_Δt" exactly as the reference wrappers do. */
int gb25_initialize(gb25_handle* h);
int gb25_update_state(gb25_handle* h);
int gb25_first_time_step(gb25_handle* h, float dt);
int gb25_time_step(gb25_handle* h, float dt);
int gb25_loop(gb25_handle* h, float dt, int nsteps); /* no host synchronisation inside */
int gb25_synchronize(gb25_handle* h);

/* Operator-level API: one entry point per *_workload! of src/precompile.jl:44-127. */
int gb25_mask_immersed_fields(gb25_handle* h);       /* precompile.jl:34  mask_immersed_model_fields!      */
int gb25_fill_halo_regions(gb25_handle* h);          /* precompile.jl:44-46 tupled_fill_halo_regions!      */
int gb25_compute_auxiliaries(gb25_handle* h);        /* precompile.jl:113-115 (w from continuity, pHY')    */
int gb25_compute_tendencies(gb25_handle* h);         /* precompile.jl:48-50                                */
int gb25_compute_momentum_tendencies(gb25_handle* h);/* precompile.jl:63-73                                */
int gb25_compute_tracer_tendencies(gb25_handle* h);  /* precompile.jl:75-111                               */
int gb25_compute_boundary_tendencies(gb25_handle* h);/* precompile.jl:52-61  (flux boundary conditions)    */
/* FluxBoundaryCondition at the bottom (side 0) or top (side 1) of u, v, T or S (field = GB25_U / GB25_V / GB25_T /
 * GB25_S): a 2-D parent-shaped host array ((Nx+2Hx) x (Ny+2Hy[+1]), x fastest) of fluxes in the units of the field times
 * m/s, or NULL for the default no-flux condition.  What the `boundary_conditions` keyword of the model constructor fixes
 * (src/data_free_ocean_climate_model.jl:12-70 builds wind stress and heat / salt fluxes this way); the benchmark model of
 * src/baroclinic_instability_model.jl has none, and gb25_compute_boundary_tendencies is then a no-op.  With a flux set,
 * gb25_compute_tendencies / gb25_update_state / the step entry points add G[i,j,1] += J_bottom Az / V and
 * G[i,j,Nz] -= J_top Az / V after the interior tendencies, as Oceananigans' apply_z_bcs! does. */
int gb25_set_flux_boundary_condition(gb25_handle* h, int field, int side, const gb25_real* flux_parent_2d);
int gb25_ab2_step(gb25_handle* h, float dt, float chi); /* precompile.jl:121-123 (incl. split-explicit)    */
int gb25_correct_velocities_and_cache_previous_tendencies(gb25_handle* h); /* precompile.jl:125-127         */

/* Measurement: CUDA-event time of the last gb25_loop / per-stage accumulated times. */
int gb25_last_loop_seconds(gb25_handle* h, double* seconds);
int gb25_kernel_launch_count(const gb25_handle* h, long* launches);
/* stage timers (CUDA events on the handle's stream): enable, then read after gb25_synchronize.
 * names/ms must hold `cap` entries; returns the number of stages written. */
int gb25_enable_stage_timers(gb25_handle* h, int enable);
int gb25_get_stage_times(gb25_handle* h, const char** names, float* ms, long* calls, int cap);

/* Debug aid: a handle created while the environment holds GB25_GUARD=1 places every device array between two 64 KiB
 * guard zones; gb25_check_guards counts the guard bytes that kernels have overwritten since (0 = no out-of-bounds store). */
int gb25_check_guards(gb25_handle* h, long* corrupted_bytes);
/* Number of CUDA kernels the library holds.  gb25_create loads every one of them on its device up front (lazy module
 * loading would otherwise load a kernel at its first launch, and that load can block the host behind a stream that waits
 * for a neighbour tile — a deadlock when one host thread drives all tiles).  No device needed; the test suite compares
 * the number with the entry points of the built device code. */
int gb25_kernel_table_size(void);

/* Multi-GPU (one process per GPU).  Neighbour tiles exchange halos over NVLink through peer-mapped
 * memory: every rank exports its exchange window (gb25_exchange_export), the host side gathers the
 * opaque blobs from all ranks (any transport: torch.distributed, MPI, a file) and hands the full set
 * back (gb25_exchange_connect).  Replaces the XLA collective-permute halo traffic of
 * Distributed(ReactantState(); partition=Partition(Rx,Ry,1)). */
int gb25_exchange_blob_size(void);
int gb25_exchange_export(gb25_handle* h, void* blob);
int gb25_exchange_connect(gb25_handle* h, const void* blobs, int nranks);
/* Multi-GPU, ONE process driving all devices of the node — the reference's configuration
 * (sharding/sharded_baroclinic_instability_simulation_run.jl:49, single_gpu_per_process=false): handles[r] is the tile of rank
 * r = rx + Rx*ry, each created on its own device (gb25_config.device).  Peer access is enabled between the devices and the
 * neighbours' allocations are used directly (no IPC).  gb25_loop_all enqueues every step on every tile in turn; per-tile
 * calls (gb25_time_step, gb25_first_time_step, ...) must likewise be issued for all tiles before any of them is
 * synchronised.  Nothing on the step path blocks the host: the kernels are loaded by gb25_create and the allocations of
 * the persistent substep kernel are made by the connect call. */
int gb25_exchange_connect_local(gb25_handle** handles, int n);
int gb25_loop_all(gb25_handle** handles, int n, float dt, int nsteps);

#ifdef __cplusplus
}
#endif
#endif /* GB25CUDA_H */
