"""Python mirror of the GordonBell25 host interface over libgb25cuda.

Same names, argument meaning and error behaviour as the reference's Julia API
(/root/reference/src/GordonBell25.jl:3-4 and the un-exported functions its scripts call):

=====================================================  ================================================
reference (Julia)                                       here
=====================================================  ================================================
``baroclinic_instability_model(arch, Nx, Ny, Nz; Δt)``  ``baroclinic_instability_model(arch, Nx, Ny, Nz, Δt=…)``
  src/baroclinic_instability_model.jl:12-85
``first_time_step!(model)`` / ``time_step!(model)``     ``first_time_step(model)`` / ``time_step(model)``
``loop!(model, Ninner)``  src/timestepping_utils.jl     ``loop(model, Ninner)``
``Oceananigans.initialize!`` / ``update_state!``         ``initialize(model)`` / ``update_state(model)``
``*_workload!(model)``  src/precompile.jl:44-127        ``*_workload(model)``
``compare_states`` / ``sync_states!`` src/correctness.jl ``compare_states`` / ``sync_states``
``set!(model, u=…, v=…)``                               ``set(model, u=…, v=…)``
=====================================================  ================================================

``arch`` is ``B200(device)`` (the analogue of ``GPU()`` / ``ReactantState()``); the model state lives in
device memory owned by libgb25cuda and only crosses to the host through ``parent`` / ``set_parent``.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

from . import grids as _grids
from .config import PhysicsConfig
from .lib import FIELD_LOC, FIELD_NAMES, Handle
from .splitexplicit import averaging_weights


@dataclasses.dataclass
class B200:
    """Architecture tag: one B200 (CUDA device ordinal).  There is no CPU() here."""
    device: int = -1


@dataclasses.dataclass
class Clock:
    time: float = 0.0
    iteration: int = 0
    last_Δt: float = math.inf


def _interior_slices(grid, loc):
    lx, ly, lz, three_d = loc
    ny = grid.Ny + (1 if (ly and grid.owns_north_wall) else 0)
    nz = grid.Nz + (1 if lz else 0)
    sx = slice(grid.Hx, grid.Hx + grid.Nx)
    sy = slice(grid.Hy, grid.Hy + ny)
    if not three_d:
        return (slice(0, 1), sy, sx)
    return (slice(grid.Hz, grid.Hz + nz), sy, sx)


class ModelBase:
    """What both backends expose: parent-shaped arrays by name and the stepping entry points."""
    grid: _grids.Grid
    physics: PhysicsConfig
    clock: Clock

    # fields(model) of Oceananigans: (u, v, w, T, S, η)
    field_names = ("u", "v", "w", "T", "S", "eta")

    def parent(self, name) -> np.ndarray:      # Array(parent(ψ))
        raise NotImplementedError

    def set_parent(self, name, a):
        raise NotImplementedError

    def interior(self, name):
        return self.parent(name)[_interior_slices(self.grid, FIELD_LOC[name])]

    def set_interior(self, name, values):
        p = self.parent(name)
        sl = _interior_slices(self.grid, FIELD_LOC[name])
        p[sl] = np.broadcast_to(np.asarray(values, dtype=p.dtype), p[sl].shape)
        self.set_parent(name, p)


class HydrostaticFreeSurfaceModel(ModelBase):
    """Device-resident model: the analogue of ``HydrostaticFreeSurfaceModel(; grid, free_surface, …)``
    (/root/reference/src/baroclinic_instability_model.jl:67-70) on the ``B200`` architecture."""

    def __init__(self, arch, grid, physics=None, partition=(1, 1, 0, 0), float_type=np.float32):
        if not isinstance(arch, B200):
            raise TypeError("arch must be B200(); libgb25cuda has no CPU architecture")
        self.arch = arch
        self.grid = grid
        self.dtype = np.dtype(float_type)       # eltype(grid): Float32 (BASELINE) or Float64 (the reference CLI's default)
        self.physics = physics or PhysicsConfig()
        self.dtau_frac, self.weights = averaging_weights(self.physics.substeps)
        self.handle = Handle(grid, self.physics, self.dtau_frac, self.weights, device=arch.device, partition=partition,
                             float_type=self.dtype)
        self.clock = Clock()

    # --- state access
    def parent(self, name):
        return self.handle.get_field(name)

    def set_parent(self, name, a):
        self.handle.set_field(name, a)

    def interior(self, name):                  # Array(interior(ψ)): one strided device-to-host copy
        return self.handle.get_interior(name)

    def set_interior(self, name, values):      # set!(model, name=values)
        self.handle.set_interior(name, values)

    def _push_clock(self):
        dt = self.clock.last_Δt
        self.handle.set_clock(self.clock.time, self.clock.iteration, 0.0 if math.isinf(dt) else dt)

    def _pull_clock(self):
        t, it, dt = self.handle.get_clock()
        self.clock.time, self.clock.iteration, self.clock.last_Δt = t, it, dt

    def synchronize(self):
        self.handle.call("gb25_synchronize")

    def close(self):
        self.handle.close()


# ------------------------------------------------------------------------------------------------
# constructors (src/baroclinic_instability_model.jl, src/model_utils.jl)
# ------------------------------------------------------------------------------------------------
def make_grid(Nx, Ny, Nz, halo=(8, 8, 8), grid_type="simple_lat_lon"):
    if grid_type in ("gaussian_islands", ":gaussian_islands"):
        return _grids.gaussian_islands_tripolar_grid(Nx, Ny, Nz, halo)
    if grid_type in ("tripolar", ":tripolar"):          # TripolarGrid without bathymetry (profiling aid)
        return _grids.tripolar_grid(Nx, Ny, Nz, halo)
    if grid_type in ("simple_lat_lon", ":simple_lat_lon"):
        return _grids.simple_latitude_longitude_grid(Nx, Ny, Nz, halo)
    raise ValueError(f"grid_type={grid_type} must be :gaussian_islands or :simple_lat_lon.")


def baroclinic_instability_model(arch, Nx=None, Ny=None, Nz=None, *, Δt=None, dt=None, resolution=None,
                                 halo=(8, 8, 8), grid_type="simple_lat_lon", physics=None, model_cls=None, **kw):
    """baroclinic_instability_model(arch, Nx, Ny, Nz; Δt, halo, grid_type)  or  (arch; resolution, Nz, Δt).
    As in the reference the initial-condition call is *not* made (it is commented out at
    src/baroclinic_instability_model.jl:74-80): T = S = u = v = 0 until the caller sets them."""
    if resolution is not None:
        Nx, Ny = _grids.resolution_to_points(resolution)
    if Δt is None:
        Δt = dt
    if Δt is None:
        raise TypeError("baroclinic_instability_model: keyword argument Δt not assigned")
    grid = make_grid(Nx, Ny, Nz, halo, grid_type)
    cls = model_cls or HydrostaticFreeSurfaceModel
    model = cls(arch, grid, physics, **kw)
    model.clock.last_Δt = float(np.float32(Δt))
    return model


def set_baroclinic_instability(model):
    """set_baroclinic_instability!(model): src/model_utils.jl:99-131."""
    T, S = _grids.baroclinic_instability_state(model.grid)
    model.set_interior("T", T)
    model.set_interior("S", S)


def set(model, **fields):
    """Oceananigans ``set!(model, u=…, v=…)``: interior values, halos untouched."""
    for name, val in fields.items():
        model.set_interior(name, val)


# ------------------------------------------------------------------------------------------------
# stepping (src/timestepping_utils.jl:21-45)
# ------------------------------------------------------------------------------------------------
def _call(model, name, *args):
    model._push_clock()
    model.handle.call(name, *args)
    model._pull_clock()


def first_time_step(model):
    Δt = model.clock.last_Δt
    _call(model, "gb25_first_time_step", Δt)


def time_step(model):
    Δt = model.clock.last_Δt + 0
    _call(model, "gb25_time_step", Δt)


def loop(model, Ninner):
    Δt = model.clock.last_Δt + 0
    _call(model, "gb25_loop", Δt, int(Ninner))


def initialize(model):
    _call(model, "gb25_initialize")


def update_state(model):
    _call(model, "gb25_update_state")


# ------------------------------------------------------------------------------------------------
# workloads (src/precompile.jl:44-127)
# ------------------------------------------------------------------------------------------------
def mask_immersed_model_fields_workload(model): _call(model, "gb25_mask_immersed_fields")
def tupled_fill_halo_regions_workload(model): _call(model, "gb25_fill_halo_regions")
def compute_tendencies_workload(model): _call(model, "gb25_compute_tendencies")
def compute_interior_momentum_tendencies_workload(model): _call(model, "gb25_compute_momentum_tendencies")
def compute_interior_tracer_tendencies_workload(model): _call(model, "gb25_compute_tracer_tendencies")
def compute_auxiliaries_workload(model): _call(model, "gb25_compute_auxiliaries")
def ab2_step_workload(model, Δt, χ=None): _call(model, "gb25_ab2_step", Δt, model.physics.chi if χ is None else χ)
def correct_velocities_and_cache_previous_tendencies_workload(model, Δt=None):
    _call(model, "gb25_correct_velocities_and_cache_previous_tendencies")


def compute_boundary_tendencies_workload(model):
    """compute_boundary_tendencies_workload! (src/precompile.jl:52-61): adds the flux boundary conditions to Gⁿ.  The
    benchmark model has default (no-flux) boundary conditions, for which this does nothing."""
    _call(model, "gb25_compute_boundary_tendencies")


def set_flux_boundary_condition(model, name, side, values):
    """``FieldBoundaryConditions(top=FluxBoundaryCondition(values))`` of ``name`` ∈ (u, v, T, S), ``side`` ∈ ("bottom",
    "top"): a 2-D parent-shaped array, or None for the default no-flux condition (how
    src/data_free_ocean_climate_model.jl passes wind stress and surface heat / salt fluxes)."""
    target = model.handle if hasattr(model.handle, "set_flux_boundary_condition") else model
    target.set_flux_boundary_condition(name, side, values)


def fill_halo_regions_workload(model):
    """No-op: closure = nothing, so there are no diffusivity fields."""


# ------------------------------------------------------------------------------------------------
# parity harness (src/correctness.jl)
# ------------------------------------------------------------------------------------------------
def _isapprox(a, b, rtol, atol):
    """Julia isapprox(::Array, ::Array): ‖a-b‖₂ <= max(atol, rtol*max(‖a‖₂, ‖b‖₂))."""
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    d = np.linalg.norm((a64 - b64).ravel())
    if not np.isfinite(d):
        return False
    return bool(d <= max(atol, rtol * max(np.linalg.norm(a64.ravel()), np.linalg.norm(b64.ravel()))))


def compare_parent(name, ψ1, ψ2, rtol=1e-8, atol=None, verbose=True, elementwise=None):
    """compare_parent (src/correctness.jl:4-15): crop ψ2 to ψ1's shape, norm-based isapprox, report max|δ|.
    ``elementwise`` (our addition): also require max|δ| <= elementwise*max|ψ|."""
    if atol is None:
        atol = math.sqrt(np.finfo(ψ1.dtype).eps)
    nz, ny, nx = ψ1.shape
    ψ2 = ψ2[:nz, :ny, :nx]
    δ = ψ1.astype(np.float64) - ψ2.astype(np.float64)
    ok = _isapprox(ψ1, ψ2, rtol, atol)
    amax = float(np.nanmax(np.abs(δ))) if δ.size else 0.0
    if elementwise is not None:
        scale = max(float(np.max(np.abs(ψ1))), float(np.max(np.abs(ψ2))))
        ok = ok and bool(amax <= elementwise * scale)
    if verbose:
        idx = np.unravel_index(np.nanargmax(np.abs(δ)), δ.shape) if δ.size else (0, 0, 0)
        print("(%8s) ψ₁ ≈ ψ₂: %-5s, max|ψ₁|, max|ψ₂|: %.15e, %.15e, max|δ|: %.15e at %d %d %d" %
              (name, ok, np.max(np.abs(ψ1)), np.max(np.abs(ψ2)), amax, idx[2] + 1, idx[1] + 1, idx[0] + 1))
    return ok


def compare_interior(name, m1, m2, fname, rtol=1e-8, atol=None, verbose=True, elementwise=None):
    return compare_parent(name, m1.interior(fname), m2.interior(fname), rtol, atol, verbose, elementwise)


def compare_states(m1, m2, rtol=None, atol=0.0, include_halos=False, throw_error=False, verbose=True,
                   elementwise=None):
    """compare_states (src/correctness.jl:28-90): every field of fields(model); Gⁿ and G⁻ for all but
    w and η; the split-explicit filtered state (U, V, η).  rtol defaults to sqrt(eps(Float32))."""
    if rtol is None:
        rtol = math.sqrt(np.finfo(np.float32).eps)
    get1 = (lambda n: m1.parent(n)) if include_halos else (lambda n: m1.interior(n))
    get2 = (lambda n: m2.parent(n)) if include_halos else (lambda n: m2.interior(n))
    ok = True
    failed = []

    def cmp(label, n):
        nonlocal ok
        r = compare_parent(label, get1(n), get2(n), rtol, atol, verbose, elementwise)
        if not r:
            failed.append(label)
        ok = ok and r

    for name in m1.field_names:
        cmp(name, name)
        if name not in ("w", "eta"):
            cmp(f"Gⁿ.{name}", f"Gn_{name}")
            cmp(f"G⁻.{name}", f"Gm_{name}")
    for label, n in (("U", "filt_U"), ("V", "filt_V"), ("η", "filt_eta")):
        cmp(label, n)
    if ok:
        if verbose:
            print(f"The two models are consistent within rtol={rtol} and atol={atol} !")
    else:
        msg = "There is a discrepancy between the models!  See the details above: " + ", ".join(failed)
        if throw_error:
            raise AssertionError(msg)
        if verbose:
            print("ERROR:", msg)
    return ok


def sync_states(m1, m2):
    """sync_states!(m1, m2) (src/correctness.jl:92-103): copy the parents of fields(m2) into m1."""
    for name in m1.field_names:
        p2 = m2.parent(name)
        s1 = m1.parent(name).shape
        m1.set_parent(name, p2[:s1[0], :s1[1], :s1[2]].astype(getattr(m1, "dtype", np.float32)))


ALL_FIELD_NAMES = FIELD_NAMES
