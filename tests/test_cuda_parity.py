"""GPU parity tests: libgb25cuda (through the C ABI) against the CPU oracle on the same seeded inputs.

Protocol mirrored from /root/reference/correctness/correctness_baroclinic_instability_simulation_run.jl
(same model on both architectures, random u, v, staged comparisons with halos included, 2-norm
rtol = sqrt(eps(Float32)) from /root/reference/src/correctness.jl:28) plus per-operator checks in the
decomposition of /root/reference/src/precompile.jl:44-127.

Tolerances (stated): halo fills and every copy-type operator are BIT-EXACT.  Floating-point stages:
rtol = sqrt(eps(Float32)) = 3.4527e-4 in the 2-norm with halos included (the reference's criterion) and
additionally an element-wise bound max|d| <= 1e-4*max|psi| on the reference's own test state (T = S = 0).
On the baroclinic state the pressure field is O(700) so Float32 round-off dominates the small zonal
tendency; there the criterion is "as close to the Float64 oracle as the Float32 oracle is" (factor 3).
"""
import math

import numpy as np
import pytest

from gb25_b200 import grids, model as M
from gb25_b200.config import PhysicsConfig
from conftest import make_models

pytestmark = pytest.mark.gpu
RTOL = math.sqrt(np.finfo(np.float32).eps)

GRIDS = [("simple_lat_lon", 64, 32, 8), ("gaussian_islands", 64, 48, 10)]
STATE_FIELDS = ("u", "v", "w", "T", "S", "eta", "p", "U", "V", "filt_eta", "filt_U", "filt_V",
                "Gn_u", "Gn_v", "Gn_T", "Gn_S", "Gm_u", "Gm_v", "Gm_T", "Gm_S", "Gn_U", "Gn_V", "Gm_U", "Gm_V")


def _index_valued(m, names):
    rng = np.random.default_rng(7)
    for n in names:
        p = m.parent(n)
        p[...] = (rng.standard_normal(p.shape) * 100).astype(np.float32)
        m.set_parent(n, p)


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS)
@pytest.mark.parametrize("fold_variant", [0, 1])
def test_fill_halo_regions_bit_exact(oracle_mod, grid_type, Nx, Ny, Nz, fold_variant):
    if grid_type == "simple_lat_lon" and fold_variant == 1:
        pytest.skip("fold variant only exists on the tripolar grid")
    ph = PhysicsConfig(fold_variant=fold_variant)
    rm, vm = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod, physics=ph)
    names = ("u", "v", "T", "S", "eta", "U", "V")
    _index_valued(vm, names)
    for n in names:
        rm.set_parent(n, vm.parent(n))
    M.tupled_fill_halo_regions_workload(rm)
    M.tupled_fill_halo_regions_workload(vm)
    for n in names:
        a, b = rm.parent(n), vm.parent(n)
        assert a.shape == b.shape
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{n}: halo fill is not bit-exact"


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS)
def test_set_get_roundtrip_all_fields_bit_exact(oracle_mod, grid_type, Nx, Ny, Nz):
    rm, vm = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod)
    rng = np.random.default_rng(11)
    for n in STATE_FIELDS:
        shp = rm.handle.field_shape(n)
        assert shp == vm.parent(n).shape
        a = rng.standard_normal(shp).astype(np.float32)
        rm.set_parent(n, a)
        assert np.array_equal(rm.parent(n), a)


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS)
def test_interior_and_batched_transfers(oracle_mod, grid_type, Nx, Ny, Nz):
    """gb25_set_interior / gb25_get_interior (set!, Array(interior(psi))) and the batched gb25_set_fields / gb25_get_fields
    move exactly the bytes the parent-shaped calls move: interiors land inside the parent, halos are untouched."""
    rm, vm = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod)
    rng = np.random.default_rng(5)
    names = [n for n in STATE_FIELDS]
    for n in names:
        par = rng.standard_normal(rm.handle.field_shape(n)).astype(np.float32)
        rm.set_parent(n, par)
        ishp = rm.handle.interior_shape(n)
        assert ishp == M.ModelBase.interior(vm, n).shape, n                  # the mirror's slice of the parent
        inner = rng.standard_normal(ishp).astype(np.float32)
        rm.handle.set_interior(n, inner)
        sl = M._interior_slices(rm.grid, M.FIELD_LOC[n])
        want = par.copy(); want[sl] = inner
        assert np.array_equal(rm.parent(n), want), n
        assert np.array_equal(rm.handle.get_interior(n), inner), n
    pars = [np.ascontiguousarray(rng.standard_normal(rm.handle.field_shape(n)).astype(np.float32)) for n in names]
    rm.handle.set_fields(names, pars)
    outs = [np.empty_like(a) for a in pars]
    rm.handle.get_fields(names, outs)
    for n, a, b in zip(names, pars, outs):
        assert np.array_equal(a, b), n
    inner = [np.ascontiguousarray(rng.standard_normal(rm.handle.interior_shape(n)).astype(np.float32)) for n in names]
    rm.handle.set_fields(names, inner, interior=True)
    outs = [np.empty_like(a) for a in inner]
    rm.handle.get_fields(names, outs, interior=True)
    for n, a, b, par in zip(names, inner, outs, pars):
        assert np.array_equal(a, b), n
        want = par.copy(); want[M._interior_slices(rm.grid, M.FIELD_LOC[n])] = a
        assert np.array_equal(rm.parent(n), want), n
    # right after a step the interior downloads of u, v, T, S, eta, U, V start behind "final since" events, under the
    # tendency kernels: they must return what the parent-shaped download returns once the step is complete
    M.sync_states(rm, vm)
    M.first_time_step(rm)
    for rep in range(3):
        M.time_step(rm)
        prog = ["u", "v", "T", "S", "eta", "U", "V", "w", "Gn_u"]
        outs = [np.empty(rm.handle.interior_shape(n), dtype=np.float32) for n in prog]
        rm.handle.get_fields(prog, outs, interior=True)          # no synchronisation in between
        for n, a in zip(prog, outs):
            assert np.array_equal(a, rm.parent(n)[M._interior_slices(rm.grid, M.FIELD_LOC[n])]), (rep, n)
        assert np.array_equal(rm.handle.get_interior("T"), outs[2])
    rm.close()


def _assert_close(rm, vm, names, rtol=RTOL, elementwise=None, halos=True):
    bad = []
    for n in names:
        a = rm.parent(n) if halos else rm.interior(n)
        b = vm.parent(n) if halos else vm.interior(n)
        if not M.compare_parent(n, a, b, rtol=rtol, atol=0.0, verbose=True, elementwise=elementwise):
            bad.append(n)
    assert not bad, f"mismatch in {bad}"


def _assert_as_close_as_f32(rm, v32, v64, names, factor=3.0, rtol=RTOL, halos=True):
    """CUDA is as close to the Float64 oracle as the Float32 oracle is (x factor), or inside rtol of it."""
    bad = []
    for n in names:
        get = (lambda m: m.parent(n)) if halos else (lambda m: m.interior(n))
        t = get(v64).astype(np.float64)
        nrm = max(np.linalg.norm(t), 1e-300)
        e32 = np.linalg.norm(get(v32).astype(np.float64) - t) / nrm
        ecu = np.linalg.norm(get(rm).astype(np.float64) - t) / nrm
        print(f"{n:8s} |cuda-f64|={ecu:.3e}  |f32-f64|={e32:.3e}")
        if not (np.isfinite(ecu) and ecu <= max(rtol, factor * e32)):
            bad.append((n, ecu, e32))
    assert not bad, bad


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS)
@pytest.mark.parametrize("state", ["zero_tracers", "baroclinic"])
def test_operators_one_by_one(oracle_mod, grid_type, Nx, Ny, Nz, state):
    """Each workload applied to identical inputs (the oracle's state is re-uploaded before every stage)."""
    rm, vm = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod, state=state)
    ew = 1e-4 if state == "zero_tracers" else None

    def resync():
        for n in STATE_FIELDS:
            rm.set_parent(n, vm.parent(n))

    M.initialize(vm); M.initialize(rm)
    _assert_close(rm, vm, ("U", "V"), rtol=1e-6)
    M.mask_immersed_model_fields_workload(vm); M.mask_immersed_model_fields_workload(rm)
    for n in ("u", "v", "T", "S"):
        assert np.array_equal(rm.parent(n), vm.parent(n)), f"mask {n} not bit-exact"
    M.tupled_fill_halo_regions_workload(vm); resync()
    M.compute_auxiliaries_workload(vm); M.compute_auxiliaries_workload(rm)
    _assert_close(rm, vm, ("w",), rtol=1e-5, elementwise=1e-5)
    _assert_close(rm, vm, ("p",), rtol=1e-5, elementwise=2e-5)
    resync()
    vm_inputs = {n: vm.parent(n) for n in STATE_FIELDS}       # the inputs of the tendency operators
    M.compute_interior_tracer_tendencies_workload(vm); M.compute_interior_tracer_tendencies_workload(rm)
    M.compute_interior_momentum_tendencies_workload(vm); M.compute_interior_momentum_tendencies_workload(rm)
    if state == "zero_tracers":
        assert not rm.parent("Gn_T").any() and not rm.parent("Gn_S").any()        # exactly 0, not NaN (A.14 item 5)
        _assert_close(rm, vm, ("Gn_u", "Gn_v"), elementwise=ew)
    else:
        _assert_close(rm, vm, ("Gn_v", "Gn_T", "Gn_S"))
        # Gn_u: Float32 round-off of the O(700) pressure divided by dx dominates the small zonal tendency, so the
        # Float32 oracle is itself percent-level away from the truth there.  Criterion: the same operator applied by
        # the Float64 oracle to the same (Float32-valued) inputs is the truth; CUDA must be as close to it as the
        # Float32 oracle is (factor 3), or inside the reference rtol.
        _, v64 = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod, dtype=np.float64, with_cuda=False, state=state)
        for n in STATE_FIELDS:
            v64.set_parent(n, vm_inputs[n])
        M.compute_interior_momentum_tendencies_workload(v64)
        _assert_as_close_as_f32(rm, vm, v64, ("Gn_u", "Gn_v"))
    resync()
    # give G- something to chew on, then the AB2 step with chi = 0.1 and the Euler variant
    for n in ("u", "v", "T", "S"):
        vm.set_parent(f"Gm_{n}", 0.7 * vm.parent(f"Gn_{n}"))
    resync()
    for chi in (0.1, -0.5):
        M.ab2_step_workload(vm, 60.0, chi); M.ab2_step_workload(rm, 60.0, chi)
        _assert_close(rm, vm, ("u", "v", "T", "S", "eta", "U", "V", "filt_eta", "filt_U", "filt_V", "Gn_U", "Gn_V"),
                      rtol=2e-5)
        resync()
    M.correct_velocities_and_cache_previous_tendencies_workload(vm)
    M.correct_velocities_and_cache_previous_tendencies_workload(rm)
    wet = lambda a: np.nan_to_num(a, nan=0.0, posinf=0.0, neginf=0.0)
    for n in ("u", "v"):        # fully solid columns hold 0/0 until the next mask: compare the rest
        a, b = rm.parent(n), vm.parent(n)
        assert np.array_equal(np.isfinite(a), np.isfinite(b))
        assert M.compare_parent(n, wet(a), wet(b), rtol=1e-6, atol=0.0, verbose=True)
    _assert_close(rm, vm, ("filt_U", "filt_V"), rtol=1e-6)
    for n in ("Gm_u", "Gm_v", "Gm_T", "Gm_S", "Gm_U", "Gm_V"):
        assert np.array_equal(rm.parent(n), vm.parent(n)), f"cache {n} not bit-exact"


def _staged(rm, vm, dt, nsteps_loop, rtol, elementwise):
    kw = dict(include_halos=True, throw_error=True, rtol=rtol, atol=0.0, elementwise=elementwise, verbose=False)
    M.compare_states(rm, vm, **kw)                                   # at the beginning
    M.initialize(rm); M.initialize(vm)
    M.update_state(rm); M.update_state(vm)
    M.compare_states(rm, vm, **kw)                                   # after initialization and update state
    M.sync_states(rm, vm)
    M.first_time_step(rm); M.first_time_step(vm)
    M.compare_states(rm, vm, **kw)                                   # after first time step
    for _ in range(2 + 10):
        M.time_step(rm); M.time_step(vm)
    M.compare_states(rm, vm, **kw)                                   # after 2 + 10 steps
    M.sync_states(rm, vm)
    M.update_state(rm)
    M.compare_states(rm, vm, **kw)                                   # after syncing and updating state again
    M.loop(rm, nsteps_loop); M.loop(vm, nsteps_loop)
    M.compare_states(rm, vm, **kw)                                   # after a loop
    assert rm.clock.iteration == vm.clock.iteration == 13 + nsteps_loop


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", [("simple_lat_lon", 112, 112, 16), ("gaussian_islands", 64, 48, 10),
                                                ("simple_lat_lon", 96, 45, 9), ("gaussian_islands", 64, 51, 9)])   # odd Ny: partial tiles / patches
@pytest.mark.parametrize("dt", [1e-9, 60.0])
def test_reference_correctness_protocol(oracle_mod, grid_type, Nx, Ny, Nz, dt):
    """The reference's own staged protocol and state (T = S = 0, random u, v), at its Δt = 1e-9
    (correctness_…_run.jl:21) and at a physical Δt."""
    rm, vm = make_models(grid_type, Nx, Ny, Nz, dt, oracle_mod, state="zero_tracers")
    _staged(rm, vm, dt, 100 if dt < 1 else 30, RTOL, 1e-4)


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS)
def test_baroclinic_state_against_float64_oracle(oracle_mod, grid_type, Nx, Ny, Nz):
    """Baroclinic-instability state, physical Δt: the CUDA result must be as close to the Float64 oracle as
    the Float32 oracle is (factor 3), and within rtol wherever Float32 itself is."""
    rm, v32 = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod)
    _, v64 = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod, dtype=np.float64, with_cuda=False)
    for m in (rm, v32, v64):
        M.first_time_step(m)
        for _ in range(5):
            M.time_step(m)
    names = ["u", "v", "w", "T", "S", "eta", "Gn_u", "Gn_v", "Gn_T", "Gn_S", "filt_U", "filt_V"]
    bad = []
    for n in names:
        t = v64.parent(n).astype(np.float64)
        nrm = max(np.linalg.norm(t), 1e-300)
        e32 = np.linalg.norm(v32.parent(n).astype(np.float64) - t) / nrm
        ecu = np.linalg.norm(rm.parent(n).astype(np.float64) - t) / nrm
        print(f"{n:8s} |cuda-f64|={ecu:.3e}  |f32-f64|={e32:.3e}")
        if not (np.isfinite(ecu) and ecu <= max(RTOL, 3 * e32)):
            bad.append((n, ecu, e32))
    assert not bad, bad


def test_error_behaviour(oracle_mod):
    from gb25_b200 import lib as L
    g = grids.simple_latitude_longitude_grid(32, 16, 4)
    m = M.HydrostaticFreeSurfaceModel(M.B200(0), g)
    with pytest.raises(ValueError):
        m.set_parent("u", np.zeros((3, 3, 3), dtype=np.float32))
    with pytest.raises(L.Gb25Error) as ei:
        m.handle.call("gb25_loop", 1.0, -1)
    assert ei.value.code == L.GB25_ERR_INVALID
    with pytest.raises(L.Gb25Error):
        m.handle.last_loop_seconds() if False else m.handle.check(m.handle.lib.gb25_get_field(m.handle.h, 999, 0))
    M.loop(m, 2)
    m.synchronize()
    assert m.handle.last_loop_seconds() > 0 and m.handle.launch_count() > 0
    m.close()


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS)
def test_fused_step_path_equals_operator_path(oracle_mod, monkeypatch, grid_type, Nx, Ny, Nz):
    """gb25_time_step runs fused kernels (AB2+mask+column sums, corrector+mask, G pointer swap, blocked
    tendency kernels); GB25_FUSED=0 selects the operator-per-kernel path.  Masks, AB2 and corrector are the
    same arithmetic (bit-exact); the blocked tendency kernels differ from the per-cell ones only in
    reciprocal/FMA details, far inside the reference tolerance."""
    import os
    rm_f, vm = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod)
    monkeypatch.setenv("GB25_FUSED", "0")
    rm_o, _ = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod)
    monkeypatch.delenv("GB25_FUSED")
    for m in (rm_f, rm_o):
        M.first_time_step(m)
        for _ in range(4):
            M.time_step(m)
    assert M.compare_states(rm_f, rm_o, include_halos=True, rtol=2e-5, atol=0.0, verbose=False, elementwise=2e-5)


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS + [("gaussian_islands", 256, 96, 12), ("tripolar", 128, 64, 9)])
@pytest.mark.parametrize("closure", [0, 2])
def test_ab2_epilogue_of_the_tendency_kernels_is_bit_identical(monkeypatch, grid_type, Nx, Ny, Nz, closure):
    """Rows A8/A9/A1 fused into A5/A6 (P1 of SURVEY F.1): the tendency kernels that end a step apply the next step's AB2
    update into the second state buffer and accumulate GU, GV and the transport sums; the next step starts with a pointer
    swap.  Same arithmetic, same summation order as the stand-alone AB2 kernels (GB25_SPECULATE=0): every array of the
    model state must agree bit for bit — after single steps, after a loop, after an upload in between (which discards the
    speculation) and after a change of dt (Euler restart)."""
    ph = PhysicsConfig(closure=closure, kappa=1e-3, nu=1e-2)
    ms = []
    for spec in ("1", "0"):
        monkeypatch.setenv("GB25_SPECULATE", spec)
        monkeypatch.setenv("GB25_ZHALO_FOLD", spec)       # the second model also fills every z halo with k_halo_bottom_top
        m = M.baroclinic_instability_model(M.B200(0), Nx, Ny, Nz, Δt=60.0, grid_type=grid_type, physics=ph)
        M.set_baroclinic_instability(m)
        rng = np.random.default_rng(3)
        M.set(m, u=1e-3 * rng.random(m.interior("u").shape), v=1e-3 * rng.random(m.interior("v").shape))
        ms.append(m)
    monkeypatch.delenv("GB25_SPECULATE")
    monkeypatch.delenv("GB25_ZHALO_FOLD")

    def same(tag):
        for n in STATE_FIELDS:
            if n in ("p",):
                continue
            a, b = ms[0].parent(n), ms[1].parent(n)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (tag, n, float(np.nanmax(np.abs(a - b))))

    for m in ms:
        M.first_time_step(m)
        M.time_step(m)
        M.time_step(m)
    same("single steps")
    for m in ms:
        M.loop(m, 4)
    same("loop")
    for m in ms:                                   # an upload discards the speculation
        u = m.parent("u"); u[...] *= np.float32(0.5); m.set_parent("u", u)
        M.update_state(m)
        M.time_step(m); M.time_step(m)
    same("after upload")
    for m in ms:                                   # another dt: Euler restart, the speculation for dt = 60 is not used
        m.clock.last_Δt = 30.0
        M.time_step(m); M.time_step(m); M.time_step(m)
    same("after dt change")
    assert np.isfinite(ms[0].parent("u")).all() and np.abs(ms[0].interior("T")).max() > 1
    for m in ms:
        m.close()


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS)
def test_tma_kernels_equal_blocked_kernels(oracle_mod, monkeypatch, grid_type, Nx, Ny, Nz):
    """The TMA-staged momentum kernels evaluate the same expressions as the register-blocked ones (GB25_TMA=0);
    only the compiler's FMA contraction differs between the two instruction streams, so the model states agree
    to a few ulp (two orders of magnitude inside the reference tolerance)."""
    rm_t, vm = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod)
    monkeypatch.setenv("GB25_TMA", "0")
    rm_b, _ = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod)
    monkeypatch.delenv("GB25_TMA")
    for m in (rm_t, rm_b):
        M.first_time_step(m)
        for _ in range(4):
            M.time_step(m)
    assert M.compare_states(rm_t, rm_b, include_halos=True, rtol=3e-6, atol=0.0, verbose=False, elementwise=1e-5)


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS + [("gaussian_islands", 130, 33, 7)])
def test_packed_pressure_kernel_is_bit_identical(monkeypatch, grid_type, Nx, Ny, Nz):
    """Row A4: k_compute_p2 evaluates TEOS-10 for two columns at once in FP32x2 (FFMA2 rounds each lane like FFMA):
    the hydrostatic pressure anomaly, halo columns 0 and Nx+1 included, must equal the one-column-per-thread kernel
    (GB25_PACKED=0) bit for bit, and nothing outside 0..Nx+1 may be touched."""
    ms = []
    for packed in ("1", "0"):
        monkeypatch.setenv("GB25_PACKED", packed)
        m = M.baroclinic_instability_model(M.B200(0), Nx, Ny, Nz, Δt=60.0, grid_type=grid_type)
        M.set_baroclinic_instability(m)
        rng = np.random.default_rng(5)
        M.set(m, T=m.interior("T") + rng.random(m.interior("T").shape).astype(np.float32))
        M.tupled_fill_halo_regions_workload(m)
        M.compute_auxiliaries_workload(m)
        m.synchronize()
        ms.append(m)
    monkeypatch.delenv("GB25_PACKED")
    a, b = ms[0].parent("p"), ms[1].parent("p")
    assert np.isfinite(a).all() and np.abs(a).max() > 0
    assert np.array_equal(a, b), float(np.abs(a - b).max())
    for m in ms:
        m.close()


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS + [("gaussian_islands", 1440, 600, 4), ("simple_lat_lon", 96, 301, 4)])
@pytest.mark.parametrize("bands", [None, "7"])
def test_persistent_barotropic_kernel_equals_substep_kernels(monkeypatch, grid_type, Nx, Ny, Nz, bands):
    """Row A10: the persistent split-explicit kernel (one CTA per SM owning a band of rows, eta/U/V in shared memory,
    band-to-band flags; gb25_baro.cu) evaluates the expressions of k_baro_eta / k_baro_uv / k_baro_finish
    (GB25_BARO_PERSISTENT=0): the barotropic state must agree bit for bit, with one row per band, with several
    rows per band (GB25_BARO_BANDS) and at the benchmark's horizontal size."""
    if bands:
        monkeypatch.setenv("GB25_BARO_BANDS", bands)
    rm_p = M.baroclinic_instability_model(M.B200(0), Nx, Ny, Nz, Δt=60.0, grid_type=grid_type)
    M.set_baroclinic_instability(rm_p)
    rng = np.random.default_rng(3)
    M.set(rm_p, u=1e-3 * rng.random(rm_p.interior("u").shape), v=1e-3 * rng.random(rm_p.interior("v").shape))
    monkeypatch.setenv("GB25_BARO_PERSISTENT", "0")
    rm_s = M.baroclinic_instability_model(M.B200(0), Nx, Ny, Nz, Δt=60.0, grid_type=grid_type)
    monkeypatch.delenv("GB25_BARO_PERSISTENT")
    M.sync_states(rm_s, rm_p)
    for m in (rm_p, rm_s):
        M.first_time_step(m)
        for _ in range(3):
            M.time_step(m)
        m.synchronize()
    for name in ("eta", "U", "V", "filt_eta", "Gn_U", "Gn_V", "u", "v", "T", "S", "w"):
        a, b = rm_p.parent(name), rm_s.parent(name)
        assert np.isfinite(a).all(), name
        assert np.array_equal(a, b), (name, float(np.abs(a - b).max()))
    rm_p.close(); rm_s.close()


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS)
@pytest.mark.parametrize("closure", [1, 2])
def test_vertical_diffusion_closures(oracle_mod, grid_type, Nx, Ny, Nz, closure):
    """Row A13: VerticalScalarDiffusivity, explicit (tendency term) and vertically implicit (Thomas solve after
    the AB2 update), with coefficients large enough to matter.  Same criterion as the baroclinic-state test:
    as close to the Float64 oracle as the Float32 oracle is (x3), and within rtol wherever Float32 itself is."""
    # strong mixing: K dt / dz^2 ~ 0.03 per step.  The Float32 oracle evaluates the smoothness indicators as sums of
    # squares here (oracle_beta_form=1): on the mixed, very smooth velocity field the expanded form of the reference
    # hits beta + eps == 0 and returns NaN in Float32 (deviation D1 of DESIGN.md; observed in this very test).
    ph = PhysicsConfig(closure=closure, kappa=10.0, nu=50.0, oracle_beta_form=1)
    rm, v32 = make_models(grid_type, Nx, Ny, Nz, 120.0, oracle_mod, physics=ph)
    _, v64 = make_models(grid_type, Nx, Ny, Nz, 120.0, oracle_mod, physics=ph, dtype=np.float64, with_cuda=False)
    _, v64_off = make_models(grid_type, Nx, Ny, Nz, 120.0, oracle_mod, dtype=np.float64, with_cuda=False)
    for m in (rm, v32, v64, v64_off):
        M.first_time_step(m)
        for _ in range(4):
            M.time_step(m)
    bad = []
    for n in ("u", "v", "w", "T", "S", "eta", "Gn_u", "Gn_v", "Gn_T", "Gn_S", "filt_U", "filt_V"):
        t = v64.parent(n).astype(np.float64)
        nrm = max(np.linalg.norm(t), 1e-300)
        e32 = np.linalg.norm(v32.parent(n).astype(np.float64) - t) / nrm
        ecu = np.linalg.norm(rm.parent(n).astype(np.float64) - t) / nrm
        print(f"{n:8s} |cuda-f64|={ecu:.3e}  |f32-f64|={e32:.3e}")
        if not (np.isfinite(ecu) and ecu <= max(RTOL, 3 * e32)):
            bad.append((n, ecu, e32))
    assert not bad, bad
    # and the closure really did something (otherwise the comparison above proves nothing)
    dT = np.linalg.norm(v64.interior("T") - v64_off.interior("T")) / np.linalg.norm(v64_off.interior("T"))
    du = np.linalg.norm(v64.interior("u") - v64_off.interior("u")) / np.linalg.norm(v64_off.interior("u"))
    assert dT > 1e-5 and du > 1e-2


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS)
def test_flux_boundary_conditions(oracle_mod, grid_type, Nx, Ny, Nz):
    """Row A7: compute_boundary_tendencies_workload! with flux boundary conditions (wind stress on u, v, a surface heat flux,
    bottom fluxes on u and S) — the operator on identical inputs, then five steps, against the oracle."""
    rm, vm = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod)
    rng = np.random.default_rng(9)
    for name, side, scale in (("u", "top", 1e-4), ("v", "top", 1e-4), ("T", "top", 1e-4), ("u", "bottom", 1e-5), ("S", "bottom", 1e-6)):
        two_d = {"u": "U", "v": "V", "T": "eta", "S": "eta"}[name]
        J = (scale * rng.standard_normal(rm.handle.field_shape(two_d)[1:])).astype(np.float32)
        for m in (rm, vm):
            M.set_flux_boundary_condition(m, name, side, J)
    M.initialize(vm); M.update_state(vm)
    for n in STATE_FIELDS:
        rm.set_parent(n, vm.parent(n))
    before = {n: vm.parent(n) for n in ("Gn_u", "Gn_v", "Gn_T", "Gn_S")}
    M.compute_boundary_tendencies_workload(vm); M.compute_boundary_tendencies_workload(rm)
    for n in before:
        a, b = rm.parent(n), vm.parent(n)
        assert np.abs(b - before[n]).max() > 0, n                                  # the operator did something
        assert np.allclose(a, b, rtol=2e-6, atol=0), (n, float(np.abs(a - b).max()))
    for m in (rm, vm):
        M.first_time_step(m)
        M.loop(m, 4)
    _, v64 = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod, dtype=np.float64, with_cuda=False)
    # the Float64 oracle with the same conditions is the truth for the "as close as the Float32 oracle" criterion
    rng = np.random.default_rng(9)
    for name, side, scale in (("u", "top", 1e-4), ("v", "top", 1e-4), ("T", "top", 1e-4), ("u", "bottom", 1e-5), ("S", "bottom", 1e-6)):
        two_d = {"u": "U", "v": "V", "T": "eta", "S": "eta"}[name]
        J = (scale * rng.standard_normal(rm.handle.field_shape(two_d)[1:])).astype(np.float32)
        M.set_flux_boundary_condition(v64, name, side, J)
    M.first_time_step(v64)
    M.loop(v64, 4)
    _assert_as_close_as_f32(rm, vm, v64, ("u", "v", "w", "T", "S", "eta", "Gn_u", "Gn_v", "Gn_T", "Gn_S"))
    rm.close()


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS)
@pytest.mark.parametrize("closure", [0, 2])
def test_float64_build_against_float64_oracle(oracle_mod, grid_type, Nx, Ny, Nz, closure):
    """Row f-4: libgb25cuda_f64.so (`--float-type Float64`, the reference CLI's default, src/arg_parsing.jl:28-31) against
    the Float64 oracle: the reference's staged protocol (src/correctness.jl:28-90, halos included) with rtol = 1e-10 instead
    of sqrt(eps(Float64)) = 1.5e-8 — the two differ by FMA contraction and, for the oracle's expanded smoothness indicators,
    by the conditioning of that form (second comparison, rtol 1e-7)."""
    ph = PhysicsConfig(closure=closure, kappa=1e-3, nu=1e-2, oracle_beta_form=1)
    rm, vm = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod, dtype=np.float64, physics=ph, cuda_dtype=np.float64)
    assert rm.parent("u").dtype == np.float64 and rm.handle.lib.gb25_real_bytes() == 8
    kw = dict(include_halos=True, throw_error=True, atol=0.0, verbose=False)
    M.initialize(rm); M.initialize(vm)
    M.update_state(rm); M.update_state(vm)
    M.compare_states(rm, vm, rtol=1e-10, elementwise=1e-9, **kw)
    M.first_time_step(rm); M.first_time_step(vm)
    for _ in range(3):
        M.time_step(rm); M.time_step(vm)
    M.loop(rm, 4); M.loop(vm, 4)
    # eight steps: round-off differences (FMA contraction) have been amplified by the nonlinear WENO weights, as in Float32;
    # the reference's criterion, rtol = sqrt(eps(Float64)) = 1.5e-8 (src/correctness.jl:28), is the bar here
    kw["verbose"] = True
    M.compare_states(rm, vm, rtol=math.sqrt(np.finfo(np.float64).eps), **kw)
    # and against the expanded-form (reference-form) Float64 oracle.  Lat-lon only: on the tripolar grid the two copies of the
    # fold line (row Ny, decision U1 keeps both) are integrated independently, and after eight 60 s steps at the surface they
    # have amplified the round-off difference between the two algebraically identical indicator forms to O(1) there — two
    # Float64 ORACLES that differ only in that form disagree by 30 % on rows Ny-1, Ny (observed), everywhere else by 1e-9.
    if grid_type != "simple_lat_lon":
        rm.close()
        return
    _, vr = make_models(grid_type, Nx, Ny, Nz, 60.0, oracle_mod, dtype=np.float64, with_cuda=False,
                        physics=PhysicsConfig(closure=closure, kappa=1e-3, nu=1e-2))
    M.first_time_step(vr)
    for _ in range(3):
        M.time_step(vr)
    M.loop(vr, 4)
    M.compare_states(rm, vr, rtol=1e-7, **kw)
    rm.close()


@pytest.mark.parametrize("grid_type,Nx,Ny,Nz", GRIDS + [("gaussian_islands", 64, 51, 9), ("simple_lat_lon", 96, 45, 9),
                                                        ("gaussian_islands", 256, 96, 12), ("simple_lat_lon", 130, 33, 7)])
@pytest.mark.parametrize("variant", ["default", "GB25_FUSED=0", "GB25_TMA=0", "GB25_PACKED=0", "GB25_TMA_TRACER=0",
                                     "GB25_BARO_PERSISTENT=0", "GB25_SPECULATE=0", "GB25_OVERLAP=0", "closure=1", "closure=2"])
def test_no_out_of_bounds_stores_under_guard_zones(monkeypatch, grid_type, Nx, Ny, Nz, variant):
    """compute-sanitizer is closed on the shared pool, so the library carries its own store checker: with GB25_GUARD=1 every
    device array sits between two 64 KiB guard zones, and after stepping through every kernel generation — fused and
    operator path, TMA / register-blocked / per-cell tendencies, persistent and per-substep barotropic kernels, with and
    without the AB2 epilogue and the second stream, both closures, sizes with partial tiles — not one guard byte may have
    changed."""
    monkeypatch.setenv("GB25_GUARD", "1")
    closure = 0
    if "=" in variant and variant.startswith("GB25_"):
        k, v = variant.split("=")
        monkeypatch.setenv(k, v)
    elif variant.startswith("closure"):
        closure = int(variant[-1])
    m = M.baroclinic_instability_model(M.B200(0), Nx, Ny, Nz, Δt=60.0, grid_type=grid_type,
                                       physics=PhysicsConfig(closure=closure, kappa=1e-3, nu=1e-2))
    M.set_baroclinic_instability(m)
    rng = np.random.default_rng(1)
    M.set(m, u=1e-3 * rng.random(m.interior("u").shape), v=1e-3 * rng.random(m.interior("v").shape))
    assert m.handle.check_guards() == 0
    M.first_time_step(m)
    M.time_step(m)
    M.loop(m, 3)
    for wl in (M.mask_immersed_model_fields_workload, M.tupled_fill_halo_regions_workload, M.compute_auxiliaries_workload,
               M.compute_tendencies_workload, M.correct_velocities_and_cache_previous_tendencies_workload):
        wl(m)
    M.ab2_step_workload(m, 60.0)
    m.synchronize()
    assert np.isfinite(m.interior("T")).all() and np.isfinite(m.interior("eta")).all()
    assert m.handle.check_guards() == 0, "a kernel stored outside its arrays"
    m.close()


def test_baseline_config_c1_latlon_128x64x8_100_steps(oracle_mod):
    """BASELINE.json configs[0]: baroclinic_instability_model on LatitudeLongitudeGrid 128x64x8, Float32,
    1 Euler + 100 AB2 steps, reference state (T = S = 0, random u, v), against the oracle."""
    rm, vm = make_models("simple_lat_lon", 128, 64, 8, 60.0, oracle_mod, state="zero_tracers")
    M.first_time_step(rm); M.first_time_step(vm)
    M.loop(rm, 100); M.loop(vm, 100)
    assert rm.clock.iteration == vm.clock.iteration == 101
    M.compare_states(rm, vm, include_halos=True, throw_error=True, rtol=RTOL, atol=0.0, elementwise=2e-4, verbose=False)


def test_full_size_properties_tripolar_1440x600x50():
    """BASELINE.json configs[1] at full size (no oracle: it would take minutes): size-independent properties.
    (i) state stays finite; (ii) masks hold: u, v, T, S vanish on every immersed / peripheral node; (iii) halos are
    consistent: periodic wrap and the zipper fold relations hold on the final state; (iv) two identical runs are
    bit-identical (no races, no uninitialised reads); (v) G halos stay identically zero."""
    import bench
    outs = []
    for rep in range(2):
        m = M.baroclinic_instability_model(M.B200(0), 1440, 600, 50, Δt=60.0, grid_type="gaussian_islands")
        bench.synthetic_state(m)
        M.first_time_step(m)
        M.loop(m, 3)
        outs.append({n: m.parent(n) for n in ("u", "v", "T", "eta", "Gn_T", "Gn_u")})
        g = m.grid
        m.close()
    a, b = outs
    for n in a:
        assert np.isfinite(a[n]).all(), n
        assert np.array_equal(a[n].view(np.uint32), b[n].view(np.uint32)), f"{n}: run-to-run difference"
    Hx, Hy, Hz, Nx, Ny, Nz = g.Hx, g.Hy, g.Hz, g.Nx, g.Ny, g.Nz
    zc = np.float32(g.z["z_c"])[Hz:Hz + Nz]
    kb = (zc[None, None, :] <= np.float32(g.bottom_height)[:, :, None]).sum(-1)
    kbi = kb[Hy:Hy + Ny, Hx:Hx + Nx]
    solid = np.arange(1, Nz + 1)[:, None, None] <= kbi[None]
    assert solid.sum() > 1e5
    T = a["T"][Hz:Hz + Nz, Hy:Hy + Ny, Hx:Hx + Nx]
    u = a["u"][Hz:Hz + Nz, Hy:Hy + Ny, Hx:Hx + Nx]
    assert (T[solid] == 0).all() and (u[solid] == 0).all()
    kbw = kb[Hy:Hy + Ny, Hx - 1:Hx + Nx - 1]
    assert (u[np.arange(1, Nz + 1)[:, None, None] <= kbw[None]] == 0).all()        # west neighbour solid => u masked
    for n in ("u", "v", "T"):
        p = a[n]
        assert np.array_equal(p[:, :, :Hx], p[:, :, Nx:Nx + Hx]) and np.array_equal(p[:, :, Nx + Hx:], p[:, :, Hx:2 * Hx])
    kk = slice(Hz, Hz + Nz)
    i = np.arange(1, Nx + 1)
    for mm in (1, 4, 8):
        assert np.array_equal(a["T"][kk, Ny + mm + Hy - 1, i + Hx - 1], a["T"][kk, Ny - mm + Hy - 1, (Nx - i + 1) + Hx - 1])
        assert np.array_equal(a["v"][kk, Ny + mm + Hy - 1, i + Hx - 1], -a["v"][kk, Ny - mm + 1 + Hy - 1, (Nx - i + 1) + Hx - 1])
    for n in ("Gn_T", "Gn_u"):
        p = a[n].copy()
        p[Hz:Hz + Nz, Hy:Hy + Ny, Hx:Hx + Nx] = 0
        assert not p.any(), f"{n}: halo of a tendency array was written"


def _three_way(grid_type, Nx, Ny, Nz, dt, nsteps, oracle_mod):
    """libgb25cuda, the Float32 oracle and the Float64 oracle from the benchmark's synthetic state
    (bench.synthetic_state), first_time_step + nsteps AB2 steps each.  At these sizes the Float32 oracle evaluates
    the smoothness indicators as sums of squares (oracle_beta_form=1): the expanded form of the recalled reference
    hits beta + eps == 0 about once per 1e8 evaluations in Float32 and returns NaN (DESIGN.md deviation D1); the
    Float64 oracle keeps the reference's expanded form."""
    import bench
    ph = PhysicsConfig(oracle_beta_form=1)
    mk = lambda dtype, physics: M.baroclinic_instability_model(oracle_mod.CPUOracle(dtype), Nx, Ny, Nz, Δt=dt, grid_type=grid_type,
                                                                 model_cls=oracle_mod.OracleModel, physics=physics)
    v32, v64 = mk(np.float32, ph), mk(np.float64, None)
    bench.synthetic_state(v32)
    M.sync_states(v64, v32)
    rm = M.baroclinic_instability_model(M.B200(0), Nx, Ny, Nz, Δt=dt, grid_type=grid_type)
    M.sync_states(rm, v32)
    for m in (rm, v32, v64):
        M.first_time_step(m)
        if m is rm:
            M.loop(m, nsteps)                     # the benchmark's entry point (gb25_loop)
        else:
            for _ in range(nsteps):
                M.time_step(m)
    return rm, v32, v64


COMPARED = ("u", "v", "w", "T", "S", "eta", "Gn_u", "Gn_v", "Gn_T", "Gn_S", "Gm_u", "Gm_v", "Gm_T", "Gm_S",
            "filt_U", "filt_V", "filt_eta")


def _parity_at_size(rm, v32, v64, tag):
    """Float32 itself is not accurate to sqrt(eps) on this state after a few steps: the hydrostatic pressure is O(700)
    m2/s2, its Float32 round-off divided by dx is a per-cent-level noise on the zonal tendency, and 60 s steps carry that
    into u, eta and the transports (the Float32 ORACLE is 1e-4 .. 1e-2 away from the Float64 oracle, printed below).
    Criterion per field, halos included: the CUDA result is as close to the Float64 oracle as the Float32 oracle is
    (factor 3) or inside the reference rtol, in the 2-norm (src/correctness.jl:28-90) AND in the max-norm; and it is
    within twice that band of the Float32 oracle itself."""
    bad, lines = [], []
    for n in COMPARED:
        t = v64.parent(n).astype(np.float64)
        a, b = rm.parent(n).astype(np.float64), v32.parent(n).astype(np.float64)
        nrm, mx = max(np.linalg.norm(t), 1e-300), max(np.abs(t).max(), 1e-300)
        e32, ecu, ed = np.linalg.norm(b - t) / nrm, np.linalg.norm(a - t) / nrm, np.linalg.norm(a - b) / nrm
        m32, mcu = np.abs(b - t).max() / mx, np.abs(a - t).max() / mx
        ok = (np.isfinite(ecu) and ecu <= max(RTOL, 3 * e32) and mcu <= max(1e-4, 3 * m32) and ed <= 2 * max(RTOL, 3 * e32))
        lines.append(f"{n:9s} 2-norm: |cuda-f64|={ecu:.3e} |f32-f64|={e32:.3e} |cuda-f32|={ed:.3e}   max-norm: "
                     f"|cuda-f64|={mcu:.3e} |f32-f64|={m32:.3e}  {'ok' if ok else 'FAIL'}")
        if not ok:
            bad.append(n)
    print("\n".join(lines))
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        open(os.path.join(out, f"parity_{tag}.txt"), "w").write("\n".join(lines) + "\n")
    assert not bad, bad


def test_headline_config_against_oracle_tripolar_1440x600x50(oracle_mod):
    """BASELINE.json configs[1] at FULL size — the configuration bench.py times — first step + 2 AB2 steps through
    gb25_first_time_step / gb25_loop against the CPU oracle in Float32 and Float64.  Here 98.5 % of the cells take the
    TMA fast-path kernels that are 64 % of the benchmarked step."""
    rm, v32, v64 = _three_way("gaussian_islands", 1440, 600, 50, 60.0, 2, oracle_mod)
    _parity_at_size(rm, v32, v64, "tripolar_1440x600x50_3steps")
    rm.close()


def test_half_size_10_steps_against_oracle_tripolar_720x300x50(oracle_mod):
    rm, v32, v64 = _three_way("gaussian_islands", 720, 300, 50, 60.0, 10, oracle_mod)
    _parity_at_size(rm, v32, v64, "tripolar_720x300x50_11steps")
    rm.close()


def test_flat_tripolar_fast_path_against_oracle_256x128x12(oracle_mod):
    """No bathymetry: every interior tile runs k_mom_tma_p2 / k_tracer_tma only (no generic list)."""
    rm, v32, v64 = _three_way("tripolar", 256, 128, 12, 60.0, 5, oracle_mod)
    _parity_at_size(rm, v32, v64, "flat_tripolar_256x128x12_6steps")
    rm.close()


def test_full_size_free_surface_volume_conservation_latlon_1440x600x50():
    """Full-size lat-lon run: the split-explicit substeps conserve sum(Az*eta) (no flow through the walls,
    periodic in x) — checked to Float32 round-off over 5 steps."""
    import bench
    m = M.baroclinic_instability_model(M.B200(0), 1440, 600, 50, Δt=60.0, grid_type="simple_lat_lon")
    bench.synthetic_state(m)
    g = m.grid
    az = np.float32(g.metrics["az_cc"])[g.Hy:g.Hy + g.Ny, g.Hx:g.Hx + g.Nx].astype(np.float64)
    M.first_time_step(m)
    v0 = (m.interior("eta")[0].astype(np.float64) * az).sum()
    M.loop(m, 5)
    eta = m.interior("eta")[0].astype(np.float64)
    v1 = (eta * az).sum()
    scale = (np.abs(eta) * az).sum()
    assert np.isfinite(eta).all() and scale > 0
    assert abs(v1 - v0) <= 2e-5 * scale
    m.close()


def test_dump_and_resume_is_bit_exact(tmp_path, oracle_mod):
    """Resume path (the reference only dumps): 3 steps + save + 3 steps == 3 steps + save -> new model + load + 3 steps."""
    from gb25_b200 import sharded_io as IO
    rm, _ = make_models("gaussian_islands", 64, 48, 10, 60.0, oracle_mod)
    M.first_time_step(rm)
    M.loop(rm, 2)
    IO.save_model_state(str(tmp_path), rm)
    M.loop(rm, 3)
    rm2 = M.baroclinic_instability_model(M.B200(0), 64, 48, 10, Δt=60.0, grid_type="gaussian_islands")
    IO.load_model_state(str(tmp_path), rm2)
    M.loop(rm2, 3)
    assert rm2.clock.iteration == rm.clock.iteration == 6
    for n in ("u", "v", "w", "T", "S", "eta", "Gn_u", "Gn_T", "Gm_u", "U", "filt_U"):
        assert np.array_equal(rm.parent(n).view(np.uint32), rm2.parent(n).view(np.uint32)), n


def test_analytic_answers_on_the_device():
    """No oracle in this one: libgb25cuda against analytic answers (tests/analytic_answers.py) — Coriolis sign and magnitude,
    u = -g d(eta)/dx tbar after one short step from a free-surface slope, the centripetal term -u^2 tan(phi)/R of solid-body
    rotation, the advection of a sine of temperature by that rotation, and constancy preservation over the islands and the
    fold.  The thresholds are Float32 ones; the Float32 oracle meets each at a third (CPU test of the same name)."""
    import analytic_answers as AA
    mk = lambda Nx, Ny, Nz, dt, gt: M.baroclinic_instability_model(M.B200(0), Nx, Ny, Nz, Δt=dt, grid_type=gt)
    err = AA.analytic_errors(mk)
    assert set(err) == set(AA.THRESHOLDS)
    bad = {k: (v, AA.THRESHOLDS[k]) for k, v in err.items() if not (np.isfinite(v) and v <= AA.THRESHOLDS[k])}
    assert not bad, bad


_FIRST_RUN = "first run on a GPU is the round-end suite: the round-2 GPU budget was spent before this test existed (XPASS = it works)"


@pytest.mark.xfail(strict=False, reason=_FIRST_RUN)
def test_analytic_answers_on_the_device_float64_build():
    """The Float64 build (libgb25cuda_f64.so, operator-per-kernel generation) against the same analytic answers."""
    import analytic_answers as AA
    mk = lambda Nx, Ny, Nz, dt, gt: M.baroclinic_instability_model(M.B200(0), Nx, Ny, Nz, Δt=dt, grid_type=gt, float_type=np.float64)
    err = AA.analytic_errors(mk)
    bad = {k: (v, AA.THRESHOLDS[k]) for k, v in err.items() if not (np.isfinite(v) and v <= AA.THRESHOLDS[k])}
    assert not bad, bad


@pytest.mark.xfail(strict=False, reason=_FIRST_RUN)
def test_c_example_runs_on_the_device(tmp_path):
    """examples/lat_lon_from_c.c, compiled with gcc against the header alone, steps the model 11 times on the device."""
    import os
    import shutil
    import subprocess
    from gb25_b200 import lib as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    libdir = os.path.dirname(L.LIB_PATH)
    exe = str(tmp_path / "lat_lon_from_c")
    subprocess.run([gcc, "-std=c99", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "lat_lon_from_c.c"),
                    "-o", exe, "-L", libdir, "-lgb25cuda", "-lm"], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=libdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and "iteration 11" in r.stdout and "all finite" in r.stdout, r.stdout + r.stderr
