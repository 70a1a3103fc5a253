"""CPU-side checks of the boundary and the host logic: the C-ABI library loads and exports every symbol
include/gb25cuda.h declares, fails loudly without a device, and the host mirror follows the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gb25_b200 import grids, lib as L, model as M, sharding
from gb25_b200.config import PhysicsConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(cuda_lib):
    hdr = open(os.path.join(ROOT, "include", "gb25cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(gb25_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(cuda_lib, name), f"{name} declared in gb25cuda.h but not exported"
    assert sorted(L.EXPORTED_SYMBOLS) == declared
    assert cuda_lib.gb25_abi_version() == 1 and cuda_lib.gb25_real_bytes() == 4
    # the Float64 build of the same sources exports the same interface
    f64 = L.load(np.float64)
    for name in declared:
        assert hasattr(f64, name), f"{name} missing from libgb25cuda_f64.so"
    assert f64.gb25_real_bytes() == 8


def test_every_kernel_is_in_the_preload_table(cuda_lib):
    """gb25_create loads every kernel up front (a lazy load behind a wait for a neighbour tile blocks the host: a deadlock
    when one host thread drives all tiles).  The tables are hand-written lists: their total must equal the number of
    kernel entry points in the device code of the built libraries."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    for lib, path in ((cuda_lib, L.LIB_PATH), (L.load(np.float64), L.LIB_PATH_F64)):
        out = subprocess.run([cuobjdump, "--dump-elf-symbols", path], capture_output=True, text=True, check=True).stdout
        entries = {ln.split()[-1] for ln in out.splitlines() if "STT_FUNC" in ln and "STO_ENTRY" in ln}
        assert len(entries) >= 20
        assert lib.gb25_kernel_table_size() == len(entries), (path, sorted(entries))


def test_header_is_plain_c_and_the_c_example_builds_its_grid_like_the_host_mirror(tmp_path, cuda_lib):
    """include/gb25cuda.h compiles as C99 (-pedantic), examples/lat_lon_from_c.c links against the library with gcc alone, its
    grid products equal those of gb25_b200.grids (checksums), and without a device it fails loudly with GB25_ERR_NO_DEVICE."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    src = os.path.join(ROOT, "examples", "lat_lon_from_c.c")
    libdir = os.path.dirname(L.LIB_PATH)
    exe = str(tmp_path / "lat_lon_from_c")
    for flags, lib in (([], "gb25cuda"), (["-DGB25_F64"], "gb25cuda_f64")):
        r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O1", *flags, "-I", os.path.join(ROOT, "include"),
                            src, "-o", exe + lib, "-L", libdir, "-l" + lib, "-lm"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    env = dict(os.environ, LD_LIBRARY_PATH=libdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe + "gb25cuda", "--checksums"], capture_output=True, text=True, env=env, check=True).stdout
    got = dict(ln.split() for ln in out.strip().splitlines())
    g = grids.simple_latitude_longitude_grid(64, 32, 8)
    f32 = lambda a: np.abs(np.asarray(a, dtype=np.float32).astype(np.float64)).sum()
    for k in ("dx_cc", "dx_cf", "dy_cc", "az_cc", "az_cf", "f_ff"):
        assert float(got[k]) == pytest.approx(f32(g.metrics[k]), rel=1e-7), k
    for k in ("z_f", "z_c", "dz_c", "dz_f"):
        assert float(got[k]) == pytest.approx(f32(g.z[k]), rel=1e-7), k
    from gb25_b200.splitexplicit import averaging_weights
    _, w = averaging_weights(30)
    assert int(got["nweights"]) == len(w) and float(got["weights"]) == pytest.approx(f32(w), rel=1e-7)
    if not _has_gpu():
        r = subprocess.run([exe + "gb25cuda"], capture_output=True, text=True, env=env)
        assert r.returncode == 2 and "no CPU path" in r.stderr, (r.returncode, r.stderr)


def test_signatures_carry_no_torch_or_cxx_types():
    hdr = open(os.path.join(ROOT, "include", "gb25cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    assert "torch" not in hdr and "std::" not in hdr and "at::" not in hdr


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device error path")
def test_create_fails_loudly_without_a_device(cuda_lib):
    g = grids.simple_latitude_longitude_grid(16, 16, 4)
    with pytest.raises(L.Gb25Error) as ei:
        M.HydrostaticFreeSurfaceModel(M.B200(0), g)
    assert ei.value.code == L.GB25_ERR_NO_DEVICE
    assert "no CPU path" in str(ei.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package or csrc may reference it."""
    pkg = os.path.join(ROOT, "gb-25_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, fn), errors="ignore").read()
                assert "gb25_oracle" not in src and "libgb25oracle" not in src, fn
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_arch_must_be_b200():
    g = grids.simple_latitude_longitude_grid(16, 16, 4)
    with pytest.raises(TypeError):
        M.HydrostaticFreeSurfaceModel("CPU", g)
    with pytest.raises(ValueError):
        M.make_grid(16, 16, 4, grid_type="cubed_sphere")
    with pytest.raises(TypeError):
        M.baroclinic_instability_model(M.B200(), 16, 16, 4)      # Δt is a required keyword, as in the reference


def test_factors_matches_reference_rule():
    # /root/reference/src/sharding_utils.jl:39-62
    assert sharding.factors(2) == (2, 1) and sharding.factors(8) == (4, 2) and sharding.factors(32) == (8, 4)
    assert sharding.factors(4) == (2, 2) and sharding.factors(16) == (4, 4) and sharding.factors(9180) == (135, 68)
    with pytest.raises(ValueError):
        sharding.factors(3)
    with pytest.raises(ValueError):
        sharding.factors(6)
    assert sharding.global_size_from_tile(1536, 768, 4, 2) == (1536 * 4 - 16, 768 * 2 - 16)
    assert grids.resolution_to_points(8) == (48, 24)      # simulations/baroclinic_instability_simulation_run.jl:12-18


def test_field_shapes_follow_oceananigans_parents():
    g = grids.simple_latitude_longitude_grid(32, 16, 8)
    assert g.field_shape((0, 0, 0)) == (48, 32, 24)
    assert g.field_shape((0, 1, 0)) == (48, 33, 24)        # Bounded Face has N+1 points
    assert g.field_shape((0, 0, 1)) == (48, 32, 25)
    t = grids.tripolar_grid(32, 16, 8)
    assert t.field_shape((0, 1, 0)) == (48, 32, 24)        # RightConnected Face has N points


def test_isapprox_is_norm_based_like_julia():
    a = np.ones((2, 3, 4), dtype=np.float32)
    b = a.copy(); b[0, 0, 0] += 1e-3
    assert M.compare_parent("x", a, b, rtol=1e-3, atol=0, verbose=False)          # 1e-3/sqrt(24) < 1e-3
    assert not M.compare_parent("x", a, b, rtol=1e-5, atol=0, verbose=False)
    assert not M.compare_parent("x", a, b, rtol=1e-3, atol=0, verbose=False, elementwise=1e-4)
    c = a.copy(); c[0, 0, 0] = np.nan
    assert not M.compare_parent("x", a, c, rtol=1.0, atol=0, verbose=False)


def test_state_dump_and_global_reassembly(tmp_path, oracle_mod):
    """sharded_io: per-rank dump + offline reassembly (format of /root/reference/src/sharded_io.jl), exercised on the
    CPU with the oracle model in the model seat (the dump code only needs parent()/interior())."""
    from gb25_b200 import sharded_io as IO
    m = M.baroclinic_instability_model(oracle_mod.CPUOracle(np.float32), 32, 16, 4, Δt=60.0, model_cls=oracle_mod.OracleModel)
    rng = np.random.default_rng(0)
    M.set(m, u=1e-3 * rng.random(m.interior("u").shape), v=1e-3 * rng.random(m.interior("v").shape))
    M.first_time_step(m)
    path = IO.save_model_state(str(tmp_path), m, label="first_loop")
    assert os.path.basename(path) == "fields_rank0.npz"
    allf = IO.load_all_fields(str(tmp_path), label="first_loop")
    assert allf["iteration"] == 1 and allf["time"] == 60.0
    for n in ("u", "v", "w", "T", "S", "eta"):
        assert np.array_equal(allf[n], m.interior(n))
    m2 = M.baroclinic_instability_model(oracle_mod.CPUOracle(np.float32), 32, 16, 4, Δt=60.0, model_cls=oracle_mod.OracleModel)
    IO.load_model_state(str(tmp_path), m2, label="first_loop")
    M.time_step(m); M.time_step(m2)
    for n in ("u", "v", "T", "eta", "Gn_u", "Gm_T"):
        assert np.array_equal(m.parent(n), m2.parent(n)), n
