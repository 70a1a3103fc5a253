import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure).  Built on demand with gcc."""
    from oracle import oracle as O
    O.build()
    O.load()
    return O


@pytest.fixture(scope="session")
def cuda_lib():
    """libgb25cuda through ctypes; GPU tests fail loudly if it is missing (no fallback)."""
    from gb25_b200 import lib
    return lib.load()


def make_models(grid_type, Nx, Ny, Nz, dt, oracle_mod, dtype=np.float32, physics=None, seed=42, with_cuda=True,
                state="baroclinic", cuda_dtype=np.float32):
    """Build the (cuda, oracle) pair of the reference's correctness protocol
    (/root/reference/correctness/correctness_baroclinic_instability_simulation_run.jl:33-43):
    same model on both architectures, random u, v, state synchronised."""
    from gb25_b200 import model as M
    vm = M.baroclinic_instability_model(oracle_mod.CPUOracle(dtype), Nx, Ny, Nz, Δt=dt, grid_type=grid_type,
                                        model_cls=oracle_mod.OracleModel, physics=physics)
    rng = np.random.default_rng(seed)
    if state == "baroclinic":
        M.set_baroclinic_instability(vm)
    M.set(vm, u=1e-3 * rng.random(vm.interior("u").shape), v=1e-3 * rng.random(vm.interior("v").shape))
    rm = None
    if with_cuda:
        rm = M.baroclinic_instability_model(M.B200(0), Nx, Ny, Nz, Δt=dt, grid_type=grid_type, physics=physics,
                                            float_type=cuda_dtype)
        M.sync_states(rm, vm)
    return rm, vm
