"""Regenerates tests/golden/oracle_states_r1.npz: the state of the Float64 CPU oracle after first_time_step! + 4 AB2 steps
on two small grids.  SELF-GENERATED regression fixtures (the reference holds no golden vectors and cannot run here:
DESIGN.md section 5): they pin the oracle against silent drift — later edits of oracle/gb25_oracle.cpp, of the grid
generators or of the averaging weights must reproduce these numbers — they do NOT pin it to Oceananigans.

    python tests/golden/make_golden.py            (from the repo root; takes a few seconds)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = {"latlon_24x16x5": ("simple_lat_lon", 24, 16, 5), "islands_24x16x6": ("gaussian_islands", 24, 16, 6)}
FIELDS = ("u", "v", "w", "T", "S", "eta", "U", "V", "Gn_u", "Gn_v", "Gn_T", "Gn_S")
NSTEPS = 4
DT = 60.0


def run_case(grid_type, Nx, Ny, Nz, dtype=np.float64):
    import gb25_b200  # noqa: F401
    from gb25_b200 import model as M
    from oracle import oracle as O
    O.build(); O.load()
    m = M.baroclinic_instability_model(O.CPUOracle(dtype), Nx, Ny, Nz, Δt=DT, grid_type=grid_type, model_cls=O.OracleModel)
    M.set_baroclinic_instability(m)
    rng = np.random.default_rng(42)
    M.set(m, u=1e-3 * rng.random(m.interior("u").shape), v=1e-3 * rng.random(m.interior("v").shape))
    M.first_time_step(m)
    for _ in range(NSTEPS):
        M.time_step(m)
    return {f: np.asarray(m.interior(f), dtype=np.float64) for f in FIELDS}


if __name__ == "__main__":
    out = {}
    for name, (gt, Nx, Ny, Nz) in CASES.items():
        for f, a in run_case(gt, Nx, Ny, Nz).items():
            out[f"{name}/{f}"] = a
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_states_r1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")
