"""libgb25cuda against the analytic answers of tests/analytic_answers.py, outside pytest (no torch import: a few seconds).
    python scripts/analytic_device_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gb25_b200  # noqa: E402,F401
from gb25_b200 import model as M  # noqa: E402
import analytic_answers as AA  # noqa: E402

err = AA.analytic_errors(lambda Nx, Ny, Nz, dt, gt: M.baroclinic_instability_model(M.B200(0), Nx, Ny, Nz, Δt=dt, grid_type=gt))
ok = True
for k, v in err.items():
    good = v <= AA.THRESHOLDS[k]
    ok = ok and good
    print(f"{k:20s} {v:.3e}  threshold {AA.THRESHOLDS[k]:.0e}  {'ok' if good else 'FAIL'}")
sys.exit(0 if ok else 1)
