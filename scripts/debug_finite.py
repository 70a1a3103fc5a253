"""Find where/when the C2 run goes non-finite."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gb25_b200
from gb25_b200 import model as M
from bench import synthetic_state
dt = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
m = M.baroclinic_instability_model(M.B200(0), 1440, 600, 50, Δt=dt, grid_type="gaussian_islands")
g = m.grid
synthetic_state(m)
def report(tag):
    bad = False
    for n in ("eta", "u", "v", "w", "T", "Gn_u", "Gn_v", "Gn_T", "U", "V"):
        a = m.interior(n)
        fin = np.isfinite(a)
        mx = np.abs(np.where(fin, a, 0)).max()
        loc = np.unravel_index(np.abs(np.where(fin, a, 0)).argmax(), a.shape)
        msg = f"  {n:5s} max|.|={mx:.3e} at k,j,i={loc}"
        if not fin.all():
            idx = np.argwhere(~fin)
            msg += f"  NONFINITE n={len(idx)} first={idx[0]} last={idx[-1]}"
            bad = True
        print(msg)
    return bad
M.first_time_step(m)
print("after first step"); report("")
for s in range(2, 14):
    M.time_step(m)
    print("after step", s)
    if report(""):
        break
kb = None
