"""Short C2 run for ncu: first step + 2 AB2 steps of the 1440x600x50 tripolar workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gb25_b200
from gb25_b200 import model as M
from bench import synthetic_state, WORKLOADS
wl = sys.argv[1] if len(sys.argv) > 1 else "tripolar_quarter_degree"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
gt, Nx, Ny, Nz, dt = WORKLOADS[wl]
m = M.baroclinic_instability_model(M.B200(0), Nx, Ny, Nz, Δt=dt, grid_type=gt)
synthetic_state(m)
M.first_time_step(m)
M.loop(m, n)
m.synchronize()
print("ok", m.handle.launch_count(), "launches;", m.handle.last_loop_seconds() / n * 1e3, "ms/step")
