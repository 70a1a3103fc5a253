"""Analytic answers for the time-step path that need no oracle: a model factory goes in, relative errors come out.
Used with the Float32 CPU oracle (tests/test_oracle_known_answers.py, CPU) and with libgb25cuda through the C ABI
(tests/test_cuda_parity.py, GPU): the same states, the same formulas, the same thresholds — so the device path is pinned to
physics directly, not only to the oracle.  Derivations: tests/test_oracle_known_answers.py (Float64 versions of the same
checks, with convergence orders)."""
import numpy as np

from gb25_b200 import grids, model as M
from gb25_b200.config import PhysicsConfig
from gb25_b200.splitexplicit import averaging_weights

# thresholds for Float32 arithmetic (the Float32 oracle sits at <= 1/3 of each; the discretisation error dominates)
THRESHOLDS = {"coriolis_Gu": 1e-3, "coriolis_Gv": 1e-3, "slope_u": 1e-4, "slope_v": 1e-4, "rotation_Gu": 1e-4,
              "rotation_metric": 2e-3, "rotation_coriolis": 1e-3, "tracer_GT": 5e-4, "tracer_GS": 1e-4, "uniform_GT": 1e-5,
              "uniform_GS": 1e-5}


def _f64(a):
    return np.asarray(a, dtype=np.float64)


def analytic_errors(make_model):
    """``make_model(Nx, Ny, Nz, dt, grid_type)`` -> model (Float32).  Returns {check: relative error}."""
    err = {}
    Omega, R = grids.OMEGA_EARTH, grids.R_EARTH

    # 1. Coriolis: vanishing uniform flow, uniform T, S  =>  Gu = +f v, Gv = -f u
    m = make_model(64, 96, 8, 1.0, "simple_lat_lon")
    g = m.grid
    U0, V0 = 1e-8, -2e-8
    M.set(m, T=10.0 + 0 * m.interior("T"), S=35.0 + 0 * m.interior("S"), u=U0 + 0 * m.interior("u"), v=V0 + 0 * m.interior("v"))
    M.update_state(m)
    phi_c = np.deg2rad(_f64(g.phi_cc)[g.Hy:g.Hy + g.Ny, g.Hx])
    f_c = 2 * Omega * np.sin(phi_c)
    js = slice(3, g.Ny - 3)
    fu = f_c[js][None, :, None]
    err["coriolis_Gu"] = np.abs(_f64(m.interior("Gn_u"))[:, js] - fu * V0).max() / np.abs(fu * V0).max()
    f_f = 0.5 * (f_c[2:g.Ny - 4] + f_c[3:g.Ny - 3])[None, :, None]
    err["coriolis_Gv"] = np.abs(_f64(m.interior("Gn_v"))[:, js] + f_f * U0).max() / np.abs(f_f * U0).max()
    m.close()

    # 2. free-surface slope: one short step from rest  =>  u = -g d(eta)/dx tbar, v = 0
    m = make_model(64, 32, 8, 1.0, "simple_lat_lon")
    g = m.grid
    lam = np.deg2rad(_f64(g.lam_cc)[g.Hy:g.Hy + g.Ny, g.Hx:g.Hx + g.Nx])
    phi = np.deg2rad(_f64(g.phi_cc)[g.Hy:g.Hy + g.Ny, g.Hx:g.Hx + g.Nx])
    A = 1e-3
    M.set(m, T=10.0 + 0 * m.interior("T"), S=35.0 + 0 * m.interior("S"), u=0 * m.interior("u"), v=0 * m.interior("v"),
          eta=(A * np.sin(lam))[None])
    M.first_time_step(m)
    dlam = 2 * np.pi / g.Nx
    deta_dx = A * np.cos(lam - 0.5 * dlam) * np.sin(0.5 * dlam) / (0.5 * dlam) / (R * np.cos(phi))
    frac, w = averaging_weights(30)
    tbar = 1.0 * frac * float((w * np.arange(1, len(w) + 1)).sum())
    u_exp = -PhysicsConfig().g * deta_dx * tbar
    err["slope_u"] = np.abs(_f64(m.interior("u")) - u_exp[None]).max() / np.abs(u_exp).max()
    err["slope_v"] = np.abs(_f64(m.interior("v"))).max() / np.abs(u_exp).max()
    m.close()

    # 3. solid-body rotation u = U0 cos(phi)  =>  Gu = 0, Gv = -f u - u^2 tan(phi) / R
    U0, Ny = 10.0, 96
    Gv = {}
    err["rotation_Gu"] = 0.0
    for sgn in (1.0, -1.0):
        m = make_model(64, Ny, 8, 1.0, "simple_lat_lon")
        g = m.grid
        phi_c = np.deg2rad(_f64(g.phi_cc)[g.Hy:g.Hy + g.Ny, g.Hx])
        M.set(m, T=10.0 + 0 * m.interior("T"), S=35.0 + 0 * m.interior("S"),
              u=sgn * U0 * np.cos(phi_c)[None, :, None] + 0 * m.interior("u"), v=0 * m.interior("v"))
        M.update_state(m)
        Gv[sgn] = _f64(m.interior("Gn_v"))[:, 1:Ny, :]
        scale = 2 * Omega * U0
        err["rotation_Gu"] = max(err["rotation_Gu"], np.abs(_f64(m.interior("Gn_u"))).max() / scale)
        m.close()
    phi_f = 0.5 * (phi_c[:-1] + phi_c[1:])
    u_f = U0 * np.cos(phi_f)
    js = slice(6, Ny - 7)
    metric = 0.5 * (Gv[1.0] + Gv[-1.0])[:, js]
    cor = 0.5 * (Gv[1.0] - Gv[-1.0])[:, js]
    e_m = (-u_f ** 2 * np.tan(phi_f) / R)[js][None, :, None]
    e_c = (-2 * Omega * np.sin(phi_f) * u_f)[js][None, :, None]
    err["rotation_metric"] = np.abs(metric - e_m).max() / np.abs(e_m).max()
    err["rotation_coriolis"] = np.abs(cor - e_c).max() / np.abs(e_c).max()

    # 4. the same rotation carries T = T0 + a sin(lambda): dT/dt = -(U0/R) a cos(lambda) dphi / (2 sin(dphi/2)); S uniform
    m = make_model(64, 24, 8, 1.0, "simple_lat_lon")
    g = m.grid
    phi_c = np.deg2rad(_f64(g.phi_cc)[g.Hy:g.Hy + g.Ny, g.Hx])
    lam_c = np.deg2rad(_f64(g.lam_cc)[g.Hy, g.Hx:g.Hx + g.Nx])
    dphi = phi_c[1] - phi_c[0]
    err["tracer_GT"] = err["tracer_GS"] = 0.0
    for sgn in (1.0, -1.0):
        M.set(m, T=10.0 + 0.5 * np.sin(lam_c)[None, None, :] + 0 * m.interior("T"), S=35.0 + 0 * m.interior("S"),
              u=sgn * U0 * np.cos(phi_c)[None, :, None] + 0 * m.interior("u"), v=0 * m.interior("v"))
        M.update_state(m)
        e = -sgn * (U0 / R) * 0.5 * np.cos(lam_c) * dphi / (2 * np.sin(dphi / 2))
        err["tracer_GT"] = max(err["tracer_GT"], np.abs(_f64(m.interior("Gn_T")) - e[None, None, :]).max() / np.abs(e).max())
        err["tracer_GS"] = max(err["tracer_GS"], np.abs(_f64(m.interior("Gn_S"))).max() / (35.0 * U0 / R))
    m.close()

    # 5. constancy preservation: uniform T, S under an arbitrary masked flow (w from continuity), islands and fold included
    m = make_model(64, 48, 8, 30.0, "gaussian_islands")
    g = m.grid
    rng = np.random.default_rng(3)
    Ur = 0.5
    M.set(m, T=12.5 + 0 * m.interior("T"), S=35.0 + 0 * m.interior("S"),
          u=Ur * (rng.random(m.interior("u").shape) - 0.5), v=Ur * (rng.random(m.interior("v").shape) - 0.5))
    M.update_state(m)
    dx = min(float(np.min(_f64(g.metrics[k])[g.Hy:g.Hy + g.Ny, g.Hx:g.Hx + g.Nx])) for k in ("dx_cc", "dy_cc"))
    err["uniform_GT"] = np.abs(_f64(m.interior("Gn_T"))).max() / (12.5 * Ur / dx)
    err["uniform_GS"] = np.abs(_f64(m.interior("Gn_S"))).max() / (35.0 * Ur / dx)
    m.close()
    return {k: float(v) for k, v in err.items()}
