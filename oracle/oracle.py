"""ctypes wrapper of the CPU oracle (oracle/gb25_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

*** PARITY UNPINNED *** (see the header of gb25_oracle.cpp and DESIGN.md): the oracle restates
SURVEY.md Appendix A and is pinned only by self-contained known answers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  ``OracleModel`` duck-types the product's ``HydrostaticFreeSurfaceModel`` so that the
reference-style stepping functions and ``compare_states`` of ``gb25_b200.model`` run on both.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

import gb25_b200  # noqa: F401  (host-side grid products and the interface being mirrored)
from gb25_b200 import grids as _grids
from gb25_b200.config import PhysicsConfig
from gb25_b200.lib import FIELD_ID, FIELD_LOC
from gb25_b200.model import Clock, ModelBase
from gb25_b200.splitexplicit import averaging_weights

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgb25oracle.so")


class OConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("Nx", "Ny", "Nz", "Hx", "Hy", "Hz", "topo_y", "immersed", "nsub",
                                       "coriolis_scheme", "fold_variant", "south_inactive", "cond_diff", "eos_r0", "beta_form", "closure")] + \
               [(n, C.c_double) for n in ("g", "rho0", "chi", "dtau_frac", "weno_eps", "kappa", "nu")]


_lib = None


def build(force=False):
    src = os.path.join(_HERE, "gb25_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def set_num_threads(n=None):
    """OpenMP threads of the oracle (default: every host core).  Returns the count in effect."""
    lib = load()
    lib.gb25o_set_num_threads(int(n or os.cpu_count() or 1))
    return int(lib.gb25o_max_threads())


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.gb25o_set_num_threads.argtypes = [C.c_int]
        _lib.gb25o_max_threads.restype = C.c_int
        for suf, ft in (("f32", C.c_float), ("f64", C.c_double)):
            P = C.POINTER(ft)
            getattr(_lib, f"gb25o_create_{suf}").restype = C.c_void_p
            getattr(_lib, f"gb25o_create_{suf}").argtypes = [C.POINTER(OConfig), C.POINTER(P), C.POINTER(P), P, P]
            getattr(_lib, f"gb25o_destroy_{suf}").argtypes = [C.c_void_p]
            getattr(_lib, f"gb25o_field_{suf}").restype = P
            getattr(_lib, f"gb25o_field_{suf}").argtypes = [C.c_void_p, C.c_int]
            getattr(_lib, f"gb25o_kbot_{suf}").restype = C.POINTER(C.c_int)
            getattr(_lib, f"gb25o_kbot_{suf}").argtypes = [C.c_void_p]
            getattr(_lib, f"gb25o_depth_{suf}").restype = P
            getattr(_lib, f"gb25o_depth_{suf}").argtypes = [C.c_void_p, C.c_int]
            getattr(_lib, f"gb25o_set_clock_{suf}").argtypes = [C.c_void_p, C.c_double, C.c_long, C.c_double]
            getattr(_lib, f"gb25o_op_{suf}").argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
            getattr(_lib, f"gb25o_fill_halo_{suf}").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
            getattr(_lib, f"gb25o_set_flux_bc_{suf}").argtypes = [C.c_void_p, C.c_int, C.c_int, P]
            getattr(_lib, f"gb25o_rho_prime_{suf}").restype = ft
            getattr(_lib, f"gb25o_rho_prime_{suf}").argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
            getattr(_lib, f"gb25o_weno_{suf}").restype = ft
            getattr(_lib, f"gb25o_weno_{suf}").argtypes = [C.c_void_p, C.c_int, C.c_int, P, P, P]
    return _lib


_OPS = {
    "gb25_initialize": 0, "gb25_update_state": 1, "gb25_first_time_step": 2, "gb25_time_step": 3,
    "gb25_mask_immersed_fields": 4, "gb25_fill_halo_regions": 5, "gb25_compute_auxiliaries": 6,
    "gb25_compute_tendencies": 7, "gb25_ab2_step": 8,
    "gb25_correct_velocities_and_cache_previous_tendencies": 9,
    "gb25_compute_momentum_tendencies": 12, "gb25_compute_tracer_tendencies": 13,
    "gb25_compute_boundary_tendencies": 14,
}


class CPUOracle:
    """Architecture tag of the oracle (the seat Oceananigans' ``CPU()`` takes in correctness/*.jl)."""
    def __init__(self, dtype=np.float32):
        self.dtype = np.dtype(dtype)


class _OracleHandle:
    def __init__(self, model):
        self.m = model

    def call(self, name, *args):
        m = self.m
        if name == "gb25_synchronize":
            return 0
        if name == "gb25_loop":
            dt, n = args
            for _ in range(n):
                m._op(3, dt, 0.0)
            return 0
        dt = args[0] if len(args) > 0 else 0.0
        chi = args[1] if len(args) > 1 else 0.0
        m._op(_OPS[name], dt, chi)
        return 0

    def set_clock(self, t, it, last_dt):
        getattr(self.m.lib, f"gb25o_set_clock_{self.m.suf}")(self.m.h, t, it, last_dt)

    def get_clock(self):
        return self.m._clock_shadow


class OracleModel(ModelBase):
    def __init__(self, arch, grid, physics=None):
        self.lib = load()
        self.arch = arch
        self.dtype = arch.dtype
        self.suf = "f32" if self.dtype == np.float32 else "f64"
        self.ft = C.c_float if self.dtype == np.float32 else C.c_double
        self.grid = grid
        self.physics = physics or PhysicsConfig()
        self.dtau_frac, self.weights = averaging_weights(self.physics.substeps)
        p = self.physics
        cfg = OConfig(grid.Nx, grid.Ny, grid.Nz, grid.Hx, grid.Hy, grid.Hz, grid.topo_y, 1 if grid.immersed else 0,
                      len(self.weights), p.coriolis_scheme, p.fold_variant, p.south_inactive, p.cond_diff, p.eos_r0, getattr(p, "oracle_beta_form", 0), p.closure,
                      float(np.float32(p.g)), float(np.float32(p.rho0)), float(np.float32(p.chi)),
                      float(np.float32(self.dtau_frac)), float(np.float32(p.weno_eps)),
                      float(np.float32(p.kappa)), float(np.float32(p.nu)))
        # inputs are rounded to Float32 first (what the host model holds), then promoted for the f64 oracle
        conv = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32).astype(self.dtype))
        self._keep = []
        P = C.POINTER(self.ft)
        g2 = (P * 13)()
        for q, name in enumerate(_grids.METRIC_NAMES):
            a = conv(grid.metrics[name]); self._keep.append(a); g2[q] = a.ctypes.data_as(P)
        gz = (P * 4)()
        for q, name in enumerate(_grids.Z_NAMES):
            a = conv(grid.z[name]); self._keep.append(a); gz[q] = a.ctypes.data_as(P)
        bh = None
        if grid.bottom_height is not None:
            a = conv(grid.bottom_height); self._keep.append(a); bh = a.ctypes.data_as(P)
        w = conv(self.weights); self._keep.append(w)
        self.h = getattr(self.lib, f"gb25o_create_{self.suf}")(C.byref(cfg), g2, gz, bh, w.ctypes.data_as(P))
        self.handle = _OracleHandle(self)
        self.clock = Clock()
        self._clock_shadow = (0.0, 0, 0.0)

    def close(self):
        if self.h:
            getattr(self.lib, f"gb25o_destroy_{self.suf}")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- raw (uniform-layout) views, zero copy
    def raw(self, name):
        g = self.grid
        ptr = getattr(self.lib, f"gb25o_field_{self.suf}")(self.h, FIELD_ID[name])
        three_d = FIELD_LOC[name][3]
        shape = (g.PZ, g.PY, g.PX) if three_d else (1, g.PY, g.PX)
        return np.ctypeslib.as_array(ptr, shape=shape)

    def _parent_shape(self, name):
        lx, ly, lz, three_d = FIELD_LOC[name]
        tx, ty, tz = self.grid.field_shape((lx, ly, lz))
        return (tz if three_d else 1, ty, tx)

    def parent(self, name):
        s = self._parent_shape(name)
        return self.raw(name)[:s[0], :s[1], :s[2]].copy()

    def set_parent(self, name, a):
        s = self._parent_shape(name)
        a = np.asarray(a)
        if a.shape != s:
            raise ValueError(f"{name}: parent shape {a.shape} != {s}")
        self.raw(name)[:s[0], :s[1], :s[2]] = a.astype(self.dtype)

    def set_flux_boundary_condition(self, name, side, values):
        """FluxBoundaryCondition(values) at the bottom / top of u, v, T or S; None = default no-flux."""
        g = self.grid
        q, sd = ("u", "v", "T", "S").index(name), ("bottom", "top").index(side)
        fn = getattr(self.lib, f"gb25o_set_flux_bc_{self.suf}")
        if values is None:
            fn(self.h, q, sd, None)
            return
        a = np.zeros((g.PY, g.PX), dtype=self.dtype)
        v = np.asarray(values, dtype=np.float32).astype(self.dtype)
        a[:v.shape[0], :v.shape[1]] = v
        fn(self.h, q, sd, a.ctypes.data_as(C.POINTER(self.ft)))

    def kbot(self):
        g = self.grid
        return np.ctypeslib.as_array(getattr(self.lib, f"gb25o_kbot_{self.suf}")(self.h), shape=(g.PY, g.PX))

    def _push_clock(self):
        dt = self.clock.last_Δt
        self.handle.set_clock(self.clock.time, self.clock.iteration, 0.0 if math.isinf(dt) else dt)

    def _pull_clock(self):
        pass

    def _op(self, op, dt=0.0, chi=0.0):
        getattr(self.lib, f"gb25o_op_{self.suf}")(self.h, op, float(dt), float(chi))
        if op in (2, 3):
            self.clock.time += float(dt)
            self.clock.iteration += 1
            self.clock.last_Δt = float(np.float32(dt))
            # keep the C++ side's clock in step for consecutive raw ops
            self.handle.set_clock(self.clock.time, self.clock.iteration, self.clock.last_Δt)

    def synchronize(self):
        pass

    # --- direct access to sub-operators used by the known-answer tests
    def rho_prime(self, T, S, Z):
        return float(getattr(self.lib, f"gb25o_rho_prime_{self.suf}")(self.h, T, S, Z))

    def weno(self, B, left, q, s1=None, s2=None):
        P = C.POINTER(self.ft)
        arr = lambda a: None if a is None else np.ascontiguousarray(a, dtype=self.dtype)
        q_, s1_, s2_ = arr(q), arr(s1), arr(s2)
        ptr = lambda a: None if a is None else a.ctypes.data_as(P)
        return float(getattr(self.lib, f"gb25o_weno_{self.suf}")(self.h, B, 1 if left else 0, ptr(q_), ptr(s1_), ptr(s2_)))

    def fill_halo(self, name, sign=1.0):
        lx, ly, lz, three_d = FIELD_LOC[name]
        getattr(self.lib, f"gb25o_fill_halo_{self.suf}")(self.h, FIELD_ID[name], lx, ly, lz, float(sign), 1 if three_d else 0)


def oracle_model(grid_or_args, physics=None, dtype=np.float32, **kw):
    """OracleModel from a Grid."""
    return OracleModel(CPUOracle(dtype), grid_or_args, physics)
