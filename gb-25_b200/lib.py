"""ctypes binding of libgb25cuda (include/gb25cuda.h).

This is the binding a maintainer of the reference would write with Julia's ``ccall``
(INTEGRATION.md); here it is Python because Julia is not available in the build image.
The library is the only compute path: if it is missing, or no CUDA device is present,
calls fail loudly (``Gb25Error``) — nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GB25_LIB selects another in-tree build of the same library (kernel experiments: gb25_b200.build --suffix); it is never
# a different implementation, and a missing file still fails loudly.
LIB_PATH = os.environ.get("GB25_LIB") or os.path.join(_HERE, "csrc", "libgb25cuda.so")
# the Float64 build of the same sources (include/gb25cuda.h: gb25_real = double); `--float-type Float64` of the reference CLI
LIB_PATH_F64 = os.environ.get("GB25_LIB_F64") or os.path.join(_HERE, "csrc", "libgb25cuda_f64.so")

GB25_OK, GB25_ERR_INVALID, GB25_ERR_NO_DEVICE, GB25_ERR_CUDA, GB25_ERR_ALLOC, GB25_ERR_COMM = 0, -1, -2, -3, -4, -5

FIELD_NAMES = ("u", "v", "w", "T", "S", "p",
               "Gn_u", "Gn_v", "Gn_T", "Gn_S", "Gm_u", "Gm_v", "Gm_T", "Gm_S",
               "eta", "U", "V", "filt_eta", "filt_U", "filt_V",
               "Gn_U", "Gn_V", "Gm_U", "Gm_V")
FIELD_ID = {n: i for i, n in enumerate(FIELD_NAMES)}
# (lx, ly, lz, three_d) — staggering of every field, 1 = Face
FIELD_LOC = {
    "u": (1, 0, 0, True), "v": (0, 1, 0, True), "w": (0, 0, 1, True), "T": (0, 0, 0, True),
    "S": (0, 0, 0, True), "p": (0, 0, 0, True),
    "Gn_u": (1, 0, 0, True), "Gn_v": (0, 1, 0, True), "Gn_T": (0, 0, 0, True), "Gn_S": (0, 0, 0, True),
    "Gm_u": (1, 0, 0, True), "Gm_v": (0, 1, 0, True), "Gm_T": (0, 0, 0, True), "Gm_S": (0, 0, 0, True),
    "eta": (0, 0, 1, False), "U": (1, 0, 0, False), "V": (0, 1, 0, False),
    "filt_eta": (0, 0, 1, False), "filt_U": (1, 0, 0, False), "filt_V": (0, 1, 0, False),
    "Gn_U": (1, 0, 0, False), "Gn_V": (0, 1, 0, False), "Gm_U": (1, 0, 0, False), "Gm_V": (0, 1, 0, False),
}

EXPORTED_SYMBOLS = (
    "gb25_abi_version", "gb25_real_bytes", "gb25_create", "gb25_destroy", "gb25_last_error", "gb25_clear_error",
    "gb25_field_shape", "gb25_set_field", "gb25_get_field", "gb25_set_clock", "gb25_get_clock",
    "gb25_interior_shape", "gb25_set_interior", "gb25_get_interior", "gb25_set_fields", "gb25_get_fields",
    "gb25_initialize", "gb25_update_state", "gb25_first_time_step", "gb25_time_step", "gb25_loop",
    "gb25_synchronize", "gb25_mask_immersed_fields", "gb25_fill_halo_regions", "gb25_compute_auxiliaries",
    "gb25_compute_tendencies", "gb25_compute_momentum_tendencies", "gb25_compute_tracer_tendencies",
    "gb25_compute_boundary_tendencies", "gb25_set_flux_boundary_condition",
    "gb25_ab2_step", "gb25_correct_velocities_and_cache_previous_tendencies",
    "gb25_last_loop_seconds", "gb25_kernel_launch_count", "gb25_enable_stage_timers", "gb25_get_stage_times",
    "gb25_check_guards", "gb25_kernel_table_size",
    "gb25_exchange_blob_size", "gb25_exchange_export", "gb25_exchange_connect",
    "gb25_exchange_connect_local", "gb25_loop_all",
)


class Gb25Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgb25cuda error {code}: {msg}")
        self.code = code


class gb25_config(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("Nx", "Ny", "Nz", "Hx", "Hy", "Hz", "topo_y", "immersed", "nsubsteps",
                                       "coriolis_scheme", "fold_variant", "south_inactive", "cond_diff", "eos_r0")] + \
               [(n, C.c_float) for n in ("g", "rho0", "chi", "dtau_frac", "weno_eps")] + \
               [(n, C.c_int) for n in ("Rx", "Ry", "rx", "ry", "device", "closure")] + \
               [(n, C.c_float) for n in ("kappa", "nu")]


_GRID_PTRS = ("dx_cc", "dx_fc", "dx_cf", "dx_ff", "dy_cc", "dy_fc", "dy_cf", "dy_ff",
              "az_cc", "az_fc", "az_cf", "az_ff", "f_ff", "z_f", "z_c", "dz_c", "dz_f",
              "bottom_height", "avg_weights")


class gb25_grid(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _GRID_PTRS]      # const gb25_real* (const float* for avg_weights)


_libs = {}


def load(float_type=np.float32):
    """Load libgb25cuda.so (Float32) or libgb25cuda_f64.so (Float64), both in-tree.  Raises if the library has not been
    built: there is no fallback."""
    ft = np.dtype(float_type)
    if ft in _libs:
        return _libs[ft]
    path = LIB_PATH if ft == np.float32 else LIB_PATH_F64
    if ft not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("float_type must be Float32 or Float64")
    if not os.path.exists(path):
        raise Gb25Error(GB25_ERR_NO_DEVICE,
                        f"{path} not found — build it with __graft_entry__.build(); there is no CPU fallback")
    lib = C.CDLL(path)
    lib.gb25_real_bytes.restype = C.c_int
    if lib.gb25_real_bytes() != ft.itemsize:
        raise Gb25Error(GB25_ERR_INVALID, f"{path} is not the {ft} build")
    H = C.c_void_p
    lib.gb25_abi_version.restype = C.c_int
    lib.gb25_create.argtypes = [C.POINTER(gb25_config), C.POINTER(gb25_grid), C.POINTER(H)]
    lib.gb25_last_error.argtypes = [H]
    lib.gb25_last_error.restype = C.c_char_p
    for name in ("gb25_destroy", "gb25_clear_error", "gb25_initialize", "gb25_update_state", "gb25_synchronize",
                 "gb25_mask_immersed_fields", "gb25_fill_halo_regions", "gb25_compute_auxiliaries",
                 "gb25_compute_tendencies", "gb25_compute_momentum_tendencies", "gb25_compute_tracer_tendencies",
                 "gb25_compute_boundary_tendencies", "gb25_correct_velocities_and_cache_previous_tendencies"):
        getattr(lib, name).argtypes = [H]
    lib.gb25_field_shape.argtypes = [H, C.c_int, C.POINTER(C.c_int)]
    lib.gb25_set_field.argtypes = [H, C.c_int, C.c_void_p]
    lib.gb25_get_field.argtypes = [H, C.c_int, C.c_void_p]
    lib.gb25_interior_shape.argtypes = [H, C.c_int, C.POINTER(C.c_int)]
    lib.gb25_set_interior.argtypes = [H, C.c_int, C.c_void_p]
    lib.gb25_get_interior.argtypes = [H, C.c_int, C.c_void_p]
    lib.gb25_set_fields.argtypes = [H, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_int]
    lib.gb25_get_fields.argtypes = [H, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_int]
    lib.gb25_set_flux_boundary_condition.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.gb25_set_clock.argtypes = [H, C.c_double, C.c_long, C.c_float]
    lib.gb25_get_clock.argtypes = [H, C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_float)]
    lib.gb25_first_time_step.argtypes = [H, C.c_float]
    lib.gb25_time_step.argtypes = [H, C.c_float]
    lib.gb25_loop.argtypes = [H, C.c_float, C.c_int]
    lib.gb25_ab2_step.argtypes = [H, C.c_float, C.c_float]
    lib.gb25_last_loop_seconds.argtypes = [H, C.POINTER(C.c_double)]
    lib.gb25_kernel_launch_count.argtypes = [H, C.POINTER(C.c_long)]
    lib.gb25_check_guards.argtypes = [H, C.POINTER(C.c_long)]
    lib.gb25_kernel_table_size.argtypes = []
    lib.gb25_kernel_table_size.restype = C.c_int
    lib.gb25_enable_stage_timers.argtypes = [H, C.c_int]
    lib.gb25_get_stage_times.argtypes = [H, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_long), C.c_int]
    lib.gb25_exchange_blob_size.restype = C.c_int
    lib.gb25_exchange_export.argtypes = [H, C.c_void_p]
    lib.gb25_exchange_connect.argtypes = [H, C.c_void_p, C.c_int]
    lib.gb25_exchange_connect_local.argtypes = [C.POINTER(H), C.c_int]
    lib.gb25_loop_all.argtypes = [C.POINTER(H), C.c_int, C.c_float, C.c_int]
    _libs[ft] = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Handle:
    """Owns one gb25_handle (one tile on one GPU)."""

    def __init__(self, grid, physics, dtau_frac, weights, device=-1, partition=(1, 1, 0, 0), float_type=np.float32):
        self.dtype = np.dtype(float_type)
        self.lib = load(self.dtype)
        self._keep = {}
        cfg = gb25_config()
        cfg.Nx, cfg.Ny, cfg.Nz, cfg.Hx, cfg.Hy, cfg.Hz = grid.Nx, grid.Ny, grid.Nz, grid.Hx, grid.Hy, grid.Hz
        cfg.topo_y = grid.topo_y
        cfg.immersed = 1 if grid.immersed else 0
        cfg.nsubsteps = len(weights)
        cfg.coriolis_scheme, cfg.fold_variant = physics.coriolis_scheme, physics.fold_variant
        cfg.south_inactive, cfg.cond_diff, cfg.eos_r0 = physics.south_inactive, physics.cond_diff, physics.eos_r0
        cfg.g, cfg.rho0, cfg.chi, cfg.dtau_frac, cfg.weno_eps = physics.g, physics.rho0, physics.chi, dtau_frac, physics.weno_eps
        cfg.Rx, cfg.Ry, cfg.rx, cfg.ry = partition
        cfg.device = device
        cfg.closure, cfg.kappa, cfg.nu = physics.closure, physics.kappa, physics.nu
        g = gb25_grid()
        arrays = dict(grid.metrics)
        arrays.update(grid.z)
        arrays["avg_weights"] = weights
        arrays["bottom_height"] = grid.bottom_height
        for name in _GRID_PTRS:
            a = arrays[name]
            if a is None:
                setattr(g, name, None)
                continue
            # grid products are Float32 values on the host (the reference model's grid type); the Float64 build receives
            # them promoted, exactly as the Float64 oracle does
            a = _f32(a) if name == "avg_weights" else np.ascontiguousarray(_f32(a).astype(self.dtype))
            self._keep[name] = a
            setattr(g, name, a.ctypes.data)
        h = C.c_void_p()
        rc = self.lib.gb25_create(C.byref(cfg), C.byref(g), C.byref(h))
        if rc != GB25_OK:
            raise Gb25Error(rc, (self.lib.gb25_last_error(None) or b"").decode())
        self.h = h
        self.cfg = cfg

    def close(self):
        if getattr(self, "h", None):
            self.lib.gb25_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc < 0:
            raise Gb25Error(rc, (self.lib.gb25_last_error(self.h) or b"").decode())
        return rc

    def call(self, name, *args):
        return self.check(getattr(self.lib, name)(self.h, *args))

    def field_shape(self, name):
        s = (C.c_int * 3)()
        self.check(self.lib.gb25_field_shape(self.h, FIELD_ID[name], s))
        return (s[2], s[1], s[0])          # NumPy (z, y, x) view of the Julia (x, y, z) parent

    def _arr(self, a):
        return np.ascontiguousarray(a, dtype=self.dtype)

    def set_field(self, name, parent):
        a = self._arr(parent)
        if a.shape != self.field_shape(name):
            raise ValueError(f"{name}: parent shape {a.shape} != {self.field_shape(name)}")
        self.check(self.lib.gb25_set_field(self.h, FIELD_ID[name], a.ctypes.data))

    def get_field(self, name, out=None):
        if out is None:
            out = np.empty(self.field_shape(name), dtype=self.dtype)
        self.check(self.lib.gb25_get_field(self.h, FIELD_ID[name], out.ctypes.data))
        return out

    def interior_shape(self, name):
        s = (C.c_int * 3)()
        self.check(self.lib.gb25_interior_shape(self.h, FIELD_ID[name], s))
        return (s[2], s[1], s[0])

    def set_interior(self, name, values):
        """set!(model, name=values): interior-shaped upload, halos untouched."""
        shp = self.interior_shape(name)
        a = self._arr(np.broadcast_to(np.asarray(values, dtype=self.dtype), shp))
        self.check(self.lib.gb25_set_interior(self.h, FIELD_ID[name], a.ctypes.data))

    def get_interior(self, name, out=None):
        if out is None:
            out = np.empty(self.interior_shape(name), dtype=self.dtype)
        self.check(self.lib.gb25_get_interior(self.h, FIELD_ID[name], out.ctypes.data))
        return out

    def _batch(self, fn, names, arrays, interior):
        n = len(names)
        ids = (C.c_int * n)(*[FIELD_ID[x] for x in names])
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
        for x, a in zip(names, arrays):
            want = self.interior_shape(x) if interior else self.field_shape(x)
            if a.shape != want or a.dtype != self.dtype or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"{x}: need a C-contiguous {self.dtype} array of shape {want}")
        self.check(fn(self.h, n, ids, ptrs, 1 if interior else 0))

    def set_fields(self, names, arrays, interior=False):
        """Batched upload: one cudaMemcpy3DAsync per field, one synchronisation."""
        self._batch(self.lib.gb25_set_fields, names, arrays, interior)

    def get_fields(self, names, arrays, interior=False):
        self._batch(self.lib.gb25_get_fields, names, arrays, interior)

    def set_flux_boundary_condition(self, name, side, values):
        """FluxBoundaryCondition(values) at the "bottom" / "top" of u, v, T or S; None = default no-flux."""
        sd = ("bottom", "top").index(side)
        if values is None:
            self.check(self.lib.gb25_set_flux_boundary_condition(self.h, FIELD_ID[name], sd, None))
            return
        two_d = {"u": "U", "v": "V", "T": "eta", "S": "eta"}[name]
        shp = self.field_shape(two_d)[1:]
        a = self._arr(values)
        if a.shape != shp:
            raise ValueError(f"flux of {name}: need a 2-D parent of shape {shp}, got {a.shape}")
        self.check(self.lib.gb25_set_flux_boundary_condition(self.h, FIELD_ID[name], sd, a.ctypes.data))

    def set_clock(self, time, iteration, last_dt):
        self.check(self.lib.gb25_set_clock(self.h, time, iteration, last_dt))

    def get_clock(self):
        t, it, dt = C.c_double(), C.c_long(), C.c_float()
        self.check(self.lib.gb25_get_clock(self.h, C.byref(t), C.byref(it), C.byref(dt)))
        return t.value, it.value, dt.value

    def last_loop_seconds(self):
        s = C.c_double()
        self.check(self.lib.gb25_last_loop_seconds(self.h, C.byref(s)))
        return s.value

    def check_guards(self):
        """Bytes of the guard zones around the device arrays that kernels overwrote (handle created under GB25_GUARD=1)."""
        n = C.c_long()
        self.check(self.lib.gb25_check_guards(self.h, C.byref(n)))
        return n.value

    def launch_count(self):
        n = C.c_long()
        self.check(self.lib.gb25_kernel_launch_count(self.h, C.byref(n)))
        return n.value

    def stage_times(self):
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        calls = (C.c_long * cap)()
        n = self.check(self.lib.gb25_get_stage_times(self.h, names, ms, calls, cap))
        return {names[i].decode(): (ms[i], calls[i]) for i in range(n)}
