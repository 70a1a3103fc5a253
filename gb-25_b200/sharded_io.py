"""Per-rank state dumps in the shape of the reference's ``src/sharded_io.jl`` — plus the resume path it lacks.

Reference: every rank serialises its addressable shards to ``<dir>/<label>/fields_rank{R}.dat`` with, per field,
``(local_arrays, local_slices, global_shape)`` and the metadata ``iteration, time, field_names``
(/root/reference/src/sharded_io.jl:70-96, 122-138); ``load_global_field`` / ``load_all_fields`` reassemble global
arrays offline (:146-213).  Here the container is ``.npz`` (Julia ``Serialization`` is not portable), the tile's
parent array is stored whole (halos included, so a reload is bit-exact), ``local_slices`` are the 1-based inclusive
global index ranges of the tile's interior, and ``load_model_state`` restores a model so that stepping continues
bit-identically (the reference has no code that loads a dump back into a model)."""
from __future__ import annotations

import glob
import os

import numpy as np

from .lib import FIELD_LOC
from .model import _interior_slices

# everything needed for a bit-exact restart: prognostic + diagnostic fields, both tendency sets, barotropic state
RESTART_FIELDS = ("u", "v", "w", "T", "S", "p", "eta", "U", "V", "filt_eta", "filt_U", "filt_V",
                  "Gn_u", "Gn_v", "Gn_T", "Gn_S", "Gm_u", "Gm_v", "Gm_T", "Gm_S", "Gn_U", "Gn_V", "Gm_U", "Gm_V")
OUTPUT_FIELDS = ("u", "v", "w", "T", "S", "eta")      # extract_model_fields of the reference: fields(model)


def _partition(model):
    return getattr(model, "partition", (1, 1, 0, 0))


def _tile_slices(model, name):
    """1-based inclusive global ranges (z, y, x) of this tile's interior for field ``name``."""
    Rx, Ry, rx, ry = _partition(model)
    g = model.grid
    sl = _interior_slices(g, FIELD_LOC[name])
    nz, ny, nx = (s.stop - s.start for s in sl)
    return ((1, nz), (ry * g.Ny + 1, ry * g.Ny + ny), (rx * g.Nx + 1, rx * g.Nx + nx))


def save_model_state(dir, model, label="checkpoint", fields=RESTART_FIELDS):
    """save_model_state(dir, model, arch; label): one file per rank, no cross-rank gather."""
    Rx, Ry, rx, ry = _partition(model)
    rank = rx + Rx * ry
    outdir = os.path.join(dir, label)
    os.makedirs(outdir, exist_ok=True)
    model.synchronize()
    payload = {"iteration": np.int64(model.clock.iteration), "time": np.float64(model.clock.time),
               "last_dt": np.float64(model.clock.last_Δt), "field_names": np.array(fields),
               "partition": np.array([Rx, Ry, rx, ry])}
    for n in fields:
        payload[f"{n}__local_array"] = model.parent(n)
        payload[f"{n}__local_slices"] = np.array(_tile_slices(model, n))
        g = model.grid
        sl = _interior_slices(g, FIELD_LOC[n])
        nz = sl[0].stop - sl[0].start
        payload[f"{n}__global_shape"] = np.array([nz, Ry * g.Ny + (sl[1].stop - sl[1].start - g.Ny), Rx * g.Nx])
    path = os.path.join(outdir, f"fields_rank{rank}.npz")
    np.savez(path, **payload)
    return path


def load_global_field(dir, field_name, label="checkpoint", ranks=None):
    """Reassemble the global interior of one field from the per-rank files (offline, any process)."""
    files = sorted(glob.glob(os.path.join(dir, label, "fields_rank*.npz")))
    if not files:
        raise FileNotFoundError(f"no fields_rank*.npz under {os.path.join(dir, label)}")
    out = None
    for f in files:
        d = np.load(f)
        if ranks is not None and int(os.path.basename(f)[11:-4]) not in ranks:
            continue
        shape = tuple(int(v) for v in d[f"{field_name}__global_shape"])
        if out is None:
            out = np.full(shape, np.nan, dtype=d[f"{field_name}__local_array"].dtype)
        (z0, z1), (y0, y1), (x0, x1) = d[f"{field_name}__local_slices"]
        p = d[f"{field_name}__local_array"]
        hz = (p.shape[0] - (z1 - z0 + 1)) // 2 if p.shape[0] > 1 else 0
        hy, hx = 8, 8
        hx = (p.shape[2] - (x1 - x0 + 1)) // 2
        hy = (p.shape[1] - (y1 - y0 + 1)) // 2
        out[z0 - 1:z1, y0 - 1:y1, x0 - 1:x1] = p[hz:hz + (z1 - z0 + 1), hy:hy + (y1 - y0 + 1), hx:hx + (x1 - x0 + 1)]
    return out


def load_all_fields(dir, label="checkpoint", fields=OUTPUT_FIELDS):
    meta = np.load(sorted(glob.glob(os.path.join(dir, label, "fields_rank*.npz")))[0])
    out = {n: load_global_field(dir, n, label) for n in fields if f"{n}__local_array" in meta}
    out["iteration"], out["time"] = int(meta["iteration"]), float(meta["time"])
    return out


def load_model_state(dir, model, label="checkpoint"):
    """Resume: restore this rank's tile (all RESTART_FIELDS that were saved) and the clock."""
    Rx, Ry, rx, ry = _partition(model)
    path = os.path.join(dir, label, f"fields_rank{rx + Rx * ry}.npz")
    d = np.load(path)
    if tuple(int(v) for v in d["partition"]) != (Rx, Ry, rx, ry):
        raise ValueError(f"{path} was written for partition {tuple(d['partition'])}, model has {(Rx, Ry, rx, ry)}")
    for n in d["field_names"]:
        model.set_parent(str(n), d[f"{n}__local_array"])
    model.clock.iteration, model.clock.time, model.clock.last_Δt = int(d["iteration"]), float(d["time"]), float(d["last_dt"])
    return model
