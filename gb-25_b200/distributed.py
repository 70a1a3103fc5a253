"""x/y domain partition across the GPUs of one box — host side.

Mirrors the sharded set-up of the reference: ``Distributed(ReactantState(); partition=Partition(Rx, Ry, 1))``
with ``(Rx, Ry) = factors(Ndev)`` (/root/reference/sharding/sharded_baroclinic_instability_simulation_run.jl:65-72,
/root/reference/src/sharding_utils.jl:39-62).  One process per GPU; rank = rx + Rx*ry.  Every rank builds the
grid products of the *global* grid on the host (they are 2-D and cheap), keeps its tile, creates its handle and
connects the peer-memory halo exchange (gb25_exchange_export / gb25_exchange_connect).  ``dist`` is an initialised
``torch.distributed`` (NCCL on the GPU box, gloo on CPU); it is used for bootstrap and barriers only — halo data
moves inside libgb25cuda over NVLink peer stores.

``host_fill_halo`` is a NumPy + send/recv restatement of the exchange protocol (same two phases, same fold
partners); it exists so that the partition logic is testable with world_size-2 gloo on a GPU-less box.
"""
from __future__ import annotations

import copy
import ctypes as C

import numpy as np

from . import grids as _grids
from . import model as M
from .sharding import factors, rank_coords


def tile_grid(gg: _grids.Grid, Rx, Ry, rx, ry) -> _grids.Grid:
    """Tile (rx, ry) of the global grid ``gg``, halos included (they overlap the neighbouring tiles)."""
    if gg.Nx % Rx or gg.Ny % Ry:
        raise ValueError(f"global size {gg.Nx}x{gg.Ny} is not divisible by the partition {Rx}x{Ry}")
    nx, ny = gg.Nx // Rx, gg.Ny // Ry
    js = slice(ry * ny, ry * ny + ny + 2 * gg.Hy + 1)
    is_ = slice(rx * nx, rx * nx + nx + 2 * gg.Hx)
    cut = lambda a: None if a is None else np.ascontiguousarray(a[js, is_])
    t = copy.copy(gg)
    t.Nx, t.Ny = nx, ny
    t.metrics = {k: cut(v) for k, v in gg.metrics.items()}
    t.bottom_height = cut(gg.bottom_height)
    t.lam_cc, t.phi_cc = cut(gg.lam_cc), cut(gg.phi_cc)
    t.wall_n = (gg.topo_y == _grids.TOPO_BOUNDED) and (ry == Ry - 1)
    return t


def connect(model, dist):
    """All-gather the IPC blobs of every rank and map the neighbours' halos."""
    h = model.handle
    n = h.lib.gb25_exchange_blob_size()
    buf = C.create_string_buffer(n)
    h.check(h.lib.gb25_exchange_export(h.h, buf))
    blobs = [None] * dist.get_world_size()
    dist.all_gather_object(blobs, buf.raw)
    allb = C.create_string_buffer(b"".join(blobs), n * len(blobs))
    h.check(h.lib.gb25_exchange_connect(h.h, allb, len(blobs)))
    dist.barrier()


def sharded_baroclinic_instability_model(arch, Nx, Ny, Nz, *, Δt, grid_type="simple_lat_lon", Rx=None, Ry=None,
                                         rank=None, dist=None, halo=(8, 8, 8), physics=None, global_size=False,
                                         float_type=np.float32):
    """The model of ``baroclinic_instability_model`` on tile ``rank`` of an (Rx, Ry) partition.  ``Nx, Ny`` are the
    PER-TILE interior sizes (the reference's scaling scripts also fix the tile, e.g. sharding/alps_scaling_test.jl:34)
    unless ``global_size`` is set."""
    world = dist.get_world_size()
    rank = dist.get_rank() if rank is None else rank
    if Rx is None:
        Rx, Ry = factors(world)
    if Rx * Ry != world:
        raise ValueError(f"partition {Rx}x{Ry} does not match world size {world}")
    rx, ry = rank_coords(rank, Rx, Ry)
    gNx, gNy = (Nx, Ny) if global_size else (Nx * Rx, Ny * Ry)
    gg = M.make_grid(gNx, gNy, Nz, halo, grid_type)
    tile = tile_grid(gg, Rx, Ry, rx, ry)
    model = M.HydrostaticFreeSurfaceModel(arch, tile, physics, partition=(Rx, Ry, rx, ry), float_type=float_type)
    model.partition = (Rx, Ry, rx, ry)
    model.global_grid = gg
    model.dist = dist
    if world > 1:
        connect(model, dist)
    model.clock.last_Δt = float(np.float32(Δt))
    return model


def scatter_interior(model, name, global_interior):
    """Every rank holds the same global interior array (Nz, gNy, gNx) and keeps its tile.  Follow with
    ``barrier(model)`` before stepping: uploads must not race with a neighbour's halo pushes."""
    Rx, Ry, rx, ry = model.partition
    g = model.grid
    sl = M._interior_slices(g, M.FIELD_LOC[name])
    nyl = sl[1].stop - sl[1].start
    a = np.asarray(global_interior)
    model.set_interior(name, a[:, ry * g.Ny: ry * g.Ny + nyl, rx * g.Nx:(rx + 1) * g.Nx])


def gather_interior(model, name):
    """Global interior on every rank (tiles' own rows/columns only: the wall row of a Bounded Face-y field
    is owned by the top tiles)."""
    Rx, Ry, rx, ry = model.partition
    g = model.grid
    a = model.interior(name)
    parts = [None] * model.dist.get_world_size()
    model.dist.all_gather_object(parts, a)
    rows = []
    for y in range(Ry):
        rows.append(np.concatenate([parts[x + Rx * y][:, :g.Ny + (1 if (y == Ry - 1 and parts[x + Rx * y].shape[1] > g.Ny) else 0)]
                                    for x in range(Rx)], axis=2))
    return np.concatenate([r[:, :g.Ny] if y < Ry - 1 else r for y, r in enumerate(rows)], axis=1)


def barrier(model):
    model.synchronize()
    model.dist.barrier()


# ------------------------------------------------------------------------------------------------
# One process driving every GPU of the node (the reference's `single_gpu_per_process=false`,
# /root/reference/sharding/sharded_baroclinic_instability_simulation_run.jl:49): a list of tile models, one per device,
# connected with gb25_exchange_connect_local; every stepping call is issued for all tiles before any is synchronised.
# ------------------------------------------------------------------------------------------------
class LocalPartition:
    def __init__(self, Nx, Ny, Nz, *, Δt, grid_type="simple_lat_lon", devices=(0, 1), Rx=None, Ry=None, halo=(8, 8, 8),
                 physics=None, global_size=False):
        n = len(devices)
        if Rx is None:
            Rx, Ry = factors(n)
        if Rx * Ry != n:
            raise ValueError(f"partition {Rx}x{Ry} does not match {n} devices")
        gNx, gNy = (Nx, Ny) if global_size else (Nx * Rx, Ny * Ry)
        self.global_grid = M.make_grid(gNx, gNy, Nz, halo, grid_type)
        self.Rx, self.Ry = Rx, Ry
        self.models = []
        for r, dev in enumerate(devices):
            rx, ry = rank_coords(r, Rx, Ry)
            m = M.HydrostaticFreeSurfaceModel(M.B200(dev), tile_grid(self.global_grid, Rx, Ry, rx, ry), physics,
                                              partition=(Rx, Ry, rx, ry))
            m.partition = (Rx, Ry, rx, ry)
            m.clock.last_Δt = float(np.float32(Δt))
            self.models.append(m)
        self.lib = self.models[0].handle.lib
        self._hs = (C.c_void_p * n)(*[m.handle.h for m in self.models])
        if n > 1:
            self.models[0].handle.check(self.lib.gb25_exchange_connect_local(self._hs, n))

    def _each(self, fn):
        for m in self.models:
            fn(m)

    def first_time_step(self): self._each(M.first_time_step)
    def time_step(self): self._each(M.time_step)
    def update_state(self): self._each(M.update_state)

    def loop(self, Ninner):
        m0 = self.models[0]
        for m in self.models:
            m._push_clock()
        m0.handle.check(self.lib.gb25_loop_all(self._hs, len(self.models), m0.clock.last_Δt, int(Ninner)))
        for m in self.models:
            m._pull_clock()

    def synchronize(self): self._each(lambda m: m.synchronize())

    def scatter_interior(self, name, a):
        a = np.asarray(a)
        for m in self.models:
            Rx, Ry, rx, ry = m.partition
            g = m.grid
            sl = M._interior_slices(g, M.FIELD_LOC[name])
            nyl = sl[1].stop - sl[1].start
            m.set_interior(name, np.ascontiguousarray(a[:, ry * g.Ny: ry * g.Ny + nyl, rx * g.Nx:(rx + 1) * g.Nx]))

    def gather_interior(self, name):
        g = self.models[0].grid
        rows = []
        for y in range(self.Ry):
            parts = [self.models[x + self.Rx * y].interior(name) for x in range(self.Rx)]
            row = np.concatenate(parts, axis=2)
            rows.append(row[:, :g.Ny] if y < self.Ry - 1 else row)
        return np.concatenate(rows, axis=1)

    def close(self):
        self.synchronize()
        self._each(lambda m: m.close())


def local_partition_check(devices=(0, 1), grid_type="gaussian_islands", tx=64, ty=48, Nz=10, nsteps=6, log=None):
    """partition_check for ONE process driving all devices (LocalPartition; the reference's single_gpu_per_process=false,
    /root/reference/sharding/sharded_baroclinic_instability_simulation_run.jl:49): the run over ``devices`` must equal the
    single-GPU run of the same global problem bit for bit.  Meant to be the first thing a fresh process does with the
    library (tests/test_multi_gpu.py runs it in a subprocess): one host thread enqueues the tiles' steps one after the other,
    so any host-side blocking behind a wait for a neighbour tile — a lazily loaded kernel, an allocation — is a deadlock."""
    P = LocalPartition(tx, ty, Nz, Δt=60.0, grid_type=grid_type, devices=tuple(devices))
    gg = P.global_grid
    rng = np.random.default_rng(42)
    T, S = _grids.baroclinic_instability_state(gg)
    ny_v = gg.Ny + (1 if gg.topo_y == _grids.TOPO_BOUNDED else 0)
    state = {"T": T.astype(np.float32), "S": S.astype(np.float32),
             "u": (1e-3 * rng.random((Nz, gg.Ny, gg.Nx))).astype(np.float32),
             "v": (1e-3 * rng.random((Nz, ny_v, gg.Nx))).astype(np.float32)}
    for n, a in state.items():
        P.scatter_interior(n, a)
    P.synchronize()
    P.first_time_step()
    P.time_step()
    P.loop(nsteps - 2)
    P.synchronize()
    ref = M.baroclinic_instability_model(M.B200(devices[0]), gg.Nx, gg.Ny, Nz, Δt=60.0, grid_type=grid_type)
    for n, a in state.items():
        ref.set_interior(n, a)
    M.first_time_step(ref)
    M.time_step(ref)
    M.loop(ref, nsteps - 2)
    ok = P.models[0].clock.iteration == ref.clock.iteration == nsteps
    for n in ("u", "v", "w", "T", "S", "eta", "Gn_u", "Gn_v", "Gn_T", "Gn_S", "Gm_u", "U", "V", "filt_U", "filt_eta"):
        r, g = ref.interior(n), P.gather_interior(n)
        g = g[:, :r.shape[1]]
        r = r[:, :g.shape[1]]
        if not np.array_equal(r.view(np.uint32), g.view(np.uint32)):
            ok = False
            if log:
                d = np.abs(r.astype(np.float64) - g.astype(np.float64))
                log(f"MISMATCH {n}: max|d|={np.nanmax(d):.3e} of max {np.abs(r).max():.3e} n={np.count_nonzero(r != g)}")
    P.close()
    ref.close()
    if log:
        log(f"LOCAL_OK grid={grid_type} devices={tuple(devices)} steps={nsteps}" if ok else "LOCAL_FAIL")
    return ok


def partition_check(dist, local_rank, grid_type="gaussian_islands", tx=64, ty=48, Nz=10, nsteps=5, log=None):
    """The reference's sharded correctness protocol
    (/root/reference/correctness/correctness_sharded_baroclinic_instability_simulation_run.jl: a sharded model against
    an unsharded one) with a stricter criterion: the (Rx, Ry)-partitioned run over all ranks of ``dist`` must equal the
    single-GPU run of the same global problem BIT FOR BIT (same kernels, same per-cell arithmetic; halos are copies).
    Collective: every rank calls it; returns the same bool on every rank.  Rank 0 runs the unpartitioned model."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    Rx, Ry = factors(world)
    gNx, gNy = tx * Rx, ty * Ry
    m = sharded_baroclinic_instability_model(M.B200(local_rank), tx, ty, Nz, Δt=60.0, grid_type=grid_type, Rx=Rx, Ry=Ry,
                                             rank=rank, dist=dist)
    gg = m.global_grid
    rng = np.random.default_rng(42)
    T, S = _grids.baroclinic_instability_state(gg)
    ny_v = gNy + (1 if gg.topo_y == _grids.TOPO_BOUNDED else 0)
    state = {"T": T.astype(np.float32), "S": S.astype(np.float32),
             "u": (1e-3 * rng.random((Nz, gNy, gNx))).astype(np.float32),
             "v": (1e-3 * rng.random((Nz, ny_v, gNx))).astype(np.float32)}
    for n, a in state.items():
        scatter_interior(m, n, a)
    barrier(m)
    M.first_time_step(m)
    M.time_step(m)
    M.loop(m, nsteps - 1)        # both entry points: gb25_time_step and gb25_loop
    barrier(m)
    names = ("u", "v", "w", "T", "S", "eta", "Gn_u", "Gn_v", "Gn_T", "Gn_S", "Gm_u", "U", "V", "filt_U", "filt_eta")
    got = {n: gather_interior(m, n) for n in names}
    ok = True
    if rank == 0:
        ref = M.baroclinic_instability_model(M.B200(local_rank), gNx, gNy, Nz, Δt=60.0, grid_type=grid_type)
        for n, a in state.items():
            ref.set_interior(n, a)
        M.first_time_step(ref)
        M.time_step(ref)
        M.loop(ref, nsteps - 1)
        for n in names:
            r = ref.interior(n)
            g = got[n][:, :r.shape[1]]
            r = r[:, :g.shape[1]]
            if not np.array_equal(r.view(np.uint32), g.view(np.uint32)):
                d = np.abs(r.astype(np.float64) - g.astype(np.float64))
                idx = np.unravel_index(np.nanargmax(d), d.shape)
                if log:
                    log(f"MISMATCH {n}: max|d|={np.nanmax(d):.3e} of max {np.abs(r).max():.3e} at k,j,i={idx} "
                        f"n={np.count_nonzero(r != g)}")
                ok = False
        if log:
            log(("DIST_OK" if ok else "DIST_FAIL") + f" {grid_type} {Rx}x{Ry} tiles of {tx}x{ty}x{Nz}, {nsteps + 1} steps, "
                f"{m.handle.launch_count()} launches on rank 0")
        ref.close()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    m.close()
    dist.barrier()
    return bool(int(flag.item()))


# ------------------------------------------------------------------------------------------------
# Host restatement of the exchange protocol (tests only; the product moves halos inside libgb25cuda)
# ------------------------------------------------------------------------------------------------
TAG_W, TAG_E, TAG_S, TAG_N, TAG_F, TAG_F2 = range(6)   # = the ExSlot numbers of gb25_exchange.cu


def host_fill_halo(dist, a, loc, sign, topo_y, Rx, Ry, rx, ry, Hx=8, Hy=8, Hz=8, wall_n=None):
    """Fill the halos of the tile-parent ``a`` (nz, ny, nx; 3-D or a single plane) in place, exchanging strips
    with the neighbouring ranks through ``dist.send/recv``.  Same order as launch_fill_halo_dist:
    wall conditions + z halos (local), phase Y (north/south strips over all planes, fold rows with the mirrored
    partner), phase X (west/east strips over the full parent extent)."""
    import torch
    lx, ly, lz = loc
    three_d = a.shape[0] > 1
    top, bottom = ry == Ry - 1, ry == 0
    wall_n = (topo_y == _grids.TOPO_BOUNDED and top) if wall_n is None else wall_n
    Nx = a.shape[2] - 2 * Hx
    Ny = a.shape[1] - 2 * Hy - (1 if (ly and wall_n) else 0)
    Nz = a.shape[0] - 2 * Hz - (1 if lz else 0) if three_d else 1
    kin = slice(Hz, Hz + Nz + (1 if lz else 0)) if three_d else slice(0, 1)
    xin = slice(Hx, Hx + Nx)
    rk = lambda x, y: (x % Rx) + Rx * y
    J = lambda j: j + Hy - 1

    def xchg(pairs):
        """pairs: list of (peer_rank, send_array, recv_shape, send_tag, recv_tag) -> list of received arrays.
        Tags name the halo the strip lands in (the flag slot of the CUDA implementation): with Rx == 2 the east
        and the west neighbour are the same rank, and only the tag tells the two strips apart."""
        reqs, outs = [], []
        for peer, snd, shp, stag, rtag in pairs:
            if snd is not None:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(snd)), peer, tag=stag))
        for peer, snd, shp, stag, rtag in pairs:
            if shp is not None:
                t = torch.empty(shp, dtype=torch.from_numpy(a[:1, :1, :1]).dtype)
                dist.recv(t, peer, tag=rtag)
                outs.append(t.numpy())
            else:
                outs.append(None)
        for r in reqs:
            r.wait()
        return outs

    # ---- local wall conditions and z halos
    if bottom:
        if ly == 0:
            for m in range(1, Hy + 1):
                a[kin, J(1 - m), xin] = a[kin, J(m), xin]
        else:
            a[kin, J(1), xin] = 0
    if wall_n:
        if ly == 0:
            for m in range(1, Hy + 1):
                a[kin, J(Ny + m), xin] = a[kin, J(Ny + 1 - m), xin]
        else:
            a[kin, J(Ny + 1), xin] = 0
    fold = topo_y == _grids.TOPO_FOLD and top
    if fold and Rx == 1:
        isrc, quirk, jsrc = _grids.fold_index_maps(Nx, Ny, Hx, Hy, lx, ly)
        sg = np.where(quirk, abs(sign), sign)
        for m in range(1, Hy + 1):
            a[kin, J(Ny + m), xin] = sg * a[kin, J(jsrc[m - 1]), isrc + Hx - 1]
    if three_d:
        jt = Ny + (1 if (ly and wall_n) else 0)
        ys = slice(Hy, Hy + jt)
        if lz == 0:
            for m in range(1, Hz + 1):
                a[Hz - m, ys, xin] = a[Hz + m - 1, ys, xin]
                a[Hz + Nz - 1 + m, ys, xin] = a[Hz + Nz - m, ys, xin]
        else:
            a[Hz, ys, xin] = 0
            a[Hz + Nz, ys, xin] = 0
    # ---- phase Y
    pairs = []
    if not top:
        pairs.append((rk(rx, ry + 1), a[:, J(Ny - Hy + 1):J(Ny) + 1, xin], (a.shape[0], Hy, Nx), TAG_S, TAG_N))
    if not bottom:
        pairs.append((rk(rx, ry - 1), a[:, J(1):J(Hy) + 1, xin], (a.shape[0], Hy, Nx), TAG_N, TAG_S))
    got = xchg(pairs)
    q = 0
    if not top:
        a[:, J(Ny + 1):J(Ny + Hy) + 1, xin] = got[q]; q += 1
    if not bottom:
        a[:, J(1 - Hy):J(0) + 1, xin] = got[q]; q += 1
    if fold and Rx > 1:
        nk = kin.stop - kin.start
        jsrc = [Ny - m if ly == 0 else Ny - m + 1 for m in range(1, Hy + 1)]
        rows = a[kin][:, [J(j) for j in jsrc], :][:, :, xin]          # (nk, Hy, Nx): my rows feeding the partner's halo
        main, second = rk(Rx - 1 - rx, ry), rk(Rx - rx, ry)
        if lx == 0:
            snd_main = sign * rows[:, :, ::-1]                        # dest col id = Nx - is + 1
            pairs = [(main, snd_main, (nk, Hy, Nx), TAG_F, TAG_F)]
            if main == rk(rx, ry):
                got = [snd_main]
            else:
                got = xchg(pairs)
            a[kin, J(Ny + 1):J(Ny + Hy) + 1, xin] = got[0]
        else:
            snd_main = sign * rows[:, :, :0:-1]                       # sources is = Nx..2 -> dest id = 2..Nx
            snd_second = (abs(sign) if rx == 0 else sign) * rows[:, :, :1]   # source is = 1 -> dest id = 1 on the second partner
            me = rk(rx, ry)
            got_main = snd_main if main == me else xchg([(main, snd_main, (nk, Hy, Nx - 1), TAG_F, TAG_F)])[0]
            got_second = snd_second if second == me else xchg([(second, snd_second, (nk, Hy, 1), TAG_F2, TAG_F2)])[0]
            a[kin, J(Ny + 1):J(Ny + Hy) + 1, Hx + 1:Hx + Nx] = got_main
            a[kin, J(Ny + 1):J(Ny + Hy) + 1, Hx:Hx + 1] = got_second
    # ---- phase X over the full parent extent
    if Rx == 1:
        a[:, :, :Hx] = a[:, :, Nx:Nx + Hx]
        a[:, :, Nx + Hx:] = a[:, :, Hx:2 * Hx]
    else:
        east, west = rk(rx + 1, ry), rk(rx - 1, ry)
        shp = (a.shape[0], a.shape[1], Hx)
        got = xchg([(east, a[:, :, Nx:Nx + Hx], shp, TAG_W, TAG_E), (west, a[:, :, Hx:2 * Hx], shp, TAG_E, TAG_W)])
        a[:, :, Nx + Hx:] = got[0]      # from the east tile: its first Hx interior columns
        a[:, :, :Hx] = got[1]           # from the west tile: its last Hx interior columns
    return a
