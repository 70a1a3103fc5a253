"""Host-side partition logic with world_size > 1 on CPU (gloo): tile geometry, neighbour / fold-partner
tables and the two-phase exchange protocol, checked bit-exactly against the single-domain halo fill of the
oracle.  The CUDA implementation of the same protocol (gb25_exchange.cu) is checked on GPUs by
tests/test_multi_gpu.py (N-GPU run == 1-GPU run, bit-exact)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, grid_type, Rx, Ry, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gb25_b200  # noqa: F401
        from gb25_b200 import distributed as D, grids, model as M
        from gb25_b200.lib import FIELD_LOC
        from gb25_b200.sharding import rank_coords
        from oracle import oracle as O
        gNx, gNy, Nz = 32 * Rx, 24 * Ry, 5
        gm = M.baroclinic_instability_model(O.CPUOracle(np.float32), gNx, gNy, Nz, Δt=1.0, grid_type=grid_type,
                                            model_cls=O.OracleModel)
        gg = gm.grid
        rx, ry = rank_coords(rank, Rx, Ry)
        tile = D.tile_grid(gg, Rx, Ry, rx, ry)
        # tile metrics are windows of the global arrays: my halo is my neighbour's interior
        assert tile.Nx == 32 and tile.Ny == 24
        assert np.array_equal(tile.metrics["dx_fc"][8:8 + 24, :8], gg.metrics["dx_fc"][ry * 24 + 8: ry * 24 + 32, rx * 32: rx * 32 + 8])
        rng = np.random.default_rng(5)
        bad = []
        for name in ("u", "v", "T", "eta", "U", "V"):
            p = gm.parent(name)
            p[...] = rng.standard_normal(p.shape).astype(np.float32)
            lx, ly, lz, three_d = FIELD_LOC[name]
            sl = M._interior_slices(gg, FIELD_LOC[name])
            interior = p[sl].copy()
            p[...] = 0
            p[sl] = interior
            gm.set_parent(name, p)
            sign = -1.0 if name in ("u", "v", "U", "V") else 1.0
            gm.fill_halo(name, sign)
            ref = gm.parent(name)
            # my tile: interior from the global interior, halos zero, then the distributed protocol
            tz, ty, tx = [tile.field_shape((lx, ly, lz))[k] for k in (2, 1, 0)]
            if not three_d:
                tz = 1
            a = np.zeros((tz, ty, tx), dtype=np.float32)
            tsl = M._interior_slices(tile, FIELD_LOC[name])
            nyl = tsl[1].stop - tsl[1].start
            a[tsl] = interior[:, ry * 24: ry * 24 + nyl, rx * 32:(rx + 1) * 32]
            D.host_fill_halo(dist, a, (lx, ly, lz), sign, gg.topo_y, Rx, Ry, rx, ry, wall_n=tile.owns_north_wall)
            want = ref[:, ry * 24: ry * 24 + ty, rx * 32: rx * 32 + tx]
            if not np.array_equal(a.view(np.uint32), want.view(np.uint32)):
                d = np.argwhere(a != want)
                bad.append((name, len(d), d[:3].tolist()))
        q.put((rank, bad))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid_type,Rx,Ry", [("simple_lat_lon", 2, 1), ("gaussian_islands", 2, 1),
                                              ("simple_lat_lon", 2, 2), ("gaussian_islands", 2, 2),
                                              ("gaussian_islands", 4, 1)])
def test_partitioned_halo_fill_equals_single_domain(oracle_mod, grid_type, Rx, Ry):
    world = Rx * Ry
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid_type, Rx, Ry, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, bad in res:
        assert not bad, f"rank {rank}: {bad}"


def test_tile_grid_rejects_uneven_partitions():
    from gb25_b200 import distributed as D, grids
    g = grids.simple_latitude_longitude_grid(30, 16, 4)
    with pytest.raises(ValueError):
        D.tile_grid(g, 4, 1, 0, 0)
