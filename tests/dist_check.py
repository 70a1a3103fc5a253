"""Run under torchrun (one rank per GPU): the N-GPU partitioned run must equal the 1-GPU run of the same global
problem BIT FOR BIT (same kernels, same per-cell arithmetic; halos are copies).  Mirrors the protocol of
/root/reference/correctness/correctness_sharded_baroclinic_instability_simulation_run.jl (sharded model vs an
unsharded one).  Prints DIST_OK on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import gb25_b200  # noqa: F401
    from gb25_b200 import distributed as D, model as M, sharding
    grid_type = sys.argv[1] if len(sys.argv) > 1 else "gaussian_islands"
    tx, ty, Nz, nsteps = (int(v) for v in (sys.argv[2:6] if len(sys.argv) > 5 else (64, 48, 10, 5)))
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Rx, Ry = sharding.factors(world)
    gNx, gNy = tx * Rx, ty * Ry
    m = D.sharded_baroclinic_instability_model(M.B200(local), tx, ty, Nz, Δt=60.0, grid_type=grid_type, Rx=Rx, Ry=Ry,
                                               rank=rank, dist=dist)
    gg = m.global_grid
    rng = np.random.default_rng(42)
    from gb25_b200 import grids
    T, S = grids.baroclinic_instability_state(gg)
    ny_v = gNy + (1 if gg.topo_y == grids.TOPO_BOUNDED else 0)
    state = {"T": T.astype(np.float32), "S": S.astype(np.float32),
             "u": (1e-3 * rng.random((Nz, gNy, gNx))).astype(np.float32),
             "v": (1e-3 * rng.random((Nz, ny_v, gNx))).astype(np.float32)}
    for n, a in state.items():
        D.scatter_interior(m, n, a)
    D.barrier(m)
    M.first_time_step(m)
    for _ in range(nsteps):
        M.time_step(m)
    D.barrier(m)
    names = ("u", "v", "w", "T", "S", "eta", "Gn_u", "Gn_v", "Gn_T", "Gn_S", "Gm_u", "U", "V", "filt_U", "filt_eta")
    got = {n: D.gather_interior(m, n) for n in names}
    ok = True
    if rank == 0:
        ref = M.baroclinic_instability_model(M.B200(local), gNx, gNy, Nz, Δt=60.0, grid_type=grid_type)
        for n, a in state.items():
            ref.set_interior(n, a)
        M.first_time_step(ref)
        for _ in range(nsteps):
            M.time_step(ref)
        for n in names:
            r = ref.interior(n)
            g = got[n][:, :r.shape[1]]
            r = r[:, :g.shape[1]]
            same = np.array_equal(r.view(np.uint32), g.view(np.uint32))
            if not same:
                d = np.abs(r.astype(np.float64) - g.astype(np.float64))
                idx = np.unravel_index(np.nanargmax(d), d.shape)
                print(f"MISMATCH {n}: max|d|={np.nanmax(d):.3e} of max {np.abs(r).max():.3e} at k,j,i={idx} n={np.count_nonzero(r != g)}")
                ok = False
        print(("DIST_OK" if ok else "DIST_FAIL"), grid_type, f"{Rx}x{Ry} tiles of {tx}x{ty}x{Nz}, {nsteps + 1} steps,",
              m.handle.launch_count(), "launches on rank 0")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
