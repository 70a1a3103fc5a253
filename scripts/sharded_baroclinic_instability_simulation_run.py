#!/usr/bin/env python
"""The reference's sharded run protocol on libgb25cuda (one process per GPU, launched by torchrun):
/root/reference/sharding/sharded_baroclinic_instability_simulation_run.jl.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        scripts/sharded_baroclinic_instability_simulation_run.py --grid-x 1536 --grid-y 768 --grid-z 50 --float-type Float32

Same argument meaning as the reference: --grid-x/--grid-y are the PER-DEVICE tile sizes including the 2H halo cells, the
global interior is Nx = grid-x * Rx - 2H (…_run.jl:82-88), (Rx, Ry) = factors(Ndev) (src/sharding_utils.jl:39-62), dt = 1,
one first_time_step! and one loop!(model, 256), each timed per rank like the reference's `@time "[rank] loop"`.
The initial state is the reference's (T = S = u = v = 0: set_baroclinic_instability! is commented out in
src/baroclinic_instability_model.jl:74-80) unless --baroclinic-state is given."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import gb25_b200  # noqa: F401
    from gb25_b200 import arg_parsing as A, distributed as D, model as M, sharding

    argv = [a for a in sys.argv[1:] if a != "--baroclinic-state"]
    parsed = A.parse_baroclinic_instability_args(grid_x_default=1536, grid_y_default=768, grid_z_default=4, argv=argv)
    FT = A.supported_float_type(parsed)          # Float32: all kernel generations; Float64: libgb25cuda_f64.so
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Rx, Ry = sharding.factors(world)
    H = 8
    Nx, Ny = sharding.global_size_from_tile(parsed["grid-x"], parsed["grid-y"], Rx, Ry, H)
    Nz = parsed["grid-z"]
    if Nx % Rx or Ny % Ry:
        raise SystemExit(f"global interior {Nx}x{Ny} does not split into {Rx}x{Ry} equal tiles")
    print(f"[{rank}] Generating model (Nx={Nx}, Ny={Ny})...", file=sys.stderr)
    if world > 1:
        model = D.sharded_baroclinic_instability_model(M.B200(local), Nx // Rx, Ny // Ry, Nz, Δt=1.0, Rx=Rx, Ry=Ry,
                                                       rank=rank, dist=dist, halo=(H, H, H), float_type=FT)
    else:
        model = M.baroclinic_instability_model(M.B200(local), Nx, Ny, Nz, Δt=1.0, halo=(H, H, H), float_type=FT)
    if "--baroclinic-state" in sys.argv:
        M.set_baroclinic_instability(model)
        rng = np.random.default_rng(42 + rank)
        M.set(model, u=1e-3 * rng.random(model.interior("u").shape, dtype=np.float32),
              v=1e-3 * rng.random(model.interior("v").shape, dtype=np.float32))
    if dist is not None:
        D.barrier(model)
    Ninner = 256
    t0 = time.perf_counter()
    M.first_time_step(model)
    model.synchronize()
    print(f"[{rank}] first time step: {time.perf_counter() - t0:.6f} seconds")
    t0 = time.perf_counter()
    M.loop(model, Ninner)
    model.synchronize()
    el = time.perf_counter() - t0
    cells = (Nx // Rx) * (Ny // Ry) * Nz
    print(f"[{rank}] loop: {el:.6f} seconds ({Ninner} steps, {cells * Ninner / el:.4g} cell-steps/s on this tile, "
          f"device time {model.handle.last_loop_seconds():.6f} s)")
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
