"""Run under torchrun (one rank per GPU): the N-GPU partitioned run must equal the 1-GPU run of the same global
problem BIT FOR BIT (same kernels, same per-cell arithmetic; halos are copies).  Mirrors the protocol of
/root/reference/correctness/correctness_sharded_baroclinic_instability_simulation_run.jl (sharded model vs an
unsharded one); the comparison itself is gb25_b200.distributed.partition_check, which bench.py --gpus N also runs
before its timed region.  Prints DIST_OK on success."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import gb25_b200  # noqa: F401
    from gb25_b200 import distributed as D
    grid_type = sys.argv[1] if len(sys.argv) > 1 else "gaussian_islands"
    tx, ty, Nz, nsteps = (int(v) for v in (sys.argv[2:6] if len(sys.argv) > 5 else (64, 48, 10, 5)))
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = D.partition_check(dist, local, grid_type, tx, ty, Nz, nsteps, log=print)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
