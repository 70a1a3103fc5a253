#!/usr/bin/env python
"""bench.py — headline benchmark of libgb25cuda: Float32 cell-steps/s of the GB-25 baroclinic-instability
HydrostaticFreeSurfaceModel time step (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm: the oracle on the host cores)

A "step" is one GordonBell25.time_step! (one AB2 step incl. 21 barotropic substeps and update_state!).
N = 1 workload: BASELINE.json configs[1] — TripolarGrid 1/4 degree (1440x600x50), gaussian-islands
bathymetry, Float32.  N > 1: the same 1440x600x50 tile per GPU (weak scaling), domain partitioned in x/y
with (Rx, Ry) = factors(N) (/root/reference/src/sharding_utils.jl:39-62).

Prints ONE JSON line (rank 0).  `value` = whole-job cell-steps/s with the state resident in HBM;
`e2e` = the same metric when every step round-trips the prognostic state through host buffers over the
C ABI (gb25_set_field / gb25_time_step / gb25_get_field); `roofline` = the dominant kernel against the
measured HBM peak; `cpu_baseline` = the CPU oracle on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (grid_type, Nx, Ny, Nz, dt)
    "tripolar_quarter_degree": ("gaussian_islands", 1440, 600, 50, 60.0),
    "latlon_128x64x8": ("simple_lat_lon", 128, 64, 8, 60.0),
    "tripolar_flat": ("tripolar", 1440, 600, 50, 60.0),              # profiling aid: no bathymetry => fast path only
    "latlon_1440x600x50": ("simple_lat_lon", 1440, 600, 50, 60.0),
    # BASELINE.json configs[4]: the per-GPU tile of the 1/8-degree weak-scaling sweep (3.5e8 cells, ~30 GB of HBM)
    "tripolar_eighth_degree_tile": ("gaussian_islands", 2880, 1200, 100, 30.0),
}
ALGORITHMIC_BYTES_PER_CELL_STEP = 104     # 26 Float32 words of compulsory 3-D traffic
# algorithmic bytes per cell of one launch of each tendency kernel (DESIGN.md "kernels")
# ("kernel:<name>" entries are single-launch CUDA-event timers of libgb25cuda)
# With the AB2 epilogue (the path gb25_loop takes) a tendency kernel also reads its G- and writes the updated state.
KERNEL_BYTES_PER_CELL = {"kernel:k_tracer_tendency_v2": (5 + 2) * 4,   # one launch, both tracers: read u,v,w,T,S, write GT,GS
                         "kernel:k_tracer_tma": (5 + 2) * 4,
                         "kernel:k_tracer_tma+ab2": (7 + 4) * 4,        # + read G-T, G-S, write T', S'
                         "kernel:k_gu_tma": (4 + 1) * 4,                # read u,v,w,p, write Gu
                         "kernel:k_gv_tma": (4 + 1) * 4,
                         "kernel:k_gu_tma+ab2": (5 + 2) * 4,            # + read G-u, write u*
                         "kernel:k_gv_tma+ab2": (5 + 2) * 4,
                         "kernel:k_ab2_fused": 16 * 4,                  # read 4 fields + 8 G, write 4 fields
                         "momentum_tendencies": 2 * (4 + 1) * 4, "tracer_tendencies": (5 + 2) * 4}
# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) comes from profiles/ncu_traffic.json, written by
# scripts/ncu_table.py from an `ncu --set full` capture; it is reported only when the capture was taken with the kernel
# sources that are being timed (sha256 over gb-25_b200/csrc), otherwise `traffic` is null.
def kernel_source_sha():
    import hashlib
    d = os.path.join(ROOT, "gb-25_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def kernel_sass_sha(lib_path=None):
    """sha of the DEVICE code of the built library (cuobjdump -sass, without the lines that carry source paths): what an ncu
    capture is really a capture of.  A change of host code in a .cu file changes kernel_source_sha but not this."""
    import hashlib
    import shutil
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    lib_path = lib_path or os.path.join(ROOT, "gb-25_b200", "csrc", "libgb25cuda.so")
    if not (os.path.exists(exe) and os.path.exists(lib_path)):
        return None
    try:
        out = subprocess.run([exe, "-sass", lib_path], capture_output=True, text=True, check=True, timeout=300).stdout
    except Exception:
        return None
    h = hashlib.sha256()
    for ln in out.splitlines():
        if ln.startswith("identifier") or "/" in ln and ".cu" in ln and "/*" not in ln:
            continue
        h.update(ln.encode()); h.update(b"\n")
    return h.hexdigest()[:16]


def ncu_traffic():
    """Per-kernel DRAM traffic of the last ncu --set full capture, if it was taken on this code: same kernel sources, or —
    when only host code in those files changed since — the same machine code."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return {}, "no capture"
    d = json.load(open(p))
    src = d.get("capture", "profiles/ncu_traffic.json")
    if d.get("kernel_source_sha") == kernel_source_sha():
        return d.get("kernels", {}), src
    if d.get("kernel_sass_sha") and d.get("kernel_sass_sha") == kernel_sass_sha():
        return d.get("kernels", {}), src + f" (device code unchanged since: sass sha {d['kernel_sass_sha']})"
    return {}, f"stale capture ({d.get('kernel_source_sha')}): kernel sources changed since"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self, t0=None, t1=None):
        """Summarise the samples whose timestamp falls inside [t0, t1] (the timed region; wall-clock seconds)."""
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                if t0 is not None:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if ts < t0 - 0.05 or ts > t1 + 0.05:
                        continue
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_state(model, seed=42):
    """Synthetic baroclinic-instability state: T, S of set_baroclinic_instability_kernel!
    (/root/reference/src/model_utils.jl:99-110) and u, v = 1e-3*U[0,1)
    (/root/reference/correctness/correctness_baroclinic_instability_simulation_run.jl:40-42)."""
    from gb25_b200 import model as M
    M.set_baroclinic_instability(model)
    rng = np.random.default_rng(seed)
    shp = model.interior("u").shape
    M.set(model, u=(1e-3 * rng.random(shp, dtype=np.float32)), v=(1e-3 * rng.random(model.interior("v").shape, dtype=np.float32)))


def _oracle_model(grid_type, nx, ny, Nz, dt):
    from gb25_b200 import model as M
    from oracle import oracle as O
    from gb25_b200.config import PhysicsConfig
    # sum-of-squares smoothness indicators: the expanded form of the recalled reference NaNs in Float32 at these sizes
    # (DESIGN.md deviation D1); same flop count to within a few per cent
    return M.baroclinic_instability_model(O.CPUOracle(np.float32), nx, ny, Nz, Δt=dt, grid_type=grid_type,
                                          model_cls=O.OracleModel, physics=PhysicsConfig(oracle_beta_form=1))


def _scaled_dims(workload, shrink):
    grid_type, Nx, Ny, Nz, dt = WORKLOADS[workload]
    return grid_type, Nx // shrink, Ny // shrink, Nz, dt * shrink


def run_reference_arm(args, workload):
    """--impl reference: the reference's CPU implementation of the path.  Oceananigans CPU() cannot run here (no
    Julia in the image), so this is the CPU oracle port (`kind: "port"`) with all host threads.  At N = 1
    it runs the SAME configuration as our arm (full size) when K + W steps fit a few minutes; otherwise, and for
    N > 1 (rank 0 alone works), a bounded sample: a reduced tile of the same grid family whose true size and dt are
    what `config` reports, with `same_config: false`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gb25_b200 import model as M
    from oracle import oracle as O
    O.build()
    cores = O.set_num_threads()      # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    grid_type, Nx, Ny, Nz, dt = WORKLOADS[workload]
    budget_s = float(os.environ.get("GB25_REFERENCE_BUDGET_S", "240"))
    nsteps_total = args.steps + max(args.warmup, 1) + 1          # first_time_step costs about two steps
    shrink = 1
    # measured on the round-1 box (16 cores): 6.5e6 cell-steps/s; pick the largest tile of the family that fits the budget
    est_rate = 4.0e5 * cores
    while Nx // shrink >= 64 and (Nx // shrink) * (Ny // shrink) * Nz * nsteps_total / est_rate > budget_s and shrink < 8:
        shrink *= 2
    if Nx * Ny * Nz < 1e6:
        shrink = 1
    grid_type, nx, ny, Nz, dts = _scaled_dims(workload, shrink)
    m = _oracle_model(grid_type, nx, ny, Nz, dts)
    synthetic_state(m)
    M.first_time_step(m)
    for _ in range(max(args.warmup - 1, 0)):
        M.time_step(m)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        M.time_step(m)
    el = time.perf_counter() - t0
    finite = bool(np.isfinite(m.interior("eta")).all())
    cells = nx * ny * Nz
    val = cells * args.steps / el
    same = shrink == 1 and args.gpus == 1
    sample = (f"{grid_type} {nx}x{ny}x{Nz}, dt = {dts:g} s, {args.steps} timed steps" +
              ("" if shrink == 1 else f" (1/{shrink * shrink} of the {Nx}x{Ny}x{Nz} per-GPU tile)") +
              f"; CPU oracle port (C++/OpenMP, {cores} threads), not Oceananigans CPU(): no Julia in the image")
    line = {"impl": "reference", "metric": "cell_steps_per_s", "value": val, "unit": "cell-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "grid": grid_type, "Nx": nx, "Ny": ny, "Nz": Nz, "dt": dts,
                       "same_config": same, "kind": "port", "state_finite": finite},
            "cpu_baseline": {"value": val, "unit": "cell-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    if not finite:
        raise SystemExit("reference arm: state is not finite")


def cpu_baseline_sample(workload, budget_s=20.0):
    """The CPU oracle on the box's host cores, on the bench's own configuration (full size), for a bounded number of
    steps (about `budget_s` of stepping after the first step)."""
    from gb25_b200 import model as M
    from oracle import oracle as O
    O.build()
    cores = O.set_num_threads()
    grid_type, Nx, Ny, Nz, dt = WORKLOADS[workload]
    shrink = 1
    while Nx // shrink * (Ny // shrink) * Nz > 6e7:      # the 1/8-degree tile: sample a quarter-size tile
        shrink *= 2
    grid_type, nx, ny, Nz, dts = _scaled_dims(workload, shrink)
    m = _oracle_model(grid_type, nx, ny, Nz, dts)
    synthetic_state(m)
    M.first_time_step(m)
    t0 = time.perf_counter()
    n = 0
    while True:
        M.time_step(m)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 50:
            break
    return {"value": nx * ny * Nz * n / el, "unit": "cell-steps/s", "cores": cores, "kind": "port",
            "sample": f"CPU oracle (C++/OpenMP restatement, not Oceananigans: no Julia in the image) on {grid_type} "
                      f"{nx}x{ny}x{Nz}, dt = {dts:g} s, {n} steps in {el:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tripolar_quarter_degree", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = the workload's tile on every GPU (default); strong = the workload's GLOBAL grid "
                         "split over (Rx, Ry) = factors(N) (BASELINE.json configs[2])")
    ap.add_argument("--no-partition-check", action="store_true")
    ap.add_argument("--float-type", default="Float32", choices=["Float32", "Float64"],
                    help="Float32 = BASELINE.json's metric (default); Float64 = libgb25cuda_f64.so, reported beside the "
                         "reference's published Float64 numbers (BASELINE.md)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args, args.workload)
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly one JSON line: anything a library prints (e.g. NCCL's version banner) goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import gb25_b200  # noqa: F401
    from gb25_b200 import model as M, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libgb25cuda has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    grid_type, Nx, Ny, Nz, dt = WORKLOADS[args.workload]
    Rx, Ry = sharding.factors(world)
    if args.scaling == "strong" and world > 1:
        if Nx % Rx or Ny % Ry:
            raise SystemExit(f"--scaling strong: {Nx}x{Ny} is not divisible by the partition {Rx}x{Ry}")
        Nx, Ny = Nx // Rx, Ny // Ry

    # N > 1: before anything is timed, the partitioned run must equal the single-GPU run of the same global problem
    # bit for bit (the reference's sharded correctness protocol on 64x48x10 tiles, both grids)
    bit_identical = None
    if world > 1 and not args.no_partition_check:
        from gb25_b200 import distributed as D
        bit_identical = all(D.partition_check(dist, local_rank, gt, 64, 48, 10, 5,
                                              log=(lambda m: print(m, file=sys.stderr)) if rank == 0 else None)
                            for gt in ("simple_lat_lon", "gaussian_islands"))

    FT = np.float64 if args.float_type == "Float64" else np.float32
    if world > 1:
        from gb25_b200 import distributed as D
        model = D.sharded_baroclinic_instability_model(M.B200(local_rank), Nx, Ny, Nz, Δt=dt, grid_type=grid_type,
                                                       Rx=Rx, Ry=Ry, rank=rank, dist=dist, float_type=FT)
    else:
        model = M.baroclinic_instability_model(M.B200(local_rank), Nx, Ny, Nz, Δt=dt, grid_type=grid_type, float_type=FT)
    synthetic_state(model, seed=42 + rank)
    if dist is not None:
        from gb25_b200 import distributed as D
        D.barrier(model)          # uploads must not race with a neighbour's halo pushes
    M.first_time_step(model)
    for _ in range(args.warmup - 1):
        M.time_step(model)
    model.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        model.synchronize()

    # ---- timed region: K steps, state resident in HBM, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)      # let nvidia-smi come up so that the timed region is covered by samples
    l0 = model.handle.launch_count()
    model.handle.call("gb25_enable_stage_timers", 1)
    barrier()
    wall0 = time.time()
    M.loop(model, args.steps)
    model.synchronize()
    wall1 = time.time()
    secs = model.handle.last_loop_seconds()
    barrier()
    stages = model.handle.stage_times()
    model.handle.call("gb25_enable_stage_timers", 0)
    launches = model.handle.launch_count() - l0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    if dist is not None:
        t = torch.tensor([secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    cells_per_rank = Nx * Ny * Nz
    value = cells_per_rank * world * args.steps / secs
    # sanity: the run must still be finite
    eta = model.interior("eta")
    finite = bool(np.isfinite(eta).all())

    # ---- e2e: every step round-trips the prognostic state through pinned host buffers over the C ABI, the way the
    # reference moves state: set!(model, u=..., v=..., ...) writes interiors, Array(interior(psi)) reads them
    e2e = None
    if not args.no_e2e:
        names = ("u", "v", "T", "S", "eta", "U", "V")
        host = []
        for n in names:
            tns = torch.empty(model.handle.interior_shape(n), dtype=torch.float64 if FT is np.float64 else torch.float32, pin_memory=True)
            host.append(tns.numpy())
        model.handle.get_fields(names, host, interior=True)
        nb = sum(a.nbytes for a in host)
        ksteps = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            model.handle.set_fields(names, host, interior=True)
            if dist is not None:
                dist.barrier()    # a tile's upload must be complete before a neighbour pushes halos into it
            M.time_step(model)             # (the uploaded state is the one just downloaded: halos and tendencies are current)
            model.handle.get_fields(names, host, interior=True)
        barrier()
        el = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        e2e = {"value": cells_per_rank * world * ksteps / el, "unit": "cell-steps/s", "h2d_bytes_per_step": nb,
               "d2h_bytes_per_step": nb, "steps": ksteps,
               "protocol": "per step: gb25_set_fields(interiors of u,v,T,S,eta,U,V) from pinned host = set!(model, ...), "
                           "gb25_time_step, gb25_get_fields(interiors) = Array(interior(psi)); "
                           "one cudaMemcpy3DAsync per field, one synchronisation per batch"}
        # the reference's own usage (sync_states!, loop!(model, Ninner), compare_states): one upload, Ninner steps in one
        # host call, one download — reported beside the per-step protocol, not instead of it
        ninner = args.steps
        phost = [torch.empty(model.handle.field_shape(n), dtype=torch.float64 if FT is np.float64 else torch.float32,
                             pin_memory=True).numpy() for n in names]
        model.handle.get_fields(names, phost)
        pb = sum(a.nbytes for a in phost)
        barrier()
        t0 = time.perf_counter()
        model.handle.set_fields(names, phost)
        if dist is not None:
            dist.barrier()
        M.loop(model, ninner)
        model.handle.get_fields(names, phost)
        barrier()
        el = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        e2e["loop_protocol"] = {"value": cells_per_rank * world * ninner / el, "unit": "cell-steps/s", "steps": ninner,
                                "h2d_bytes": pb, "d2h_bytes": pb,
                                "protocol": "sync_states! (parents uploaded once), loop!(model, Ninner), download once; wall clock"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    hbm, peak_src = measured_peaks()
    # dominant kernel = the stage with the largest share of the step
    kernels = {k: v for k, v in stages.items() if k.startswith("kernel:")}
    exch = {k: v for k, v in stages.items() if k.startswith("exchange:")}
    stages = {k: v for k, v in stages.items() if not k.startswith(("kernel:", "exchange:"))}
    if "ab2_step_fields" in stages and "kernel:k_ab2_fused" not in kernels:      # the AB2 stage was a pointer swap: epilogue path
        for k in ("kernel:k_tracer_tma", "kernel:k_gu_tma", "kernel:k_gv_tma"):
            if k in kernels:
                kernels[k + "+ab2"] = kernels.pop(k)
    tot_ms = sum(ms for ms, _ in stages.values()) or 1.0
    pool = kernels if any(k in KERNEL_BYTES_PER_CELL for k in kernels) else stages
    dom = max((k for k in pool if k in KERNEL_BYTES_PER_CELL), key=lambda k: pool[k][0], default=None)
    roofline = None
    traffic_tab, traffic_src = ncu_traffic()
    if dom:
        ms, calls = pool[dom]
        per_call_s = ms * 1e-3 / max(calls, 1)
        achieved = KERNEL_BYTES_PER_CELL[dom] * cells_per_rank / per_call_s / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                    "frac": achieved / hbm,
                    "traffic": (traffic_tab.get(dom[7:].replace("+ab2", ""), {}).get("dram_bytes")
                                if args.workload == "tripolar_quarter_degree" else None),
                    "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": KERNEL_BYTES_PER_CELL[dom] * cells_per_rank, "peak_source": peak_src,
                    "avg_launch_ms": per_call_s * 1e3, "share_of_step": ms / tot_ms,
                    "whole_step": {"algorithmic_bytes_per_cell_step": ALGORITHMIC_BYTES_PER_CELL_STEP,
                                   "traffic": (traffic_tab.get("__step__", {}).get("dram_bytes")
                                               if args.workload == "tripolar_quarter_degree" else None),
                                   "achieved": ALGORITHMIC_BYTES_PER_CELL_STEP * value / world / 1e9,
                                   "frac": ALGORITHMIC_BYTES_PER_CELL_STEP * value / world / 1e9 / hbm},
                    "stage_ms_per_step": {k: v[0] / args.steps for k, v in stages.items()},
                    "kernel_ms_per_launch": {k[7:]: v[0] / max(v[1], 1) for k, v in kernels.items()},
                    "exchange_ms_per_step": {k[9:]: v[0] / args.steps for k, v in exch.items()}}
    field_mb = (Nx + 16) * (Ny + 17) * (Nz + 17) * (8 if FT is np.float64 else 4) / 1e6
    l2_note = (f"inputs larger than L2: 21 3-D arrays of {field_mb:.0f} MB each are streamed every step (L2 = 126 MB)"
               if 21 * field_mb > 2 * 126 else "working set fits in L2: latency-bound config")
    line = {"metric": "cell_steps_per_s", "value": value, "unit": "cell-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f64" if FT is np.float64 else "f32", "data": "synthetic",
            "config": {"workload": args.workload, "grid": grid_type, "Nx_per_gpu": Nx, "Ny_per_gpu": Ny, "Nz": Nz, "dt": dt,
                       "partition": [Rx, Ry], "halo": 8, "substeps": 30, "l2": l2_note,
                       "state_finite": finite},
            "multi_gpu_bit_identical": bit_identical,
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(args.workload)
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()
    if not finite:
        raise SystemExit("bench.py: the model state is not finite after the timed region")
    if bit_identical is False:
        raise SystemExit("bench.py: the partitioned run differs from the single-GPU run (partition_check)")


if __name__ == "__main__":
    main()
