"""Key metrics of an `ncu -i X.ncu-rep --page raw --csv` dump as a markdown table, and (with --traffic OUT.json) the per-kernel
DRAM traffic file that bench.py reads for `roofline.traffic` (keyed by the sha of the kernel sources and of the device code it was captured with).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python scripts/ncu_table.py raw.csv [--traffic profiles/ncu_traffic.json --capture "<how it was taken>"] > table.md
"""
import csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

args = sys.argv[1:]
rows = list(csv.reader(open(args[0])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue-slot active %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long-scoreboard stall / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "fixed-latency (wait) stall / issue"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "math-pipe throttle / issue")]
names = [re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "") for r in rows[2:]]
print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
print("|---|" + "---:|" * len(names))
for key, label in want:
    if key in idx:
        print(f"| {label} ({units[idx[key]]}) | " + " | ".join(r[idx[key]] for r in rows[2:]) + " |")


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


if "--traffic" in args:
    import bench
    out = args[args.index("--traffic") + 1]
    capture = args[args.index("--capture") + 1] if "--capture" in args else args[0]
    # timer names of libgb25cuda ("kernel:<name>") for the kernels bench.py reports on
    alias = {"k_mom_tma_p2<0, 1>": "k_gu_tma", "k_mom_tma_p2<1, 1>": "k_gv_tma", "k_mom_tma_p2<0, 0>": "k_gu_tma",
             "k_mom_tma_p2<1, 0>": "k_gv_tma", "k_tracer_tma<1>": "k_tracer_tma", "k_tracer_tma<0>": "k_tracer_tma",
             "k_aux_columns_vec<4>": "k_aux_columns"}
    kern, step = {}, 0.0
    for n, r in zip(names, rows[2:]):
        b = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
            to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        step += b
        key = alias.get(n, n)
        e = kern.setdefault(key, {"dram_bytes": 0.0, "launches": 0})
        e["dram_bytes"] += b; e["launches"] += 1
    for e in kern.values():      # per launch, like roofline.achieved
        e["dram_bytes"] = e["dram_bytes"] / e["launches"]
    kern["__step__"] = {"dram_bytes": step, "launches": len(names)}
    json.dump({"kernel_source_sha": bench.kernel_source_sha(), "kernel_sass_sha": bench.kernel_sass_sha(), "capture": capture, "kernels": kern}, open(out, "w"), indent=1)
