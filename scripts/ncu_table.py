"""Key metrics of an `ncu -i X.ncu-rep --page raw --csv` dump as a markdown table."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue-slot active %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long-scoreboard stall / issue")]
names = [re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "") for r in rows[2:]]
print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
print("|---|" + "---:|" * len(names))
for key, label in want:
    if key in idx:
        print(f"| {label} ({units[idx[key]]}) | " + " | ".join(r[idx[key]] for r in rows[2:]) + " |")
