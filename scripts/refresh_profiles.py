"""Turn the files a `gpurun` profiling call left under gpurun_out/ into the tracked evidence under profiles/r2/:
the ncu raw CSV, the per-kernel table, profiles/ncu_traffic.json (keyed by the kernel sources' sha), the launch-list summary,
the bench line and the SASS mix.  Run from the repo root after
    python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
    ncu --set full --clock-control none --import-source on --launch-skip 62 --launch-count 15 -o gpurun_out/r2_final_prof -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline
"""
import collections, csv, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
out = "profiles/r2"
os.makedirs(out, exist_ok=True)
raw = f"{out}/ncu_r2_final_raw.csv"
with open(raw, "w") as f:
    subprocess.run(["ncu", "-i", "gpurun_out/r2_final_prof.ncu-rep", "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=True)
cap = ("ncu --set full --clock-control none, 15 consecutive launches of a steady-state step of bench.py --steps 2 --warmup 3 "
       f"(tripolar 1440x600x50), {raw}")
with open(f"{out}/ncu_r2_final_kernels.md", "w") as f:
    subprocess.run([sys.executable, "scripts/ncu_table.py", raw, "--traffic", "profiles/ncu_traffic.json", "--capture", cap], stdout=f, check=True)
shutil.copy("gpurun_out/r2_launches.csv", f"{out}/launches_r2_final.csv")
shutil.copy("gpurun_out/r2_final_bench.json", f"{out}/bench_r2_final.json")
with open(f"{out}/sass_r2_mnemonics.md", "w") as f:
    subprocess.run([sys.executable, "scripts/sass_mnemonics.py"], stdout=f, check=True)
rows = list(csv.reader(open("gpurun_out/r2_launches.csv")))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ix = {h: j for j, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    n = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ix["Metric Unit"]], 1)
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
lines = ["| kernel | launches | total µs | avg µs | share |", "|---|---:|---:|---:|---:|"]
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{n}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f} % |")
lines.append(f"| **total** | {sum(a[0] for a in agg.values())} | {tot:.1f} | | |")
open(f"{out}/launches_r2_final_summary.md", "w").write(
    "ncu launch list of `bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline` (first step + 4 steps; cold-cache, serialised: "
    "compare shares)\n\n" + "\n".join(lines) + "\n")
print("\n".join(lines[:8]))
