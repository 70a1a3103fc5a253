"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share)."""
import csv, re, sys
from collections import defaultdict
path, skip = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
rows = rows[skip:]
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    tot[name] += float(r[14]) * 1e-3; cnt[name] += 1
total = sum(tot.values())
print(f"| kernel | launches | total µs | avg µs | share |\n|---|---:|---:|---:|---:|")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"| `{k}` | {cnt[k]} | {tot[k]:.1f} | {tot[k]/cnt[k]:.1f} | {100*tot[k]/total:.1f} % |")
print(f"| **total** | {sum(cnt.values())} | {total:.1f} | | |")
