"""Static SASS instruction mix of the hot kernels (cuobjdump -sass on the in-tree objects; no GPU needed).
usage: python scripts/sass_mnemonics.py > profiles/r2/sass_r2_mnemonics.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gb-25_b200", "csrc")
KERNELS = [("gb25_tend_tma.o", "k_tracer_tmaILb1E"), ("gb25_tend_tma.o", "k_mom_tma_p2ILi0ELb1E"), ("gb25_tend_tma.o", "k_mom_tma_p2ILi1ELb1E"),
           ("gb25_tend_tma.o", "k_tracer_tmaILb0E"), ("gb25_tend_tma.o", "k_mom_tma_p2ILi0ELb0E"),
           ("gb25_baro.o", "k_baro_persistentILi5ELb0"), ("gb25_baro.o", "k_baro_persistentILi5ELb1"),
           ("gb25_kernels.o", "k_compute_p2"), ("gb25_kernels.o", "k_ab2_uv"), ("gb25_kernels.o", "k_ab2_ts_3d"),
           ("gb25_kernels.o", "k_correct_3dILb1E"), ("gb25_tend_v2.o", "k_aux_columns_vecILi4E"), ("gb25_tend_v2.o", "k_generic_list"),
           ("gb25_exchange.o", "k_push_cols_packed"), ("gb25_exchange.o", "k_push_rows")]
COLS = ["UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "FSEL", "FMNMX", "MUFU", "FCHK", "CALL", "LDS", "LDG", "STG", "LD", "ST", "BAR"]


def mix(obj, pat):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    cnt, on, total = collections.Counter(), False, 0
    for ln in out.splitlines():
        if "Function :" in ln:
            on = pat in ln
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            cnt[m.group(2)] += 1
            total += 1
    return cnt, total


print("Static SASS instruction counts (sm_100a, `cuobjdump -sass`; `UTMALDG` = TMA tensor load, `SYNCS` = mbarrier,")
print("`FFMA2/FADD2/FMUL2` = packed FP32x2, `FCHK`/`CALL` = range-checked IEEE division slow paths).\n")
print("| kernel | total | " + " | ".join(COLS) + " |")
print("|---|---:|" + "---:|" * len(COLS))
for obj, pat in KERNELS:
    c, t = mix(obj, pat)
    if t:
        print(f"| `{pat}` | {t} | " + " | ".join(str(c.get(k, 0)) for k in COLS) + " |")
