#!/usr/bin/env python
"""Hot source lines of one kernel from an ncu report captured with --import-source on:
    ncu -i x.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:NAME > src.csv
    python tools/ncu_lines.py src.csv [top]
Rows of the CUDA-C view carry the per-line totals (warp instructions executed, stall samples)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
lines = []
for r in rows[hi + 1:]:
    if len(r) > ie and r[0].isdigit() and r[ie].replace(",", "").isdigit():
        lines.append((int(r[0]), r[1], int(r[ie].replace(",", "")), int(r[isamp].replace(",", "") or 0)))
    elif "Instructions Executed" in r:
        break
tot = sum(x[2] for x in lines) or 1
ts = sum(x[3] for x in lines) or 1
print("warp instructions", tot, "samples", ts)
for ln, src, n, s in sorted(sorted(lines, key=lambda x: -x[2])[:top]):
    print("%5d %6.2f%% instr %6.2f%% samples | %s" % (ln, 100.0 * n / tot, 100.0 * s / ts, src.strip()[:140]))
