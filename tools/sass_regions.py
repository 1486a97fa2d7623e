#!/usr/bin/env python
"""Break an `ncu --page source --csv` dump into hot code regions (runs of SASS with equal execution count):
share of warp instructions / samples, stall mix, opcode mix of the hottest region.
    ncu -i x.ncu-rep --page source --csv > src.csv ; python tools/sass_regions.py src.csv"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, ins, iti, isamp = (hdr.index(k) for k in ('Address', 'Source', 'Instructions Executed', 'Thread Instructions Executed', '# Samples'))
names = ['stall_long_sb', 'stall_math', 'stall_no_inst', 'stall_not_selected', 'stall_selected', 'stall_short_sb', 'stall_wait', 'stall_lg', 'stall_dispatch', 'stall_branch_resolving']
idx = {n: hdr.index(n) for n in names}
data = [r for r in rows[2:] if len(r) > max(idx.values()) and r[ins].isdigit()]
base = int(data[0][ia], 16)
tot = sum(int(r[ins]) for r in data); tots = sum(int(r[isamp]) for r in data)
full = sum(int(r[ins]) for r in data if int(r[iti]) == 32 * int(r[ins]))
print('warp instr', tot, 'samples', tots, 'full-warp frac %.3f' % (full / tot))
seg = []; cur = None
for r in data:
    a, n = int(r[ia], 16) - base, int(r[ins])
    if cur and cur['n'] == n: cur['b'] = a; cur['k'] += 1; cur['rows'].append(r)
    else:
        cur = {'a': a, 'b': a, 'n': n, 'k': 1, 'rows': [r]}; seg.append(cur)
S = {n: 0 for n in names}
for r in data:
    for n in names: S[n] += int(r[idx[n]] or 0)
T = sum(S.values())
print('ALL   ', {n[6:]: round(100 * S[n] / T, 1) for n in names})
for x in sorted(seg, key=lambda x: -x['n'] * x['k'])[:int(sys.argv[2]) if len(sys.argv) > 2 else 8]:
    M = {n: 0 for n in names}
    for r in x['rows']:
        for n in names: M[n] += int(r[idx[n]] or 0)
    TT = max(1, sum(M.values()))
    c = Counter()
    for r in x['rows']:
        toks = r[isrc].split(); op = toks[1] if toks[0].startswith('@') else toks[0]; c[op.rstrip(';')] += 1
    print('%#x-%#x exec %d x %d instr = %.1f%% of warp instr, %.1f%% of samples' % (x['a'], x['b'], x['n'], x['k'], 100 * x['n'] * x['k'] / tot, 100 * sum(int(r[isamp]) for r in x['rows']) / tots))
    print('      stalls', {n[6:]: round(100 * M[n] / TT, 1) for n in names})
    print('      ops', dict(c.most_common(9)))
