#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: warp-stall samples and executed instructions per
contiguous code region (regions = runs of SASS instructions with a similar execution count, i.e. the
warp-specialised roles and their loops), plus the hottest instructions.
usage: tools/ncu_regions.py gpurun_out/src.csv [min_samples]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
h = rows[1]
idx = {n: i for i, n in enumerate(h)}
data = rows[2:]
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
S = [int(r[idx['# Samples']] or 0) for r in data]
E = [int(r[idx['Instructions Executed']] or 0) for r in data]
print(rows[0][1][:100], 'samples', sum(S), 'warp-instr', sum(E))


def agg(a, b):
    c = collections.Counter()
    for r in data[a:b + 1]:
        for n in stalls:
            v = int(r[idx[n]] or 0)
            if v:
                c[n[6:]] += v
    return c


seg, cur, start = [], None, 0
for i, e in enumerate(E):
    if cur is None:
        cur, start = e, i
    elif e > 0 and abs(e - cur) > 0.3 * max(cur, 1):
        seg.append((start, i - 1, cur))
        cur, start = e, i
seg.append((start, len(E) - 1, cur))
for a, b, e in seg:
    s = sum(S[a:b + 1])
    if s >= minS or sum(E[a:b + 1]) > 0.02 * sum(E):
        print(f"[{a:5d}-{b:5d}] exec/instr {e:8d} warp-instr {sum(E[a:b+1]):9d} samples {s:5d}", agg(a, b).most_common(5))
print('hottest instructions:')
for i in sorted(sorted(range(len(data)), key=lambda i: -S[i])[:25]):
    r = data[i]
    st = {n[6:]: int(r[idx[n]] or 0) for n in stalls}
    print(i, S[i], E[i], r[idx['Source']][:64], {k: v for k, v in st.items() if v})
