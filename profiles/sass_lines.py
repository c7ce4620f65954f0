#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS dump per CUDA source line.

usage: sass_lines.py <nvdisasm -g output> <mangled kernel name> <ncu source csv> [top N]
ncu's CSV source page is SASS-only; nvdisasm -g gives the line of every SASS address.  Inlined code is attributed to
the innermost line (`inlined at` chains are ignored)."""
import csv, re, sys
from collections import defaultdict

sass, kern, src = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
addr2line = {}
cur = None; inside = False
for ln in open(sass):
    if ln.startswith('.text.'):
        inside = ln.strip() == f'.text.{kern}:'
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m and 'inlined at' not in ln.split('line')[0]:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
h = rows[1]
ia, isamp, iinst = h.index('Address'), h.index('# Samples'), h.index('Instructions Executed')
istall = {n: i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n}
base = None
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot_s = tot_i = 0
for r in rows[2:]:
    if len(r) <= iinst: continue
    if r[ia] == 'Address': break   # a second launch of the same kernel follows: the first one is enough
    a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    if base is None: base = a
    line = addr2line.get(a - base)
    s = int(r[isamp] or 0); n = int(r[iinst] or 0)
    e = agg[line]; e[0] += s; e[1] += n; tot_s += s; tot_i += n
    for k, i in istall.items():
        v = int(r[i] or 0)
        if v: e[2][k] += v
print(f'total samples {tot_s}, warp instructions {tot_i}')
for line, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ', '.join(f'{k[6:]}={v}' for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f'{str(line):32s} samples {s:7d} ({100*s/tot_s:5.1f}%)  inst {n:11d} ({100*n/tot_i:5.1f}%)  {tops}')
