#!/usr/bin/env python
"""Stall-reason and per-opcode sample breakdown of one kernel from an ncu source-page CSV.
Usage: ncu -i REP --page source --csv --print-source sass --kernel-name regex:X --launch-skip N
       --launch-count 1 > k.csv;  tools/ncu_stalls.py k.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
print(rows[0][1][:100])
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 40 and r[4].isdigit()]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter(); total = 0
byop = collections.Counter(); byop_n = collections.Counter()
for r in data:
    s = int(r[ix['# Samples']]); total += s
    for h in stalls:
        tot[h] += int(r[ix[h]])
    t = r[ix['Source']].split()
    if not t:
        continue
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    byop[op] += s; byop_n[op] += int(r[ix['Instructions Executed']])
print("samples", total, "instructions", sum(byop_n.values()))
for h, c in tot.most_common(10):
    print(f"{h:28s} {c:8d} {100*c/total:5.1f}%")
print("--- by opcode: samples, share, executed, samples per 1000 executed")
for op, c in byop.most_common(24):
    print(f"{op:10s} {c:8d} {100*c/total:5.1f}%  exec {byop_n[op]:10d}  {1000*c/max(byop_n[op],1):7.2f}")
