#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: total stall samples by reason, and the hottest SASS lines."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter(); lines = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == 'Address': continue
    n = int(r[col['# Samples']] or 0)
    for h in stall_cols:
        tot[h] += int(r[col[h]] or 0)
    lines.append((n, r[col['Source']][:110], int(r[col['Instructions Executed']] or 0)))
s = sum(tot.values())
print('total samples', s)
for h, v in tot.most_common(12): print(f'  {h:28s} {v:8d} {100*v/max(s,1):5.1f}%')
print('hottest instructions:')
for n, src, ie in sorted(lines, reverse=True)[:top]: print(f'  {n:7d} {100*n/max(s,1):5.1f}%  exec={ie:9d}  {src}')
