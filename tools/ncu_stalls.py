"""Warp-stall reason totals of one kernel from an ncu source page CSV (see ncu_sass_hist.py)."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
cols = [i for i, n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n]
tot = {hdr[i]: 0 for i in cols}
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    for i in cols:
        tot[hdr[i]] += int(r[i] or 0)
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda t: -t[1]):
    print(f'{k:28s} {v:8d} {v / max(s, 1):6.3f}')
