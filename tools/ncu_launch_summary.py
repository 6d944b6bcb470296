"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
  python tools/ncu_launch_summary.py launches.csv [first_row_fraction_to_skip=0]"""
import collections
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
ci = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hi + 2:] if len(r) >= len(hdr)]
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
data = data[int(len(data) * skip):]
agg = collections.OrderedDict()
for r in data:
    k = r[ci['Kernel Name']][:70]
    v = float(r[ci['Metric Value']].replace(',', ''))
    if r[ci['Metric Unit']] in ('ns', 'nsecond'):
        v /= 1000.0
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print(f'{len(data)} launches, {tot:.1f} us (cold-cache, serialised)')
for k, v in sorted(agg.items(), key=lambda t: -t[1][1]):
    print(f'{v[1]:10.1f} us {v[1] / tot:6.3f} {v[0]:5d}  {k}')
