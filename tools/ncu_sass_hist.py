"""Opcode histogram / hottest SASS lines of one kernel from an ncu report's source page.
  ncu -i X.ncu-rep --page source --csv > x.csv ; python tools/ncu_sass_hist.py x.csv [N]
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
ci = {n: i for i, n in enumerate(hdr)}
ops, samp = Counter(), Counter()
lines = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[ci['Source']].strip()
    ex = int(r[ci['Instructions Executed']] or 0)
    sm = int(r[ci['# Samples']] or 0)
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    op = op.split('.')[0]
    ops[op] += ex
    samp[op] += sm
    lines.append((ex, sm, src))
tot = sum(ops.values())
tots = sum(samp.values())
print(f'total warp-instructions executed {tot}, stall samples {tots}')
print('opcode            executed   share   samples share')
for op, n in ops.most_common(topn):
    print(f'{op:16s} {n:10d}  {n / tot:6.3f}  {samp[op]:8d} {samp[op] / max(tots, 1):6.3f}')
print('--- hottest lines by stall samples')
for ex, sm, src in sorted(lines, key=lambda t: -t[1])[:topn]:
    print(f'{sm:8d} {ex:10d}  {src}')
