"""SASS opcode census of a shared library's sm_100a cubins, per kernel: the mnemonics that prove the Blackwell paths
(UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor loads, UBLKCP = bulk copies, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, FFMA2/FADD2/FMUL2 = packed fp32).   python tools/sass_census.py lib.so > profiles/rNN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pat = re.compile(r'\b(UTC[A-Z]*MMA[.\w]*|LDTM[.\w]*|UTMALDG[.\w]*|UBLKCP[.\w]*|UTCBAR[.\w]*|UTCCP[.\w]*|FFMA2|FADD2|FMUL2|SYNCS[.\w]*|ELECT|UTMAPF[.\w]*|REDUX[.\w]*)\b')
fn = None
counts = collections.defaultdict(collections.Counter)
arch = collections.Counter()
for line in out.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'arch = (sm_\w+)', line)
    if m:
        arch[m.group(1)] += 1
    if fn:
        for op in pat.findall(line.split('/*')[1] if '/*' in line and line.count('/*') > 1 else line):
            counts[fn][op] += 1
print(f'# {lib}: cubin architectures {dict(arch)}')
for f in sorted(counts):
    name = subprocess.run(['c++filt', f], capture_output=True, text=True).stdout.strip()
    print(f'\n{name}')
    for op, n in sorted(counts[f].items(), key=lambda t: (-t[1], t[0])):
        print(f'  {n:6d}  {op}')
