"""Runs the wide conv layers of the FFHQ step once each at N=32 (for an ncu metrics pass)."""
import sys
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo')
from conv_sweep import run
for name in ['g8.conv2', 'g7.conv2', 'g5.conv2', 'g8.deconv']:
    r, by, fl = run(name, 32, None, 0, 'fp16')
    print(name, 'algorithmic_bytes', int(by), 'algorithmic_flops', int(fl), r['plan'], flush=True)
