"""Times single shift-GEMM convolution layers through the C ABI's gsx_op_conv hook, optionally sweeping
plan overrides.  Usage (on the GPU box):
  python tools/conv_sweep.py --layer g10.conv2 --n 32 [--sweep] [--once]
"""
import argparse
import itertools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gan_segmentation_b200 import _lib as L, ops  # noqa: E402

# name: (mode, cin0, cin1, cout, H, W, flags, extras)
LAYERS = {
    'g10.conv2': ('CONV3', 16, 0, 16, 1024, 1024, 'gen'),
    'g10.conv2.nonoise': ('CONV3', 16, 0, 16, 1024, 1024, 'gen_nonoise'),
    'g10.conv2.nostats': ('CONV3', 16, 0, 16, 1024, 1024, 'gen_nostats'),
    'g9.conv2': ('CONV3', 32, 0, 32, 512, 512, 'gen'),
    'g8.conv2': ('CONV3', 64, 0, 64, 256, 256, 'gen'),
    'g7.conv2': ('CONV3', 128, 0, 128, 128, 128, 'gen'),
    'g6.conv2': ('CONV3', 256, 0, 256, 64, 64, 'gen'),
    'g5.conv2': ('CONV3', 512, 0, 512, 32, 32, 'gen'),
    'g10.deconv': ('DECONV4', 32, 0, 16, 512, 512, 'raw'),
    'g9.deconv': ('DECONV4', 64, 0, 32, 256, 256, 'raw'),
    'g8.deconv': ('DECONV4', 128, 0, 64, 128, 128, 'raw'),
    'g10.deconvb': ('DECONV4B', 32, 0, 16, 512, 512, 'gen'),
    'g9.deconvb': ('DECONV4B', 64, 0, 32, 256, 256, 'gen'),
    'g8.deconvb': ('DECONV4B', 128, 0, 64, 128, 128, 'gen'),
    'g7.deconv': ('DECONV4', 256, 0, 128, 64, 64, 'raw'),
    'g6.upconv': ('UPCONV3', 512, 0, 256, 32, 32, 'raw'),
    'g5.upconv': ('UPCONV3', 512, 0, 512, 16, 16, 'raw'),
    'd8.cvt': ('CONV3', 16, 0, 16, 1024, 1024, 'dec'),
    'd7.cvt': ('CONV3', 32, 0, 32, 512, 512, 'dec'),
    'd7.conv_a': ('UPCONV3', 32, 32, 16, 512, 512, 'dec'),
    'd7.conv_b': ('CONV3', 16, 0, 16, 1024, 1024, 'res'),
    'd6.conv_b': ('CONV3', 32, 0, 32, 512, 512, 'res'),
    'd7.shortcut': ('CONV1', 32, 32, 16, 512, 512, 'lin'),
    'd6.conv_a': ('UPCONV3', 32, 32, 32, 256, 256, 'dec'),
    'd8.final': ('CONV3', 16, 16, 2, 1024, 1024, 'argmax'),
}


def run(name, n, override, repeat, dtype):
    mode, c0, c1, co, h, w, kind = LAYERS[name]
    g = torch.Generator().manual_seed(0)
    x = torch.randn((n, c0 + c1, h, w), generator=g).cuda()
    k = {'CONV3': 3, 'UPCONV3': 3, 'DECONV4': 4, 'DECONV4B': 4, 'CONV1': 1}[mode]
    wt = (torch.randn((c0 + c1, co, k, k) if mode in ('DECONV4', 'DECONV4B') else (co, c0 + c1, k, k), generator=g) / np.sqrt((c0 + c1) * k * k)).numpy()
    up = mode in ('UPCONV3', 'DECONV4', 'DECONV4B')
    ho, wo = (2 * h, 2 * w) if up else (h, w)
    kw = {}
    if kind == 'gen':
        kw = dict(bias=torch.randn(co).cuda(), nscale=torch.randn(co).cuda(), noise=torch.randn((n, 1, ho, wo)).cuda(),
                  flags=L.EPI_LRELU | L.EPI_STATS)
    elif kind == 'gen_nonoise':
        kw = dict(bias=torch.randn(co).cuda(), flags=L.EPI_LRELU | L.EPI_STATS)
    elif kind == 'gen_nostats':
        kw = dict(bias=torch.randn(co).cuda(), nscale=torch.randn(co).cuda(), noise=torch.randn((n, 1, ho, wo)).cuda(),
                  flags=L.EPI_LRELU)
    elif kind == 'dec':
        kw = dict(bias=torch.randn(co).cuda(), flags=L.EPI_LRELU)
    elif kind == 'res':
        kw = dict(bias=torch.randn(co).cuda(), flags=L.EPI_LRELU, addsrc=torch.randn((n, co, ho // 2, wo // 2)).cuda())
    elif kind == 'lin':
        kw = dict(bias=torch.randn(co).cuda())
    elif kind == 'argmax':
        kw = dict(bias=torch.randn(16).cuda(), flags=L.EPI_ARGMAX, num_classes=co)
    r = ops.conv(getattr(L, mode), x[:, :c0], wt, x1=x[:, c0:] if c1 else None, override=override, repeat=repeat,
                 dtype=dtype, **kw)
    by = 2.0 * n * (c0 + c1) * h * w + (n * ho * wo if kind == 'argmax' else 2.0 * n * co * ho * wo)
    if kind == 'gen':
        by += 4.0 * n * ho * wo
    if kind == 'res':
        by += 2.0 * n * co * ho * wo / 4
    taps = {'CONV3': 9, 'UPCONV3': 9, 'DECONV4': 4, 'DECONV4B': 4, 'CONV1': 1}[mode]
    fl = 2.0 * n * ho * wo * taps * (c0 + c1) * co
    return r, by, fl


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--layer', default='g10.conv2')
    ap.add_argument('--n', type=int, default=32)
    ap.add_argument('--repeat', type=int, default=20)
    ap.add_argument('--dtype', default='fp16')
    ap.add_argument('--sweep', action='store_true')
    ap.add_argument('--override', default='', help='json dict of plan overrides')
    args = ap.parse_args()
    names = list(LAYERS) if args.layer == 'all' else args.layer.split(',')
    for name in names:
        ovs = [json.loads(args.override) if args.override else None]
        if args.sweep:
            mode = LAYERS[name][0]
            ths = [7, 15, 23, 31, 47] if mode in ('CONV3', 'CONV1') else [3, 5, 7, 11, 15]
            ovs = [None] + [dict(TH=th, acc_bufs=ab) for th, ab in itertools.product(ths, [1, 2])]
        for ov in ovs:
            try:
                r, by, fl = run(name, args.n, ov, args.repeat, args.dtype)
            except Exception as e:  # infeasible override
                print(f'{name} override={ov}: {e}')
                continue
            ms = r['ms']
            p = r['plan']
            print(f"{name} n={args.n} ov={ov} ms={ms:.4f} {by / ms / 1e6:.0f} GB/s {fl / ms / 1e9:.0f} TFLOP/s | TH={p['TH']} TW={p['TW']} "
                  f"NB={p['NB']} CBK={p['CBK']} Nt={p['N_tile']} st={p['stages']} mt={p['n_mtiles']} nk={p['n_k']} "
                  f"tmem={p['tmem_cols']} smem={p['smem_bytes']} ctas={p['tiles']}", flush=True)
