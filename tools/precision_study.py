"""CPU emulation of the rounding points of the CUDA generate path (16-bit storage of weights and activations, fp32
accumulation and statistics) on top of the oracle, to localise which stage produces the image-error tail at a given
size.  Test/tuning infrastructure only (imports oracle/).

  python tools/precision_study.py --res 10 --seed 41 --psi 0.7 [--skip raw,t1,u1,t2,x,w] [--from-res R] [--dtype fp16|bf16]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import generate_oracle as O          # noqa: E402
from parity_util import make_case, psnr           # noqa: E402


def emulate(gp, gc, z, noise, psi, on, qdt=torch.float16, from_res=2, fold_res=(10,), new_fold=False, only_res=None):
    """on: set of rounding points that are active: w (weights), raw (conv1 output before blur), t1, u1, t2, x."""
    def q(x, key, r=99):
        lo = on.get(key, 99) if isinstance(on, dict) else (from_res if key in on else 99)   # dict: key -> first res rounded
        if r >= lo and (only_res is None or r == only_res):
            return x.to(qdt).float()
        return x
    P = {k: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in gp.items()}
    z = torch.from_numpy(z).float()
    noise = [torch.from_numpy(a).float() for a in noise]
    L = gc['max_res_log2']
    n = z.shape[0]
    w = O.mapping(P, gc, z)
    tp = torch.broadcast_to(torch.tensor(np.asarray(psi, np.float32)), (2 * (L - 1),)) if psi is not None else P['truncation_psi']
    avg = P['latent_avg'].reshape(1, -1)
    y = q(P['constant_tensor'], 'x', 2).expand(n, -1, -1, -1)

    def coef(t, wl, p, k):
        st = O.dense_w(wl, P[f'{p}.adain{k}.affine.weight'], P[f'{p}.adain{k}.affine.bias'], P[f'{p}.adain{k}.affine.std'], 1.0)
        c = t.shape[1]
        st = st.reshape(n, 2, c)
        ys, yb = st[:, 0].reshape(n, c, 1, 1), st[:, 1].reshape(n, c, 1, 1)
        mean = t.mean(dim=(2, 3), keepdim=True)
        var = (t * t).mean(dim=(2, 3), keepdim=True) - mean * mean
        a = torch.rsqrt(var.clamp_min(0) + 1e-5) * (ys + 1)
        return a, yb - mean * a

    feats = []
    for r in range(2, L + 1):
        i = 2 * (r - 2)
        p = f'net{r}'
        w1 = O.lerp(tp[i], avg, w)
        w2 = O.lerp(tp[i + 1], avg, w)
        if r > 2:
            wgt = P[f'{p}.block0.weight'] * P[f'{p}.block0.std']
            if r >= 7:
                t = F.conv_transpose2d(y, q(wgt, 'w', r), None, stride=2, padding=1)
            else:
                t = F.conv2d(F.interpolate(y, scale_factor=2, mode='nearest'), q(wgt, 'w', r), None, 1, 1)
            if r not in fold_res:
                t = q(t, 'raw', r)
            t = O.blur(t)
        else:
            t = y
        t = F.leaky_relu(t + P[f'{p}.block1.0.scale_factors'] * noise[i] + P[f'{p}.block1.1.bias'], 0.2)
        a, b = coef(t, w1, p, 1)
        t = q(t, 't1', r)
        w2c = P[f'{p}.block2.0.weight'] * P[f'{p}.block2.0.std']
        if new_fold and r >= new_fold:
            # AdaIN folded into the consumer: per-sample modulated weights (rounded), bias term exact in fp32
            outs = []
            for s in range(n):
                wm = q(w2c * a[s].reshape(1, -1, 1, 1), 'w', r)
                bb = F.conv2d(b[s:s + 1].expand(1, -1, t.shape[2], t.shape[3]), w2c, None, 1, 1)
                outs.append(F.conv2d(t[s:s + 1], wm, None, 1, 1) + bb)
            t = torch.cat(outs)
        else:
            u = q(a * t + b, 'u1', r)
            t = F.conv2d(u, q(w2c, 'w', r), None, 1, 1)
        t = F.leaky_relu(t + P[f'{p}.block2.1.scale_factors'] * noise[i + 1] + P[f'{p}.block2.2.bias'], 0.2)
        a, b = coef(t, w2, p, 2)
        t = q(t, 't2', r)
        xf = a * t + b
        y = q(xf, 'x', r)
        feats.append(y)
    wr = P[f'to_rgb{L}.0.weight'] * P[f'to_rgb{L}.0.std']
    img = F.conv2d(xf, wr, P[f'to_rgb{L}.0.bias'])
    return img, feats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--res', type=int, default=10)
    ap.add_argument('--seed', type=int, default=41)
    ap.add_argument('--n', type=int, default=1)
    ap.add_argument('--psi', type=float, default=0.7)
    ap.add_argument('--dtype', default='fp16')
    ap.add_argument('--variants', default='all')
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    gc, dc, gp, dp, z, noise = make_case(args.res, args.n, seed=args.seed)
    qdt = torch.float16 if args.dtype == 'fp16' else torch.bfloat16
    with torch.no_grad():
        t0 = time.time()
        ref, rfeats = O.generator_forward(gp, gc, z, noise, args.psi)
        print(f'oracle {time.time() - t0:.1f}s', flush=True)
        ALL = {'w', 'raw', 't1', 'u1', 't2', 'x'}
        variants = [('all', ALL, 2, False)]
        if args.variants == 'all':
            variants += [('only ' + k, {k}, 2, False) for k in sorted(ALL)]
            variants += [(f'all from res {r}', ALL, r, False) for r in (5, 7, 8, 9, 10)]
            variants += [('all, apply folded (r>=8)', ALL, 2, 8), ('all, apply folded (r>=2)', ALL, 2, 2)]
        for name, on, fr, nf in variants:
            img, feats = emulate(gp, gc, z, noise, args.psi, on, qdt, fr, new_fold=nf)
            a, b = img.clamp(-1, 1).numpy(), ref.clamp(-1, 1).numpy()
            d = np.abs(a - b)
            fe = [float(((f - g) ** 2).mean().sqrt() / (g ** 2).mean().sqrt()) for f, g in zip(feats, rfeats)]
            print(f'{name:28s} max-abs {d.max():.4f} psnr {psnr(a, b):.1f} frac>2e-2 {float((d > 2e-2).mean()):.2e} '
                  f'frac>1e-2 {float((d > 1e-2).mean()):.2e} feat rel-rms last3 {["%.5f" % e for e in fe[-3:]]}', flush=True)


if __name__ == '__main__':
    main()
