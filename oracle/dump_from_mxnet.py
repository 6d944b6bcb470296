#!/usr/bin/env python
"""Pins the oracle against the REAL reference -- to be run wherever Apache MXNet 1.5.x and the reference checkout are
available (neither is installable in the offline build container, which is why oracle/ says "parity unpinned").
Test infrastructure only; nothing in the product imports it.

    python oracle/dump_from_mxnet.py --reference /path/to/GAN-segmentation --out tests/golden/mxnet_res5_seed0.npz
    python -m pytest tests/test_mxnet_pin.py          # now compares oracle/ with the dump instead of skipping

What it does: rebuilds the seeded case of tests/parity_util.make_case(res, n, seed) (pure NumPy: this repo's
random_init / config modules), builds the reference's own ``Generator`` (fix_noise=True, so that the ``AddNoise._noise``
hook of networks_stylegan.py:275-298 can be pre-assigned) and ``Decoder``, sets every parameter by its structural name
(``_collect_params_with_prefix``, the names ``save_parameters`` writes), runs the reference forward passes on the CPU
context and ONE training step of the reference's loop body (seg_solver.py:386-421, dropout off so that no MXNet RNG
stream is involved), and writes image, features, logits, argmax mask, loss, gradients and updated weights.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def set_params(net, params, mx, ctx):
    table = net._collect_params_with_prefix()
    missing = [k for k in table if k not in params]
    if missing:
        raise KeyError(f'no value for reference parameters {missing[:5]} ...')
    for name, p in table.items():
        a = np.asarray(params[name], np.float32).reshape(p.shape)
        p.initialize(ctx=ctx, force_reinit=True)
        p.set_data(mx.nd.array(a, ctx=ctx))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', required=True, help='checkout of author-hidden-name/GAN-segmentation')
    ap.add_argument('--out', required=True)
    ap.add_argument('--res', type=int, default=5)
    ap.add_argument('--n', type=int, default=2)
    ap.add_argument('--seed', type=int, default=0)
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    import mxnet as mx
    from networks_stylegan import Generator            # the reference's own classes
    from networks_seg import Decoder
    import gan_segmentation_b200  # noqa: F401  (import shim of this repo; NumPy-only modules are used below)
    from parity_util import make_case

    gc, dc, gp, dp, z, noise = make_case(args.res, args.n, seed=args.seed)
    ctx = mx.cpu()
    gcfg = dict(gc, fix_noise=True, use_wscale=True)
    netG = Generator(gcfg)
    set_params(netG, gp, mx, ctx)
    for r in range(2, args.res + 1):                       # the AddNoise blocks: net{r}.block1[0] / net{r}.block2[1]
        blk = getattr(netG, f'net{r}')
        blk.block1[0]._noise = mx.nd.array(noise[2 * (r - 2)], ctx=ctx)
        blk.block2[1]._noise = mx.nd.array(noise[2 * (r - 2) + 1], ctx=ctx)
    img, feats = netG(mx.nd.array(z, ctx=ctx))
    out = {'img_f32': img.asnumpy()}
    for i, f in enumerate(feats):
        out[f'feat{i}'] = f.asnumpy()

    dcfg = dict(dc)
    netD = Decoder(dcfg, num_devices=1)
    set_params(netD, dp, mx, ctx)
    logits = netD(*feats)
    out['logits'] = logits.asnumpy()
    out['mask'] = mx.nd.argmax(logits, axis=1, keepdims=True).asnumpy().transpose(0, 2, 3, 1)     # seg_solver.py:326-327

    # one iteration of the reference's training loop body (seg_solver.py:386-421), dropout off
    tcfg = dict(dc, use_dropout=False)
    netT = Decoder(tcfg, num_devices=1)
    set_params(netT, dp, mx, ctx)
    loss_fn = mx.gluon.loss.SoftmaxCELoss(axis=1)
    trainer = mx.gluon.Trainer(netT.collect_params(), 'adam', {'learning_rate': 1e-3, 'wd': 0.0}, kvstore=None)
    rs = np.random.RandomState(args.seed + 9)
    h, w = out['logits'].shape[2:]
    mask = rs.randint(-1, dc['features'][-1], size=(args.n, 1, h, w)).astype(np.float32)
    mask_s = mx.nd.array(mask, ctx=ctx)
    feats_d = [mx.nd.array(f.asnumpy(), ctx=ctx) for f in feats]
    with mx.autograd.record():
        pred = netT(*[f.detach() for f in feats_d])
        l_ones = mx.nd.ones(mask_s.shape, ctx=ctx, dtype=np.float32)
        l_zeros = mx.nd.zeros(mask_s.shape, ctx=ctx, dtype=np.float32)
        l_w = 1.0 * mx.nd.ones(mask_s.shape, ctx=ctx, dtype=np.float32)
        sample_weight = mx.nd.where(mask_s > -1, l_ones, l_zeros)
        sample_weight = mx.nd.where(mask_s >= 0.5, l_w, sample_weight)
        err = loss_fn(pred, mask_s, sample_weight)
    err.backward()
    out['train_mask'] = mask
    out['train_loss'] = err.asnumpy()
    table = netT._collect_params_with_prefix()
    for name, p in table.items():
        if p.grad_req != 'null':
            out['grad:' + name] = p.grad(ctx).asnumpy()
    trainer.step(args.n)
    for name, p in table.items():
        out['new:' + name] = p.data(ctx).asnumpy()
    out['meta'] = np.array([args.res, args.n, args.seed], np.int64)
    out['mxnet_version'] = np.array(mx.__version__)
    np.savez_compressed(args.out, **out)
    print('written', args.out, 'with MXNet', mx.__version__)


if __name__ == '__main__':
    main()
