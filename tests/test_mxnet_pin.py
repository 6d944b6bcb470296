"""Pins the oracle against outputs of the REAL reference, when somebody has produced them.

``oracle/dump_from_mxnet.py`` (run wherever MXNet 1.5 and the reference checkout exist) writes
``tests/golden/mxnet_res{R}_seed{S}.npz``; these tests then hold both oracles (generate + train) to it with fp32
tolerances and "parity unpinned" can be struck from oracle/.  Offline (no dump present) they skip."""
import glob
import os

import numpy as np
import pytest
import torch

from parity_util import make_case

DUMPS = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'mxnet_res*_seed*.npz')))
pytestmark = pytest.mark.skipif(not DUMPS, reason='no MXNet dump (tests/golden/mxnet_*.npz): run oracle/dump_from_mxnet.py '
                                                  'where MXNet 1.5 is installed')


@pytest.mark.parametrize('path', DUMPS or ['-'])
def test_generate_oracle_matches_the_mxnet_reference(path):
    from oracle import generate_oracle as O
    d = np.load(path)
    res, n, seed = (int(v) for v in d['meta'])
    gc, dc, gp, dp, z, noise = make_case(res, n, seed=seed)
    ref = O.generate(gp, gc, dp, dc, z, noise)
    assert np.abs(ref['img_f32'] - d['img_f32']).max() < 1e-4
    for i, f in enumerate(ref['features']):
        assert np.abs(f - d[f'feat{i}']).max() < 1e-3 * max(1.0, np.abs(d[f'feat{i}']).max())
    assert np.abs(ref['logits'] - d['logits']).max() < 1e-3 * max(1.0, np.abs(d['logits']).max())
    assert (ref['mask'] == d['mask']).mean() > 0.9999


@pytest.mark.parametrize('path', DUMPS or ['-'])
def test_train_oracle_matches_the_mxnet_reference(path):
    from oracle import generate_oracle as O
    from oracle import train_oracle as T
    d = np.load(path)
    res, n, seed = (int(v) for v in d['meta'])
    gc, dc, gp, dp, z, noise = make_case(res, n, seed=seed)
    with torch.no_grad():
        _, feats = O.generator_forward(gp, gc, z, noise)
    cfg = dict(dc, use_dropout=False, base_lr=1e-3, wd=0.0)
    new, st, loss, grads = T.train_step(dp, cfg, [f.numpy() for f in feats], d['train_mask'].astype(np.int64))
    assert np.abs(loss - d['train_loss']).max() < 1e-5
    for k, g in grads.items():
        ref = d['grad:' + k]
        assert np.abs(g - ref).max() < 1e-4 * max(1e-3, np.abs(ref).max()) + 1e-7, k
    for k, v in new.items():
        assert np.abs(v - d['new:' + k]).max() < 1e-5, k
