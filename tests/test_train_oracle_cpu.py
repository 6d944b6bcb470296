"""Known-answer tests of the training-step oracle (oracle/train_oracle.py; seg_solver.py:351-421) -- the checker the
decoder-training path will be held against.  CPU only."""
import numpy as np
import torch

from gan_segmentation_b200.config import decoder_config
from gan_segmentation_b200.random_init import init_decoder_params
from oracle import train_oracle as T
from oracle.generate_oracle import decoder_forward


def _case(res=4, n=2, seed=0):
    cfg = decoder_config(res)
    params = init_decoder_params(cfg, seed=2)
    rs = np.random.RandomState(seed)
    feats = [rs.randn(n, c, 4 << i, 4 << i).astype(np.float32) for i, c in enumerate(cfg['in_channels'][:res - 1])]
    h = 4 << (res - 2)
    mask = rs.randint(-1, cfg['num_classes'], (n, 1, h, h))
    return cfg, params, feats, mask


def test_loss_weighting_known_answer():
    logits = torch.tensor([[[[2.0, 0.0]], [[0.0, 0.0]]]])                 # [1,2,1,2]
    mask = torch.tensor([[[[0, -1]]]])
    # pixel 0: -log softmax([2,0])[0] = log(1+e^-2); pixel 1 ignored but stays in the denominator (mean over 2)
    want = np.log1p(np.exp(-2.0)) / 2
    assert abs(T.softmax_ce(logits, mask).item() - want) < 1e-6


def test_train_forward_matches_inference_forward_when_stats_agree():
    """With running stats set to the batch statistics and no dropout, the train-mode forward equals the inference
    forward of generate_oracle (consistency of the two restatements)."""
    cfg, params, feats, mask = _case()
    cfg = dict(cfg, use_dropout=False)
    P = {k: torch.tensor(v) for k, v in params.items()}
    logits, stats = T.decoder_forward_train(P, cfg, [torch.tensor(f) for f in feats])
    p2 = dict(params)
    for k, v in stats.items():      # stats = .9*old + .1*batch  ->  batch = (stats - .9*old)/.1
        p2[k] = ((v.numpy() - 0.9 * params[k]) / 0.1).astype(np.float32)
    ref = decoder_forward(p2, cfg, feats)
    assert torch.allclose(logits, ref, rtol=1e-3, atol=1e-3)


def test_adam_step_formula_and_loss_decreases():
    cfg, params, feats, mask = _case()
    cfg = dict(cfg, use_dropout=False, base_lr=1e-2)
    p, st, loss0, grads = T.train_step(params, cfg, feats, mask)
    # first Adam step: m = .1 g, v = .001 g^2, lr_1 = lr*sqrt(.001)/.1  ->  |dw| = lr * |g|/(|g| + eps') ~ lr
    k = 'main_block_2.0.weight'
    g = grads[k] / mask.shape[0]
    step = params[k].astype(np.float64) - p[k]
    want = 1e-2 * np.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-8)
    assert np.allclose(step, want, rtol=1e-4, atol=1e-9)
    assert st['t'] == 1 and set(st['m']) == set(grads)
    losses = [loss0.mean()]
    for _ in range(6):
        p, st, l, _ = T.train_step(p, cfg, feats, mask, st)
        losses.append(l.mean())
    assert losses[-1] < 0.8 * losses[0], losses
    # running statistics moved towards the batch statistics, learnables changed, shapes kept
    assert all(p[q].shape == np.asarray(params[q]).shape for q in params)
    assert not np.allclose(p['cvt_block_0.1.running_mean'], params['cvt_block_0.1.running_mean'])


def test_gradient_sum_over_contexts_equals_full_batch_when_bn_is_per_context():
    """The reference splits the batch over contexts, back-propagates each shard and lets the KVStore sum the gradients
    (seg_solver.py:386-421); with use_sync_bn=False every context normalises with its own shard statistics.  The oracle's
    allreduce hook reproduces that: sum of shard gradients == what train_step applies."""
    cfg, params, feats, mask = _case(n=2)
    cfg = dict(cfg, use_dropout=False)
    shard_grads = []
    for i in range(2):
        _, _, _, g = T.train_step(params, cfg, [f[i:i + 1] for f in feats], mask[i:i + 1])
        shard_grads.append(g)
    summed = {k: shard_grads[0][k] + shard_grads[1][k] for k in shard_grads[0]}
    pa, _, _, _ = T.train_step(params, cfg, [f[0:1] for f in feats], mask[0:1], allreduce=lambda g: summed)
    k = 'cvt_block_0.0.weight'
    w, _, _ = T.adam_update(params[k].astype(np.float64), summed[k], 0.0, 0.0, 1, cfg['base_lr'], 1)
    assert np.allclose(pa[k], w.astype(np.float32))
