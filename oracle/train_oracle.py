"""ORACLE (test infrastructure only -- never imported by the product path): CPU restatement of ONE decoder-training step
of the reference (seg_solver.py:351-421, generator frozen) in PyTorch fp32 with autograd.  PARITY UNPINNED: MXNet is
not installable offline, no golden vectors exist; the restatement follows the reference source line by line.

  forward   Decoder.hybrid_forward in train mode (networks_seg.py:64-113): BatchNorm with batch statistics
            (eps 1e-5, moving stats <- 0.9*moving + 0.1*batch, MXNet keeps the BIASED batch variance there),
            LeakyReLU(0.2), Dropout(0.5) after every cvt block (:77-78; masks passed in explicitly, scale 2)
  loss      SoftmaxCELoss(axis=1)(pred, mask, sample_weight) with sample_weight = (mask > -1) (:399-404; the second
            `where` with l_w == 1 is a no-op): per sample, mean over ALL pixels of -w * log_softmax(pred)[mask]
  backward  err.backward() on the per-sample loss vector == gradient of the SUM over the batch (:411-412)
  update    trainer.step(batch) (:421): Adam(lr=base_lr, beta1=.9, beta2=.999, eps=1e-8, wd=cfg['wd']),
            rescale_grad = 1/batch; MXNet folds the bias correction into the step size:
            lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t); g = grad*rescale + wd*w; m,v EMA; w -= lr_t*m/(sqrt(v)+eps)
The multi-GPU form of the reference sums the per-context gradients through KVStore('nccl') (:55-56) before the same
update: `allreduce` below is the hook for that (identity on one rank).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

LEARNABLE_SUFFIXES = ('.weight', '.bias', '.gamma', '.beta')


def _bn_train(P, new_stats, prefix, x, eps=1e-5, momentum=0.9):
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    with torch.no_grad():
        new_stats[f'{prefix}.running_mean'] = momentum * P[f'{prefix}.running_mean'] + (1 - momentum) * mean
        new_stats[f'{prefix}.running_var'] = momentum * P[f'{prefix}.running_var'] + (1 - momentum) * var
    xh = (x - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + eps)
    return xh * P[f'{prefix}.gamma'][None, :, None, None] + P[f'{prefix}.beta'][None, :, None, None]


def decoder_forward_train(P, cfg, feats, dropout_masks=None):
    """P: dict of torch tensors (learnable ones with requires_grad).  dropout_masks[i]: {0,1} tensor shaped like the
    cvt_i output, or None for no dropout.  Returns (logits, new_running_stats)."""
    nf = len(cfg['in_channels'])
    s0 = cfg['start_res']
    use_bn = cfg['use_bn']
    stats = {}
    prev = None
    for i in range(s0, nf):
        x = F.conv2d(feats[i], P[f'cvt_block_{i}.0.weight'], P[f'cvt_block_{i}.0.bias'], 1, 1)
        if use_bn:
            x = _bn_train(P, stats, f'cvt_block_{i}.1', x)
        x = F.leaky_relu(x, 0.2)
        if cfg.get('use_dropout', False) and dropout_masks is not None and dropout_masks[i] is not None:
            x = x * dropout_masks[i] * 2.0                                # Dropout(0.5), networks_seg.py:77-78
        if i > s0:
            x = torch.cat([prev, x], dim=1)
        if i < nf - 1:
            p = f'main_block_{i}.1'
            x = F.interpolate(x, scale_factor=2, mode='nearest')
            j = 0
            y = F.conv2d(x, P[f'{p}.base_layers.{j}.weight'], P[f'{p}.base_layers.{j}.bias'], 1, 1)
            j += 1
            if use_bn:
                y = _bn_train(P, stats, f'{p}.base_layers.{j}', y)
                j += 1
            y = F.leaky_relu(y, 0.2)
            j += 1
            y = F.conv2d(y, P[f'{p}.base_layers.{j}.weight'], P[f'{p}.base_layers.{j}.bias'], 1, 1)
            j += 1
            if use_bn:
                y = _bn_train(P, stats, f'{p}.base_layers.{j}', y)
            y = F.leaky_relu(y, 0.2)
            sc = F.conv2d(x, P[f'{p}.shortcut.0.weight'], P[f'{p}.shortcut.0.bias']) if f'{p}.shortcut.0.weight' in P else x
            prev = sc + y
        else:
            prev = F.conv2d(x, P[f'main_block_{i}.0.weight'], P[f'main_block_{i}.0.bias'], 1, 1)
    return prev, stats


def softmax_ce(logits, mask):
    """seg_solver.py:399-406.  logits [N,K,H,W], mask [N,1,H,W] int in {-1,0..K-1} -> per-sample loss [N]."""
    w = (mask > -1).to(logits.dtype)
    lp = F.log_softmax(logits, dim=1)
    picked = -torch.gather(lp, 1, mask.clamp(min=0).long()) * w
    return picked.mean(dim=(1, 2, 3))


def adam_update(w, g, m, v, t, lr, batch, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0):
    """MXNet Adam with rescale_grad = 1/batch (numpy, float64 internally); returns (w, m, v)."""
    g = g.astype(np.float64) / batch + wd * w.astype(np.float64)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return (w - lr_t * m / (np.sqrt(v) + eps)), m, v


def train_step(params, cfg, feats, mask, state=None, dropout_masks=None, allreduce=None):
    """One fit-loop iteration.  params: name -> numpy; state: {'t', 'm', 'v'} or None.  Returns
    (new_params, new_state, per-sample loss numpy, grads dict)."""
    P = {}
    for k, v in params.items():
        t = torch.tensor(np.asarray(v, np.float32))
        if k.endswith(LEARNABLE_SUFFIXES):
            t.requires_grad_(True)
        P[k] = t
    feats = [torch.tensor(np.asarray(f, np.float32)) for f in feats]     # detached: generator frozen (:393)
    mask_t = torch.tensor(np.asarray(mask)).long()
    logits, stats = decoder_forward_train(P, cfg, feats, dropout_masks)
    loss = softmax_ce(logits, mask_t)
    loss.sum().backward()                                                 # :411-412
    grads = {k: P[k].grad.numpy().astype(np.float64) for k in P if P[k].requires_grad and P[k].grad is not None}
    if allreduce is not None:
        grads = allreduce(grads)
    state = state or {'t': 0, 'm': {k: np.zeros_like(g) for k, g in grads.items()}, 'v': {k: np.zeros_like(g) for k, g in grads.items()}}
    t = state['t'] + 1
    new_params = {k: np.asarray(v, np.float32).copy() for k, v in params.items()}
    new_state = {'t': t, 'm': {}, 'v': {}}
    batch = mask_t.shape[0]
    for k, g in grads.items():
        w, m, v = adam_update(np.asarray(params[k], np.float64), g, state['m'][k], state['v'][k], t, cfg['base_lr'], batch,
                              wd=cfg.get('wd', 0.0) or 0.0)
        new_params[k] = w.astype(np.float32)
        new_state['m'][k], new_state['v'][k] = m, v
    for k, v in stats.items():
        new_params[k] = v.numpy().astype(np.float32)
    return new_params, new_state, loss.detach().numpy(), grads
