"""ORACLE, second line of defence (test infrastructure only -- never imported by the product path).

Independent NumPy restatements of the five MXNet operators whose semantics the PyTorch oracle has to ASSUME
(MXNet 1.5 is not installable offline, so neither restatement is pinned against MXNet itself):
written from the operators' documented formulas with explicit index arithmetic, sharing no code with
``generate_oracle.py`` / ``train_oracle.py`` (which lean on torch.nn.functional).  ``tests/test_numpy_ops_cpu.py``
checks that the two restatements agree, so that a wrong assumption would have to be made twice, in two different
formulations, to go unnoticed.  PARITY UNPINNED, like the rest of ``oracle/``.

  deconvolution   mx.nd.Deconvolution(kernel, stride, pad, no_bias): weight (Cin, Cout, kh, kw); scatter form
                  out[n, co, s*i + ky - p, s*j + kx - p] += x[n, ci, i, j] * w[ci, co, ky, kx];
                  output size (in-1)*s - 2p + k                          (reference networks_stylegan.py:460-476)
  instance_norm   gluon.nn.InstanceNorm(center=False, scale=False), eps 1e-5: per (n, c) over H*W, biased variance
                                                                        (networks_stylegan.py:246, 261)
  batch_norm_train  gluon.nn.BatchNorm in training mode: per c over (N, H, W), biased variance, eps 1e-5,
                  moving <- 0.9*moving + 0.1*batch                       (networks_seg.py:16-21, 70-75)
  softmax_ce      gluon.loss.SoftmaxCrossEntropyLoss(axis=1, sparse_label=True)(pred, label, sample_weight):
                  -pick(log_softmax(pred, 1), label) * weight, then mean over every non-batch axis
                                                                        (seg_solver.py:54, 399-407)
  adam_update     mx.optimizer.Adam.update: bias correction folded into lr, grad*rescale_grad + wd*weight
                                                                        (seg_solver.py:51-58, 421)
"""
from __future__ import annotations

import numpy as np


def deconvolution(x, w, stride=2, pad=1):
    n, cin, h, wd = x.shape
    cin2, cout, kh, kw = w.shape
    assert cin == cin2
    ho, wo = (h - 1) * stride - 2 * pad + kh, (wd - 1) * stride - 2 * pad + kw
    full = np.zeros((n, cout, (h - 1) * stride + kh, (wd - 1) * stride + kw), np.float64)
    xd, wdbl = x.astype(np.float64), w.astype(np.float64)
    for ky in range(kh):
        for kx in range(kw):
            contrib = np.einsum('ncij,cd->ndij', xd, wdbl[:, :, ky, kx])
            full[:, :, ky:ky + (h - 1) * stride + 1:stride, kx:kx + (wd - 1) * stride + 1:stride] += contrib
    return full[:, :, pad:pad + ho, pad:pad + wo]


def instance_norm(x, eps=1e-5):
    xd = x.astype(np.float64)
    n, c, h, w = xd.shape
    flat = xd.reshape(n, c, h * w)
    mean = flat.sum(axis=2) / (h * w)
    var = ((flat - mean[:, :, None]) ** 2).sum(axis=2) / (h * w)         # biased
    return ((flat - mean[:, :, None]) / np.sqrt(var[:, :, None] + eps)).reshape(n, c, h, w)


def batch_norm_train(x, gamma, beta, moving_mean, moving_var, eps=1e-5, momentum=0.9):
    xd = x.astype(np.float64)
    n, c, h, w = xd.shape
    m = n * h * w
    mean = np.array([xd[:, k].sum() / m for k in range(c)])
    var = np.array([((xd[:, k] - mean[k]) ** 2).sum() / m for k in range(c)])     # biased
    y = np.empty_like(xd)
    for k in range(c):
        y[:, k] = (xd[:, k] - mean[k]) / np.sqrt(var[k] + eps) * gamma[k] + beta[k]
    return y, momentum * moving_mean + (1 - momentum) * mean, momentum * moving_var + (1 - momentum) * var


def softmax_ce(pred, label, weight):
    """pred [N,K,H,W], label [N,1,H,W] int (already clipped to >= 0 where weight is 0), weight [N,1,H,W]."""
    p = pred.astype(np.float64)
    n, k, h, w = p.shape
    mx_ = p.max(axis=1, keepdims=True)
    lse = mx_ + np.log(np.exp(p - mx_).sum(axis=1, keepdims=True))
    logsm = p - lse
    loss = np.zeros(n)
    for b in range(n):
        acc = 0.0
        for i in range(h):
            for j in range(w):
                lab = int(label[b, 0, i, j])
                lab = min(max(lab, 0), k - 1)                                # pick() clips the index
                acc += -logsm[b, lab, i, j] * float(weight[b, 0, i, j])
        loss[b] = acc / (h * w)                                              # mean over ALL pixels
    return loss


def softmax_ce_grad(pred, label, weight):
    """d(sum_n loss_n) / d pred of the loss above."""
    p = pred.astype(np.float64)
    n, k, h, w = p.shape
    e = np.exp(p - p.max(axis=1, keepdims=True))
    sm = e / e.sum(axis=1, keepdims=True)
    lab = np.clip(label[:, 0].astype(np.int64), 0, k - 1)
    onehot = np.zeros_like(sm)
    for c in range(k):
        onehot[:, c] = (lab == c)
    return (sm - onehot) * weight.astype(np.float64) / (h * w)


def adam_update(w, g, m, v, t, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0, rescale_grad=1.0):
    coef1, coef2 = 1.0 - beta1 ** t, 1.0 - beta2 ** t
    lr_t = lr * np.sqrt(coef2) / coef1
    g = g.astype(np.float64) * rescale_grad + wd * w
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    return w - lr_t * m / (np.sqrt(v) + eps), m, v
