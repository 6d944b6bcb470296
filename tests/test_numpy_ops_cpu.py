"""The independent NumPy restatements of the five assumption-laden MXNet operators (oracle/numpy_ops.py) against the
PyTorch oracle's formulations: two restatements, no shared code."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import numpy_ops as NP
from oracle import train_oracle as T


def test_deconvolution_matches_conv_transpose_and_the_reference_shapes():
    rs = np.random.RandomState(0)
    x = rs.randn(2, 5, 6, 7).astype(np.float32)
    w = rs.randn(5, 3, 4, 4).astype(np.float32)
    a = NP.deconvolution(x, w, 2, 1)
    b = F.conv_transpose2d(torch.from_numpy(x), torch.from_numpy(w), None, stride=2, padding=1).numpy()
    assert a.shape == (2, 3, 12, 14) == b.shape             # (in-1)*2 - 2 + 4 = 2*in  (networks_stylegan.py:16)
    assert np.abs(a - b).max() < 1e-4


def test_instance_norm_biased_variance_eps():
    rs = np.random.RandomState(1)
    x = (rs.randn(3, 4, 5, 6) * 3 + 1).astype(np.float32)
    a = NP.instance_norm(x)
    b = F.instance_norm(torch.from_numpy(x), eps=1e-5).numpy()
    assert np.abs(a - b).max() < 1e-5
    # tiny plane: the eps and the biased variance are visible
    x2 = np.array([[[[0.0, 1e-3]]]], np.float32)
    assert np.abs(NP.instance_norm(x2) - F.instance_norm(torch.from_numpy(x2), eps=1e-5).numpy()).max() < 1e-6


def test_batch_norm_train_and_moving_stats():
    rs = np.random.RandomState(2)
    x = (rs.randn(3, 4, 5, 6) * 2 - 0.5).astype(np.float32)
    gamma, beta = rs.rand(4).astype(np.float32) + 0.5, rs.randn(4).astype(np.float32)
    mm, mv = rs.randn(4).astype(np.float32), rs.rand(4).astype(np.float32) + 0.5
    y, nm, nv = NP.batch_norm_train(x, gamma, beta, mm, mv)
    P = {'p.gamma': torch.from_numpy(gamma), 'p.beta': torch.from_numpy(beta), 'p.running_mean': torch.from_numpy(mm),
         'p.running_var': torch.from_numpy(mv)}
    stats = {}
    yt = T._bn_train(P, stats, 'p', torch.from_numpy(x)).numpy()
    assert np.abs(y - yt).max() < 1e-5
    assert np.abs(nm - stats['p.running_mean'].numpy()).max() < 1e-6
    assert np.abs(nv - stats['p.running_var'].numpy()).max() < 1e-6


def test_softmax_ce_loss_and_gradient():
    rs = np.random.RandomState(3)
    pred = rs.randn(2, 3, 4, 5).astype(np.float32) * 2
    mask = rs.randint(-1, 3, size=(2, 1, 4, 5))
    weight = (mask > -1).astype(np.float32)
    a = NP.softmax_ce(pred, np.maximum(mask, 0), weight)
    pt = torch.from_numpy(pred).requires_grad_(True)
    loss = T.softmax_ce(pt, torch.from_numpy(mask))
    assert np.abs(a - loss.detach().numpy()).max() < 1e-6
    loss.sum().backward()
    g = NP.softmax_ce_grad(pred, np.maximum(mask, 0), weight)
    assert np.abs(g - pt.grad.numpy()).max() < 1e-7


def test_adam_update_bias_correction_form():
    rs = np.random.RandomState(4)
    w, g = rs.randn(50), rs.randn(50)
    m, v = np.zeros(50), np.zeros(50)
    w2, m2, v2 = w.copy(), m.copy(), v.copy()
    for t in range(1, 4):
        w, m, v = NP.adam_update(w, g, m, v, t, lr=1e-3, wd=0.01, rescale_grad=0.5)
        w2, m2, v2 = T.adam_update(w2, g, m2, v2, t, 1e-3, 2, wd=0.01)
    assert np.abs(w - w2).max() < 1e-12
    # first step of Adam moves every weight by ~lr against the gradient sign (bias correction makes m/sqrt(v) = +-1)
    w1, _, _ = NP.adam_update(np.zeros(3), np.array([1.0, -2.0, 0.5]), np.zeros(3), np.zeros(3), 1, lr=1e-3)
    assert np.allclose(w1, [-1e-3, 1e-3, -1e-3], rtol=1e-4)
