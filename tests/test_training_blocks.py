"""Decoder-training building blocks: loss + flat-bucket all-reduce / Adam."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.nn.functional as F


@pytest.mark.gpu
def test_softmax_ce_matches_reference_semantics(gsx_lib):
    """SoftmaxCELoss(axis=1) with sample_weight=(mask>-1): per-sample mean over ALL pixels (ignored ones in the
    denominator), gradient of the summed loss."""
    from gan_segmentation_b200.training import softmax_ce
    g = torch.Generator().manual_seed(0)
    for k in (2, 5):
        n, h, w = 3, 37, 53
        lg = (torch.randn((n, k, h, w), generator=g) * 3).cuda().requires_grad_(True)
        lab = torch.randint(-1, k, (n, 1, h, w), generator=g).cuda()
        wgt = (lab > -1).float()
        lp = F.log_softmax(lg, dim=1)
        picked = -torch.gather(lp, 1, lab.clamp(min=0)) * wgt
        ref = picked.mean(dim=(1, 2, 3))
        ref.sum().backward()
        loss, dl = softmax_ce(lg.detach(), lab.int())
        assert torch.allclose(loss, ref.detach(), rtol=1e-5, atol=1e-6)
        assert torch.allclose(dl, lg.grad, rtol=1e-4, atol=1e-8)


@pytest.mark.gpu
def test_flat_adam_matches_mxnet_formula(gsx_lib):
    from gan_segmentation_b200.training import FlatAdam
    shapes = {'a.weight': (7, 5, 3, 3), 'a.bias': (7,), 'b.gamma': (11,)}
    opt = FlatAdam(shapes, lr=1e-2, wd=1e-3)
    rs = np.random.RandomState(0)
    params = {k: rs.randn(*s).astype(np.float32) for k, s in shapes.items()}
    opt.load(params)
    w = np.concatenate([params[k].ravel() for k in shapes]).astype(np.float64)
    m = np.zeros_like(w); v = np.zeros_like(w)
    for t in range(1, 4):
        g = rs.randn(w.size).astype(np.float32)
        opt.g.copy_(torch.from_numpy(g))
        opt.step(global_batch=4)
        gg = g.astype(np.float64) / 4 + 1e-3 * w
        m = 0.9 * m + 0.1 * gg
        v = 0.999 * v + 0.001 * gg * gg
        lr_t = 1e-2 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        w = w - lr_t * m / (np.sqrt(v) + 1e-8)
    torch.cuda.synchronize()
    assert np.allclose(opt.w.cpu().numpy(), w, rtol=2e-5, atol=1e-6)
    assert opt.state()['a.bias'].shape == (7,)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from gan_segmentation_b200.training import FlatAdam
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    opt = FlatAdam({'w': (4, 3), 'b': (5,)}, device='cpu')
    opt.g.copy_(torch.arange(opt.count, dtype=torch.float32) * (rank + 1))
    opt.allreduce_grads()                         # the only collective of a training step
    if rank == 0:
        q.put(opt.g.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_gloo_two_ranks():
    """world_size-2 on CPU: one all-reduce over the flat gradient bucket sums the ranks' gradients."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    g = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(g, np.arange(17, dtype=np.float32) * 3)


@pytest.mark.gpu
def test_conv_dgrad_runs_on_the_forward_kernel(gsx_lib):
    """dX of a 3x3 / 1x1 'same' conv == the forward shift-GEMM kernel applied to dY with transposed, flipped weights
    (training.dgrad_weights), checked against autograd."""
    from gan_segmentation_b200 import _lib as L, ops
    from gan_segmentation_b200.training import dgrad_weights
    g = torch.Generator().manual_seed(3)
    for (mode, k, n, cin, cout, h, w) in [(L.CONV3, 3, 2, 32, 16, 40, 56), (L.CONV3, 3, 1, 64, 64, 16, 16), (L.CONV1, 1, 2, 64, 16, 24, 24)]:
        x = torch.randn((n, cin, h, w), generator=g).cuda().requires_grad_(True)
        wt = (torch.randn((cout, cin, k, k), generator=g) / np.sqrt(cin * k * k)).half().float()
        dy = torch.randn((n, cout, h, w), generator=g).half().float().cuda()
        y = F.conv2d(x, wt.cuda(), None, 1, k // 2)
        y.backward(dy)
        r = ops.conv(mode, dy, dgrad_weights(wt.numpy()))
        ref = x.grad
        err = (r['out'] - ref).abs().max().item()
        assert err <= 2e-3 * ref.abs().max().item() + 1e-3, (mode, err)


@pytest.mark.gpu
def test_conv_wgrad_matches_autograd(gsx_lib):
    """dW / db of the decoder's convs (3x3 and the 1x1 shortcut) against autograd on the same fp16-rounded operands;
    bit-reproducible (per-tile partials summed in a fixed order)."""
    from gan_segmentation_b200.training import conv_wgrad
    g = torch.Generator().manual_seed(4)
    for (k, n, cin, cout, h, w) in [(3, 2, 16, 16, 40, 72), (3, 1, 64, 32, 16, 16), (1, 2, 64, 16, 24, 40), (3, 3, 8, 8, 5, 7)]:
        x = torch.randn((n, cin, h, w), generator=g).half().float().cuda()
        dy = torch.randn((n, cout, h, w), generator=g).half().float().cuda()
        wt = torch.zeros((cout, cin, k, k), device='cuda', requires_grad=True)
        b = torch.zeros((cout,), device='cuda', requires_grad=True)
        F.conv2d(x, wt, b, 1, k // 2).backward(dy)
        dw, db = conv_wgrad(x, dy, k)
        scale = wt.grad.abs().max().item()
        assert (dw - wt.grad).abs().max().item() <= 2e-4 * scale + 1e-3, (k, n, cin, cout, h, w)
        assert torch.allclose(db, b.grad, rtol=1e-4, atol=1e-2)
        dw2, db2 = conv_wgrad(x, dy, k)
        assert torch.equal(dw, dw2) and torch.equal(db, db2)
