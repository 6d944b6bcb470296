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
        # the scaled form the 16-bit backward pass uses: grad_scale = H*W -> w * (softmax - onehot)
        _, dls = softmax_ce(lg.detach(), lab.int(), grad_scale=float(h * w))
        assert torch.allclose(dls, lg.grad * (h * w), rtol=1e-4, atol=1e-6)


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


@pytest.mark.gpu
def test_bn_lrelu_and_upsample_kernels(gsx_lib):
    """Train-mode BatchNorm + LeakyReLU (+ dropout mask) forward / backward and the nearest-x2 pair against the
    hand-written torch formulas of the test backend (which the CPU suite checks against autograd)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from torch_backend import TorchBackend
    from gan_segmentation_b200.decoder_training import CudaBackend
    cb, tb = CudaBackend(), TorchBackend()
    g = torch.Generator().manual_seed(9)
    for (n, c, h, w, use_drop) in [(2, 32, 12, 20, False), (1, 16, 33, 17, True), (3, 512, 4, 4, True)]:
        z = torch.randn((n, c, h, w), generator=g) * 2 + 0.5
        dy = torch.randn((n, c, h, w), generator=g)
        gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
        drop = (torch.rand((n, c, h, w), generator=g) > 0.5).float() if use_drop else None
        y_ref, c_ref = tb.bn_lrelu_fwd(z, gamma, beta, drop)
        dz_ref, dg_ref, db_ref = tb.bn_lrelu_bwd(dy, z, c_ref, gamma, beta, drop)
        zc, dyc, gc_, bc = z.cuda(), dy.cuda(), gamma.cuda(), beta.cuda()
        dc = drop.cuda() if drop is not None else None
        y, cc = cb.bn_lrelu_fwd(zc, gc_, bc, dc)
        assert torch.allclose(cc['mean'].cpu(), c_ref['mean'], atol=1e-5) and torch.allclose(cc['var'].cpu(), c_ref['var'], rtol=1e-4, atol=1e-6)
        assert torch.allclose(y.cpu(), y_ref, rtol=1e-4, atol=1e-5)
        dz, dg, db = cb.bn_lrelu_bwd(dyc, zc, cc, gc_, bc, dc)
        assert torch.allclose(dg.cpu(), dg_ref, rtol=1e-4, atol=1e-3) and torch.allclose(db.cpu(), db_ref, rtol=1e-4, atol=1e-3)
        assert torch.allclose(dz.cpu(), dz_ref, rtol=1e-3, atol=1e-5)
        assert torch.equal(cb.upsample2(zc).cpu(), tb.upsample2(z))
        z2 = torch.randn((n, c, 2 * h, 2 * w), generator=g)
        assert torch.allclose(cb.sumpool2(z2.cuda()).cpu(), tb.sumpool2(z2), atol=1e-5)


@pytest.mark.gpu
def test_decoder_training_step_on_cuda_kernels(gsx_lib):
    """One and several steps of DecoderTrainer on the CUDA backend: loss and gradients agree with the same trainer on the
    fp32 torch test backend up to the 16-bit conv operands, and the loss goes down."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from torch_backend import TorchBackend
    from gan_segmentation_b200.config import decoder_config
    from gan_segmentation_b200.decoder_training import CudaBackend, DecoderTrainer
    from gan_segmentation_b200.random_init import init_decoder_params
    res, n = 5, 2
    cfg = dict(decoder_config(res), use_dropout=True, base_lr=2e-3)
    params = init_decoder_params(cfg, seed=2)
    rs = np.random.RandomState(0)
    feats = [rs.randn(n, c, 4 << i, 4 << i).astype(np.float32) for i, c in enumerate(cfg['in_channels'][:res - 1])]
    mask = rs.randint(-1, 2, (n, 1, 32, 32))
    drops = [(rs.rand(n, cfg['features'][i], 4 << i, 4 << i) > 0.5).astype(np.float32) for i in range(res - 1)]
    ref = DecoderTrainer(cfg, params, TorchBackend())
    loss_ref, g_ref = ref.loss_and_grads(feats, mask, drops)
    tr = DecoderTrainer(cfg, params, CudaBackend())
    loss, grads = tr.loss_and_grads(feats, mask, drops)
    assert tr._grad_scale == 32 * 32 and ref._grad_scale == 1.0
    grads = {k: v / tr._grad_scale for k, v in grads.items()}
    assert np.allclose(loss.cpu().numpy(), loss_ref.numpy(), rtol=2e-2, atol=1e-3)
    # (the bias of a conv that feeds a BatchNorm has an exactly-zero gradient: compare on an absolute scale there)
    gscale = max(float(np.abs(v.numpy()).max()) for v in g_ref.values())
    for k in g_ref:
        a, b = grads[k].cpu().numpy(), g_ref[k].numpy()
        err = np.abs(a - b).max() / max(float(np.abs(b).max()), 2e-3 * gscale)
        assert err < 6e-2, (k, err, float(np.abs(b).max()), gscale)
    losses = [float(loss.mean())]
    for _ in range(8):
        losses.append(float(tr.step(feats, mask, drops).mean()))
    assert losses[-1] < 0.85 * losses[0], losses
    st = tr.state()
    assert set(st) == set(params) and all(st[k].shape == np.asarray(params[k]).shape for k in params)


@pytest.mark.gpu
def test_training_step_at_ffhq_size_matches_the_oracle(gsx_lib):
    """BASELINE config 4's own size: one decoder-training step at max_res_log2 = 10 (1024^2 logits), batch 1, on the CUDA
    backend against oracle/train_oracle.py (fp32 autograd).  At this size the loss gradient is 1/(H*W) = 9.5e-7 per
    pixel -- below the fp16 normal range -- so the step only works with the gradient scaling of gsx_softmax_ce; every
    parameter's gradient must agree with the oracle to a relative error (L2) of 3e-2, no entry off by more than 6e-2 of
    the tensor's largest entry.  Measured on B200 (r02): worst tensor 2.5e-2 (a BatchNorm beta at the 16^2 level, nine
    16-bit conv levels below the loss), 1e-2 or better from level 6 upwards; the error is set by the 16-bit operands of
    the FORWARD pass (the logits differ from the fp32 oracle's by ~3e-3 relative, LeakyReLU signs flip near zero), not by
    gradient underflow: without the scaling the same test is off by tens of percent."""
    from oracle import train_oracle as T
    from gan_segmentation_b200.config import decoder_config
    from gan_segmentation_b200.decoder_training import CudaBackend, DecoderTrainer
    from gan_segmentation_b200.random_init import init_decoder_params
    res, n = 10, 1
    cfg = dict(decoder_config(res), use_dropout=False, base_lr=1e-4)
    params = init_decoder_params(cfg, seed=2)
    rs = np.random.RandomState(5)
    feats = [rs.randn(n, c, 4 << i, 4 << i).astype(np.float32) for i, c in enumerate(cfg['in_channels'])]
    yy, xx = np.mgrid[0:1024, 0:1024]
    rr = np.hypot(yy - 500, xx - 540)
    mask = np.where(rr < 260, 1, np.where(rr < 420, 0, -1)).astype(np.int64)[None, None]      # disk / ring / ignore
    p_ref, st, loss_ref, g_ref = T.train_step(params, cfg, feats, mask)
    tr = DecoderTrainer(cfg, params, CudaBackend())
    loss, grads = tr.loss_and_grads(feats, mask, None)
    assert tr._grad_scale == 1024 * 1024
    assert abs(float(loss[0]) - float(loss_ref[0])) < 2e-2 * abs(float(loss_ref[0]))
    gmax = max(float(np.abs(g).max()) for g in g_ref.values())
    rows = []
    for k, b in g_ref.items():
        a = grads[k].cpu().numpy().astype(np.float64) / tr._grad_scale
        # relative error of the tensor: ||a - b|| / ||b||  (the bias of a conv that feeds a BatchNorm has an exactly-zero
        # gradient: absolute floor there), and the largest entry-wise deviation relative to the largest entry
        floor = 1e-3 * gmax
        l2 = float(np.linalg.norm(a - b)) / max(float(np.linalg.norm(b)), floor * np.sqrt(b.size))
        mx = float(np.abs(a - b).max()) / max(float(np.abs(b).max()), floor)
        rows.append((l2, mx, k))
    rows.sort(reverse=True)
    print('largest gradient errors (relative L2, relative max):', [(k, round(l2, 4), round(mx, 4)) for l2, mx, k in rows[:12]])
    for l2, mx, k in rows:
        assert l2 <= 3e-2, (k, l2, mx)
        assert mx <= 6e-2, (k, l2, mx)
    tr.step(feats, mask, None)
    new = tr.state()
    for k in ('main_block_8.0.weight', 'cvt_block_8.0.weight', 'cvt_block_0.0.weight', 'main_block_3.1.base_layers.1.gamma'):
        # Adam's first step moves every weight by lr * g/(|g| + eps): the sign pattern is the check
        moved = new[k] - np.asarray(params[k], np.float32)
        ref_moved = p_ref[k] - np.asarray(params[k], np.float32)
        big = np.abs(g_ref[k]) > 0.05 * np.abs(g_ref[k]).max()
        assert np.mean(np.sign(moved[big]) == np.sign(ref_moved[big])) > 0.999, k


@pytest.mark.gpu
@pytest.mark.parametrize('k,n,h,w,cin,cout', [(3, 2, 37, 53, 16, 16), (3, 1, 64, 200, 64, 32), (1, 2, 32, 32, 64, 32),
                                               (3, 2, 16, 16, 256, 32), (3, 1, 130, 300, 32, 2), (3, 3, 8, 8, 512, 32),
                                               (3, 1, 4, 4, 32, 32), (3, 1, 256, 256, 16, 16)])
def test_tensor_core_wgrad_matches_autograd(gsx_lib, k, n, h, w, cin, cout):
    """csrc/wgrad.cu (tcgen05, pixels as the GEMM K dimension, MN-major operands straight from the blocked layout) against
    autograd on the same 16-bit-rounded operands; ragged tiles, 1x1, many-channel and 2-class (final conv) shapes."""
    import ctypes as C
    from gan_segmentation_b200 import _lib as L
    lib = L.lib()
    g = torch.Generator().manual_seed(k * 1000 + h)
    x = torch.randn((n, cin, h, w), generator=g).half().float().cuda().requires_grad_(False)
    dy = torch.randn((n, cout, h, w), generator=g).half().float().cuda()
    wt = torch.zeros((cout, cin, k, k), device='cuda', requires_grad=True)
    F.conv2d(x, wt, None, 1, k // 2).backward(dy)
    dw = torch.empty((cout, cin, k, k), device='cuda')
    rc = lib.gsx_op_conv_wgrad_tc(k, n, h, w, cin, cout, L.ptr(x), L.ptr(dy), L.ptr(dw), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    L.check(rc, 'gsx_op_conv_wgrad_tc')
    ref = wt.grad
    err = float((dw - ref).abs().max() / ref.abs().max())
    assert err < 2e-3, err


def _resident_case(res, n, use_dropout, seed=3):
    from gan_segmentation_b200.config import decoder_config
    from gan_segmentation_b200.random_init import init_decoder_params
    cfg = dict(decoder_config(res), use_dropout=use_dropout, base_lr=1e-3)
    params = init_decoder_params(cfg, seed=2)
    rs = np.random.RandomState(seed)
    feats = [rs.randn(n, c, 4 << i, 4 << i).astype(np.float32) for i, c in enumerate(cfg['in_channels'])]
    hw = 4 << (res - 2)
    mask = rs.randint(-1, 2, (n, 1, hw, hw)).astype(np.int64)
    return cfg, params, feats, mask


def _grad_errors(grads, g_ref):
    gmax = max(float(np.abs(g).max()) for g in g_ref.values())
    rows = []
    for k, b in g_ref.items():
        a = grads[k].astype(np.float64)
        floor = 1e-3 * gmax
        l2 = float(np.linalg.norm(a - b)) / max(float(np.linalg.norm(b)), floor * np.sqrt(b.size))
        rows.append((l2, k))
    return sorted(rows, reverse=True)


@pytest.mark.gpu
@pytest.mark.parametrize('res,n,use_dropout', [(5, 2, False), (6, 1, True), (7, 3, True)])
def test_resident_train_step_matches_the_oracle(gsx_lib, res, n, use_dropout):
    """gsx_train_step (one C-ABI call: forward, loss, backward on resident blocked 16-bit tensors) against
    oracle/train_oracle.py with the same dropout masks (exported Philox bits): loss, every gradient tensor, the moving
    statistics, and the parameters after the Adam step."""
    from oracle import train_oracle as T
    from gan_segmentation_b200.decoder_training import ResidentTrainer
    cfg, params, feats, mask = _resident_case(res, n, use_dropout)
    tr = ResidentTrainer(cfg, params, n)
    seed = 1234
    drops = [tr.dropout_mask(i, seed).cpu() for i in range(res - 1)] if use_dropout else None
    if use_dropout:
        assert 0.4 < float(drops[-1].mean()) < 0.6 and set(np.unique(drops[0].numpy())) <= {0.0, 1.0}
    p_ref, st, loss_ref, g_ref = T.train_step(params, cfg, feats, mask, dropout_masks=drops)
    loss = tr.step(feats, mask, dropout_seed=seed)
    torch.cuda.synchronize()
    assert np.allclose(loss.cpu().numpy(), loss_ref, rtol=2e-2, atol=1e-3), (loss, loss_ref)
    rows = _grad_errors(tr.grads(), g_ref)
    print('largest relative gradient errors:', [(k, round(e, 4)) for e, k in rows[:6]])
    for e, k in rows:
        assert e <= 4e-2, (k, e)
    new = tr.state()
    for k in p_ref:
        if k.endswith(('running_mean', 'running_var')):
            assert np.allclose(new[k], p_ref[k], rtol=2e-2, atol=2e-3), k
    # the first Adam step moves a weight by ~lr * sign(g): compare the signs where the gradient is not tiny
    for k in ('main_block_%d.0.weight' % (res - 2), 'cvt_block_1.0.weight', 'main_block_1.1.base_layers.3.weight'):
        big = np.abs(g_ref[k]) > 0.1 * np.abs(g_ref[k]).max()
        a = np.sign(new[k] - np.asarray(params[k], np.float32))[big]
        b = np.sign(p_ref[k] - np.asarray(params[k], np.float32))[big]
        assert np.mean(a == b) > 0.995, k


@pytest.mark.gpu
def test_resident_training_reduces_the_loss_and_is_reproducible(gsx_lib):
    from gan_segmentation_b200.decoder_training import ResidentTrainer
    cfg, params, feats, mask = _resident_case(6, 2, True)
    runs = []
    for _ in range(2):
        tr = ResidentTrainer(cfg, params, 2, base_lr=2e-3)
        losses = [float(tr.step(feats, mask, dropout_seed=7 + i).mean()) for i in range(10)]
        runs.append((losses, tr.state()))
    assert runs[0][0][-1] < 0.85 * runs[0][0][0], runs[0][0]
    assert runs[0][0] == runs[1][0]                                         # no atomics anywhere: bit-reproducible
    assert all(np.array_equal(runs[0][1][k], runs[1][1][k]) for k in runs[0][1])


@pytest.mark.gpu
def test_resident_train_step_at_ffhq_size(gsx_lib):
    """BASELINE config 4's own size for the product's training step: max_res_log2 = 10 (1024^2 logits), batch 1, against the
    fp32 autograd oracle; same bounds as the hook-based step above (relative L2 error of every gradient tensor)."""
    from oracle import train_oracle as T
    from gan_segmentation_b200.config import decoder_config
    from gan_segmentation_b200.decoder_training import ResidentTrainer
    from gan_segmentation_b200.random_init import init_decoder_params
    cfg = dict(decoder_config(10), use_dropout=False, base_lr=1e-4)
    params = init_decoder_params(cfg, seed=2)
    rs = np.random.RandomState(5)
    feats = [rs.randn(1, c, 4 << i, 4 << i).astype(np.float32) for i, c in enumerate(cfg['in_channels'])]
    yy, xx = np.mgrid[0:1024, 0:1024]
    rr = np.hypot(yy - 500, xx - 540)
    mask = np.where(rr < 260, 1, np.where(rr < 420, 0, -1)).astype(np.int64)[None, None]
    p_ref, st, loss_ref, g_ref = T.train_step(params, cfg, feats, mask)
    tr = ResidentTrainer(cfg, params, 1)
    loss = tr.step(feats, mask)
    torch.cuda.synchronize()
    assert tr.grad_scale == 1024 * 1024
    assert abs(float(loss[0]) - float(loss_ref[0])) < 2e-2 * abs(float(loss_ref[0]))
    rows = _grad_errors(tr.grads(), g_ref)
    print('largest relative gradient errors:', [(k, round(e, 4)) for e, k in rows[:6]])
    for e, k in rows:
        assert e <= 4e-2, (k, e)
    pred = tr.pred.cpu().numpy()
    assert pred.shape == (1, 1024, 1024) and set(np.unique(pred)) <= {0, 1}


def _nccl_worker(rank, world, port, q):
    import os
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from gan_segmentation_b200.decoder_training import ResidentTrainer
    cfg, params, feats, mask = _resident_case(6, 2, False)
    tr = ResidentTrainer(cfg, params, 1, device=f'cuda:{rank}', base_lr=1e-3)
    tr.step([f[rank:rank + 1] for f in feats], mask[rank:rank + 1], global_batch=world)
    torch.cuda.synchronize()
    st = tr.state()
    q.put((rank, {k: st[k] for k in ('cvt_block_0.0.weight', 'main_block_4.0.bias', 'main_block_1.1.base_layers.1.gamma')}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_data_parallel_resident_step_two_gpus_nccl(gsx_lib):
    """Two ranks, one sample each, ONE NCCL all-reduce of the flat gradient bucket, identical Adam step on both ranks: the
    weights agree bit for bit across the ranks and match the single-rank batch-2 step (same summed gradient, 1/2 rescale;
    BatchNorm statistics stay per rank, use_sync_bn = False in the reference, so only near-equality there)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, 29731, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for k in got[0]:
        assert np.array_equal(got[0][k], got[1][k]), k
    from gan_segmentation_b200.decoder_training import ResidentTrainer
    cfg, params, feats, mask = _resident_case(6, 2, False)
    one = ResidentTrainer(cfg, params, 2, base_lr=1e-3)
    one.step(feats, mask)
    st = one.state()
    for k in got[0]:
        assert np.abs(got[0][k] - st[k]).max() < 2.5e-3, k            # first Adam step: |dw| = lr, sign flips where g ~ 0
