"""DecoderTrainer (forward, hand-derived backward, Adam) against the autograd oracle of the training step, on the CPU
through the test-only torch backend: checks the orchestration the CUDA backend shares."""
import numpy as np
import pytest

from gan_segmentation_b200.config import decoder_config
from gan_segmentation_b200.decoder_training import DecoderTrainer
from gan_segmentation_b200.random_init import init_decoder_params
from oracle import train_oracle as T
from torch_backend import TorchBackend


def _case(res, n, seed=0, base=(4, 4)):
    cfg = decoder_config(res)
    params = init_decoder_params(cfg, seed=2)
    rs = np.random.RandomState(seed)
    feats = [rs.randn(n, c, base[0] << i, base[1] << i).astype(np.float32) for i, c in enumerate(cfg['in_channels'][:res - 1])]
    mask = rs.randint(-1, cfg['num_classes'], (n, 1, base[0] << (res - 2), base[1] << (res - 2)))
    drops = [(rs.rand(n, cfg['features'][i], base[0] << i, base[1] << i) > 0.5).astype(np.float32) for i in range(res - 1)]
    return cfg, params, feats, mask, drops


@pytest.mark.parametrize('res,n,use_dropout', [(4, 2, False), (5, 1, True), (3, 3, True)])
def test_gradients_and_update_match_the_autograd_oracle(res, n, use_dropout):
    import torch
    cfg, params, feats, mask, drops = _case(res, n)
    cfg = dict(cfg, use_dropout=use_dropout, base_lr=1e-3)
    dm = drops if use_dropout else None
    tr = DecoderTrainer(cfg, params, TorchBackend())
    loss, grads = tr.loss_and_grads(feats, mask, dm)
    p_ref, st, loss_ref, g_ref = T.train_step(params, cfg, feats, mask,
                                              dropout_masks=[torch.tensor(d) for d in dm] if dm else None)
    assert np.allclose(loss.numpy(), loss_ref, rtol=1e-5, atol=1e-6)
    assert set(grads) == set(g_ref)
    for k in g_ref:
        a, b = grads[k].numpy(), g_ref[k]
        assert np.allclose(a, b, rtol=2e-3, atol=2e-5 * max(1.0, np.abs(b).max())), (k, np.abs(a - b).max(), np.abs(b).max())
    tr.step(feats, mask, dm)
    new = tr.state()
    for k in p_ref:
        tol = 5e-4 if k.endswith(('.weight', '.bias', '.gamma', '.beta')) else 1e-5   # Adam's first step ~ lr * sign(g)
        assert np.allclose(new[k], p_ref[k], rtol=1e-4, atol=tol * 1.0), (k, np.abs(new[k] - p_ref[k]).max())


def _dp_worker(rank, world, port, q):
    import os, sys
    sys.path.insert(0, os.path.dirname(__file__))
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    cfg, params, feats, mask, _ = _case(4, 2)
    cfg = dict(cfg, use_dropout=False, base_lr=1e-3)
    tr = DecoderTrainer(cfg, params, TorchBackend())
    tr.step([f[rank:rank + 1] for f in feats], mask[rank:rank + 1], None, global_batch=world)
    st = tr.state()
    q.put((rank, {k: st[k] for k in ('cvt_block_0.0.weight', 'main_block_2.0.bias', 'main_block_1.1.base_layers.1.gamma')}))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_step_two_ranks_gloo():
    """world_size 2 on CPU: each rank back-propagates its shard (BatchNorm statistics per rank, use_sync_bn=False), the
    gradients are summed by one all-reduce and both ranks apply the same Adam step with rescale 1/global batch --
    the reference's split_and_load + KVStore + trainer.step(batch) (seg_solver.py:386-421)."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg, params, feats, mask, _ = _case(4, 2)
    cfg = dict(cfg, use_dropout=False, base_lr=1e-3)
    g = [T.train_step(params, cfg, [f[r:r + 1] for f in feats], mask[r:r + 1])[3] for r in range(2)]
    for k in got[0]:
        assert np.array_equal(got[0][k], got[1][k]), k                       # both ranks hold the same parameters
        w, _, _ = T.adam_update(np.asarray(params[k], np.float64), g[0][k] + g[1][k], 0.0, 0.0, 1, 1e-3, 2)
        assert np.allclose(got[0][k], w.astype(np.float32), rtol=1e-4, atol=5e-4), k
