"""Drop-in surface on the GPU: the loop of the reference's ``main.py generate`` (main.py:94-103) runs against
the mirrored ImageGenerator / SegSolver classes, single- and (when 2 GPUs are visible) multi-context."""
import os

import numpy as np
import pytest
import torch

from gan_segmentation_b200.config import generator_config, decoder_config
from gan_segmentation_b200.random_init import init_generator_params, init_decoder_params

pytestmark = pytest.mark.gpu


def _run(gpu_ids, tmp_path, batch):
    from gan_segmentation_b200.image_generator import ImageGenerator
    from gan_segmentation_b200.seg_solver import SegSolver
    from gan_segmentation_b200.params_io import save_params
    gc, dc = generator_config(8), decoder_config(8)
    gan_dir, ckpt = tmp_path / 'stylegan-models', tmp_path / 'checkpoints'
    gan_dir.mkdir(exist_ok=True); ckpt.mkdir(exist_ok=True)
    save_params(str(gan_dir / 'stylegan-bedrooms.params'), init_generator_params(gc, seed=0))      # the reference's file names
    save_params(str(ckpt / 'checkpoint_last.params'), init_decoder_params(dc, seed=2))
    solver = SegSolver(8, str(tmp_path / 'data'), str(ckpt), gpu_ids=gpu_ids[:1], keep_weights=False, verbose=False)
    assert solver.is_trained and solver.params_file == 'checkpoint_last.params'
    netG = ImageGenerator(gpu_ids=gpu_ids, gan_dir=str(gan_dir), gan='bedrooms', batch_size=batch)
    assert netG.max_res_log2 == 8 and netG.latent_size == 512
    out = []
    data_iter = netG.get_images(5, seed=3)
    for index in range(5):                                       # main.py:97-99
        img, features = next(data_iter)
        mask = solver.predict(features)[0].astype(np.uint8)
        assert img.shape == (256, 256, 3) and img.dtype == np.uint8
        assert len(features) == 7 and features[0].shape == (512, 4, 4) and features[-1].shape == (64, 256, 256)
        assert mask.shape == (256, 256, 1) and set(np.unique(mask)) <= {0, 1}
        out.append((img, mask))
    return out


def test_main_generate_loop_single_gpu(tmp_path):
    _run([0], tmp_path, batch=2)


def test_forward_is_independent_of_earlier_work_in_the_process(tmp_path):
    """A forward pass gives bit-identical features whatever ran on the device before it.  Regression test: kernels chained
    by programmatic dependent launch read producer-written data (statistics, AdaIN coefficients, activations) through the
    non-coherent load path, and a res-8 generate loop before a res-6 forward left stale lines behind -> NaN features."""
    from gan_segmentation_b200.networks import Generator
    gc = generator_config(6)

    def features():
        G = Generator(gc)
        G.set_parameters(init_generator_params(gc, seed=0))
        outs = []
        for _ in range(3):
            out = G.forward(n=3, seed=5, return_u8=True, return_features=True)
            outs.append([f.float().cpu().numpy() for f in out['features']] + [out['img_u8'].cpu().numpy()])
        return outs

    before = features()
    _run([0], tmp_path, batch=2)
    after = features()
    for rep in before + after:
        for a, b in zip(before[0], rep):
            assert np.isfinite(b).all() and np.array_equal(a, b)


@pytest.mark.parametrize('res', [6, 8])
def test_dependent_launch_does_not_change_results(res):
    """Programmatic dependent launch (gsx_set_option("pdl", ...)) only overlaps launches: a sequence of forward passes with
    DIFFERENT latents / noise per pass gives bit-identical images, features, logits and masks with it off and in every
    mode.  Catches any read of a buffer before (or through a cache line from before) its producer's write of this pass."""
    import gan_segmentation_b200._lib as L
    from gan_segmentation_b200.networks import Generator, Decoder
    gc, dc = generator_config(res), decoder_config(res)
    lib = L.lib()

    def run(mode):
        assert lib.gsx_set_option(b'pdl', mode) == 0
        G = Generator(gc)
        G.set_parameters(init_generator_params(gc, seed=0))
        D = Decoder(dc)
        D.set_parameters(init_decoder_params(dc, seed=2))
        outs = []
        for seed in (1, 2, 3, 4, 5, 6):
            n = 1 + seed % 3
            out = G.forward(n=n, seed=seed, return_u8=True, return_features=True)
            dec = D.forward(generator=G, n=n, return_logits=True)
            outs.append([f.float().cpu().numpy() for f in out['features']] +
                        [out['img_u8'].cpu().numpy(), dec['logits'].cpu().numpy(), dec['mask'].cpu().numpy()])
        return outs

    try:
        ref = run(0)
        for mode in (1, 2, 3, 4):
            got = run(mode)
            for i, (a, b) in enumerate(zip(ref, got)):
                for j, (x, y) in enumerate(zip(a, b)):
                    assert np.isfinite(y).all() and np.array_equal(x, y), (mode, i, j)
    finally:
        lib.gsx_set_option(b'pdl', 3)


def test_multi_context_matches_single(tmp_path):
    """split_and_load over two contexts (image_generator.py:95-101) gives the same samples as one context."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    a = _run([0], tmp_path, batch=4)
    b = _run([0, 1], tmp_path, batch=4)
    # z from the seeded host RNG, noise from the Philox stream keyed by the running sample index: every sample is
    # bit-identical however the batch is split over contexts
    for (ia, ma), (ib, mb) in zip(a, b):
        assert np.array_equal(ia, ib) and np.array_equal(ma, mb)
    # and consecutive samples do not share noise
    assert not np.array_equal(a[0][0], a[1][0])


def test_evaluate_annotated_samples(tmp_path):
    """SegSolver.evaluate (seg_solver.py:222-305) over annotated samples written in the reference's format: the
    returned accuracy / mean-IoU / total-loss equal the reference formulas (metrics.py:549-606, SoftmaxCELoss with
    weight (mask > -1)) evaluated in numpy/torch on the decoder's own logits, and the per-sample files appear."""
    import torch.nn.functional as F
    from gan_segmentation_b200.seg_solver import SegSolver
    from gan_segmentation_b200.seg_datasets import save_sample
    from gan_segmentation_b200.params_io import save_params
    gc, dc = generator_config(6), decoder_config(6)
    gan_dir, ckpt, data = tmp_path / 'stylegan-models', tmp_path / 'checkpoints', tmp_path / 'data'
    for d in (gan_dir, ckpt, data):
        d.mkdir(exist_ok=True)
    save_params(str(ckpt / 'checkpoint_last.params'), init_decoder_params(dc, seed=2))
    solver = SegSolver(6, str(data), str(ckpt), gpu_ids=[0], keep_weights=False, verbose=False)
    from gan_segmentation_b200.networks import Generator
    G = Generator(gc)
    G.set_parameters(init_generator_params(gc, seed=0))
    out = G.forward(n=3, seed=5, return_u8=True, return_features=True)
    imgs = out['img_u8'].cpu().numpy()
    feats_all = [f.cpu().numpy() for f in out['features']]
    rs = np.random.RandomState(1)
    samples = [(imgs[i], [f[i] for f in feats_all]) for i in range(3)]
    labels = []
    for i, (img, feats) in enumerate(samples):
        lab = rs.randint(-1, 2, img.shape[:2])
        labels.append(lab)
        save_sample(str(data), i, img, feats, lab)
    res = dict(solver.evaluate(str(data), output_dir=str(tmp_path / 'eval_out')))
    assert set(res) == {'accuracy', 'mean-iou', 'total-loss'}
    # reference formulas on the decoder's logits
    tot_c = tot_l = 0
    inter = np.zeros(2); union = np.zeros(2); losses = []
    for (img, feats), lab in zip(samples, labels):
        lg = solver.net.forward([f[None] for f in feats], return_logits=True)['logits'].float().cpu()
        pred = lg.argmax(1)[0].numpy()
        valid = lab > -1
        tot_c += int(((pred == lab) & valid).sum()); tot_l += int(valid.sum())
        for c in range(2):
            p_c = (pred == c) & valid; l_c = lab == c
            inter[c] += (p_c & l_c).sum(); union[c] += (p_c | l_c).sum()
        lp = F.log_softmax(lg, dim=1)
        t = torch.as_tensor(lab)[None, None]
        picked = -torch.gather(lp, 1, t.clamp(min=0)) * (t > -1).float()
        losses.append(picked.mean().item())
    assert abs(res['accuracy'] - tot_c / tot_l) < 1e-9
    assert abs(res['mean-iou'] - (inter / union)[1:].mean()) < 1e-9
    assert abs(res['total-loss'] - np.mean(losses)) < 1e-5
    names = sorted(p.name for p in (tmp_path / 'eval_out').iterdir())
    assert names == sorted([f'{k}_{i:06d}.{e}' for i in range(3) for k, e in
                            (('img', 'jpg'), ('mask', 'png'), ('gt_mask', 'png'), ('metrics', 'txt'))])


def test_fit_trains_the_decoder(tmp_path):
    """SegSolver.fit (seg_solver.py:351-465) on annotated samples: runs on the CUDA kernels, lowers the evaluation loss,
    marks the solver trained and writes checkpoint_last.params that a fresh solver loads."""
    from gan_segmentation_b200.seg_solver import SegSolver
    from gan_segmentation_b200.seg_datasets import save_sample
    from gan_segmentation_b200.networks import Generator
    gc = generator_config(6)
    ckpt, data = tmp_path / 'checkpoints', tmp_path / 'data'
    ckpt.mkdir(); data.mkdir()
    G = Generator(gc)
    G.set_parameters(init_generator_params(gc, seed=0))
    out = G.forward(n=4, seed=5, return_u8=True, return_features=True)
    imgs = out['img_u8'].cpu().numpy()
    feats_all = [f.cpu().numpy() for f in out['features']]
    for i in range(4):
        lab = (feats_all[-1][i, 0] > np.median(feats_all[-1][i, 0])).astype(np.int64)      # a learnable target
        save_sample(str(data), i, imgs[i], [f[i] for f in feats_all], lab)
    solver = SegSolver(6, str(data), str(ckpt), gpu_ids=[0], keep_weights=True, verbose=False)
    assert not solver.is_trained
    solver.cfg['train_epochs'] = 6
    solver.cfg['base_lr'] = 2e-3
    before = dict(solver.evaluate(str(data)))['total-loss']
    calls = []
    assert solver.fit(epoch_end_callback=lambda: calls.append(1)) == []
    after = dict(solver.evaluate(str(data)))
    assert len(calls) == 6 and solver.is_trained and solver.params_file == 'checkpoint_last.params'
    assert after['total-loss'] < 0.8 * before, (before, after)
    solver2 = SegSolver(6, str(data), str(ckpt), gpu_ids=[0], keep_weights=True, verbose=False)
    assert solver2.is_trained
    assert abs(dict(solver2.evaluate(str(data)))['total-loss'] - after['total-loss']) < 1e-5


def test_cli_generate_action_writes_the_dataset(tmp_path):
    """``python -m gan_segmentation_b200.main generate`` (reference main.py:75-104): GENERATE_NUM pairs of
    img_XXXXXX.jpg / mask_XXXXXX.png under BASE_DIR/dataset/train_generated, read from the reference's config keys."""
    import cv2
    from gan_segmentation_b200 import main as M
    cfg = tmp_path / 'config.yml'
    cfg.write_text('BASE_DIR: "%s"\nGAN: "bedrooms"\nGAN_DIR: "none"\nGAN_GPU_IDS: [0]\nGAN_BATCH_SIZE_PER_GPU: 2\n'
                   'SOLVER_GPU_IDS: [0]\nANNOTATION: "segmentation"\nGENERATE_NUM: 5\n' % tmp_path)
    assert M.main(['generate', '--config', str(cfg), '--random-init', '--psi', '0.7']) == 0
    out = tmp_path / 'dataset' / 'train_generated'
    names = sorted(os.listdir(out))
    assert names == [f'img_{i:06d}.jpg' for i in range(5)] + [f'mask_{i:06d}.png' for i in range(5)]
    m = cv2.imread(str(out / 'mask_000004.png'), cv2.IMREAD_UNCHANGED)
    assert m.shape == (256, 256) and set(np.unique(m)) <= {0, 1}
    assert cv2.imread(str(out / 'img_000000.jpg')).shape == (256, 256, 3)
