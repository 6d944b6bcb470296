"""Drop-in surface on the GPU: the loop of the reference's ``main.py generate`` (main.py:94-103) runs against
the mirrored ImageGenerator / SegSolver classes, single- and (when 2 GPUs are visible) multi-context."""
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
