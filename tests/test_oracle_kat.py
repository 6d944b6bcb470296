"""Known-answer tests of the CPU oracle: the reference-derived invariants of SURVEY.md section 4.
(The reference holds no tests or golden vectors and MXNet is not installable offline, so these KATs
and the committed fixture under tests/golden/ are what pins the oracle.)"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gan_segmentation_b200.config import generator_config, decoder_config, noise_shapes, num_features
from gan_segmentation_b200.random_init import init_generator_params, init_decoder_params
from oracle import generate_oracle as O
from parity_util import make_case, psnr

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def test_blur_kernel_is_121_over_16():
    x = torch.zeros(1, 2, 5, 5)
    x[0, :, 2, 2] = 16.0
    y = O.blur(x)
    assert torch.equal(y[0, 0, 1:4, 1:4], torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]))
    # zero padding: a corner impulse loses the taps that fall outside
    x = torch.zeros(1, 1, 4, 4)
    x[0, 0, 0, 0] = 16.0
    assert float(O.blur(x).sum()) == 9.0


def test_pixel_norm_unit_mean_square():
    x = torch.randn(4, 512)
    y = O.pixel_norm(x)
    assert torch.allclose((y * y).mean(dim=1), torch.ones(4), atol=1e-5)


def test_num_features_table():
    cfg = generator_config(10)
    assert [num_features(cfg, r) for r in range(2, 11)] == [512, 512, 512, 512, 256, 128, 64, 32, 16]
    assert decoder_config(10)['features'] == [32] * 8 + [16, 2]
    assert decoder_config(9)['features'] == [32] * 8 + [2]
    assert decoder_config(8)['features'] == [32] * 7 + [2]
    assert decoder_config(10)['in_channels'] == [512, 512, 512, 512, 256, 128, 64, 32, 16]


def _small(max_res_log2=4, n=2, base=(4, 4), seed=0):
    return make_case(max_res_log2, n, base, seed)


def test_zero_noise_scale_makes_output_independent_of_noise():
    gc, dc, gp, dp, z, noise = _small()
    for k in gp:
        if k.endswith('scale_factors'):
            gp[k] = np.zeros_like(gp[k])
    with torch.no_grad():
        a, _ = O.generator_forward(gp, gc, z, noise)
        b, _ = O.generator_forward(gp, gc, z, [3 * p - 1 for p in noise])
    assert torch.equal(a, b)


def test_truncation_psi_semantics():
    gc, dc, gp, dp, z, noise = _small()
    with torch.no_grad():
        a, _ = O.generator_forward(gp, gc, z, noise, psi=0.0)           # w_l = latent_avg for every layer
        b, _ = O.generator_forward(gp, gc, z[::-1].copy(), noise, psi=0.0)
        c, _ = O.generator_forward(gp, gc, z, noise, psi=1.0)
        gp2 = dict(gp)
        gp2['latent_avg'] = gp['latent_avg'] + 5.0                      # psi=1: w_l = w, avg irrelevant
        d, _ = O.generator_forward(gp2, gc, z, noise, psi=1.0)
    assert torch.equal(a, b)
    assert torch.allclose(c, d, atol=1e-5)
    # per-layer vector: only layer 0 truncated differs from none truncated
    psi = np.ones(2 * (gc['max_res_log2'] - 1), np.float32)
    psi[0] = 0.0
    with torch.no_grad():
        e, _ = O.generator_forward(gp, gc, z, noise, psi=psi)
    assert not torch.allclose(e, c)


def test_instance_norm_and_adain_identity():
    x = torch.randn(2, 8, 6, 6) * 3 + 1
    w = torch.randn(2, 512)
    y = O.adain(x, w, torch.zeros(16, 512), torch.zeros(16), torch.tensor([1.0]), )
    assert torch.allclose(y.mean(dim=(2, 3)), torch.zeros(2, 8), atol=1e-5)
    var = x.var(dim=(2, 3), unbiased=False)
    assert torch.allclose(y.var(dim=(2, 3), unbiased=False), var / (var + 1e-5), atol=1e-5)
    # scale = first C units, shift = last C
    b = torch.zeros(16)
    b[:8] = 1.0          # scale -> 2x
    b[8:] = 0.5          # shift
    y2 = O.adain(x, w, torch.zeros(16, 512), b, torch.tensor([1.0]))
    assert torch.allclose(y2, 2 * y + 0.5, atol=1e-5)


def test_output_shapes_square_and_nonsquare():
    for L, base in ((5, (4, 4)), (5, (3, 4))):
        gc, dc, gp, dp, z, noise = make_case(L, 1, base)
        with torch.no_grad():
            img, feats = O.generator_forward(gp, gc, z, noise)
        assert tuple(img.shape) == (1, 3, base[0] << (L - 2), base[1] << (L - 2))
        assert [f.shape[1] for f in feats] == [512] * 4
        assert tuple(feats[0].shape[2:]) == base
        assert tuple(gp['constant_tensor'].shape) == (1, 512, base[0], base[1])


def test_deconv_and_upconv_double_resolution():
    x = torch.randn(1, 4, 5, 7)
    assert F.conv_transpose2d(x, torch.randn(4, 3, 4, 4), stride=2, padding=1).shape[2:] == (10, 14)
    assert F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), torch.randn(3, 4, 3, 3), None, 1, 1).shape[2:] == (10, 14)


def test_decoder_fresh_bn_and_wiring():
    dc = decoder_config(5)
    dp = init_decoder_params(dc, seed=1, mode='reference')       # fresh BN: x / sqrt(1 + 1e-5)
    feats = [np.random.RandomState(i).randn(1, c, 4 << i, 4 << i).astype(np.float32) for i, c in enumerate(dc['in_channels'])]
    with torch.no_grad():
        lg = O.decoder_forward(dp, dc, feats)
    assert tuple(lg.shape) == (1, 2, 32, 32)
    assert 'main_block_0.1.shortcut.0.weight' not in dp           # 32 -> 32: identity shortcut (networks_seg.py:35)
    assert dp['main_block_1.1.shortcut.0.weight'].shape == (32, 64, 1, 1)
    assert dp['main_block_3.0.weight'].shape == (2, 64, 3, 3)


def test_argmax_first_max_and_layout():
    lg = np.zeros((1, 3, 2, 2), np.float32)
    lg[0, 2, 0, 0] = 1.0
    lg[0, 1, 0, 1] = 1.0
    lg[0, 2, 0, 1] = 1.0      # tie between classes 1 and 2 -> 1
    m = O.argmax_mask(lg)
    assert m.shape == (1, 2, 2, 1) and m.dtype == np.float32
    assert m[0, :, :, 0].tolist() == [[2.0, 1.0], [0.0, 0.0]]


def test_uint8_transform_truncates():
    img = np.array([[[[-1.0, 0.0, 1.0, 2.0, -3.0, 0.999]]]], np.float32).repeat(3, axis=1)
    u8 = O.transform_gan_back(img)
    assert u8.shape == (1, 1, 6, 3)
    assert u8[0, 0, :, 0].tolist() == [0, 127, 255, 255, 0, 254]


def test_golden_fixture():
    """Committed oracle outputs (tests/golden/make_golden.py) -- guards the oracle against silent drift
    (torch / oneDNN version changes, edits)."""
    path = os.path.join(GOLDEN, 'oracle_res5_seed0.npz')
    g = np.load(path)
    gc, dc, gp, dp, z, noise = make_case(5, 2, seed=0)
    out = O.generate(gp, gc, dp, dc, z, noise)
    assert np.allclose(out['img_f32'], g['img_f32'], atol=2e-4)
    assert np.allclose(out['logits'], g['logits'], atol=2e-4)
    assert (out['img_u8'].astype(int) - g['img_u8'].astype(int)).__abs__().max() <= 1
    assert (out['mask'] == g['mask']).mean() > 0.999
    for a, b in zip(out['features'], [g[f'feat{i}'] for i in range(4)]):
        assert np.allclose(a[:, :8], b, atol=2e-4)


def test_fp64_oracle_agrees_with_fp32():
    gc, dc, gp, dp, z, noise = make_case(5, 1, seed=2)
    a = O.generate(gp, gc, dp, dc, z, noise)
    b = O.generate(gp, gc, dp, dc, z, noise, dtype=torch.float64)
    assert psnr(np.clip(a['img_f32'], -1, 1), np.clip(b['img_f32'], -1, 1)) > 90
    assert (a['mask'] == b['mask']).mean() > 0.9995


def test_bf16_storage_error_floor():
    """Why the default build stores fp16: rounding only the conv WEIGHTS of the oracle to bf16 already
    breaks the north star's max-abs 2e-2 image bound; fp16 rounding does not."""
    gc, dc, gp, dp, z, noise = make_case(6, 1, seed=0)
    with torch.no_grad():
        ref, _ = O.generator_forward(gp, gc, z, noise)

        def rounded(dt):
            q = dict(gp)
            for k, v in gp.items():
                if k.endswith('.weight') and ('block0' in k or 'block2' in k):
                    std = gp[k[:-7] + '.std']
                    q[k] = (torch.from_numpy(v * std).to(dt).float() / float(std[0])).numpy()
            img, _ = O.generator_forward(q, gc, z, noise)
            return float((img.clamp(-1, 1) - ref.clamp(-1, 1)).abs().max())
        e_bf16, e_fp16 = rounded(torch.bfloat16), rounded(torch.float16)
    assert e_fp16 < 2e-2 < e_bf16, (e_fp16, e_bf16)
