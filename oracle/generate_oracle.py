"""CPU ORACLE (test infrastructure, NOT product code) for the generate hot path.

A plain PyTorch-CPU restatement of the reference's algorithm: StyleGAN-v1 synthesis
(reference networks_stylegan.py) + segmentation decoder (networks_seg.py) + the uint8 image
transform (image_generator.py:76-84) + the argmax label map (seg_solver.py:326-327).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module, and only as the checker / the timed CPU baseline.
The product path (gan-segmentation_b200/) never imports it.

PARITY UNPINNED: the arithmetic of the reference lives in Apache MXNet (README pins v1.5.1 by
link; not in requirements.txt, not vendored, not installable offline) and the reference holds
no tests, golden vectors or sample outputs for this path.  This restatement follows the
reference call sites line by line under the MXNet-1.5 operator semantics listed below; it has
been checked against the reference-derived known-answer invariants in tests/test_oracle_kat.py
but not against MXNet output.  ``dump_npz``/``load_npz`` use the reference's parameter names so
a later MXNet run can pin it.

Assumed MXNet operator semantics:
 (1) Convolution = cross-correlation, weight (Cout,Cin,kh,kw), zero pad        == F.conv2d
 (2) Deconvolution weight (Cin,Cout,kh,kw), out=(in-1)*s-2p+k                   == F.conv_transpose2d
 (3) UpSampling(nearest, 2)                                                     == F.interpolate(nearest)
 (4) gluon InstanceNorm eps=1e-5, biased variance over HxW, gamma=1/beta=0      == F.instance_norm(eps=1e-5)
 (5) LeakyReLU(0.2): x>0 ? x : 0.2x
 (6) FullyConnected: x W^T + b, W (units, in)
 (7) gluon BatchNorm eps=1e-5; inference uses running stats
 (8) Dropout(0.5) is the identity at inference
 (9) argmax returns float32 indices, first occurrence wins ties
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _t(a, dtype):
    if isinstance(a, torch.Tensor):
        return a.to(dtype)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def _nf(cfg, r):
    # networks_stylegan.py:114-116
    return min(int(cfg['fmap_base'] / (2.0 ** ((r - 1) * cfg['fmap_decay']))), cfg['fmap_max'])


def pixel_norm(x, eps=1e-8):
    """networks_stylegan.py:558-565."""
    return x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + eps)


def dense_w(x, weight, bias, std, lr_mult):
    """DenseW.hybrid_forward, networks_stylegan.py:513-524."""
    w = weight * std * lr_mult
    b = None if bias is None else bias * lr_mult
    return F.linear(x, w, b)


def mapping(P, cfg, z):
    """build_mapping, networks_stylegan.py:128-139: PixelNorm then 8x(DenseW gain sqrt2 lr_mult .01 + lrelu)."""
    x = pixel_norm(z)
    for i in range(8):
        k = 2 * i + 1
        x = dense_w(x, P[f'mapping.{k}.weight'], P[f'mapping.{k}.bias'], P[f'mapping.{k}.std'], 0.01)
        x = F.leaky_relu(x, 0.2)
    return x


def lerp(coeff, latent_avg, w):
    """Generator.lerp, networks_stylegan.py:158-163: avg*(1-psi) + w*psi."""
    return latent_avg * (1 - coeff) + w * coeff


def blur(x):
    """Blur, networks_stylegan.py:200-236: depthwise [1,2,1]x[1,2,1]/16, zero pad 1."""
    c = x.shape[1]
    k = torch.tensor([1., 2., 1.], dtype=x.dtype)
    k = torch.outer(k, k)
    k = (k / k.sum()).reshape(1, 1, 3, 3).repeat(c, 1, 1, 1)
    return F.conv2d(x, k, None, 1, 1, 1, c)


def adain(x, wl, weight, bias, std):
    """AdaIN.hybrid_forward, networks_stylegan.py:250-264."""
    y = dense_w(wl, weight, bias, std, 1.0)            # N x 2C
    n, c = x.shape[0], x.shape[1]
    y = y.reshape(n, 2, c)
    ys = y[:, 0, :].reshape(n, c, 1, 1)
    yb = y[:, 1, :].reshape(n, c, 1, 1)
    xn = F.instance_norm(x, eps=1e-5)
    return xn * (ys + 1) + yb


def style_block(P, cfg, r, x, w1, w2, noise1, noise2):
    """StyleGeneratorBlock.hybrid_forward, networks_stylegan.py:56-73 (block built at :141-156)."""
    p = f'net{r}'
    y = x
    if r > 2:
        wgt = P[f'{p}.block0.weight'] * P[f'{p}.block0.std']
        if r >= 7:
            y = F.conv_transpose2d(y, wgt, None, stride=2, padding=1)    # :15-17
        else:
            y = F.interpolate(y, scale_factor=2, mode='nearest')         # :27, :315
            y = F.conv2d(y, wgt, None, 1, 1)                             # :24
        y = blur(y)                                                      # :34
    y = y + P[f'{p}.block1.0.scale_factors'] * noise1                    # AddNoise :302-304
    y = y + P[f'{p}.block1.1.bias']                                      # Bias :544
    y = F.leaky_relu(y, 0.2)
    y = adain(y, w1, P[f'{p}.adain1.affine.weight'], P[f'{p}.adain1.affine.bias'], P[f'{p}.adain1.affine.std'])
    y = F.conv2d(y, P[f'{p}.block2.0.weight'] * P[f'{p}.block2.0.std'], None, 1, 1)
    y = y + P[f'{p}.block2.1.scale_factors'] * noise2
    y = y + P[f'{p}.block2.2.bias']
    y = F.leaky_relu(y, 0.2)
    y = adain(y, w2, P[f'{p}.adain2.affine.weight'], P[f'{p}.adain2.affine.bias'], P[f'{p}.adain2.affine.std'])
    return y


def generator_forward(params, cfg, z, noise, psi=None, dtype=torch.float32):
    """Generator.hybrid_forward, networks_stylegan.py:165-197.

    z [N,512]; noise: list of 2*(L-1) arrays [N,1,h,w] (the reference samples them inside
    AddNoise :300; parity runs always pass them explicitly); psi: None -> the
    ``truncation_psi`` parameter, else scalar or per-layer vector overriding it.
    Returns (img [N,3,H,W], [features]) as torch tensors of ``dtype``.
    """
    P = {k: _t(v, dtype) for k, v in params.items()}
    z = _t(z, dtype)
    noise = [_t(a, dtype) for a in noise]
    L = cfg['max_res_log2']
    n = z.shape[0]
    w = mapping(P, cfg, z)                                               # :168
    tp = P['truncation_psi'] if psi is None else torch.broadcast_to(_t(np.asarray(psi, np.float32), dtype), (2 * (L - 1),))
    avg = P['latent_avg'].reshape(1, -1)                                 # :171
    const = P['constant_tensor'].expand(n, -1, -1, -1)                   # :173-178
    feats = []
    y = const
    for r in range(2, L + 1):
        i = 2 * (r - 2)
        w1 = lerp(tp[i], avg, w)                                         # :180-189
        w2 = lerp(tp[i + 1], avg, w)
        y = style_block(P, cfg, r, y, w1, w2, noise[i], noise[i + 1])
        feats.append(y)
    wr = P[f'to_rgb{L}.0.weight'] * P[f'to_rgb{L}.0.std']                # :118-126, gain 1
    img = F.conv2d(y, wr, P[f'to_rgb{L}.0.bias'])
    return img, feats


def transform_gan_back(img):
    """ImageGenerator._transform_gan_back, image_generator.py:76-84 (imrange (-1,1));
    float32 arithmetic and C-style truncation to uint8, as numpy does there."""
    a = np.asarray(img, np.float32)
    a = np.transpose(a, (0, 2, 3, 1))
    a = (a - np.float32(-1)) / np.float32(2)
    a = np.clip(a, 0.0, 1.0)
    a = np.float32(255.) * a
    return a.astype(np.uint8)


def _bn(P, prefix, x, eps=1e-5):
    return F.batch_norm(x, P[f'{prefix}.running_mean'], P[f'{prefix}.running_var'],
                        P[f'{prefix}.gamma'], P[f'{prefix}.beta'], False, 0.0, eps)


def decoder_forward(params, cfg, feats, dtype=torch.float32):
    """Decoder.hybrid_forward (inference), networks_seg.py:97-113; blocks built at :64-94,
    DecoderResBlock :7-46.  feats[i]: [N,C_i,H_i,W_i].  Returns logits [N,num_classes,H,W]."""
    P = {k: _t(v, dtype) for k, v in params.items()}
    feats = [_t(f, dtype) for f in feats]
    nf = len(cfg['in_channels'])
    s0 = cfg['start_res']
    use_bn = cfg['use_bn']
    prev = None
    for i in range(s0, nf):
        x = F.conv2d(feats[i], P[f'cvt_block_{i}.0.weight'], P[f'cvt_block_{i}.0.bias'], 1, 1)
        if use_bn:
            x = _bn(P, f'cvt_block_{i}.1', x)
        x = F.leaky_relu(x, 0.2)                     # Dropout(0.5) :78 is identity at inference
        if i > s0:
            x = torch.cat([prev, x], dim=1)          # :108-109
        if i < nf - 1:
            p = f'main_block_{i}.1'
            x = F.interpolate(x, scale_factor=2, mode='nearest')     # :87
            j = 0
            y = F.conv2d(x, P[f'{p}.base_layers.{j}.weight'], P[f'{p}.base_layers.{j}.bias'], 1, 1)
            j += 1
            if use_bn:
                y = _bn(P, f'{p}.base_layers.{j}', y)
                j += 1
            y = F.leaky_relu(y, 0.2)
            j += 1
            y = F.conv2d(y, P[f'{p}.base_layers.{j}.weight'], P[f'{p}.base_layers.{j}.bias'], 1, 1)
            j += 1
            if use_bn:
                y = _bn(P, f'{p}.base_layers.{j}', y)
            y = F.leaky_relu(y, 0.2)
            if f'{p}.shortcut.0.weight' in P:
                sc = F.conv2d(x, P[f'{p}.shortcut.0.weight'], P[f'{p}.shortcut.0.bias'])
            else:
                sc = x
            prev = sc + y                            # :46
        else:
            prev = F.conv2d(x, P[f'main_block_{i}.0.weight'], P[f'main_block_{i}.0.bias'], 1, 1)
    return prev


def argmax_mask(logits):
    """seg_solver.py:326-327: argmax(axis=1, keepdims) -> transpose (0,2,3,1); float32 class ids,
    first maximum wins.  (torch.argmax does not document its tie rule, so spell it out.)"""
    lg = np.asarray(logits, np.float32)
    best = lg[:, 0]
    idx = np.zeros(best.shape, np.float32)
    for c in range(1, lg.shape[1]):
        m = lg[:, c] > best
        idx[m] = c
        best = np.where(m, lg[:, c], best)
    return idx[:, :, :, None]


def generate(gen_params, gen_cfg, dec_params, dec_cfg, z, noise, psi=None, dtype=torch.float32):
    """What ``main.py generate`` computes per latent (main.py:97-99): uint8 image, features, mask."""
    with torch.no_grad():
        img, feats = generator_forward(gen_params, gen_cfg, z, noise, psi, dtype)
        logits = decoder_forward(dec_params, dec_cfg, feats, dtype)
    img_np = img.float().numpy()
    return {
        'img_f32': img_np,
        'img_u8': transform_gan_back(img_np),
        'features': [f.float().numpy() for f in feats],
        'logits': logits.float().numpy(),
        'mask': argmax_mask(logits.float().numpy()),
    }


def dump_npz(path, params):
    """Parameters keyed by the reference's structural names, for pinning against a real MXNet run."""
    np.savez(path, **{k: np.asarray(v) for k, v in params.items()})


def load_npz(path):
    with np.load(path) as f:
        return {k: f[k] for k in f.files}
