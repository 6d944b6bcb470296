"""Parameter names and shapes of the reference's Gluon blocks.

Two spellings are accepted everywhere a name crosses the boundary, because Gluon's
``load_parameters`` accepts both:
  * structural names -- what ``save_parameters`` writes (attribute path joined by '.'),
  * legacy prefix names -- what the explicit ``prefix=`` strings produce
    (networks_stylegan.py:16-54,125,136,244-247; Generator prefix is '' at :79).
``canonical()`` maps either spelling to the structural one.
"""
from __future__ import annotations

import re
from .config import num_features


def generator_param_shapes(cfg):
    """Ordered {structural name: shape} for ``Generator(cfg)`` (networks_stylegan.py:76-156)."""
    L = cfg['max_res_log2']
    Z = cfg['latent_size']
    nc = cfg['channels']
    shapes = {}
    c2 = num_features(cfg, 2)
    shapes['constant_tensor'] = (1, c2, cfg['base_scale_y'], cfg['base_scale_x'])   # :94-96
    shapes['latent_avg'] = (512,)                                                   # :97
    shapes['truncation_psi'] = ((L - 1) * 2,)                                       # :99
    for i in range(8):                                                              # :133-137
        k = 2 * i + 1
        shapes[f'mapping.{k}.weight'] = (Z, Z)
        shapes[f'mapping.{k}.bias'] = (Z,)
        shapes[f'mapping.{k}.std'] = (1,)
    for r in range(2, L + 1):
        c = num_features(cfg, r)
        cin = num_features(cfg, r - 1) if r > 2 else c
        p = f'net{r}'
        if r > 2:
            if r >= 7:      # fused upscale: Deconvolution weight is (Cin, Cout, 4, 4)  (:15-17, :154)
                shapes[f'{p}.block0.weight'] = (cin, c, 4, 4)
            else:
                shapes[f'{p}.block0.weight'] = (c, cin, 3, 3)
            shapes[f'{p}.block0.std'] = (1,)
            shapes[f'{p}.blur.w_kernel'] = (c, 1, 3, 3)
        shapes[f'{p}.block1.0.scale_factors'] = (1, c, 1, 1)
        shapes[f'{p}.block1.1.bias'] = (1, c, 1, 1)
        shapes[f'{p}.adain1.affine.weight'] = (2 * c, Z)
        shapes[f'{p}.adain1.affine.bias'] = (2 * c,)
        shapes[f'{p}.adain1.affine.std'] = (1,)
        shapes[f'{p}.adain1.instance.gamma'] = (c,)
        shapes[f'{p}.adain1.instance.beta'] = (c,)
        shapes[f'{p}.block2.0.weight'] = (c, c, 3, 3)
        shapes[f'{p}.block2.0.std'] = (1,)
        shapes[f'{p}.block2.1.scale_factors'] = (1, c, 1, 1)
        shapes[f'{p}.block2.2.bias'] = (1, c, 1, 1)
        shapes[f'{p}.adain2.affine.weight'] = (2 * c, Z)
        shapes[f'{p}.adain2.affine.bias'] = (2 * c,)
        shapes[f'{p}.adain2.affine.std'] = (1,)
        shapes[f'{p}.adain2.instance.gamma'] = (c,)
        shapes[f'{p}.adain2.instance.beta'] = (c,)
    cl = num_features(cfg, L)
    shapes[f'to_rgb{L}.0.weight'] = (nc, cl, 1, 1)
    shapes[f'to_rgb{L}.0.bias'] = (nc,)
    shapes[f'to_rgb{L}.0.std'] = (1,)
    return shapes


def decoder_param_shapes(cfg):
    """Ordered {structural name: shape} for ``Decoder(cfg)`` (networks_seg.py:49-95)."""
    feats = cfg['features']
    inch = cfg['in_channels']
    nf = len(inch)
    s0 = cfg['start_res']
    shapes = {}

    def bn(prefix, c):
        for k in ('gamma', 'beta', 'running_mean', 'running_var'):
            shapes[f'{prefix}.{k}'] = (c,)

    for i in range(s0, nf):
        c = feats[i]
        shapes[f'cvt_block_{i}.0.weight'] = (c, inch[i], 3, 3)
        shapes[f'cvt_block_{i}.0.bias'] = (c,)
        if cfg['use_bn']:
            bn(f'cvt_block_{i}.1', c)
    for i in range(s0, nf):
        cout = feats[i + 1]
        cin = feats[i] * (2 if i > s0 else 1)
        if i < nf - 1:
            p = f'main_block_{i}.1'
            j = 0
            shapes[f'{p}.base_layers.{j}.weight'] = (cout, cin, 3, 3)
            shapes[f'{p}.base_layers.{j}.bias'] = (cout,)
            j += 1
            if cfg['use_bn']:
                bn(f'{p}.base_layers.{j}', cout)
                j += 1
            j += 1  # LeakyReLU
            shapes[f'{p}.base_layers.{j}.weight'] = (cout, cout, 3, 3)
            shapes[f'{p}.base_layers.{j}.bias'] = (cout,)
            j += 1
            if cfg['use_bn']:
                bn(f'{p}.base_layers.{j}', cout)
            if cout != cin:
                shapes[f'{p}.shortcut.0.weight'] = (cout, cin, 1, 1)
                shapes[f'{p}.shortcut.0.bias'] = (cout,)
        else:
            shapes[f'main_block_{i}.0.weight'] = (cout, cin, 3, 3)
            shapes[f'main_block_{i}.0.bias'] = (cout,)
    return shapes


_LEGACY = [
    (re.compile(r'^mp_dense_(\d+)_(weight|bias|std)$'),
     lambda m: f'mapping.{2 * int(m.group(1)) + 1}.{m.group(2)}'),
    (re.compile(r'^(\d+)_(?:conv|deconv)_1_(weight|std)$'),
     lambda m: f'net{_log2(m.group(1))}.block0.{m.group(2)}'),
    (re.compile(r'^(\d+)_blur_1_w_kernel$'),
     lambda m: f'net{_log2(m.group(1))}.blur.w_kernel'),
    (re.compile(r'^(\d+)_noise_1_scale_factors$'),
     lambda m: f'net{_log2(m.group(1))}.block1.0.scale_factors'),
    (re.compile(r'^(\d+)_bias_1_bias$'),
     lambda m: f'net{_log2(m.group(1))}.block1.1.bias'),
    (re.compile(r'^(\d+)_adain_([12])_dense_affine_(weight|bias|std)$'),
     lambda m: f'net{_log2(m.group(1))}.adain{m.group(2)}.affine.{m.group(3)}'),
    (re.compile(r'^(\d+)_adain_([12])_norm_(gamma|beta)$'),
     lambda m: f'net{_log2(m.group(1))}.adain{m.group(2)}.instance.{m.group(3)}'),
    (re.compile(r'^(\d+)_conv_2_(weight|std)$'),
     lambda m: f'net{_log2(m.group(1))}.block2.0.{m.group(2)}'),
    (re.compile(r'^(\d+)_noise_2_scale_factors$'),
     lambda m: f'net{_log2(m.group(1))}.block2.1.scale_factors'),
    (re.compile(r'^(\d+)_bias_2_bias$'),
     lambda m: f'net{_log2(m.group(1))}.block2.2.bias'),
    (re.compile(r'^(\d+)_conv_to_rgb_(weight|bias|std)$'),
     lambda m: f'to_rgb{_log2(m.group(1))}.0.{m.group(2)}'),
]


def _log2(s):
    v = int(s)
    r = v.bit_length() - 1
    if (1 << r) != v:
        raise KeyError(f'scale prefix {s} is not a power of two')
    return r


def canonical(name):
    """Map a legacy prefix name (or an 'arg:'/'aux:' decorated one) to the structural name."""
    if name.startswith('arg:') or name.startswith('aux:'):
        name = name[4:]
    if '.' in name or name in ('constant_tensor', 'latent_avg', 'truncation_psi'):
        return name
    for rx, fn in _LEGACY:
        m = rx.match(name)
        if m:
            return fn(m)
    return name


def legacy_name(structural, cfg=None):
    """Inverse of ``canonical`` for generator names (used by the .params writer tests)."""
    m = re.match(r'^mapping\.(\d+)\.(\w+)$', structural)
    if m:
        return f'mp_dense_{(int(m.group(1)) - 1) // 2}_{m.group(2)}'
    m = re.match(r'^net(\d+)\.(.+)$', structural)
    if m:
        r = int(m.group(1))
        s = 2 ** r
        rest = m.group(2)
        first = 'deconv_1' if r >= 7 else 'conv_1'
        table = {
            'block0.weight': f'{first}_weight', 'block0.std': f'{first}_std',
            'blur.w_kernel': 'blur_1_w_kernel',
            'block1.0.scale_factors': 'noise_1_scale_factors', 'block1.1.bias': 'bias_1_bias',
            'block2.0.weight': 'conv_2_weight', 'block2.0.std': 'conv_2_std',
            'block2.1.scale_factors': 'noise_2_scale_factors', 'block2.2.bias': 'bias_2_bias',
        }
        if rest in table:
            return f'{s}_{table[rest]}'
        m2 = re.match(r'^adain([12])\.affine\.(\w+)$', rest)
        if m2:
            return f'{s}_adain_{m2.group(1)}_dense_affine_{m2.group(2)}'
        m2 = re.match(r'^adain([12])\.instance\.(\w+)$', rest)
        if m2:
            return f'{s}_adain_{m2.group(1)}_norm_{m2.group(2)}'
    m = re.match(r'^to_rgb(\d+)\.0\.(\w+)$', structural)
    if m:
        return f'{2 ** int(m.group(1))}_conv_to_rgb_{m.group(2)}'
    return structural
