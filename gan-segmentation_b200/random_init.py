"""Random-init weights of the named architecture, in the reference's parameter layout.

The pretrained ``stylegan-*.params`` / ``checkpoint_last.params`` are not available
offline, so benchmarks and parity tests use random weights of the same architecture.
With the reference's *declared* initialisers the style path is dead (mapping weights are
multiplied by 0.01*sqrt(2)/sqrt(512) (networks_stylegan.py:514-516) so w ~ 1e-16 after 8
layers; noise scale_factors, biases and latent_avg start at 0; psi at 1), which would
leave half the path untested.  ``mode='nondegenerate'`` therefore draws every parameter
from a documented distribution that exercises each term; ``mode='reference'`` reproduces
the declared initialisers (networks_stylegan.py:94-100,279-281; seg_solver.py:38).
"""
from __future__ import annotations

import numpy as np

from .naming import generator_param_shapes, decoder_param_shapes


def _blur_kernel(c):
    """networks_stylegan.py:211-226: outer([1,2,1])/16 repeated per channel."""
    k = np.array([1., 2., 1.], np.float32)
    k = np.outer(k, k)
    k = k / k.sum()
    return np.tile(k.reshape(1, 1, 3, 3), (c, 1, 1, 1)).astype(np.float32)


def wscale_std(name, shape):
    """Equalised-LR constant of a DenseW/_ConvW layer (networks_stylegan.py:397-403, 507-511)."""
    if name.startswith('mapping.'):
        return float(np.sqrt(2) / np.sqrt(shape[1]))                 # gain sqrt2, fan_in=in_units
    if '.affine.' in name:
        return float(1.0 / np.sqrt(shape[1]))                        # gain 1 (:244)
    if name.startswith('to_rgb'):
        return float(1.0 / np.sqrt(shape[1] * shape[2] * shape[3]))  # gain 1 (:123)
    if '.block0.' in name and shape[2] == 4:
        # Deconvolution weight is (Cin, Cout, 4, 4); fan_in = kh*kw*in_channels (:399-401)
        return float(np.sqrt(2) / np.sqrt(shape[2] * shape[3] * shape[0]))
    return float(np.sqrt(2) / np.sqrt(shape[1] * shape[2] * shape[3]))


def init_generator_params(cfg, seed=0, mode='nondegenerate', psi=None):
    """Return {structural name: float32 ndarray}.

    nondegenerate (documented test distribution):
      conv / dense / affine / to_rgb weights ~ N(0,1) (mapping: N(0,1)/lr_mult, as the official
      StyleGAN's 1/lrmul init, so the effective mapping weight is He-scaled N(0,1));
      mapping bias ~ N(0,10^2) (effective N(0,0.1^2) after lr_mult);  affine bias, Bias layers,
      to_rgb bias ~ N(0,0.1^2);  noise scale_factors ~ N(0,0.2^2);  latent_avg ~ N(0,0.1^2);
      to_rgb weight ~ N(0,0.3^2) so the image fills the [-1,1] range like a trained generator's output
      (std ~0.5) instead of saturating;
      constant_tensor ~ N(0,1);  truncation_psi = 0.7 on the first 8 layers and 1.0 after
      (truncation_cutoff=8 convention) unless ``psi`` is given.
    """
    rng = np.random.RandomState(seed)
    shapes = generator_param_shapes(cfg)
    out = {}
    for name, shape in shapes.items():
        if name.endswith('.std'):
            wname = name[:-4] + '.weight'
            out[name] = np.array([wscale_std(wname, shapes[wname])], np.float32)
        elif name.endswith('.w_kernel'):
            out[name] = _blur_kernel(shape[0])
        elif name.endswith('.gamma'):
            out[name] = np.ones(shape, np.float32)
        elif name.endswith('.beta'):
            out[name] = np.zeros(shape, np.float32)
        elif name == 'constant_tensor':
            out[name] = rng.randn(*shape).astype(np.float32)
        elif name == 'latent_avg':
            out[name] = (0.1 * rng.randn(*shape)).astype(np.float32) if mode == 'nondegenerate' \
                else np.zeros(shape, np.float32)
        elif name == 'truncation_psi':
            v = np.ones(shape, np.float32)
            if mode == 'nondegenerate':
                v[:8] = 0.7
            if psi is not None:
                v[:] = np.asarray(psi, np.float32)
            out[name] = v
        elif name.endswith('.weight'):
            w = rng.randn(*shape).astype(np.float32)
            if name.startswith('mapping.') and mode == 'nondegenerate':
                w = w * 100.0
            if name.startswith('to_rgb') and mode == 'nondegenerate':
                w = w * 0.3
            out[name] = w
        elif name.endswith('scale_factors'):
            out[name] = (0.2 * rng.randn(*shape)).astype(np.float32) if mode == 'nondegenerate' \
                else np.zeros(shape, np.float32)
        elif name.endswith('.bias'):
            if mode != 'nondegenerate':
                out[name] = np.zeros(shape, np.float32)
            elif name.startswith('mapping.'):
                out[name] = (10.0 * rng.randn(*shape)).astype(np.float32)
            else:
                out[name] = (0.1 * rng.randn(*shape)).astype(np.float32)
        else:
            raise KeyError(name)
    return out


def init_decoder_params(cfg, seed=2, mode='nondegenerate'):
    """Return {structural name: float32 ndarray}.

    Weights: Xavier(factor_type='in', magnitude=2.34), uniform (seg_solver.py:38):
    U(+-sqrt(2.34/fan_in)), fan_in = Cin*kh*kw.  reference mode: biases 0, fresh BatchNorm
    (gamma 1, beta 0, mean 0, var 1).  nondegenerate: conv bias ~ N(0,0.05^2), gamma ~
    U(0.5,1.5), beta, running_mean ~ N(0,0.1^2), running_var ~ U(0.5,1.5) so that the
    BN fold is exercised.
    """
    rng = np.random.RandomState(seed)
    out = {}
    nd = mode == 'nondegenerate'
    for name, shape in decoder_param_shapes(cfg).items():
        if name.endswith('.weight'):
            fan_in = shape[1] * shape[2] * shape[3]
            b = np.sqrt(2.34 / fan_in)
            out[name] = rng.uniform(-b, b, size=shape).astype(np.float32)
        elif name.endswith('.bias'):
            out[name] = (0.05 * rng.randn(*shape)).astype(np.float32) if nd else np.zeros(shape, np.float32)
        elif name.endswith('.gamma'):
            out[name] = rng.uniform(0.5, 1.5, size=shape).astype(np.float32) if nd else np.ones(shape, np.float32)
        elif name.endswith('.beta') or name.endswith('.running_mean'):
            out[name] = (0.1 * rng.randn(*shape)).astype(np.float32) if nd else np.zeros(shape, np.float32)
        elif name.endswith('.running_var'):
            out[name] = rng.uniform(0.5, 1.5, size=shape).astype(np.float32) if nd else np.ones(shape, np.float32)
        else:
            raise KeyError(name)
    return out
