"""Drop-in mirror of the reference's generator driver (image_generator.py:6-124) on the C-ABI path."""
from __future__ import annotations

import numpy as np
import torch

from .config import generator_config, MAX_RES_LOG2
from .networks import Generator


class ImageGenerator:
    """``ImageGenerator(gpu_ids, gan_dir, gan, batch_size, return_latents)`` as in image_generator.py:8.

    Differences, all additive: ``params=`` passes a {name: array} dict instead of reading
    ``{gan_dir}/stylegan-{gan}.params`` (the pretrained files are not available offline);
    ``base_scale`` exposes Generator's non-square base (networks_stylegan.py:84-85);
    ``get_images`` takes keyword-only ``psi, noise, seed, return_features, device_outputs``.
    ``gpu_ids=[]`` meant CPU in the reference (:17); here it raises -- there is no CPU path.
    """

    def __init__(self, gpu_ids, gan_dir, gan='ffhq', batch_size=4, return_latents=False, *, params=None,
                 base_scale=(4, 4)):
        self.max_res_log2 = MAX_RES_LOG2[gan]
        self.latent_size = 512
        self.return_latents = return_latents
        self.batch_size = batch_size
        if len(gpu_ids) == 0:
            raise RuntimeError('ImageGenerator needs at least one GPU id: the B200 path has no CPU fallback')
        self.ctx = [torch.device('cuda', i) for i in gpu_ids]
        self.cfg = self._get_config(max_res_log2=self.max_res_log2, base_scale=base_scale)
        self.netG = self._get_G(self.cfg, self.ctx)
        if params is None:
            from .params_io import load_params
            params = load_params(f'{gan_dir}/stylegan-{gan}.params')
        for g in self.netG:
            g.set_parameters(params)            # ignore_extra=True semantics (:22)

    def _get_G(self, config, ctx, initialize=False):
        return [Generator(config, device=d) for d in ctx]

    def _get_config(self, max_res_log2=9, base_scale=(4, 4)):
        return generator_config(max_res_log2, base_scale[0], base_scale[1])

    def _split(self, n):
        """Contiguous slices of the batch, one per context, sizes differing by at most one.  (The reference's
        split_and_load(even_split=False) (:95) gives the whole remainder to the LAST slice instead; the results are
        the same either way because latents and noise are keyed by the global sample index, not by the split.)"""
        k = len(self.ctx)
        base, rem = divmod(n, k)
        sizes = [base + (1 if i < rem else 0) for i in range(k)]
        out, o = [], 0
        for s in sizes:
            out.append((o, o + s))
            o += s
        return out

    def get_images(self, n, *, psi=None, noise=None, seed=None, return_features=True, device_outputs=False):
        """Python generator yielding ``n`` tuples ``(img uint8 [H,W,3], [feat_i float32 [C_i,H_i,W_i]])``
        (plus the batch's latents when ``return_latents``, the reference's quirk at :121-122).
        ``device_outputs=True`` yields cuda tensors instead of numpy arrays."""
        n_batches = n // self.batch_size + (1 if n % self.batch_size > 0 else 0)
        n_generated = 0
        rng = np.random if seed is None else np.random.RandomState(seed)
        # the noise planes AddNoise samples per call (networks_stylegan.py:300) come from the device Philox stream
        # keyed by (noise_seed, running sample index): fresh per sample, independent of the context split
        noise_seed = int(rng.randint(0, 2 ** 31 - 1))
        for _ in range(n_batches):
            bs = min(self.batch_size, n - n_generated)
            latent_z = rng.standard_normal((bs, self.latent_size)).astype(np.float32)
            outs = []
            for g, (a, b) in zip(self.netG, self._split(bs)):
                if b == a:
                    continue
                nz = None if noise is None else [p[n_generated + a:n_generated + b] for p in noise]
                outs.append(g.forward(latent_z[a:b], psi=psi, noise=nz, seed=noise_seed, first_sample=n_generated + a,
                                      return_features=return_features, return_image=False, return_u8=True))
            for g in self.netG:
                torch.cuda.synchronize(g.device)                  # mx.nd.waitall() (:102)
            if device_outputs:
                imgs = torch.cat([o['img_u8'].to(self.ctx[0]) for o in outs], 0)
                feats = [torch.cat([o['features'][i].to(self.ctx[0]) for o in outs], 0)
                         for i in range(len(outs[0]['features']))] if return_features else []
            else:
                imgs = np.concatenate([o['img_u8'].cpu().numpy() for o in outs], 0)
                feats = [np.concatenate([o['features'][i].cpu().numpy() for o in outs], 0)
                         for i in range(len(outs[0]['features']))] if return_features else []
            n_generated += imgs.shape[0]
            for i in range(imgs.shape[0]):
                f = [ft[i] for ft in feats]
                if self.return_latents:
                    yield imgs[i], f, latent_z
                else:
                    yield imgs[i], f
