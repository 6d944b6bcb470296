"""Host mirrors of the reference's two network objects over the C ABI.

``Generator(config)(z)``   <- networks_stylegan.py:76-197  (returns (img [N,3,H,W], [features]))
``Decoder(cfg, nd)(*feats)`` <- networks_seg.py:49-113      (returns logits [N,classes,H,W])

plus the fused device-resident call ``generate`` (image uint8 + mask uint8, features never leave HBM).
torch owns device memory / streams only; all arithmetic happens in libgsx.so.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
import re

import numpy as np
import torch

from . import _lib as L
from .config import num_features, block_hw, num_style_layers
from .naming import canonical, generator_param_shapes, decoder_param_shapes


def _stream(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _set_params(setter, handle, params, known, what, dtype=None):
    extra = []
    for name, arr in params.items():
        cname = canonical(name)
        if cname not in known:
            extra.append(name)          # load_parameters(ignore_extra=True), image_generator.py:22
            continue
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float32))
        if tuple(a.shape) != tuple(known[cname]) and a.size != int(np.prod(known[cname])):
            raise ValueError(f'{what}: parameter {name} has shape {a.shape}, expected {known[cname]}')
        shape = (C.c_int64 * a.ndim)(*a.shape)
        L.check(setter(handle, cname.encode(), L.np_ptr(a), shape, a.ndim), f'{what}.set_param({name})', dtype)
    return extra


class Generator:
    """StyleGAN-v1 generator (mapping + truncation + synthesis + ToRGB) on one GPU."""

    def __init__(self, config, device=None, dtype=None):
        self.cfg = dict(config)
        self.dtype = dtype or L.DEFAULT_DTYPE
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_res_log2 = config['max_res_log2']
        self.latent_size = config['latent_size']
        self.num_layers = num_style_layers(config)
        self._lib = L.lib(self.dtype)
        c = L.SynthCfg(config['max_res_log2'], config['base_scale_y'], config['base_scale_x'], config['fmap_base'],
                       float(config['fmap_decay']), config['fmap_max'], config['latent_size'], config['channels'])
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self._lib.gsx_synth_create(C.byref(c), C.byref(h)), 'gsx_synth_create', self.dtype)
        self._h = h
        self._ws = None
        self._ws_n = 0
        self._shapes = generator_param_shapes(config)
        self.out_hw = block_hw(config, self.max_res_log2)

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                self._lib.gsx_synth_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- parameters ------------------------------------------------------------------------------
    def set_parameters(self, params):
        """params: {reference name (structural or legacy): float32 array}.  Extra names are ignored."""
        with torch.cuda.device(self.device):
            extra = _set_params(self._lib.gsx_synth_set_param, self._h, params, self._shapes, 'Generator', self.dtype)
            L.check(self._lib.gsx_synth_finalize(self._h), 'gsx_synth_finalize', self.dtype)
        return extra

    def load_parameters(self, filename, ignore_extra=True, ctx=None):
        """Gluon-compatible entry (image_generator.py:22): reads an MXNet .params file."""
        from .params_io import load_params
        extra = self.set_parameters(load_params(filename))
        if extra and not ignore_extra:
            raise ValueError(f'unexpected parameters in {filename}: {extra[:5]}')

    # -- workspace -------------------------------------------------------------------------------
    def workspace(self, n):
        if self._ws is None or self._ws_n < n:
            sz = C.c_size_t()
            L.check(self._lib.gsx_synth_workspace_bytes(self._h, n, C.byref(sz)), 'workspace_bytes', self.dtype)
            self._ws = torch.empty(sz.value, dtype=torch.uint8, device=self.device)
            self._ws_n = n
        return self._ws

    def feature_shapes(self, n):
        return [(n, num_features(self.cfg, r)) + block_hw(self.cfg, r) for r in range(2, self.max_res_log2 + 1)]

    # -- forward ---------------------------------------------------------------------------------
    def forward(self, z=None, n=None, psi=None, noise=None, seed=0, first_sample=0, return_features=True,
                return_image=True, return_u8=False, stream=None):
        """z: [N,512] float32 (cuda tensor / numpy) or None (Philox(seed, first_sample+i), then ``n`` is
        required).  psi: None (parameter), scalar or per-layer vector.  noise: None (Philox) or a list of
        num_layers [N,1,h,w] float32 arrays.  Returns dict(img, img_u8, features) of cuda tensors."""
        with torch.cuda.device(self.device):
            if z is not None:
                z = torch.as_tensor(z, dtype=torch.float32).to(self.device).contiguous()
                n = z.shape[0]
            if n is None:
                raise ValueError('need z or n')
            ws = self.workspace(n)
            # the workspace layout depends on the batch the buffer is carved for
            psi_arr = None
            if psi is not None:
                psi_arr = np.ascontiguousarray(np.broadcast_to(np.asarray(psi, np.float32), (self.num_layers,)))
            noise_ptrs = None
            keep = []
            if noise is not None:
                if len(noise) != self.num_layers:
                    raise ValueError(f'expected {self.num_layers} noise planes')
                noise_ptrs = (C.c_void_p * self.num_layers)()
                for i, a in enumerate(noise):
                    t = torch.as_tensor(a, dtype=torch.float32).to(self.device).contiguous()
                    keep.append(t)
                    noise_ptrs[i] = t.data_ptr()
            H, W = self.out_hw
            nc = self.cfg['channels']
            img = torch.empty((n, nc, H, W), dtype=torch.float32, device=self.device) if return_image else None
            u8 = torch.empty((n, H, W, nc), dtype=torch.uint8, device=self.device) if return_u8 else None
            feats = None
            feat_ptrs = None
            if return_features:
                feats = [torch.empty(s, dtype=torch.float32, device=self.device) for s in self.feature_shapes(n)]
                feat_ptrs = (C.c_void_p * len(feats))(*[f.data_ptr() for f in feats])
            rc = self._lib.gsx_synth_forward(self._h, n, L.ptr(z), L.np_ptr(psi_arr), noise_ptrs, seed, first_sample,
                                             L.ptr(img), L.ptr(u8), feat_ptrs, L.ptr(ws), ws.numel(), _stream(stream))
            L.check(rc, 'gsx_synth_forward', self.dtype)
            self._last_n = n
            return dict(img=img, img_u8=u8, features=feats)

    def __call__(self, z, **kw):
        out = self.forward(z, **kw)
        return out['img'], out['features']

    def export_noise(self, n):
        """The noise planes the last forward used (device-generated when none were passed)."""
        ws = self.workspace(n)
        planes = []
        for l in range(self.num_layers):
            h, w = block_hw(self.cfg, 2 + l // 2)
            t = torch.empty((n, 1, h, w), dtype=torch.float32, device=self.device)
            L.check(self._lib.gsx_synth_export_noise(self._h, n, l, L.ptr(t), L.ptr(ws), _stream()), 'export_noise', self.dtype)
            planes.append(t)
        return planes

    def export_latents(self, n):
        ws = self.workspace(n)
        t = torch.empty((n, self.latent_size), dtype=torch.float32, device=self.device)
        L.check(self._lib.gsx_synth_export_latents(self._h, n, L.ptr(t), L.ptr(ws), _stream()), 'export_latents', self.dtype)
        return t


class Decoder:
    """Segmentation decoder over the generator's feature pyramid, fused with the argmax label map."""

    def __init__(self, cfg, num_devices=1, base_hw=(4, 4), device=None, dtype=None):
        self.cfg = dict(cfg)
        self.dtype = dtype or L.DEFAULT_DTYPE
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = L.lib(self.dtype)
        if cfg.get('use_sync_bn', False):
            raise NotImplementedError('SyncBatchNorm (off in the reference config, seg_solver.py:120)')
        # start_res = s (networks_seg.py:55,63,80,107): levels below s have no blocks and their feature maps are ignored.
        # The library sees the decoder of the remaining levels (its level 0 = reference level s, at base_hw << s);
        # parameter names keep the reference's level numbers and are renumbered on the way in.
        s0 = self.start_res = int(cfg.get('start_res', 0))
        nf_all = len(cfg['in_channels'])
        if not 0 <= s0 < nf_all:
            raise ValueError(f'start_res {s0} outside the {nf_all} feature levels')
        nf = nf_all - s0
        c = L.DecCfg()
        c.num_levels = nf
        for i, v in enumerate(cfg['in_channels'][s0:]):
            c.in_channels[i] = v
        for i, v in enumerate(cfg['features'][s0:]):
            c.features[i] = v
        c.use_bn = int(bool(cfg['use_bn']))
        c.base_y, c.base_x = base_hw[0] << s0, base_hw[1] << s0
        self.base_hw = (base_hw[0] << s0, base_hw[1] << s0)
        self.num_levels = nf
        self.num_levels_all = nf_all
        self.num_classes = cfg['features'][-1]
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self._lib.gsx_dec_create(C.byref(c), C.byref(h)), 'gsx_dec_create', self.dtype)
        self._h = h
        self._ws = None
        self._ws_n = 0
        self._shapes = decoder_param_shapes(cfg)
        self.out_hw = (self.base_hw[0] << (nf - 1), self.base_hw[1] << (nf - 1))

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                self._lib.gsx_dec_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _lib_name(self, cname):
        # reference level number -> the library's (levels renumbered from start_res)
        return re.sub(r'^(cvt_block|main_block)_(\d+)', lambda m: f'{m.group(1)}_{int(m.group(2)) - self.start_res}', cname)

    def set_parameters(self, params):
        with torch.cuda.device(self.device):
            setter = self._lib.gsx_dec_set_param
            if self.start_res:
                setter = lambda h, name, *rest: self._lib.gsx_dec_set_param(h, self._lib_name(name.decode()).encode(), *rest)
            extra = _set_params(setter, self._h, params, self._shapes, 'Decoder', self.dtype)
            L.check(self._lib.gsx_dec_finalize(self._h), 'gsx_dec_finalize', self.dtype)
        self._params = {canonical(k): np.asarray(v, np.float32) for k, v in params.items() if canonical(k) in self._shapes}
        return extra

    def get_parameters(self):
        return dict(self._params)

    def load_parameters(self, filename, ctx=None):
        from .params_io import load_params
        self.set_parameters(load_params(filename))

    def save_parameters(self, filename):
        from .params_io import save_params
        save_params(filename, self._params)

    def workspace(self, n):
        if self._ws is None or self._ws_n < n:
            sz = C.c_size_t()
            L.check(self._lib.gsx_dec_workspace_bytes(self._h, n, C.byref(sz)), 'dec workspace_bytes', self.dtype)
            self._ws = torch.empty(sz.value, dtype=torch.uint8, device=self.device)
            self._ws_n = n
        return self._ws

    def forward(self, features=None, generator=None, n=None, return_logits=True, stream=None):
        """features: list of [N,C_i,H_i,W_i] float32 (cuda/numpy), or None with ``generator`` = the Generator
        whose last forward(n) left its features in HBM.  Returns dict(logits, mask[N,H,W] uint8)."""
        with torch.cuda.device(self.device):
            keep = []
            feat_ptrs = None
            gh, gws = None, None
            if features is not None:
                if len(features) != self.num_levels_all:
                    raise ValueError(f'expected {self.num_levels_all} feature maps')
                feat_ptrs = (C.c_void_p * self.num_levels)()
                for i, f in enumerate(features[self.start_res:]):
                    t = torch.as_tensor(f, dtype=torch.float32).to(self.device)
                    if t.dim() == 3:
                        t = t.unsqueeze(0)                         # seg_solver.py:313-314
                    t = t.contiguous()
                    keep.append(t)
                    feat_ptrs[i] = t.data_ptr()
                n = keep[0].shape[0]
            else:
                if generator is None:
                    raise ValueError('need features or generator')
                if self.start_res:
                    raise NotImplementedError('start_res != 0 with generator-resident features: pass the feature list')
                n = generator._last_n if n is None else n
                gh, gws = generator._h, L.ptr(generator.workspace(n))
            ws = self.workspace(n)
            H, W = self.out_hw
            logits = torch.empty((n, self.num_classes, H, W), dtype=torch.float32, device=self.device) if return_logits else None
            mask = torch.empty((n, H, W), dtype=torch.uint8, device=self.device)
            rc = self._lib.gsx_dec_forward(self._h, n, feat_ptrs, gh, gws, L.ptr(logits), L.ptr(mask), L.ptr(ws), ws.numel(),
                                           _stream(stream))
            L.check(rc, 'gsx_dec_forward', self.dtype)
            return dict(logits=logits, mask=mask)

    def __call__(self, *features):
        return self.forward(list(features))['logits']


class GeneratePipeline:
    """z (host) -> uint8 image + uint8 mask (host) through gsx_generate_host: the whole ``main.py generate``
    inner loop (main.py:97-99) for one batch, H2D/D2H included, features resident in HBM.  With
    ``overlap=True`` the device-to-host copies run on a second stream into alternating host slots and
    overlap the next batch's kernels."""

    def __init__(self, generator, decoder, n, overlap=True):
        self.g, self.d, self.n = generator, decoder, n
        dev = generator.device
        H, W = generator.out_hw
        nc = generator.cfg['channels']
        self.gws = generator.workspace(n)
        self.dws = decoder.workspace(n)
        self.overlap = overlap
        slots = 2 if overlap else 1
        per_slot = n * (generator.latent_size * 4 + H * W * (nc + 1)) + 4096
        self.stage = torch.empty(per_slot * slots, dtype=torch.uint8, device=dev)
        # image and mask of a slot share ONE pinned buffer (mask right after the image, 1 KiB-aligned like the device
        # staging area), so that they leave the device as a single copy
        ib = (n * H * W * nc + 1023) // 1024 * 1024
        self._out_host = [torch.empty(ib + n * H * W, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.img_host = [b[:n * H * W * nc].view(n, H, W, nc) for b in self._out_host]
        self.mask_host = [b[ib:ib + n * H * W].view(n, H, W) for b in self._out_host]
        self.z_host = [torch.empty((n, generator.latent_size), dtype=torch.float32).pin_memory() for _ in range(slots)]
        with torch.cuda.device(dev):
            self.copy_stream = torch.cuda.Stream() if overlap else None
        self.h2d_bytes = self.z_host[0].numel() * 4
        self.d2h_bytes = self.img_host[0].numel() + self.mask_host[0].numel()
        self._calls = 0

    def run(self, z=None, psi=None, seed=0, first_sample=0, stream=None):
        """Enqueues one batch and returns the slot index; ``img_host[slot]`` / ``mask_host[slot]`` are valid
        after ``wait()`` (or once the same slot comes round again)."""
        lib = self.g._lib
        slot = self._calls % len(self.img_host)
        self._calls += 1
        zp = None
        if z is not None:
            self.z_host[slot].copy_(torch.as_tensor(z, dtype=torch.float32))
            zp = L.ptr(self.z_host[slot])
        psi_arr = None
        if psi is not None:
            psi_arr = np.ascontiguousarray(np.broadcast_to(np.asarray(psi, np.float32), (self.g.num_layers,)))
        cs = C.c_void_p(self.copy_stream.cuda_stream) if self.copy_stream is not None else None
        with torch.cuda.device(self.g.device):
            rc = lib.gsx_generate_host(self.g._h, self.d._h, self.n, zp, L.np_ptr(psi_arr), seed, first_sample,
                                       L.ptr(self.img_host[slot]), L.ptr(self.mask_host[slot]), L.ptr(self.gws),
                                       self.gws.numel(), L.ptr(self.dws), self.dws.numel(), L.ptr(self.stage),
                                       self.stage.numel(), _stream(stream), cs, slot)
        L.check(rc, 'gsx_generate_host', self.g.dtype)
        self.g._last_n = self.n
        return slot

    def wait(self):
        if self.copy_stream is not None:
            self.copy_stream.synchronize()
        torch.cuda.current_stream(self.g.device).synchronize()


class GraphedGenerate:
    """One device-resident generate step (Philox latents -> uint8 image + uint8 mask in HBM) captured in a CUDA graph.

    At small batches the ~100 launches of a step are launch-bound (256^2, batch 1: 1.2 ms of host-paced launches for
    ~0.4 ms of GPU work); replaying the captured step removes the host from the loop.  Fresh samples per replay come
    from the device-resident sample counter (``gsx_synth_device_counter``): replay k generates the samples with
    global index ``first_sample + k*n ...``, bit-identical to ``Generator.forward(n=n, seed=seed, first_sample=...)``.
    """

    def __init__(self, generator, decoder, n, seed=0, first_sample=0):
        self.g, self.d, self.n = generator, decoder, n
        dev = generator.device
        H, W = generator.out_hw
        nc = generator.cfg['channels']
        lib = generator._lib
        self.gws = generator.workspace(n)
        self.dws = decoder.workspace(n)
        self.img = torch.empty((n, H, W, nc), dtype=torch.uint8, device=dev)
        self.mask = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
        self.seed = seed

        def enqueue():
            sp = _stream(None)
            # the fused call: the image pass and the decoder's cvt / shortcut branches become branches of the captured graph
            L.check(lib.gsx_generate_dev(generator._h, decoder._h, n, None, None, seed, 0, L.ptr(self.img), L.ptr(self.mask),
                                         L.ptr(self.gws), self.gws.numel(), L.ptr(self.dws), self.dws.numel(), sp),
                    'gsx_generate_dev', generator.dtype)

        with torch.cuda.device(dev):
            L.check(lib.gsx_synth_device_counter(generator._h, 1, first_sample), 'gsx_synth_device_counter', generator.dtype)
            enqueue()                                  # warm-up outside the capture (per-device kernel attributes)
            torch.cuda.synchronize(dev)
            L.check(lib.gsx_synth_device_counter(generator._h, 1, first_sample), 'gsx_synth_device_counter', generator.dtype)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                enqueue()
        generator._last_n = n

    def replay(self):
        """Enqueues one step on the current stream; returns (img uint8 [N,H,W,C], mask uint8 [N,H,W]) device tensors
        that the next replay overwrites."""
        self.graph.replay()
        return self.img, self.mask

    def close(self):
        """Back to the ``first_sample`` argument of the plain calls."""
        L.check(self.g._lib.gsx_synth_device_counter(self.g._h, 0, 0), 'gsx_synth_device_counter', self.g.dtype)
