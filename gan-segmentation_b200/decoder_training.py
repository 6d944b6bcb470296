"""Decoder training step (reference seg_solver.py:351-421, generator frozen) on this repo's own kernels.

First, functional version of SURVEY row a19: forward in train mode (BatchNorm with batch statistics, LeakyReLU,
Dropout after every cvt block), SoftmaxCE with the reference's sample weights, hand-derived backward pass, one
flat-bucket gradient all-reduce, fused MXNet-Adam step.  Every tensor operation goes through a small backend interface:

  ``CudaBackend``  -- the C-ABI kernels: the tcgen05 shift-GEMM conv for forward and data gradients
                      (``training.dgrad_weights``), ``gsx_op_conv_wgrad``, the BatchNorm / upsample kernels of
                      ``csrc/train.cu``, ``gsx_softmax_ce``, ``gsx_adam_step``.  It goes through the single-operator
                      hooks (fp32 NCHW in and out, temporaries per call): correct, not yet fast.
  a torch backend  -- lives in ``tests/`` only; it exists to check the orchestration (and the backward formulas) against
                      the autograd oracle ``oracle/train_oracle.py`` on the CPU.  The product never falls back to it.
"""
from __future__ import annotations

import numpy as np
import torch

from .naming import decoder_param_shapes

LEARNABLE = ('.weight', '.bias', '.gamma', '.beta')


class DecoderTrainer:
    """``step(feats, mask)`` = one iteration of the reference fit loop on one rank."""

    def __init__(self, cfg, params, backend, base_lr=None, wd=None):
        self.cfg = cfg
        self.be = backend
        self.nf = len(cfg['in_channels'])
        self.use_bn = bool(cfg['use_bn'])
        self.use_dropout = bool(cfg.get('use_dropout', False))
        self.lr = cfg['base_lr'] if base_lr is None else base_lr
        self.wd = (cfg.get('wd', 0.0) or 0.0) if wd is None else wd
        self.P = {k: backend.tensor(np.asarray(v, np.float32)) for k, v in params.items()}
        self.learnable = [k for k in self.P if k.endswith(LEARNABLE)]
        self.opt = backend.make_adam({k: tuple(self.P[k].shape) for k in self.learnable}, self.P, self.lr, self.wd)

    # ---- helpers ---------------------------------------------------------------------------------------------
    def _bn(self, prefix, z, drop, cache, key):
        be, P = self.be, self.P
        if self.use_bn:
            y, c = be.bn_lrelu_fwd(z, P[f'{prefix}.gamma'], P[f'{prefix}.beta'], drop)
            # moving statistics: 0.9 * old + 0.1 * batch (MXNet BatchNorm momentum 0.9, biased batch variance)
            self._new_stats[f'{prefix}.running_mean'] = 0.9 * P[f'{prefix}.running_mean'] + 0.1 * c['mean']
            self._new_stats[f'{prefix}.running_var'] = 0.9 * P[f'{prefix}.running_var'] + 0.1 * c['var']
        else:
            y, c = be.lrelu_fwd(z, drop)
        cache[key] = dict(z=z, c=c, drop=drop, prefix=prefix)
        return y

    def _bn_bwd(self, dy, cache, key, grads):
        be, P = self.be, self.P
        e = cache[key]
        if self.use_bn:
            p = e['prefix']
            dz, dg, db = be.bn_lrelu_bwd(dy, e['z'], e['c'], P[f'{p}.gamma'], P[f'{p}.beta'], e['drop'])
            grads[f'{p}.gamma'] = dg
            grads[f'{p}.beta'] = db
            return dz
        return be.lrelu_bwd(dy, e['z'], e['drop'])

    # ---- forward + backward ----------------------------------------------------------------------------------
    def loss_and_grads(self, feats, mask, dropout_masks=None):
        """feats: list of [N,C_i,H_i,W_i]; mask [N,1,H,W] int in {-1,0..K-1}.  Returns (per-sample loss, grads dict);
        the gradients are ``self._grad_scale`` times d(sum of the per-sample losses)/d(parameter)."""
        be, P, nf = self.be, self.P, self.nf
        self._new_stats = {}
        feats = [be.tensor(f) for f in feats]
        cache, grads = {}, {}
        lv = []                                       # per level: what backward needs
        prev = None
        for i in range(nf):
            cv = f'cvt_block_{i}'
            z = be.conv([feats[i]], P[f'{cv}.0.weight'], P[f'{cv}.0.bias'], 3)
            drop = None
            if self.use_dropout and dropout_masks is not None and dropout_masks[i] is not None:
                drop = be.tensor(dropout_masks[i])
            c = self._bn(f'{cv}.1', z, drop, cache, ('cvt', i))
            xin = [prev, c] if i > 0 else [c]         # concat(prev, cvt) (networks_seg.py:108-109)
            if i < nf - 1:
                p = f'main_block_{i}.1'
                jb = 3 if self.use_bn else 2
                za = be.upconv(xin, P[f'{p}.base_layers.0.weight'], P[f'{p}.base_layers.0.bias'])
                a = self._bn(f'{p}.base_layers.1', za, None, cache, ('a', i))
                zb = be.conv([a], P[f'{p}.base_layers.{jb}.weight'], P[f'{p}.base_layers.{jb}.bias'], 3)
                bb = self._bn(f'{p}.base_layers.{jb + 1}', zb, None, cache, ('b', i))
                has_sc = f'{p}.shortcut.0.weight' in P
                if has_sc:                            # 1x1 conv commutes with the nearest upsampling in front of it
                    sc_lo = be.conv(xin, P[f'{p}.shortcut.0.weight'], P[f'{p}.shortcut.0.bias'], 1)
                else:
                    sc_lo = xin[0]
                prev_new = be.upsample2(sc_lo) + bb
                lv.append(dict(xin=xin, a=a, has_sc=has_sc, jb=jb, p=p))
                prev = prev_new
            else:
                p = f'main_block_{i}.0'
                logits = be.conv(xin, P[f'{p}.weight'], P[f'{p}.bias'], 3)
                lv.append(dict(xin=xin, p=p))
        # the backend may return the gradient multiplied by a power-of-two-ish scale (H*W on the CUDA backend: its data
        # gradients travel through 16-bit tensors); every later step is linear in it, the optimizer divides it out
        loss, dlog, self._grad_scale = be.softmax_ce(logits, mask)

        # backward
        d_prev = None
        for i in reversed(range(nf)):
            e = lv[i]
            xin = e['xin']
            c_split = [t.shape[1] for t in xin]
            if i == nf - 1:
                w = P[f"{e['p']}.weight"]
                dxin = be.conv_dgrad(dlog, w, 3)
                self._wgrad(grads, f"{e['p']}", xin, dlog, 3)
            else:
                p, jb = e['p'], e['jb']
                d_out = d_prev                                        # gradient of sc_up + bb
                dzb = self._bn_bwd(d_out, cache, ('b', i), grads)
                self._wgrad(grads, f'{p}.base_layers.{jb}', [e['a']], dzb, 3)
                da = be.conv_dgrad(dzb, P[f'{p}.base_layers.{jb}.weight'], 3)
                dza = self._bn_bwd(da, cache, ('a', i), grads)
                up_in = [be.upsample2(t) for t in xin]                # the conv saw the upsampled concat
                self._wgrad(grads, f'{p}.base_layers.0', up_in, dza, 3)
                dxin = be.sumpool2(be.conv_dgrad(dza, P[f'{p}.base_layers.0.weight'], 3))
                d_sc_lo = be.sumpool2(d_out)
                if e['has_sc']:
                    self._wgrad(grads, f'{p}.shortcut.0', xin, d_sc_lo, 1)
                    dxin = dxin + be.conv_dgrad(d_sc_lo, P[f'{p}.shortcut.0.weight'], 1)
                else:
                    dxin = dxin + d_sc_lo
            if len(xin) == 2:
                d_prev, d_c = dxin[:, :c_split[0]], dxin[:, c_split[0]:]
            else:
                d_prev, d_c = None, dxin
            cv = f'cvt_block_{i}'
            dz = self._bn_bwd(d_c.contiguous(), cache, ('cvt', i), grads)
            self._wgrad(grads, f'{cv}.0', [feats[i]], dz, 3)          # no gradient into the (frozen) generator
            if d_prev is not None:
                d_prev = d_prev.contiguous()
        return loss, grads

    def _wgrad(self, grads, prefix, xs, dy, k):
        dws, db = [], None
        for x in xs:
            dw, db = self.be.conv_wgrad(x, dy, k)
            dws.append(dw)
        grads[f'{prefix}.weight'] = dws[0] if len(dws) == 1 else self.be.cat(dws, 1)
        grads[f'{prefix}.bias'] = db

    def step(self, feats, mask, dropout_masks=None, global_batch=None, group=None):
        """Forward, backward, gradient all-reduce (sum over ranks), Adam with rescale_grad = 1/global batch
        (trainer.step(batch), seg_solver.py:421).  Returns the per-sample loss of this rank."""
        loss, grads = self.loss_and_grads(feats, mask, dropout_masks)
        n = int(np.asarray(mask).shape[0]) if not torch.is_tensor(mask) else int(mask.shape[0])
        self.opt.apply(grads, global_batch or n, group, self._grad_scale)
        for k, v in self._new_stats.items():
            self.P[k] = v
        return loss

    def state(self):
        """name -> numpy float32, the reference's structural parameter names."""
        out = self.opt.export(self.P)
        return {k: np.asarray(self.be.numpy(v), np.float32) for k, v in out.items()}


class CudaBackend:
    """The backend of the product: every operation is one of this repo's CUDA kernels, reached through the C ABI's
    single-operator hooks (fp32 NCHW device tensors; 16-bit tensor-core operands inside the conv kernels)."""

    def __init__(self, device='cuda', dtype=None):
        from . import _lib as L, ops, training
        self.L, self.ops, self.tr = L, ops, training
        self.device = torch.device(device)
        self.dtype = dtype
        self.lib = L.lib(dtype)                       # raises without the CUDA extension: no CPU fallback

    # -- plumbing
    def tensor(self, a):
        t = a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a, np.float32))
        return t.to(self.device, torch.float32).contiguous()

    def numpy(self, t):
        return t.detach().cpu().numpy()

    def cat(self, ts, dim):
        return torch.cat(ts, dim)

    def _stream(self):
        import ctypes as C
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _check(self, rc, what):
        self.L.check(rc, what, self.dtype)

    # -- convolutions: the tcgen05 shift-GEMM kernel, forward and (with repacked weights) data gradient
    def _w(self, w):
        return w.detach().cpu().numpy()

    def conv(self, xs, w, b, k):
        L = self.L
        mode = L.CONV3 if k == 3 else L.CONV1
        cout = w.shape[0]
        x1 = xs[1] if len(xs) > 1 else None
        if cout % 16:                                  # the final conv: logits come out of the argmax epilogue
            b16 = torch.zeros(16, device=self.device)
            b16[:cout] = b
            r = self.ops.conv(mode, xs[0], self._w(w), x1=x1, bias=b16, flags=L.EPI_ARGMAX, num_classes=cout, dtype=self.dtype)
            self.last_mask = r['mask']                 # first-max argmax of the same logits (train metric, seg_solver.py:415-419)
            return r['logits']
        return self.ops.conv(mode, xs[0], self._w(w), x1=x1, bias=b, dtype=self.dtype)['out']

    def upconv(self, xs, w, b):
        x1 = xs[1] if len(xs) > 1 else None
        return self.ops.conv(self.L.UPCONV3, xs[0], self._w(w), x1=x1, bias=b, dtype=self.dtype)['out']

    def conv_dgrad(self, dy, w, k):
        wd = self.tr.dgrad_weights(self._w(w))         # [Cin, Cout, k, k]
        cout = dy.shape[1]
        if cout % 16:                                  # pad the gradient's channels (final conv: num_classes)
            pad = 16 - cout % 16
            dy = torch.cat([dy, torch.zeros((dy.shape[0], pad) + tuple(dy.shape[2:]), device=dy.device)], 1)
            wd = np.concatenate([wd, np.zeros((wd.shape[0], pad) + wd.shape[2:], np.float32)], 1)
        return self.ops.conv(self.L.CONV3 if k == 3 else self.L.CONV1, dy.contiguous(), wd, dtype=self.dtype)['out']

    def conv_wgrad(self, x, dy, k):
        cout = dy.shape[1]
        if cout % 8:
            pad = 8 - cout % 8
            dyp = torch.cat([dy, torch.zeros((dy.shape[0], pad) + tuple(dy.shape[2:]), device=dy.device)], 1)
            dw, db = self.tr.conv_wgrad(x, dyp, k, dtype=self.dtype)
            return dw[:cout].contiguous(), db[:cout].contiguous()
        return self.tr.conv_wgrad(x, dy, k, dtype=self.dtype)

    # -- elementwise / reductions (csrc/train.cu)
    def upsample2(self, x):
        n, c, h, w = x.shape
        y = torch.empty((n, c, 2 * h, 2 * w), dtype=torch.float32, device=self.device)
        self._check(self.lib.gsx_op_upsample2(self.L.ptr(x.contiguous()), self.L.ptr(y), n, c, h, w, self._stream()), 'gsx_op_upsample2')
        return y

    def sumpool2(self, dy):
        n, c, h2, w2 = dy.shape
        dx = torch.empty((n, c, h2 // 2, w2 // 2), dtype=torch.float32, device=self.device)
        self._check(self.lib.gsx_op_sumpool2(self.L.ptr(dy.contiguous()), self.L.ptr(dx), n, c, h2 // 2, w2 // 2, self._stream()), 'gsx_op_sumpool2')
        return dx

    def bn_lrelu_fwd(self, z, gamma, beta, drop):
        n, c, h, w = z.shape
        y = torch.empty_like(z)
        stats = torch.empty((3, c), dtype=torch.float32, device=self.device)
        self._check(self.lib.gsx_op_bn_lrelu_fwd(self.L.ptr(z), self.L.ptr(gamma), self.L.ptr(beta), self.L.ptr(drop), self.L.ptr(y),
                                                 self.L.ptr(stats), n, c, h * w, self._stream()), 'gsx_op_bn_lrelu_fwd')
        return y, dict(mean=stats[0], var=stats[1], rstd=stats[2], stats=stats)

    def bn_lrelu_bwd(self, dy, z, c, gamma, beta, drop):
        n, ch, h, w = z.shape
        dz = torch.empty_like(z)
        dparam = torch.empty((2, ch), dtype=torch.float32, device=self.device)
        self._check(self.lib.gsx_op_bn_lrelu_bwd(self.L.ptr(dy.contiguous()), self.L.ptr(z), self.L.ptr(c['stats']), self.L.ptr(gamma),
                                                 self.L.ptr(beta), self.L.ptr(drop), self.L.ptr(dz), self.L.ptr(dparam), n, ch, h * w,
                                                 self._stream()), 'gsx_op_bn_lrelu_bwd')
        return dz, dparam[1], dparam[0]

    def lrelu_fwd(self, z, drop):
        raise NotImplementedError('use_bn=False is not built for the CUDA training backend')

    def lrelu_bwd(self, dy, z, drop):
        raise NotImplementedError('use_bn=False is not built for the CUDA training backend')

    def softmax_ce(self, logits, mask):
        m = torch.as_tensor(np.asarray(mask) if not torch.is_tensor(mask) else mask).to(self.device)
        scale = float(logits.shape[2] * logits.shape[3])
        loss, dl = self.tr.softmax_ce(logits, m.int(), want_grad=True, dtype=self.dtype, grad_scale=scale)
        return loss, dl, scale

    # -- optimizer: one flat bucket, one all-reduce, one fused Adam kernel
    def make_adam(self, shapes, P, lr, wd):
        return _FlatAdamAdapter(self, shapes, P, lr, wd)


class _FlatAdamAdapter:
    def __init__(self, be, shapes, P, lr, wd):
        self.flat = be.tr.FlatAdam(shapes, lr=lr, wd=wd, device=be.device, dtype=be.dtype)
        self.flat.load({k: be.numpy(P[k]) for k in shapes})
        for k in shapes:                                 # the trainer's parameters become views of the bucket
            P[k] = self.flat.view(self.flat.w, k)
        self.names = list(shapes)

    def apply(self, grads, batch, group=None, grad_scale=1.0):
        for k in self.names:
            self.flat.view(self.flat.g, k).copy_(grads[k])
        self.flat.step(batch, group, grad_scale)

    def export(self, P):
        return dict(P)


class ResidentTrainer:
    """The decoder training step of the product: ONE C-ABI call per iteration (``gsx_train_step``, csrc/train_step.cu) --
    blocked 16-bit activations resident in a workspace, tcgen05 convolutions / data gradients / weight gradients, fused
    BatchNorm+LeakyReLU+Dropout kernels, no host synchronisation and no library kernels -- followed by the single all-reduce
    of the flat gradient bucket (``torch.distributed``: plumbing) and the fused Adam kernel (``gsx_adam_step``).

    ``step(feats, mask, dropout_seed)`` = one iteration of the reference fit loop on one rank (seg_solver.py:386-421)."""

    def __init__(self, cfg, params, n, device='cuda', base_hw=(4, 4), base_lr=None, wd=None, dtype=None,
                 beta1=0.9, beta2=0.999, eps=1e-8, use_graph=True):
        import ctypes as C
        from . import _lib as L
        self.L, self.C = L, C
        self.cfg = cfg
        self.n = int(n)
        self.device = torch.device(device)
        self.dtype = dtype
        self.lib = L.lib(dtype)
        self.lr = cfg['base_lr'] if base_lr is None else base_lr
        self.wd = (cfg.get('wd', 0.0) or 0.0) if wd is None else wd
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        self.t = 0
        if cfg.get('start_res', 0) != 0:
            raise NotImplementedError('training with start_res != 0 (0 in the reference config, seg_solver.py:118)')
        nf = len(cfg['in_channels'])
        c = L.DecCfg()
        c.num_levels = nf
        for i, v in enumerate(cfg['in_channels']):
            c.in_channels[i] = v
        for i, v in enumerate(cfg['features']):
            c.features[i] = v
        c.use_bn = int(bool(cfg['use_bn']))
        c.base_y, c.base_x = base_hw
        self.num_levels, self.num_classes = nf, cfg['features'][-1]
        self.out_hw = (base_hw[0] << (nf - 1), base_hw[1] << (nf - 1))
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.gsx_train_create(C.byref(c), self.n, int(bool(cfg.get('use_dropout', False))), C.byref(h)),
                    'gsx_train_create', dtype)
        self._h = h
        nl, nt = C.c_size_t(), C.c_size_t()
        count = self.lib.gsx_train_param_count(h, C.byref(nl), C.byref(nt))
        self.n_learn, self.n_total = nl.value, nt.value
        self.layout = {}
        for i in range(count):
            name, off, cnt = C.c_char_p(), C.c_size_t(), C.c_size_t()
            L.check(self.lib.gsx_train_param_info(h, i, C.byref(name), C.byref(off), C.byref(cnt)), 'gsx_train_param_info', dtype)
            self.layout[name.value.decode()] = (off.value, cnt.value)
        self.p = torch.zeros(self.n_total, dtype=torch.float32, device=self.device)
        self.g = torch.zeros(self.n_learn, dtype=torch.float32, device=self.device)
        self.m = torch.zeros_like(self.g)
        self.v = torch.zeros_like(self.g)
        self.shapes = {k: tuple(np.asarray(v).shape) for k, v in params.items()}
        self.load(params)
        sz = C.c_size_t()
        L.check(self.lib.gsx_train_workspace_bytes(h, C.byref(sz)), 'gsx_train_workspace_bytes', dtype)
        self.ws = torch.empty(sz.value, dtype=torch.uint8, device=self.device)
        self.loss = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.pred = torch.zeros((self.n,) + self.out_hw, dtype=torch.uint8, device=self.device)
        self.grad_scale = 1.0
        # The ~380 launches of a step are launch-bound at batch 1: after two eager steps the call is captured once in a
        # CUDA graph and replayed (static input buffers; the dropout seed is read from device memory).
        self.use_graph = bool(use_graph)
        self._graph = None
        self._calls = 0
        self._static = None
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=self.device)

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                self.lib.gsx_train_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def view(self, buf, name):
        off, cnt = self.layout[name]
        return buf[off:off + cnt]

    def load(self, params):
        missing = [k for k in self.layout if k not in params]
        if missing:
            raise KeyError(f'missing decoder parameters: {missing[:4]}')
        host = np.zeros(self.n_total, np.float32)
        for k, (off, cnt) in self.layout.items():
            a = np.asarray(params[k], np.float32).reshape(-1)
            if a.size != cnt:
                raise ValueError(f'parameter {k} has {a.size} elements, expected {cnt}')
            host[off:off + cnt] = a
        self.p.copy_(torch.from_numpy(host))

    def state(self):
        """name -> numpy float32 in the reference's shapes (what checkpoint_last.params holds)."""
        host = self.p.detach().cpu().numpy()
        return {k: host[off:off + cnt].reshape(self.shapes.get(k, (cnt,))).copy() for k, (off, cnt) in self.layout.items()}

    def grads(self):
        """Learnable-parameter gradients of the last step (true scale), name -> numpy."""
        host = self.g.detach().cpu().numpy() / self.grad_scale
        return {k: host[off:off + cnt].reshape(self.shapes.get(k, (cnt,))).copy()
                for k, (off, cnt) in self.layout.items() if off < self.n_learn}

    def dropout_mask(self, level, seed):
        f = self.cfg['features'][level]
        h, w = self.out_hw[0] >> (self.num_levels - 1 - level), self.out_hw[1] >> (self.num_levels - 1 - level)
        out = torch.empty((self.n, f, h, w), dtype=torch.float32, device=self.device)
        self.L.check(self.lib.gsx_train_dropout_mask(self._h, level, int(seed), self.L.ptr(out), self._stream()), 'dropout_mask', self.dtype)
        return out

    def _stream(self):
        return self.C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _enqueue(self, feats, lab, dropout_seed):
        L, C = self.L, self.C
        ptrs = (C.c_void_p * self.num_levels)(*[t.data_ptr() for t in feats])
        gs = C.c_float(1.0)
        rc = self.lib.gsx_train_step(self._h, L.ptr(self.p), L.ptr(self.g), ptrs, None, None, L.ptr(lab), int(dropout_seed),
                                     L.ptr(self.loss), L.ptr(self.pred), C.byref(gs), L.ptr(self.ws), self.ws.numel(), self._stream())
        L.check(rc, 'gsx_train_step', self.dtype)
        self.grad_scale = float(gs.value)

    def forward_backward(self, feats, mask, dropout_seed=0):
        """Enqueues forward + backward; returns the per-sample loss (device tensor).  Gradients (scaled) land in ``self.g``."""
        L = self.L
        with torch.cuda.device(self.device):
            keep = [torch.as_tensor(f, dtype=torch.float32).to(self.device).contiguous() for f in feats]
            if len(keep) != self.num_levels or keep[0].shape[0] != self.n:
                raise ValueError('expected %d feature maps of batch %d' % (self.num_levels, self.n))
            lab = torch.as_tensor(np.asarray(mask) if not torch.is_tensor(mask) else mask).to(self.device)
            lab = lab.reshape((self.n,) + self.out_hw).to(torch.int32).contiguous()
            self._calls += 1
            if not self.use_graph or self._calls <= 2:
                L.check(self.lib.gsx_train_set_seed_buffer(self._h, None), 'set_seed_buffer', self.dtype)
                self._enqueue(keep, lab, dropout_seed)
                self._keep = (keep, lab)
                return self.loss
            if self._static is None:
                self._static = ([torch.empty_like(t) for t in keep], torch.empty_like(lab))
            for d, t in zip(self._static[0], keep):
                d.copy_(t)
            self._static[1].copy_(lab)
            self._seed_dev.fill_(int(dropout_seed) & 0x7FFFFFFFFFFFFFFF)
            if self._graph is None:
                L.check(self.lib.gsx_train_set_seed_buffer(self._h, L.ptr(self._seed_dev)), 'set_seed_buffer', self.dtype)
                torch.cuda.synchronize(self.device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._enqueue(self._static[0], self._static[1], 0)
                self._graph = graph
            self._graph.replay()
        return self.loss

    def step(self, feats, mask, dropout_seed=0, global_batch=None, group=None):
        loss = self.forward_backward(feats, mask, dropout_seed)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.g, op=dist.ReduceOp.SUM, group=group)          # the step's only collective
        self.t += 1
        L = self.L
        with torch.cuda.device(self.device):
            L.check(self.lib.gsx_adam_step(L.ptr(self.p), L.ptr(self.g), L.ptr(self.m), L.ptr(self.v), self.n_learn, self.t, self.lr,
                                           self.beta1, self.beta2, self.eps, self.wd,
                                           1.0 / (float(global_batch or self.n) * self.grad_scale), self._stream()),
                    'gsx_adam_step', self.dtype)
        return loss
