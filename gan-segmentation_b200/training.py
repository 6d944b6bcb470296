"""Decoder-training building blocks over the C ABI (reference seg_solver.py:351-465): the loss, the conv weight
gradient, and the flat-bucket gradient all-reduce + Adam step that replaces the reference's per-parameter
KVStore('nccl') push/pull (seg_solver.py:55-56,421).  ``decoder_training.DecoderTrainer`` composes them into the step.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def softmax_ce(logits, labels, want_grad=True, dtype=None, grad_scale=1.0):
    """logits [N,K,H,W] fp32 cuda, labels [N,1,H,W] or [N,H,W] int32 (-1 = ignore).
    Returns (loss [N], grad_scale * dlogits or None) with the reference's semantics (mean over all pixels, ignored
    ones weigh 0).  Training passes grad_scale = H*W (see gsx.h) and divides it out again in the optimizer step."""
    lib = L.lib(dtype)
    n, k, h, w = logits.shape
    logits = logits.contiguous()
    labels = labels.reshape(n, h, w).to(torch.int32).contiguous()
    loss = torch.empty(n, dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits) if want_grad else None
    scratch = torch.empty(n * 256, dtype=torch.float32, device=logits.device)
    L.check(lib.gsx_softmax_ce(L.ptr(logits), L.ptr(labels), n, k, h, w, L.ptr(loss), L.ptr(dl), float(grad_scale), L.ptr(scratch),
                               scratch.numel(), _stream()), 'gsx_softmax_ce', dtype)
    return loss, dl


class FlatAdam:
    """All parameters of the decoder in ONE fp32 bucket (942 562 floats = 3.6 MiB at FFHQ): one all-reduce per
    step (sum over ranks), then one fused Adam kernel.  MXNet semantics: Adam(lr, beta1=.9, beta2=.999, eps=1e-8),
    rescale_grad = 1 / global batch."""

    def __init__(self, shapes, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0, device='cuda', dtype=None):
        self.names = list(shapes)
        self.shapes = dict(shapes)
        self.offsets = {}
        off = 0
        for k, s in shapes.items():
            self.offsets[k] = off
            off += int(np.prod(s))
        self.count = off
        self.device = torch.device(device)
        self.w = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.g = torch.zeros_like(self.w)
        self.m = torch.zeros_like(self.w)
        self.v = torch.zeros_like(self.w)
        self.lr, self.beta1, self.beta2, self.eps, self.wd = lr, beta1, beta2, eps, wd
        self.t = 0
        self.dtype = dtype

    def view(self, buf, name):
        o = self.offsets[name]
        return buf[o:o + int(np.prod(self.shapes[name]))].view(self.shapes[name])

    def load(self, params):
        for k in self.names:
            self.view(self.w, k).copy_(torch.as_tensor(np.asarray(params[k], np.float32)))

    def state(self):
        return {k: self.view(self.w, k).detach().cpu().numpy() for k in self.names}

    def allreduce_grads(self, group=None):
        """The step's only collective: sum of the flat gradient bucket over the ranks."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.g, op=dist.ReduceOp.SUM, group=group)

    def step(self, global_batch, group=None, grad_scale=1.0):
        """``grad_scale``: the factor the loss gradient was multiplied by for the 16-bit backward pass (divided out
        here, through Adam's rescale_grad)."""
        self.allreduce_grads(group)
        self.t += 1
        if self.device.type != 'cuda':
            raise RuntimeError('FlatAdam.step needs a CUDA device: there is no CPU fallback')
        lib = L.lib(self.dtype)
        L.check(lib.gsx_adam_step(L.ptr(self.w), L.ptr(self.g), L.ptr(self.m), L.ptr(self.v), self.count, self.t, self.lr,
                                  self.beta1, self.beta2, self.eps, self.wd, 1.0 / (float(global_batch) * float(grad_scale)), _stream()),
                'gsx_adam_step', self.dtype)


def dgrad_weights(weight):
    """Weights that turn the forward 3x3 / 1x1 conv kernel into its own data gradient:
    ``dX = conv(dY, dgrad_weights(W))`` for ``Y = conv(X, W)`` (stride 1, 'same' padding) -- channels swapped,
    taps flipped.  The decoder's dgrad therefore runs on the existing tcgen05 shift-GEMM kernel; the weight gradient and
    the BatchNorm backward have their own kernels (csrc/train.cu).  weight: [Cout,Cin,k,k] numpy."""
    w = np.asarray(weight, np.float32)
    return np.ascontiguousarray(np.transpose(w, (1, 0, 2, 3))[:, :, ::-1, ::-1])


def conv_wgrad(x, dy, k, dtype=None):
    """Weight and bias gradient of a stride-1 'same' k x k conv (k = 1 or 3): x [N,Cin,H,W], dy [N,Cout,H,W] fp32 cuda
    -> (dW [Cout,Cin,k,k], db [Cout]).  First (CUDA-core, deterministic) version of the decoder's wgrad."""
    lib = L.lib(dtype)
    n, cin, h, w = x.shape
    cout = dy.shape[1]
    x = x.contiguous().float()
    dy = dy.contiguous().float()
    dw = torch.empty((cout, cin, k, k), dtype=torch.float32, device=x.device)
    db = torch.empty((cout,), dtype=torch.float32, device=x.device)
    L.check(lib.gsx_op_conv_wgrad(k, n, h, w, cin, cout, L.ptr(x), L.ptr(dy), L.ptr(dw), L.ptr(db), _stream()),
            'gsx_op_conv_wgrad', dtype)
    return dw, db
