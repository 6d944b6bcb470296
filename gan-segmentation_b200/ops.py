"""Single-operator wrappers over the C ABI's gsx_op_* hooks (tests and tuning sweeps).
torch is used only to own device memory."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L

PLAN_FIELDS = ('TH', 'TW', 'NB', 'CBK', 'N_tile', 'stages', 'phase_grid', 'n_mtiles', 'n_k', 'tmem_cols',
               'smem_bytes', 'tiles', 'n_ntiles', 'groups_bufs_epi', 'n_slots', 'BW')


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def conv(mode, x0, weight, x1=None, bias=None, nscale=None, noise=None, flags=0, addsrc=None,
         num_classes=0, override=None, repeat=0, dtype=None):
    """Run one shift-GEMM convolution.  x0/x1: [N,C,H,W] fp32 cuda; weight: fp32 (numpy or cpu tensor) in
    the reference layout of the mode.  Returns dict(out, stats, mask, logits, plan, ms)."""
    lib = L.lib(dtype)
    n, cin0, h, w = x0.shape
    cin1 = 0 if x1 is None else x1.shape[1]
    wnp = np.ascontiguousarray(np.asarray(weight, np.float32))
    if mode in (L.DECONV4, L.DECONV4B):
        cout = wnp.shape[1]
    else:
        cout = wnp.shape[0]
    up = mode in (L.UPCONV3, L.DECONV4, L.DECONV4B)
    ho, wo = (2 * h, 2 * w) if up else (h, w)
    dev = x0.device
    argmax = bool(flags & L.EPI_ARGMAX)
    out = None if argmax else torch.empty((n, cout, ho, wo), dtype=torch.float32, device=dev)
    stats = torch.zeros((n, cout, 2), dtype=torch.float32, device=dev) if flags & L.EPI_STATS else None
    mask = torch.empty((n, ho, wo), dtype=torch.uint8, device=dev) if argmax else None
    logits = torch.empty((n, num_classes, ho, wo), dtype=torch.float32, device=dev) if argmax else None
    ov = None
    if override:
        ov = L.PlanOverride(TH=0, TW=0, NB=0, CBK=0, N_tile=0, stages=0, phase_grid=-1, epi_groups=0, acc_bufs=0, max_mtiles=0, hstack=-1, s2d=-1)
        for k, v in override.items():
            setattr(ov, k, v)
    plan = (C.c_int * 16)()
    ms = C.c_float(0)
    x0 = x0.contiguous()
    x1 = None if x1 is None else x1.contiguous()
    rc = lib.gsx_op_conv(mode, n, h, w, cin0, cin1, cout, L.ptr(x0), L.ptr(x1), L.np_ptr(wnp), L.ptr(bias),
                         L.ptr(nscale), L.ptr(noise), flags, L.ptr(addsrc), L.ptr(out), L.ptr(stats), L.ptr(mask),
                         L.ptr(logits), num_classes, C.byref(ov) if ov else None, plan, repeat, C.byref(ms), _stream())
    L.check(rc, 'gsx_op_conv', dtype)
    return dict(out=out, stats=stats, mask=mask, logits=logits, plan=dict(zip(PLAN_FIELDS, list(plan))), ms=ms.value)


def pass1(x, n, blur, nscale, bias, noise, in_broadcast=False, want_stats=True, dtype=None):
    lib = L.lib(dtype)
    _, c, h, w = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    stats = torch.zeros((n, c, 2), dtype=torch.float32, device=x.device) if want_stats else None
    rc = lib.gsx_op_pass1(n, c, h, w, L.ptr(x.contiguous()), int(blur), int(in_broadcast), L.ptr(nscale), L.ptr(bias),
                          L.ptr(noise), L.ptr(out), L.ptr(stats), _stream())
    L.check(rc, 'gsx_op_pass1', dtype)
    return out, stats


def apply(x, stats, styles, wrgb=None, brgb=None, dtype=None):
    lib = L.lib(dtype)
    n, c, h, w = x.shape
    out = torch.empty_like(x)
    nc = 0 if wrgb is None else wrgb.shape[0]
    img = torch.empty((n, nc, h, w), dtype=torch.float32, device=x.device) if nc else None
    u8 = torch.empty((n, h, w, nc), dtype=torch.uint8, device=x.device) if nc else None
    rc = lib.gsx_op_apply(n, c, h, w, L.ptr(x.contiguous()), L.ptr(stats), L.ptr(styles), L.ptr(wrgb), L.ptr(brgb), nc,
                          L.ptr(out), L.ptr(img), L.ptr(u8), _stream())
    L.check(rc, 'gsx_op_apply', dtype)
    return out, img, u8


def fill_normal(per_sample, n, seed, first_sample, stream_id, device='cuda'):
    lib = L.lib()
    out = torch.empty((n, per_sample), dtype=torch.float32, device=device)
    L.check(lib.gsx_op_fill_normal(L.ptr(out), per_sample, n, seed, first_sample, stream_id, _stream()), 'fill_normal')
    return out
