import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import torch, torch.nn.functional as F
from gan_segmentation_b200 import _lib as L
lib = L.lib()
for mode in (0, 1, 2):
    lib.gsx_set_option(b'wgrad_m64', mode)
    for (k, n, h, w, cin, cout) in [(3, 1, 64, 200, 64, 32), (3, 2, 37, 53, 16, 16), (1, 2, 32, 32, 64, 32)]:
        g = torch.Generator().manual_seed(1)
        x = torch.randn((n, cin, h, w), generator=g).half().float().cuda()
        dy = torch.randn((n, cout, h, w), generator=g).half().float().cuda()
        wt = torch.zeros((cout, cin, k, k), device='cuda', requires_grad=True)
        F.conv2d(x, wt, None, 1, k // 2).backward(dy)
        dw = torch.zeros((cout, cin, k, k), device='cuda')
        rc = lib.gsx_op_conv_wgrad_tc(k, n, h, w, cin, cout, L.ptr(x), L.ptr(dy), L.ptr(dw), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        err = float((dw - wt.grad).abs().max() / wt.grad.abs().max())
        # which input channels are right?
        per_ci = ((dw - wt.grad).abs().amax(dim=(0, 2, 3)) / wt.grad.abs().max()).cpu().numpy()
        good = [i for i, e in enumerate(per_ci) if e < 2e-3]
        print('mode', mode, (k, n, h, w, cin, cout), 'rc', rc, 'err %.4g' % err, 'good ci:', good[:20], len(good))
