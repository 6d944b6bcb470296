"""GPU parity of the single CUDA operators (through the C ABI's gsx_op_* hooks) against plain PyTorch
fp32 on the same bf16-rounded operands: what differs is accumulation order and the bf16 output rounding."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


_DT = 'fp16'     # set per test from the ``dtype`` fixture


def use(dtype):
    global _DT
    _DT = dtype


def bf(x):
    """Round to the 16-bit storage type of the library under test."""
    return x.to(torch.bfloat16 if _DT == 'bf16' else torch.float16).to(torch.float32)


def close(out, ref, what, rel=None, abs_frac=None):
    if rel is None:
        rel = 2.0 ** -7 if _DT == 'bf16' else 2.0 ** -10
    if abs_frac is None:
        abs_frac = 2e-3 if _DT == 'bf16' else 3e-4
    scale = ref.abs().max().item() + 1e-6
    err = (out - ref).abs()
    tol = rel * ref.abs() + abs_frac * scale
    bad = (err > tol).sum().item()
    assert bad == 0, f'{what}: {bad} elements out of tolerance, max err {err.max().item():.4g} (scale {scale:.4g})'


def ref_conv(mode, x, w):
    from gan_segmentation_b200 import _lib as L
    if mode == L.CONV3:
        return F.conv2d(x, w, None, 1, 1)
    if mode == L.CONV1:
        return F.conv2d(x, w)
    if mode == L.UPCONV3:
        return F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), w, None, 1, 1)
    if mode == L.DECONV4:
        return F.conv_transpose2d(x, w, None, stride=2, padding=1)
    raise ValueError(mode)


def make_w(mode, cin, cout, g):
    from gan_segmentation_b200 import _lib as L
    k = {L.CONV3: 3, L.UPCONV3: 3, L.DECONV4: 4, L.CONV1: 1}[mode]
    shape = (cin, cout, k, k) if mode == L.DECONV4 else (cout, cin, k, k)
    return bf(torch.randn(shape, generator=g) / np.sqrt(cin * k * k))


CASES = [
    # mode, N, cin0, cin1, cout, H, W
    ('CONV3', 2, 16, 0, 16, 64, 64),
    ('CONV3', 3, 32, 0, 32, 20, 36),
    ('CONV3', 2, 512, 0, 512, 4, 4),
    ('CONV3', 5, 512, 0, 512, 8, 8),
    ('CONV3', 1, 256, 0, 256, 16, 16),
    ('CONV3', 2, 64, 0, 32, 128, 128),
    ('CONV3', 1, 16, 0, 16, 192, 256),
    ('CONV3', 2, 32, 32, 32, 32, 32),
    ('CONV3', 3, 512, 0, 32, 6, 8),
    ('UPCONV3', 2, 512, 0, 512, 4, 4),
    ('UPCONV3', 2, 64, 0, 32, 16, 16),
    ('UPCONV3', 2, 32, 32, 32, 24, 32),
    ('UPCONV3', 1, 32, 32, 16, 64, 64),
    ('DECONV4', 2, 256, 0, 128, 16, 16),
    ('DECONV4', 2, 32, 0, 16, 64, 64),
    ('DECONV4', 1, 64, 0, 32, 48, 64),
    ('CONV1', 2, 32, 32, 32, 16, 16),
    ('CONV1', 1, 32, 32, 16, 64, 64),
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: '-'.join(map(str, c)))
def test_conv_raw(gsx_lib, dtype, case):
    use(dtype)
    from gan_segmentation_b200 import _lib as L, ops
    mname, n, cin0, cin1, cout, h, w = case
    mode = getattr(L, mname)
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    cin = cin0 + cin1
    x = bf(torch.randn((n, cin, h, w), generator=g))
    wt = make_w(mode, cin, cout, g)
    xd = x.cuda()
    r = ops.conv(mode, xd[:, :cin0], wt.numpy(), x1=xd[:, cin0:] if cin1 else None, dtype=dtype)
    ref = ref_conv(mode, xd, wt.cuda())
    close(r['out'], ref, f'{case} plan={r["plan"]}')


def test_conv_generator_epilogue(gsx_lib, dtype):
    """conv_2 epilogue of a synthesis block: + scale*noise + bias, LeakyReLU(0.2), InstanceNorm sums."""
    use(dtype)
    from gan_segmentation_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(5)
    for (n, c, h, w) in [(2, 32, 64, 64), (2, 512, 8, 8), (1, 128, 32, 32), (4, 512, 4, 4)]:
        x = bf(torch.randn((n, c, h, w), generator=g)).cuda()
        wt = make_w(L.CONV3, c, c, g)
        ns = torch.randn(c, generator=g).cuda() * 0.3
        b = torch.randn(c, generator=g).cuda() * 0.2
        nz = torch.randn((n, 1, h, w), generator=g).cuda()
        r = ops.conv(L.CONV3, x, wt.numpy(), bias=b, nscale=ns, noise=nz, flags=L.EPI_LRELU | L.EPI_STATS, dtype=dtype)
        v = F.conv2d(x, wt.cuda(), None, 1, 1) + ns.view(1, -1, 1, 1) * nz + b.view(1, -1, 1, 1)
        v = F.leaky_relu(v, 0.2)
        close(r['out'], v, f'epilogue {n, c, h, w}')
        s1 = v.sum(dim=(2, 3))
        s2 = (v * v).sum(dim=(2, 3))
        # multi-sample tiles (tiny images) take the statistics from the bf16-rounded output instead of the
        # fp32 accumulators, hence the looser bound there
        rt = 2e-3 if r['plan']['NB'] == 1 else 8e-3
        assert torch.allclose(r['stats'][:, :, 0], s1, rtol=rt, atol=rt * s2.sqrt().max().item()), r['plan']
        assert torch.allclose(r['stats'][:, :, 1], s2, rtol=rt), r['plan']


def test_deconv_blur_folded(gsx_lib, dtype):
    """First half of a high-resolution synthesis block in ONE kernel (networks_stylegan.py:16-20): 4x4 stride-2
    transposed conv -> 3x3 blur (zero padding on the CROPPED deconv output) -> + scale*noise + bias -> LeakyReLU,
    with the InstanceNorm sums.  The blur is folded into the conv weights; the 1-pixel border gets a correction."""
    use(dtype)
    from gan_segmentation_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(11)
    blur = torch.tensor([1., 2., 1.])
    blur = (blur[:, None] * blur[None, :] / 16.0).cuda()
    for (n, cin, cout, h, w) in [(2, 32, 16, 64, 64), (1, 64, 32, 48, 40), (3, 32, 16, 6, 8), (1, 32, 16, 130, 72)]:
        x = bf(torch.randn((n, cin, h, w), generator=g)).cuda()
        wt = bf(torch.randn((cin, cout, 4, 4), generator=g) / np.sqrt(cin * 4.0))
        ns = torch.randn(cout, generator=g).cuda() * 0.3
        b = torch.randn(cout, generator=g).cuda() * 0.2
        nz = torch.randn((n, 1, 2 * h, 2 * w), generator=g).cuda()
        r = ops.conv(L.DECONV4B, x, wt.numpy(), bias=b, nscale=ns, noise=nz, flags=L.EPI_LRELU | L.EPI_STATS, dtype=dtype)
        d = F.conv_transpose2d(x, wt.cuda(), None, stride=2, padding=1)
        d = F.conv2d(d, blur.expand(cout, 1, 3, 3).contiguous(), None, 1, 1, groups=cout)
        v = F.leaky_relu(d + ns.view(1, -1, 1, 1) * nz + b.view(1, -1, 1, 1), 0.2)
        rel = 2.0 ** -6 if dtype == 'bf16' else 2.0 ** -9         # the composite weights are rounded once more
        close(r['out'], v, f'deconv+blur {n, cin, cout, h, w} plan={r["plan"]}', rel=rel)
        # the border on its own (the correction path), and without noise / bias / activation
        r0 = ops.conv(L.DECONV4B, x, wt.numpy(), dtype=dtype)
        for name, sl in (('top', (slice(None), slice(None), 0)), ('bottom', (slice(None), slice(None), -1)),
                         ('left', (slice(None), slice(None), slice(None), 0)), ('right', (slice(None), slice(None), slice(None), -1))):
            close(r0['out'][sl], d[sl], f'deconv+blur border {name} {n, cin, cout, h, w}', rel=rel)
        s1 = v.sum(dim=(2, 3))
        s2 = (v * v).sum(dim=(2, 3))
        rt = (2e-3 if r['plan']['NB'] == 1 else 8e-3) * (4 if dtype == 'bf16' else 1)   # bf16: composite-weight rounding
        assert torch.allclose(r['stats'][:, :, 0], s1, rtol=rt, atol=rt * s2.sqrt().max().item()), r['plan']
        assert torch.allclose(r['stats'][:, :, 1], s2, rtol=rt), r['plan']


def test_conv_decoder_residual(gsx_lib, dtype):
    """conv_b of a DecoderResBlock: bias, LeakyReLU, + nearest-upsampled shortcut (networks_seg.py:44-46)."""
    use(dtype)
    from gan_segmentation_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(6)
    n, c, h, w = 2, 32, 32, 48
    x = bf(torch.randn((n, c, h, w), generator=g)).cuda()
    sc = bf(torch.randn((n, c, h // 2, w // 2), generator=g)).cuda()
    wt = make_w(L.CONV3, c, c, g)
    b = torch.randn(c, generator=g).cuda() * 0.2
    r = ops.conv(L.CONV3, x, wt.numpy(), bias=b, flags=L.EPI_LRELU, addsrc=sc, dtype=dtype)
    ref = F.leaky_relu(F.conv2d(x, wt.cuda(), b, 1, 1), 0.2) + F.interpolate(sc, scale_factor=2, mode='nearest')
    close(r['out'], ref, 'residual')


def test_conv_space_to_depth_plan(gsx_lib, dtype):
    """The space-to-depth plan of thin 3x3 layers (GEMM rows = 2x2 pixel blocks, 4 input phase planes, 4 output
    phases as column blocks; default only for the final conv) forced on for the raw, generator and residual
    epilogues, against the same references as the dense plan -- and against the dense plan itself."""
    use(dtype)
    from gan_segmentation_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(12)
    s2d = dict(s2d=1)
    for (n, cin, cout, h, w) in [(2, 16, 16, 64, 64), (1, 32, 32, 20, 36), (3, 16, 16, 18, 130), (1, 32, 16, 66, 64)]:
        x = bf(torch.randn((n, cin, h, w), generator=g)).cuda()
        wt = make_w(L.CONV3, cin, cout, g)
        r = ops.conv(L.CONV3, x, wt.numpy(), override=s2d, dtype=dtype)
        assert r['plan']['n_slots'] == 16, r['plan']
        close(r['out'], F.conv2d(x, wt.cuda(), None, 1, 1), f's2d raw {n, cin, cout, h, w}')
        r0 = ops.conv(L.CONV3, x, wt.numpy(), override=dict(s2d=0), dtype=dtype)
        assert r0['plan']['n_slots'] == 9
        close(r['out'], r0['out'], 's2d vs dense plan', rel=2.0 ** -7 if dtype == 'bf16' else 2.0 ** -10)
    # generator conv_2 epilogue
    n, c, h, w = 2, 16, 64, 96
    x = bf(torch.randn((n, c, h, w), generator=g)).cuda()
    wt = make_w(L.CONV3, c, c, g)
    ns = torch.randn(c, generator=g).cuda() * 0.3
    b = torch.randn(c, generator=g).cuda() * 0.2
    nz = torch.randn((n, 1, h, w), generator=g).cuda()
    r = ops.conv(L.CONV3, x, wt.numpy(), bias=b, nscale=ns, noise=nz, flags=L.EPI_LRELU | L.EPI_STATS, override=s2d, dtype=dtype)
    v = F.leaky_relu(F.conv2d(x, wt.cuda(), None, 1, 1) + ns.view(1, -1, 1, 1) * nz + b.view(1, -1, 1, 1), 0.2)
    close(r['out'], v, 's2d generator epilogue')
    s1, s2 = v.sum(dim=(2, 3)), (v * v).sum(dim=(2, 3))
    rt = 2e-3 if r['plan']['NB'] == 1 else 8e-3
    assert torch.allclose(r['stats'][:, :, 0], s1, rtol=rt, atol=rt * s2.sqrt().max().item()), r['plan']
    assert torch.allclose(r['stats'][:, :, 1], s2, rtol=rt), r['plan']
    # decoder conv_b: residual at block resolution
    sc = bf(torch.randn((n, c, h // 2, w // 2), generator=g)).cuda()
    r = ops.conv(L.CONV3, x, wt.numpy(), bias=b, flags=L.EPI_LRELU, addsrc=sc, override=s2d, dtype=dtype)
    ref = F.leaky_relu(F.conv2d(x, wt.cuda(), b, 1, 1), 0.2) + F.interpolate(sc, scale_factor=2, mode='nearest')
    close(r['out'], ref, 's2d residual')


def test_conv_argmax(gsx_lib, dtype):
    """Final decoder conv + argmax: the mask must equal the first-max argmax of the logits the kernel
    itself produced (bit-exact), and the logits must match the fp32 reference."""
    use(dtype)
    from gan_segmentation_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(7)
    for nc in (2, 5):
        n, c0, c1, h, w = 2, 16, 16, 64, 96
        x = bf(torch.randn((n, c0 + c1, h, w), generator=g)).cuda()
        wt = make_w(L.CONV3, c0 + c1, nc, g)
        b = torch.randn(nc, generator=g) * 0.1
        b16 = torch.zeros(16)
        b16[:nc] = b
        r = ops.conv(L.CONV3, x[:, :c0], wt.numpy(), x1=x[:, c0:], bias=b16.cuda(), flags=L.EPI_ARGMAX, num_classes=nc, dtype=dtype)
        ref = F.conv2d(x, wt.cuda(), b.cuda(), 1, 1)
        close(r['logits'], ref, 'logits', rel=1e-4, abs_frac=1e-4)
        lg = r['logits'].cpu().numpy()
        best = lg[:, 0].copy()
        idx = np.zeros(best.shape, np.uint8)
        for c in range(1, nc):
            m = lg[:, c] > best
            idx[m] = c
            best = np.where(m, lg[:, c], best)
        assert np.array_equal(r['mask'].cpu().numpy(), idx)


def test_argmax_ties_first_max(gsx_lib, dtype):
    """All-zero weights + equal biases: every logit ties, class 0 must win (seg_solver.py:326)."""
    use(dtype)
    from gan_segmentation_b200 import _lib as L, ops
    x = torch.randn((1, 32, 16, 16)).cuda()
    wt = np.zeros((3, 32, 3, 3), np.float32)
    b16 = torch.zeros(16).cuda()
    r = ops.conv(L.CONV3, x[:, :16], wt, x1=x[:, 16:], bias=b16, flags=L.EPI_ARGMAX, num_classes=3, dtype=dtype)
    assert int(r['mask'].max()) == 0


def test_pass1_and_apply(gsx_lib, dtype):
    use(dtype)
    from gan_segmentation_b200 import ops
    g = torch.Generator().manual_seed(8)
    for (n, c, h, w, blur) in [(2, 16, 64, 64, True), (3, 64, 12, 16, True), (2, 512, 4, 4, False), (1, 32, 96, 128, True)]:
        x = bf(torch.randn((n, c, h, w), generator=g)).cuda()
        ns = torch.randn(c, generator=g).cuda() * 0.3
        b = torch.randn(c, generator=g).cuda() * 0.2
        nz = torch.randn((n, 1, h, w), generator=g).cuda()
        out, stats = ops.pass1(x, n, blur, ns, b, nz, dtype=dtype)
        v = x
        if blur:
            k = torch.tensor([1., 2., 1.])
            k = (torch.outer(k, k) / 16).view(1, 1, 3, 3).repeat(c, 1, 1, 1).cuda()
            v = F.conv2d(v, k, None, 1, 1, 1, c)
        v = F.leaky_relu(v + ns.view(1, -1, 1, 1) * nz + b.view(1, -1, 1, 1), 0.2)
        close(out, v, f'pass1 {n, c, h, w}')
        assert torch.allclose(stats[:, :, 0], v.sum(dim=(2, 3)), rtol=1e-3, atol=1e-2)
        assert torch.allclose(stats[:, :, 1], (v * v).sum(dim=(2, 3)), rtol=1e-3)
        # apply: InstanceNorm (eps 1e-5, biased var) + (scale+1), shift
        t = bf(v)
        st = torch.stack([t.sum(dim=(2, 3)), (t * t).sum(dim=(2, 3))], dim=2).contiguous()
        sty = torch.randn((n, 2 * c), generator=g).cuda()
        o2, _, _ = ops.apply(t, st, sty, dtype=dtype)
        ref = F.instance_norm(t, eps=1e-5) * (sty[:, :c].view(n, c, 1, 1) + 1) + sty[:, c:].view(n, c, 1, 1)
        close(o2, ref, f'apply {n, c, h, w}', rel=2.0 ** -7, abs_frac=4e-3)


def test_pass1_broadcast_const(gsx_lib, dtype):
    use(dtype)
    from gan_segmentation_b200 import ops
    g = torch.Generator().manual_seed(9)
    n, c, h, w = 3, 512, 4, 4
    x = bf(torch.randn((1, c, h, w), generator=g)).cuda()
    ns = torch.randn(c, generator=g).cuda() * 0.3
    b = torch.randn(c, generator=g).cuda() * 0.2
    nz = torch.randn((n, 1, h, w), generator=g).cuda()
    out, _ = ops.pass1(x, n, False, ns, b, nz, in_broadcast=True, dtype=dtype)
    v = F.leaky_relu(x + ns.view(1, -1, 1, 1) * nz + b.view(1, -1, 1, 1), 0.2)
    close(out, v, 'pass1 broadcast')


def test_apply_rgb_u8(gsx_lib, dtype):
    """ToRGB + uint8 transform fused in the last apply pass; uint8 must equal the reference transform
    (image_generator.py:76-84) of the fp32 image the same kernel wrote (bit-exact)."""
    use(dtype)
    from gan_segmentation_b200 import ops
    g = torch.Generator().manual_seed(10)
    n, c, h, w = 2, 16, 64, 64
    t = bf(torch.randn((n, c, h, w), generator=g)).cuda()
    st = torch.stack([t.sum(dim=(2, 3)), (t * t).sum(dim=(2, 3))], dim=2).contiguous()
    sty = torch.randn((n, 2 * c), generator=g).cuda() * 0.5
    wr = (torch.randn((3, c), generator=g) / 4).cuda()
    br = (torch.randn(3, generator=g) * 0.1).cuda()
    o, img, u8 = ops.apply(t, st, sty, wr, br, dtype=dtype)
    ref = F.instance_norm(t, eps=1e-5) * (sty[:, :c].view(n, c, 1, 1) + 1) + sty[:, c:].view(n, c, 1, 1)
    close(o, ref, 'apply rgb feature', abs_frac=4e-3)
    ref_img = F.conv2d(ref, wr.view(3, c, 1, 1), br)
    close(img, ref_img, 'rgb', rel=1e-3, abs_frac=2e-3)
    a = img.cpu().numpy().transpose(0, 2, 3, 1)
    a = (a - np.float32(-1)) / np.float32(2)
    a = (np.float32(255.) * np.clip(a, 0.0, 1.0)).astype(np.uint8)
    assert np.array_equal(u8.cpu().numpy(), a)


def test_philox_normal(gsx_lib, dtype):
    use(dtype)
    from gan_segmentation_b200 import ops
    a = ops.fill_normal(1 << 16, 4, 123, 10, 3)
    b = ops.fill_normal(1 << 16, 2, 123, 12, 3)       # samples 12,13 == rows 2,3 of a: split-invariant
    assert torch.equal(a[2:], b)
    assert abs(a.mean().item()) < 0.01 and abs(a.std().item() - 1) < 0.01
    c = ops.fill_normal(1 << 16, 4, 123, 10, 4)
    assert not torch.equal(a, c)
