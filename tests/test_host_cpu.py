"""CPU-side tests: the C-ABI library loads and exports every symbol include/gsx.h declares, the planner's
invariants hold for every layer of the three GAN configs, parameter naming / .params codec / error paths."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gan_segmentation_b200 import _lib as L
from gan_segmentation_b200.config import generator_config, decoder_config, num_features, MAX_RES_LOG2
from gan_segmentation_b200.naming import canonical, legacy_name, generator_param_shapes, decoder_param_shapes
from gan_segmentation_b200.params_io import save_params, load_params
from gan_segmentation_b200.random_init import init_generator_params, init_decoder_params
from gan_segmentation_b200.ops import PLAN_FIELDS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'gsx.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gsx_[a-z0-9_]+)\s*\(', src)))


@pytest.mark.parametrize('dtype', ['fp16', 'bf16'])
def test_library_exports_every_declared_symbol(gsx_lib, dtype):
    lib = L.lib(dtype)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/gsx.h but not exported'
    assert set(syms) == set(L.EXPORTS)
    assert lib.gsx_abi_version() == 2


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setitem(L.LIB_PATHS, 'fp16', '/nonexistent/libgsx.so')
    monkeypatch.setattr(L, '_libs', {})
    with pytest.raises(L.GsxError):
        L.lib('fp16')


def test_no_cpu_fallback_in_drivers():
    from gan_segmentation_b200.image_generator import ImageGenerator
    from gan_segmentation_b200.seg_solver import SegSolver
    with pytest.raises(RuntimeError):
        ImageGenerator([], 'stylegan-models', 'bedrooms', params={})
    with pytest.raises(RuntimeError):
        SegSolver(8, 'data', 'ckpt', [])


def plan(mode, h, w, c0, c1, co, nc=0):
    out = (C.c_int * 16)()
    rc = L.lib().gsx_plan_query(mode, h, w, c0, c1, co, nc, None, out)
    assert rc == 0, L.lib().gsx_last_error()
    return dict(zip(PLAN_FIELDS, list(out)))


def all_layers(gan, base=(4, 4)):
    Lr = MAX_RES_LOG2[gan]
    gc, dc = generator_config(Lr, *base), decoder_config(Lr)
    for r in range(2, Lr + 1):
        c = num_features(gc, r)
        h, w = base[0] << (r - 2), base[1] << (r - 2)
        if r > 2:
            yield (L.DECONV4 if r >= 7 else L.UPCONV3, h // 2, w // 2, num_features(gc, r - 1), 0, c, 0)
        yield (L.CONV3, h, w, c, 0, c, 0)
    f, inc = dc['features'], dc['in_channels']
    for i in range(len(inc)):
        h, w = base[0] << i, base[1] << i
        yield (L.CONV3, h, w, inc[i], 0, f[i], 0)
        c0, c1 = f[i], (f[i] if i > 0 else 0)
        if i < len(inc) - 1:
            yield (L.UPCONV3, h, w, c0, c1, f[i + 1], 0)
            yield (L.CONV1, h, w, c0, c1, f[i + 1], 0)
            yield (L.CONV3, 2 * h, 2 * w, f[i + 1], 0, f[i + 1], 0)
        else:
            yield (L.CONV3, h, w, c0, c1, f[i + 1], f[i + 1])


@pytest.mark.parametrize('gan,base', [('ffhq', (4, 4)), ('cars', (3, 4)), ('cars', (4, 4)), ('bedrooms', (4, 4))])
def test_planner_invariants(gsx_lib, gan, base):
    for (mode, h, w, c0, c1, co, nc) in all_layers(gan, base):
        p = plan(mode, h, w, c0, c1, co, nc)
        tag = f'{gan}{base} mode={mode} {h}x{w} {c0}+{c1}->{co}: {p}'
        assert p['smem_bytes'] <= 227 * 1024, tag
        assert p['tmem_cols'] <= 512 and p['tmem_cols'] >= 32 and p['tmem_cols'] & (p['tmem_cols'] - 1) == 0, tag
        groups, bufs, epi = p['groups_bufs_epi'] // 100, p['groups_bufs_epi'] // 10 % 10, p['groups_bufs_epi'] % 10
        assert bufs * groups * p['n_mtiles'] * p['N_tile'] <= p['tmem_cols'], tag
        assert bufs in (1, 2) and epi in (1, 2, 4), tag
        assert p['N_tile'] % 16 == 0 and p['N_tile'] <= 256, tag
        assert p['CBK'] % 2 == 0 and (c0 // 8) % p['CBK'] == 0 and (c1 // 8) % p['CBK'] == 0, tag
        assert p['BW'] == p['TW'] + 2 and p['BW'] <= 128, tag            # TMA box: 256 8-byte units
        assert p['TH'] + 2 <= 256 and p['NB'] <= 256, tag
        assert 1 <= p['stages'] <= 8, tag
        # every output pixel of a tile has an MMA row
        rows = ((p['NB'] - 1) * (p['TH'] + 2) + p['TH'] - 1) * p['BW'] + p['TW']
        assert rows <= p['n_mtiles'] * 128, tag
        # tiles cover the image
        s2d = p['n_slots'] == 16                 # space-to-depth plans tile the grid of 2x2 pixel blocks
        tiles_x = -(-(w // 2 if s2d else w) // p['TW'])
        tiles_y = -(-(h // 2 if s2d else h) // p['TH'])
        assert p['tiles'] == tiles_x * tiles_y, tag


def test_parameter_names_and_counts():
    gc = generator_config(10)
    gs = generator_param_shapes(gc)
    n_gen = sum(int(np.prod(s)) for k, s in gs.items()
                if not (k.endswith('.std') or k.endswith('w_kernel') or k.endswith('gamma') or k.endswith('beta')))
    assert abs(n_gen - 26.5e6) < 0.3e6                      # SURVEY: generator 26.5 M parameters
    ds = decoder_param_shapes(decoder_config(10))
    learn = sum(int(np.prod(s)) for k, s in ds.items() if 'running' not in k)
    stats = sum(int(np.prod(s)) for k, s in ds.items() if 'running' in k)
    assert learn == 942562 and stats == 1504
    assert sum(int(np.prod(s)) for k, s in decoder_param_shapes(decoder_config(9)).items() if 'running' not in k) == 928130
    assert sum(int(np.prod(s)) for k, s in decoder_param_shapes(decoder_config(8)).items() if 'running' not in k) == 888898
    assert gs['net7.block0.weight'] == (256, 128, 4, 4)     # Deconvolution weight is (Cin, Cout, 4, 4)
    assert gs['net6.block0.weight'] == (256, 512, 3, 3)
    for k in gs:
        assert canonical(legacy_name(k)) == k
    assert canonical('1024_conv_to_rgb_weight') == 'to_rgb10.0.weight'
    assert canonical('mp_dense_7_bias') == 'mapping.15.bias'
    assert canonical('arg:128_deconv_1_weight') == 'net7.block0.weight'
    assert canonical('8_adain_2_dense_affine_bias') == 'net3.adain2.affine.bias'


def test_params_codec_roundtrip_and_tolerant_reader(tmp_path):
    dp = init_decoder_params(decoder_config(8), seed=1)
    f = tmp_path / 'checkpoint_last.params'
    save_params(str(f), dp)
    q = load_params(str(f))
    assert list(q) == list(dp) and all(np.array_equal(q[k], dp[k]) for k in dp)
    # V1 header (no storage type) and legacy header (no magic, u32 dims) of the same array
    import struct
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    v1 = struct.pack('<QQQ', 0x112, 0, 1) + struct.pack('<II', 0xF993FAC8, 2) + struct.pack('<2q', 2, 3) + \
        struct.pack('<iii', 1, 0, 0) + a.tobytes() + struct.pack('<Q', 1) + struct.pack('<Q', 5) + b'arg:w'
    (tmp_path / 'v1.params').write_bytes(v1)
    assert np.array_equal(load_params(str(tmp_path / 'v1.params'))['w'], a)
    legacy = struct.pack('<QQQ', 0x112, 0, 1) + struct.pack('<I', 2) + struct.pack('<2I', 2, 3) + \
        struct.pack('<iii', 1, 0, 0) + a.tobytes() + struct.pack('<Q', 0)
    (tmp_path / 'legacy.params').write_bytes(legacy)
    assert np.array_equal(load_params(str(tmp_path / 'legacy.params'))['0'], a)
    (tmp_path / 'bad.params').write_bytes(b'\x00' * 64)
    with pytest.raises(ValueError):
        load_params(str(tmp_path / 'bad.params'))


def test_create_without_gpu_reports_error(gsx_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('needs a box without a GPU')
    cfg = L.SynthCfg(8, 4, 4, 8192, 1.0, 512, 512, 3)
    h = C.c_void_p()
    assert gsx_lib.gsx_synth_create(C.byref(cfg), C.byref(h)) < 0
    assert b'no CPU fallback' in gsx_lib.gsx_last_error()


def test_batch_split_matches_split_and_load():
    from gan_segmentation_b200.image_generator import ImageGenerator
    g = ImageGenerator.__new__(ImageGenerator)
    g.ctx = [0, 1, 2]
    assert g._split(8) == [(0, 3), (3, 6), (6, 8)]
    assert g._split(2) == [(0, 1), (1, 2), (2, 2)]


def test_cli_reads_the_reference_config_keys(tmp_path):
    """main.py:15-43: same actions, same config.yml keys; the GUI action is declined, compute actions need a GPU."""
    from gan_segmentation_b200 import main as M
    cfg = tmp_path / 'config.yml'
    cfg.write_text('BASE_DIR: "%s"\nGAN: "bedrooms"\nGAN_DIR: "stylegan-models"\nGAN_GPU_IDS: [0]\n'
                   'GAN_BATCH_SIZE_PER_GPU: 8\nSOLVER_GPU_IDS: [0]\nANNOTATION: "segmentation"\nGENERATE_NUM: 10000\n' % tmp_path)
    a = M.parse_args(['generate', '--config', str(cfg)])
    assert a.action == 'generate' and M.parse_args([]).action == 'annotation'          # main.py:18-20 default
    c = M.load_config_file(str(cfg))
    assert c['GAN'] == 'bedrooms' and c['GENERATE_NUM'] == 10000 and c['GAN_GPU_IDS'] == [0]
    assert M.MAX_RES_LOG2 == {'ffhq': 10, 'cars': 9, 'bedrooms': 8}
    assert M.main(['annotation', '--config', str(cfg)]) == 2
    import torch
    if not torch.cuda.is_available():
        import pytest
        with pytest.raises(Exception):                                                 # no CPU fallback
            M.main(['generate', '--config', str(cfg), '--random-init'])


def test_every_runtime_option_is_documented_in_the_header():
    """gsx_set_option names accepted by csrc/api.cu all appear (quoted) in include/gsx.h."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, 'gan-segmentation_b200', 'csrc', 'api.cu')).read()
    hdr = open(os.path.join(root, 'include', 'gsx.h')).read()
    names = set(re.findall(r'std::strcmp\(name, "([a-z_0-9]+)"\)', src))
    assert len(names) >= 8
    missing = sorted(n for n in names if f'"{n}"' not in hdr)
    assert not missing, missing
