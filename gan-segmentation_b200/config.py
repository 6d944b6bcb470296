"""Hyper-parameter dictionaries of the generate hot path.

Mirrors the two hard-coded dicts of the reference so callers see the same keys:
  * generator: ``ImageGenerator._get_config``  (reference image_generator.py:46-74)
  * decoder:   ``SegSolver.get_config``        (reference seg_solver.py:83-132)
plus the derived shape tables every other module (host mirror, oracle, CUDA plan)
is built from.
"""
from __future__ import annotations

MAX_RES_LOG2 = {'ffhq': 10, 'cars': 9, 'bedrooms': 8}     # image_generator.py:11


def generator_config(max_res_log2=9, base_scale_y=4, base_scale_x=4):
    """image_generator.py:46-74 (same keys, same values). ``base_scale_*`` is exposed
    because Generator itself is non-square capable (networks_stylegan.py:84-85,94)."""
    cfg = {}
    cfg['use_wscale'] = True
    cfg['fmap_base'] = 8192
    cfg['fmap_decay'] = 1.0
    cfg['fmap_max'] = 512
    cfg['max_res_log2'] = max_res_log2
    cfg['fix_noise'] = False
    cfg['base_scale_x'] = base_scale_x
    cfg['base_scale_y'] = base_scale_y
    cfg['init'] = 'normal'
    cfg['init_normal_std'] = 1.0
    cfg['init_xavier_magnitude'] = 1.0
    cfg['latent_size'] = 512
    cfg['latent_prior'] = 'normal'
    cfg['channels'] = 3
    cfg['imrange'] = (-1, 1)
    cfg['dtype'] = 'fp32'
    return cfg


def decoder_config(max_res_log2=9):
    """seg_solver.py:83-132 (same keys, same values)."""
    cfg = {}
    cfg['seed'] = 1
    cfg['kvstore'] = 'nccl'
    cfg['cache_max_size'] = 4
    cfg['plot_graph'] = True
    cfg['num_classes'] = 2
    cfg['not_ignore_classes'] = None
    cfg['cls_type'] = 'hair'
    cfg['train_epochs'] = 24
    cfg['base_lr'] = 1e-4
    cfg['factor_d'] = 0.1
    cfg['wd'] = 0.0
    cfg['optimizer'] = 'adam'
    cfg['momentum'] = None
    cfg['scheduler'] = None
    cfg['preprocess_mask'] = True
    cfg['train_display_iters'] = 4
    cfg['train_batch_size'] = 1
    cfg['val_batch_size'] = 1
    cfg['val_loader_workers'] = 0
    cfg['train_loader_workers'] = 0
    cfg['train_show_images'] = 1
    cfg['val_show_images'] = 1
    cfg['val_report_intermediate'] = False
    cfg['val_report_interval'] = 0.34
    cfg['use_bn'] = True
    cfg['use_sync_bn'] = False
    cfg['use_dropout'] = True
    cfg['start_res'] = 0
    cfg['features'] = [32, 32, 32, 32, 32, 32, 32, 32, 16]
    cfg['in_channels'] = [512, 512, 512, 512, 256, 128, 64, 32, 16]
    cfg['features'] = cfg['features'][:max_res_log2 - 1] + [cfg['num_classes']]
    cfg['in_channels'] = cfg['in_channels'][:max_res_log2 - 1]
    cfg['dtype'] = 'fp32'
    return cfg


def num_features(cfg, res_log2):
    """networks_stylegan.py:114-116."""
    fmaps = int(cfg['fmap_base'] / (2.0 ** ((res_log2 - 1) * cfg['fmap_decay'])))
    return min(fmaps, cfg['fmap_max'])


def num_style_layers(cfg):
    """Length of truncation_psi (networks_stylegan.py:99)."""
    return (cfg['max_res_log2'] - 1) * 2


def block_hw(cfg, res_log2):
    """Spatial size of block ``res_log2``'s output: base * 2^(res_log2-2)."""
    s = 2 ** (res_log2 - 2)
    return cfg['base_scale_y'] * s, cfg['base_scale_x'] * s


def noise_shapes(cfg, n):
    """The 2*(L-1) noise planes [n,1,h,w], in forward order (two per block)."""
    out = []
    for r in range(2, cfg['max_res_log2'] + 1):
        h, w = block_hw(cfg, r)
        out += [(n, 1, h, w), (n, 1, h, w)]
    return out
