"""ctypes binding of libgsx.so (include/gsx.h).  There is no CPU fallback: if the shared library is
missing or fails to load, every product entry point raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# One shared library per 16-bit storage / tensor-core operand type (csrc/Makefile):
#   fp16 -> libgsx.so (default: meets the image tolerance), bf16 -> libgsx_bf16.so (the north star's type)
LIB_PATHS = {'fp16': os.path.join(_HERE, 'libgsx.so'), 'bf16': os.path.join(_HERE, 'libgsx_bf16.so')}
if os.environ.get('GSX_LIB'):                      # development hook: an alternative build of the fp16 library
    LIB_PATHS['fp16'] = os.environ['GSX_LIB']
LIB_PATH = LIB_PATHS['fp16']
DEFAULT_DTYPE = os.environ.get('GSX_DTYPE', 'fp16')
CSRC = os.path.join(_HERE, 'csrc')

# enums of gsx_internal.h / gsx.h
CONV3, UPCONV3, DECONV4, CONV1, DECONV4B = 0, 1, 2, 3, 4
EPI_LRELU, EPI_STATS, EPI_ARGMAX = 1, 2, 4


class SynthCfg(C.Structure):
    _fields_ = [('max_res_log2', C.c_int), ('base_scale_y', C.c_int), ('base_scale_x', C.c_int),
                ('fmap_base', C.c_int), ('fmap_decay', C.c_float), ('fmap_max', C.c_int),
                ('latent_size', C.c_int), ('channels', C.c_int)]


class DecCfg(C.Structure):
    _fields_ = [('num_levels', C.c_int), ('in_channels', C.c_int * 16), ('features', C.c_int * 17),
                ('use_bn', C.c_int), ('base_y', C.c_int), ('base_x', C.c_int)]


class PlanOverride(C.Structure):
    _fields_ = [('TH', C.c_int), ('TW', C.c_int), ('NB', C.c_int), ('CBK', C.c_int), ('N_tile', C.c_int),
                ('stages', C.c_int), ('phase_grid', C.c_int), ('epi_groups', C.c_int), ('acc_bufs', C.c_int),
                ('max_mtiles', C.c_int), ('hstack', C.c_int), ('s2d', C.c_int)]


class GsxError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libgsx.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(['make', '-C', CSRC, '-j8'], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode:
        raise GsxError('building libgsx.so failed')
    return LIB_PATH


_libs = {}

_vp, _fp, _i, _u64, _sz = C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_size_t

_SIGS = {
    'gsx_last_error': (C.c_char_p, []),
    'gsx_abi_version': (_i, []),
    'gsx_launch_count': (_u64, []),
    'gsx_set_option': (_i, [C.c_char_p, _i]),
    'gsx_synth_create': (_i, [C.POINTER(SynthCfg), C.POINTER(_vp)]),
    'gsx_synth_destroy': (None, [_vp]),
    'gsx_synth_set_param': (_i, [_vp, C.c_char_p, _fp, C.POINTER(C.c_int64), _i]),
    'gsx_synth_finalize': (_i, [_vp]),
    'gsx_synth_workspace_bytes': (_i, [_vp, _i, C.POINTER(_sz)]),
    'gsx_synth_num_layers': (_i, [_vp]),
    'gsx_synth_feature_shape': (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    'gsx_synth_forward': (_i, [_vp, _i, _fp, _fp, C.POINTER(_vp), _u64, _u64, _fp, _vp, C.POINTER(_vp), _vp, _sz, _vp]),
    'gsx_synth_export_noise': (_i, [_vp, _i, _i, _fp, _vp, _vp]),
    'gsx_synth_export_latents': (_i, [_vp, _i, _fp, _vp, _vp]),
    'gsx_dec_create': (_i, [C.POINTER(DecCfg), C.POINTER(_vp)]),
    'gsx_dec_destroy': (None, [_vp]),
    'gsx_dec_set_param': (_i, [_vp, C.c_char_p, _fp, C.POINTER(C.c_int64), _i]),
    'gsx_dec_finalize': (_i, [_vp]),
    'gsx_dec_workspace_bytes': (_i, [_vp, _i, C.POINTER(_sz)]),
    'gsx_dec_forward': (_i, [_vp, _i, C.POINTER(_vp), _vp, _vp, _fp, _vp, _vp, _sz, _vp]),
    'gsx_generate_dev': (_i, [_vp, _vp, _i, _fp, _fp, _u64, _u64, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    'gsx_generate_host': (_i, [_vp, _vp, _i, _fp, _fp, _u64, _u64, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp, _i]),
    'gsx_synth_device_counter': (_i, [_vp, _i, C.c_uint64]),
    'gsx_op_conv_wgrad': (_i, [_i, _i, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _vp]),
    'gsx_op_conv_wgrad_tc': (_i, [_i, _i, _i, _i, _i, _i, _fp, _fp, _fp, _vp]),
    'gsx_op_release_cache': (None, []),
    'gsx_op_upsample2': (_i, [_fp, _fp, _i, _i, _i, _i, _vp]),
    'gsx_op_sumpool2': (_i, [_fp, _fp, _i, _i, _i, _i, _vp]),
    'gsx_op_bn_lrelu_fwd': (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _vp]),
    'gsx_op_bn_lrelu_bwd': (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _vp]),
    'gsx_softmax_ce': (_i, [_fp, _vp, _i, _i, _i, _i, _fp, _fp, C.c_float, _fp, _sz, _vp]),
    'gsx_adam_step': (_i, [_fp, _fp, _fp, _fp, _sz, _i, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _vp]),
    'gsx_train_create': (_i, [C.POINTER(DecCfg), _i, _i, C.POINTER(_vp)]),
    'gsx_train_destroy': (None, [_vp]),
    'gsx_train_param_count': (_i, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    'gsx_train_param_info': (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_sz), C.POINTER(_sz)]),
    'gsx_train_workspace_bytes': (_i, [_vp, C.POINTER(_sz)]),
    'gsx_train_dropout_mask': (_i, [_vp, _i, _u64, _fp, _vp]),
    'gsx_train_set_seed_buffer': (_i, [_vp, _vp]),
    'gsx_train_step': (_i, [_vp, _fp, _fp, C.POINTER(_vp), _vp, _vp, _vp, _u64, _fp, _vp, C.POINTER(C.c_float), _vp, _sz, _vp]),
    'gsx_profile_enable': (_i, [_i]),
    'gsx_profile_dump': (_i, [C.c_char_p, _sz]),
    'gsx_op_conv': (_i, [_i] * 7 + [_fp, _fp, _fp, _fp, _fp, _fp, _i, _fp, _fp, _fp, _vp, _fp, _i,
                         C.POINTER(PlanOverride), C.POINTER(_i), _i, C.POINTER(C.c_float), _vp]),
    'gsx_plan_query': (_i, [_i] * 7 + [C.POINTER(PlanOverride), C.POINTER(_i)]),
    'gsx_op_pass1': (_i, [_i, _i, _i, _i, _fp, _i, _i, _fp, _fp, _fp, _fp, _fp, _vp]),
    'gsx_op_apply': (_i, [_i, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _i, _fp, _fp, _vp, _vp]),
    'gsx_op_fill_normal': (_i, [_fp, _sz, _i, _u64, _u64, _i, _vp]),
}

EXPORTS = tuple(_SIGS)


def lib(dtype=None):
    """Load the library for ``dtype`` (once).  Raises GsxError when it is missing -- there is no fallback."""
    dtype = dtype or DEFAULT_DTYPE
    if dtype not in LIB_PATHS:
        raise GsxError(f'unknown dtype {dtype!r} (fp16 or bf16)')
    if dtype not in _libs:
        path = LIB_PATHS[dtype]
        if not os.path.exists(path):
            raise GsxError(f'{path} not found: run __graft_entry__.build() (make -C {CSRC}); '
                           'the generate path has no CPU fallback')
        l = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        # A/B switches for whole test / bench runs: GSX_OPTIONS="pdl=1,varn=0" -> gsx_set_option at load time
        for kv in filter(None, os.environ.get('GSX_OPTIONS', '').split(',')):
            k, v = kv.split('=')
            if l.gsx_set_option(k.strip().encode(), int(v)) < 0:
                raise GsxError(f'GSX_OPTIONS: {l.gsx_last_error().decode()}')
        _libs[dtype] = l
    return _libs[dtype]


def check(rc, what='', dtype=None):
    if rc < 0:
        msg = lib(dtype).gsx_last_error()
        raise GsxError(f'{what}: {msg.decode() if msg else rc}')
    return rc


def launch_count():
    """Kernels launched by every loaded library in this process."""
    return sum(int(l.gsx_launch_count()) for l in _libs.values())


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
