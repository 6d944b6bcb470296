"""Codec for MXNet's NDArray-list container -- the on-disk format of ``stylegan-*.params``
(image_generator.py:21-22) and ``checkpoint_last.params`` (seg_solver.py:331-349, written by
Gluon ``save_parameters`` = ``mx.nd.save`` of {structural name: array}).

MXNet is not installable offline and no sample file exists in the reference tree, so the layout
below is restated from MXNet 1.x's ``NDArray::Save/Load`` (src/ndarray/ndarray.cc) and is only
round-trip tested here; the reader is tolerant (V1/V2/V3 headers, legacy no-magic arrays).

  file   := u64 0x112 | u64 reserved | u64 n_arrays | array* | u64 n_names | (u64 len | bytes)*
  array  := u32 magic (V2 0xF993FAC9, V3 0xF993FACA, V1 0xF993FAC8) | [V2/V3: i32 stype (0 = dense)]
            | shape (u32 ndim | i64 dim*; V3: i32 ndim) | i32 dev_type | i32 dev_id | i32 type_flag | raw data
  legacy := u32 ndim | u32 dim* | i32 dev_type | i32 dev_id | i32 type_flag | raw data
All little-endian.
"""
from __future__ import annotations

import struct

import numpy as np

LIST_MAGIC = 0x112
V1_MAGIC, V2_MAGIC, V3_MAGIC = 0xF993FAC8, 0xF993FAC9, 0xF993FACA
_DTYPES = {0: np.float32, 1: np.float64, 2: np.float16, 3: np.uint8, 4: np.int32, 5: np.int8, 6: np.int64}
_FLAGS = {np.dtype(v): k for k, v in _DTYPES.items()}


class _Reader:
    def __init__(self, buf):
        self.b, self.o = buf, 0

    def take(self, fmt):
        v = struct.unpack_from('<' + fmt, self.b, self.o)
        self.o += struct.calcsize('<' + fmt)
        return v if len(v) > 1 else v[0]

    def raw(self, n):
        if self.o + n > len(self.b):
            raise ValueError('truncated .params file')
        v = self.b[self.o:self.o + n]
        self.o += n
        return v


def _read_array(r):
    magic = r.take('I')
    if magic in (V2_MAGIC, V3_MAGIC):
        stype = r.take('i')
        if stype != 0:
            raise ValueError(f'sparse storage type {stype} is not supported')
        ndim = r.take('i' if magic == V3_MAGIC else 'I')
        if ndim < 0:
            return None
        shape = [r.take('q') for _ in range(ndim)]
    elif magic == V1_MAGIC:
        ndim = r.take('I')
        shape = [r.take('q') for _ in range(ndim)]
    else:                                   # legacy: the word just read was ndim, dims are u32
        ndim = magic
        if ndim > 32:
            raise ValueError('not an MXNet NDArray (bad magic)')
        shape = [r.take('I') for _ in range(ndim)]
    if ndim == 0 and magic != V3_MAGIC:
        return None                         # "none" array: nothing else is stored
    r.take('i')                             # dev_type
    r.take('i')                             # dev_id
    flag = r.take('i')
    if flag not in _DTYPES:
        raise ValueError(f'unknown dtype flag {flag}')
    dt = np.dtype(_DTYPES[flag]).newbyteorder('<')
    n = int(np.prod(shape)) if shape else 1
    data = np.frombuffer(r.raw(n * dt.itemsize), dtype=dt).reshape(shape)
    return np.array(data, dtype=_DTYPES[flag])


def load_params(filename):
    """-> {name: ndarray}.  'arg:' / 'aux:' prefixes of symbol-era checkpoints are stripped."""
    with open(filename, 'rb') as f:
        r = _Reader(f.read())
    magic, _ = r.take('Q'), r.take('Q')
    if magic != LIST_MAGIC:
        raise ValueError(f'{filename}: not an MXNet NDArray list (magic {magic:#x})')
    n = r.take('Q')
    arrays = [_read_array(r) for _ in range(n)]
    n_names = r.take('Q')
    if n_names not in (0, n):
        raise ValueError('name count does not match array count')
    names = []
    for _ in range(n_names):
        ln = r.take('Q')
        names.append(bytes(r.raw(ln)).decode('utf-8'))
    if not names:
        names = [str(i) for i in range(n)]
    out = {}
    for k, a in zip(names, arrays):
        if a is None:
            continue
        if k.startswith('arg:') or k.startswith('aux:'):
            k = k[4:]
        out[k] = a
    return out


def save_params(filename, params):
    """Write {name: ndarray} as an NDArray list with V2 headers on cpu(0) (what ``mx.nd.save`` writes)."""
    chunks = [struct.pack('<QQQ', LIST_MAGIC, 0, len(params))]
    for a in params.values():
        a = np.ascontiguousarray(a)
        if a.dtype not in _FLAGS:
            a = a.astype(np.float32)
        chunks.append(struct.pack('<Ii', V2_MAGIC, 0))
        chunks.append(struct.pack('<I', a.ndim))
        chunks.append(struct.pack(f'<{a.ndim}q', *a.shape))
        chunks.append(struct.pack('<iii', 1, 0, _FLAGS[a.dtype]))          # cpu(0)
        chunks.append(a.astype(a.dtype.newbyteorder('<'), copy=False).tobytes())
    chunks.append(struct.pack('<Q', len(params)))
    for k in params:
        kb = k.encode('utf-8')
        chunks.append(struct.pack('<Q', len(kb)))
        chunks.append(kb)
    with open(filename, 'wb') as f:
        f.write(b''.join(chunks))
