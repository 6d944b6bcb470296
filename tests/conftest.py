import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def gsx_lib():
    """The C-ABI libraries (fp16 default + bf16), built in-tree if they are not there yet."""
    from gan_segmentation_b200 import _lib
    if not all(os.path.exists(p) for p in _lib.LIB_PATHS.values()):
        _lib.build()
    return _lib.lib()


@pytest.fixture(params=['fp16', 'bf16'])
def dtype(request):
    """16-bit storage type of the build under test: fp16 = libgsx.so (default), bf16 = libgsx_bf16.so."""
    return request.param
