"""
The C-ABI library: builds for sm_100a, loads without a GPU and exports exactly the symbols
include/va_b200.h declares.  No compute calls here (CPU only).
"""

import ctypes
import os
import re

import numpy as np
import pytest

from oracle import ops
from video_analysis_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    build.build(quiet=True)
    return _lib.load()


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'va_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(va_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = header_symbols()
    assert len(syms) >= 24
    for name in syms:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, 'no ctypes signature for %s' % name
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_status_strings(lib):
    assert lib.va_version() >= 100
    assert lib.va_status_string(0) == b'ok'
    assert b'invalid' in lib.va_status_string(-1)


def test_host_side_gaussian_taps_equal_opencv(lib):
    buf = (ctypes.c_int * 256)()
    for s in np.arange(0.1, 21.0, 0.05):
        n = lib.va_gauss_taps(float(s), buf, 256)
        assert n > 0
        assert np.array_equal(np.array(buf[:n]), ops.gauss_kernel_u8(float(s))), s
    assert lib.va_gauss_taps(-1.0, buf, 256) == _lib.VA_ERR_INVALID
    assert lib.va_gauss_taps(30.0, buf, 16) == _lib.VA_ERR_CAPACITY


def test_null_ctx_is_an_error_not_a_crash(lib):
    assert lib.va_luma_u8(None, None, 0, 0, 0, 0, 0, 0, 1, 1, 1, -1) == _lib.VA_ERR_INVALID
    assert lib.va_destroy(None) == 0
    assert lib.va_last_error(None) == b'null ctx'


def test_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    h = ctypes.c_void_p()
    assert lib.va_create(ctypes.byref(h), 0, 64, 64, 1) == _lib.VA_ERR_CUDA
    assert not h.value
    from video_analysis_b200.device import get_runtime
    with pytest.raises(_lib.VAError):
        get_runtime()


def test_product_never_imports_the_oracle_or_the_emulator():
    pkg = os.path.join(ROOT, 'video_analysis_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+(oracle|tests)\b', src, flags=re.M), f
                assert 'libva_b200_emu' not in src, f
