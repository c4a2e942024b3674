import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


_backends = {}


def _backend(name):
    if name not in _backends:
        from tests import harness
        if name == 'emu':
            _backends[name] = harness.EmuBackend()
        else:
            import torch
            if not torch.cuda.is_available():
                pytest.skip('no CUDA device')
            _backends[name] = harness.CudaBackend()
    return _backends[name]


@pytest.fixture(params=['emu', pytest.param('cuda', marks=pytest.mark.gpu)])
def be(request):
    """ kernel backend: `cuda` = the product library on the GPU (the parity gate),
    `emu` = the same kernel sources under the CPU thread-emulation shim (logic check) """
    return _backend(request.param)


@pytest.fixture
def cuda_be():
    return _backend('cuda')
