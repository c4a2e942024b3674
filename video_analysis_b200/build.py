"""
Builds csrc/libva_b200.so for sm_100a with nvcc (cross-compiles without a GPU).

    python -m video_analysis_b200.build [--force] [--ptxas-v]

The library is kept in-tree next to its sources so that it travels with the
repository snapshot to the GPU box.
"""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
LIB = os.path.join(CSRC, 'libva_b200.so')
SOURCES = ['va_api.cu', 'va_pointwise.cu', 'va_gauss.cu', 'va_gauss_mma.cu', 'va_ema.cu', 'va_morph.cu', 'va_label.cu', 'va_extra.cu', 'va_export.cu']
HEADERS = ['va_common.cuh', 'va_device.cuh', 'va_mma.cuh', os.path.join(INCLUDE, 'va_b200.h')]

NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-I', INCLUDE, '-I', CSRC]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, ptxas_verbose=False, quiet=False):
    """ compile every CUDA source into one shared library; returns its path """
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if ptxas_verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip() and (not quiet or p.returncode != 0):
            print('--- nvcc %s\n%s' % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, ptxas_verbose='--ptxas-v' in sys.argv)
    print(path)
