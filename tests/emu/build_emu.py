"""
TEST INFRASTRUCTURE ONLY: compiles the csrc/*.cu kernel sources with g++ against
tests/emu/cuda_emu.h (threads-as-CUDA-threads shim) into
tests/emu/libva_b200_emu.so, so kernel logic can be checked on a GPU-less box.
Never imported by the product package.
"""

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'video_analysis_b200', 'csrc')
INCLUDE = os.path.join(ROOT, 'include')
LIB = os.path.join(HERE, 'libva_b200_emu.so')
SOURCES = ['va_api.cu', 'va_pointwise.cu', 'va_gauss.cu', 'va_gauss_mma.cu', 'va_ema.cu', 'va_morph.cu', 'va_label.cu', 'va_extra.cu', 'va_export.cu']


def build(force=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh'))]
    deps += [os.path.join(HERE, 'cuda_emu.h'), os.path.join(INCLUDE, 'va_b200.h'), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(HERE, src.replace('.cu', '.emu.o'))
        cmd = ['g++', '-std=c++20', '-O2', '-g', '-fPIC', '-DVA_EMU', '-ffp-contract=off', '-x', 'c++',
               '-I', HERE, '-I', INCLUDE, '-I', CSRC, '-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            print('--- g++ %s\n%s' % (src, out))
            failed = True
    if failed:
        raise RuntimeError('emulation build failed')
    subprocess.run(['g++', '-shared', '-o', LIB] + objs + ['-lpthread'], check=True)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
