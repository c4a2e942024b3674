"""
Gaussian blur kernels on device-resident 1080p video: the tensor-core kernel (va_gauss_mma.cu) against the dot-product
kernels (VA_GAUSS_MMA=0), plain and fused with the monochrome conversion, several sigmas; checks that both paths
produce identical bytes.      python tools/blur_bench.py [--batch 64] [--sigmas 1,2,3,5,15]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analysis_b200 import synth  # noqa: E402
from video_analysis_b200.device import get_runtime  # noqa: E402


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--w', type=int, default=1920)
    ap.add_argument('--h', type=int, default=1080)
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--sigmas', default='1,2,3,5,15')
    ap.add_argument('--peak', type=float, default=6538.9)
    a = ap.parse_args()
    W, H, B = a.w, a.h, a.batch
    N = W * H
    rt = get_runtime(0)
    rt.ensure(W, H, B)
    rgbs = [synth.generate(rt, 0, i * B, B, W, H) for i in range(2)]
    monos = [rt.luma(r) for r in rgbs]
    out = [rt.empty_u8(B, H, W) for _ in range(2)]
    lib, h = rt.lib, rt._h
    state = {'i': 0}
    for sigma in [float(s) for s in a.sigmas.split(',')]:
        for fused in (False, True):
            res = {}
            for mma in (1, 0):
                os.environ['VA_GAUSS_MMA'] = str(mma)
                o = out[mma]

                def k():
                    state['i'] ^= 1
                    if fused:
                        s = rgbs[state['i']]
                        rt._check(lib.va_luma_gauss_u8(h, rt.stream, s.ptr, s.pitch, s.fstride, o.ptr, o.pitch, o.fstride, W, H, B, -1, sigma))
                    else:
                        s = monos[state['i']]
                        rt._check(lib.va_gauss_u8(h, rt.stream, s.ptr, s.pitch, s.fstride, o.ptr, o.pitch, o.fstride, W, H, 1, B, sigma))
                res[mma] = timeit(k, a.iters)
            state['i'] = 0
            for mma in (1, 0):          # same input for the comparison
                os.environ['VA_GAUSS_MMA'] = str(mma)
                o = out[mma]
                if fused:
                    s = rgbs[0]
                    rt._check(lib.va_luma_gauss_u8(h, rt.stream, s.ptr, s.pitch, s.fstride, o.ptr, o.pitch, o.fstride, W, H, B, -1, sigma))
                else:
                    s = monos[0]
                    rt._check(lib.va_gauss_u8(h, rt.stream, s.ptr, s.pitch, s.fstride, o.ptr, o.pitch, o.fstride, W, H, 1, B, sigma))
            torch.cuda.synchronize()
            same = bool(torch.equal(out[0].t, out[1].t))
            by = (4 if fused else 2) * N * B
            print(json.dumps({'sigma': sigma, 'fused': fused, 'batch': B, 'mma_ms': round(res[1], 4), 'dot_ms': round(res[0], 4),
                              'speedup': round(res[0] / res[1], 2), 'mma_GBps': round(by / res[1] / 1e6, 1),
                              'mma_frac': round(by / res[1] / 1e6 / a.peak, 3), 'dot_frac': round(by / res[0] / 1e6 / a.peak, 3),
                              'identical': same}), flush=True)
    os.environ.pop('VA_GAUSS_MMA', None)


if __name__ == '__main__':
    main()
