"""
Tuning aid for the blur kernels: times va_luma_gauss_u8 / va_gauss_u8 under a list of environment
settings (read by the launcher at every call) and checks every variant against the tile kernel's
output.
    python tools/gbench.py [--w 1920 --h 1080 --batch 64 --sigma 2] VAR=VAL,VAR=VAL ...
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analysis_b200 import synth  # noqa: E402
from video_analysis_b200.device import get_runtime  # noqa: E402
from kbench import timeit  # noqa: E402

KEYS = ('VA_GAUSS_STREAM', 'VA_GS_NT', 'VA_GS_SEGS', 'VA_GAUSS_TH')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--w', type=int, default=1920)
    ap.add_argument('--h', type=int, default=1080)
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--sigma', type=float, default=2.0)
    ap.add_argument('variants', nargs='*')
    a = ap.parse_args()
    W, H, B = a.w, a.h, a.batch
    rt = get_runtime(0)
    rt.ensure(W, H, B)
    rgbs = [synth.generate(rt, 0, i * B, B, W, H) for i in range(2)]
    monos = [rt.luma(r) for r in rgbs]
    lib, h = rt.lib, rt._h
    outs = [rt.empty_u8(B, H, W) for _ in range(2)]
    state = {'i': 0}

    def setenv(spec):
        for k in KEYS:
            os.environ.pop(k, None)
        for kv in filter(None, spec.split(',')):
            k, v = kv.split('=')
            os.environ[k] = v

    def run(fused, src, dst):
        if fused:
            rt._check(lib.va_luma_gauss_u8(h, rt.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch, dst.fstride,
                                           W, H, B, -1, a.sigma))
        else:
            rt._check(lib.va_gauss_u8(h, rt.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch, dst.fstride,
                                      W, H, 1, B, a.sigma))

    for fused in (True, False):
        srcs = rgbs if fused else monos
        setenv('VA_GAUSS_STREAM=0')
        run(fused, srcs[0], outs[0])
        torch.cuda.synchronize()
        ref = outs[0].t.clone() if hasattr(outs[0], 't') else None
        for spec in ['VA_GAUSS_STREAM=0'] + a.variants:
            setenv(spec)
            outs[1].t.zero_()
            run(fused, srcs[0], outs[1])
            torch.cuda.synchronize()
            same = bool(torch.equal(outs[1].t, ref))

            def fn():
                state['i'] ^= 1
                run(fused, srcs[state['i']], outs[1])
            med, best = timeit(fn, a.iters)
            print(json.dumps({'fused': fused, 'env': spec, 'ms': round(med, 4), 'ms_min': round(best, 4), 'same': same}), flush=True)


if __name__ == '__main__':
    main()
