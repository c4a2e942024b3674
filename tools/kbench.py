"""
Per-kernel timing on device-resident synthetic video: CUDA events on the launch stream,
inputs far larger than L2.  Prints one line per kernel with achieved algorithmic GB/s.
    python tools/kbench.py [--w 1920 --h 1080 --batch 64 --iters 10 --sigma 2]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analysis_b200 import synth  # noqa: E402
from video_analysis_b200.chain import SegmentChain  # noqa: E402
from video_analysis_b200.device import DeviceBatch, get_runtime  # noqa: E402


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
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--w', type=int, default=1920)
    ap.add_argument('--h', type=int, default=1080)
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--sigma', type=float, default=2.0)
    ap.add_argument('--k', type=int, default=3)
    ap.add_argument('--peak', type=float, default=6538.9)
    a = ap.parse_args()
    W, H, B = a.w, a.h, a.batch
    N = W * H
    rt = get_runtime(0)
    rt.ensure(W, H, B)
    # two alternating input batches so that nothing is L2 resident between iterations
    rgbs = [synth.generate(rt, 0, i * B, B, W, H) for i in range(2)]
    torch.cuda.synchronize()
    state = {'i': 0}

    def rgb():
        state['i'] ^= 1
        return rgbs[state['i']]

    res = []

    def report(name, bytes_per_frame, fn):
        med, best = timeit(fn, a.iters)
        gbs = bytes_per_frame * B / (med * 1e-3) / 1e9
        res.append({'kernel': name, 'ms': round(med, 4), 'ms_min': round(best, 4), 'alg_GBps': round(gbs, 1),
                    'frac': round(gbs / a.peak, 3), 'fps': round(B / (med * 1e-3))})
        print(json.dumps(res[-1]), flush=True)

    monos = [rt.luma(r) for r in rgbs]
    blurs = [rt.gauss(m, a.sigma) for m in monos]
    bg = rt.empty_f32(H, W)
    masks = [rt.ema_diff_thresh(bl, bg, 0.05, 25, first) for bl, first in zip(blurs, (True, False))]
    morphs = [rt.morph(m, 'open', 'rect', a.k) for m in masks]
    labs = [rt.empty_i32(B, H, W) for _ in range(2)]
    counts = torch.empty((B,), dtype=torch.int32, device=rt.device)
    lib, h = rt.lib, rt._h

    out_u8 = rt.empty_u8(B, H, W)
    out_bits = rt.empty_bits(B, H, W)

    def k_luma():
        s = rgb()
        rt._check(lib.va_luma_u8(h, rt.stream, s.ptr, s.pitch, s.fstride, out_u8.ptr, out_u8.pitch, out_u8.fstride, W, H, B, -1))
    report('K1 luma', 4 * N, k_luma)

    def k_gauss():
        s = monos[state['i']]; state['i'] ^= 1
        rt._check(lib.va_gauss_u8(h, rt.stream, s.ptr, s.pitch, s.fstride, out_u8.ptr, out_u8.pitch, out_u8.fstride, W, H, 1, B, a.sigma))
    report('K2 gauss s=%g' % a.sigma, 2 * N, k_gauss)

    def k_lg():
        s = rgb()
        rt._check(lib.va_luma_gauss_u8(h, rt.stream, s.ptr, s.pitch, s.fstride, out_u8.ptr, out_u8.pitch, out_u8.fstride, W, H, B, -1, a.sigma))
    report('K1+K2 luma_gauss s=%g' % a.sigma, 4 * N, k_lg)

    def k_ema():
        s = blurs[state['i']]; state['i'] ^= 1
        rt._check(lib.va_ema_diff_thresh(h, rt.stream, s.ptr, s.pitch, s.fstride, bg.data_ptr(), bg.stride(0),
                                         out_bits.ptr, out_bits.pitch, out_bits.fstride, W, H, B, 0.05, 25.0, 0))
    report('K3 ema_diff_thresh', N + N / 8 + 8 * N / B, k_ema)

    def k_morph():
        s = masks[state['i']]; state['i'] ^= 1
        rt._check(lib.va_morph_bits(h, rt.stream, s.ptr, s.pitch, s.fstride, out_bits.ptr, out_bits.pitch, out_bits.fstride,
                                    W, H, B, 2, 0, a.k, a.k))
    report('K4 open %dx%d' % (a.k, a.k), N / 4, k_morph)

    def k_label():
        s = morphs[state['i']]; l = labs[state['i']]; state['i'] ^= 1
        rt._check(lib.va_label_bits(h, rt.stream, s.ptr, s.pitch, s.fstride, l.ptr, l.pitch, l.fstride, counts.data_ptr(), W, H, B, 4))
    report('K5 label', 4.125 * N, k_label)

    def k_half():
        s = monos[state['i']]; state['i'] ^= 1
        rt._check(lib.va_resize_half_u8(h, rt.stream, s.ptr, s.pitch, s.fstride, out_u8.ptr, out_u8.pitch, out_u8.fstride, W, H, 1, B))
    report('K2b resize_half', 1.25 * N, k_half)

    for fuse in (False, True):
        ch = SegmentChain((W, H), sigma=a.sigma, morph_ksize=a.k, batch=B, fuse=fuse)

        def k_chain():
            s = rgb(); l = labs[state['i']]
            ch.run_device(s, l, counts)
        k_chain()
        report('chain fuse=%s' % fuse, (9.5 if fuse else 11.5) * N + 8 * N / B, k_chain)
    ch = SegmentChain((W, H), sigma=a.sigma, morph_ksize=a.k, batch=B, fuse=True)

    def k_pipe():
        for _ in range(4):
            s = rgb(); l = labs[state['i']]
            ch.run_device_pipelined(s, l, counts)
        ch.pipeline_sync()
    k_pipe()
    med, best = timeit(k_pipe, a.iters)
    print(json.dumps({'kernel': 'chain two-stream overlap (per step)', 'ms': round(med / 4, 4), 'fps': round(4 * B / (med * 1e-3))}), flush=True)
    print('counts', counts[:8].tolist(), 'launches', rt.launches)


if __name__ == '__main__':
    main()
