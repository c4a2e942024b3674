#!/usr/bin/env python
"""
Device-resident throughput of the BASELINE.json configurations that are parity-test cases rather than
the bench line (configs[0], [2] per GPU, [3], [4]); CUDA events on the launch stream, inputs alternate
between resident batches larger than L2, median of `--iters` timed passes after 3 warm-ups.
One JSON line per configuration (frames/s, ms per batch, algorithmic GB/s of the whole chain).

    python tools/configs_bench.py [--iters 10] [--only vga,4k,streams,stencil]
"""

import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(t, fn, iters):
    for _ in range(3):
        fn(0)
    t.cuda.synchronize()
    ms = []
    for i in range(iters):
        a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        a.record()
        fn(i + 1)
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), float(np.min(ms))


def synth_host(seed, stream, n, w, h):
    """ n frames of one camera in host memory (a few distinct frames, cycled by the source) """
    rng = np.random.default_rng(seed * 1000 + stream)
    base = rng.integers(60, 120, (h, w, 3), dtype=np.uint8)
    out = np.empty((n, h, w, 3), np.uint8)
    yy, xx = np.mgrid[:h, :w]
    for k in range(n):
        f = base.copy()
        cx, cy = (200 + 97 * stream + 40 * k) % w, (150 + 53 * stream + 25 * k) % h
        f[(xx - cx) ** 2 + (yy - cy) ** 2 < 60 ** 2] = 200
        out[k] = f
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--only', default='vga,4k,streams,stencil')
    args = ap.parse_args()
    import torch as t
    from video_analysis_b200 import synth
    from video_analysis_b200.chain import SegmentChain
    from video_analysis_b200.device import get_runtime, DeviceBatch
    rt = get_runtime(0)
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
    only = set(args.only.split(','))

    def report(name, frames, med, mn, alg_bytes, note):
        print(json.dumps({'config': name, 'frames_per_batch': frames, 'ms': round(med, 4), 'ms_min': round(mn, 4),
                          'fps': round(frames / med * 1e3, 1), 'alg_GBps': round(alg_bytes / med / 1e6, 1),
                          'frac_of_measured_hbm': round(alg_bytes / med / 1e6 / peak, 3), 'note': note}), flush=True)

    def chain_case(name, w, h, batch, note):
        # configs[0] / configs[2]: mono -> blur s=2 -> EMA -> |diff| > 25 -> 3x3 open -> label(4), fused luma+blur
        n_batches = max(2, int(np.ceil(300e6 / (batch * w * h * 3))))          # > 2 x L2 of input
        vids = [synth.generate(rt, 0, k * batch, batch, w, h, 8) for k in range(n_batches)]
        ch = SegmentChain((w, h), batch=batch)
        labels = rt.empty_i32(batch, h, w)
        counts = t.empty((batch,), dtype=t.int32, device=rt.device)
        med, mn = timed(t, lambda i: ch.run_device(vids[i % n_batches], labels, counts), args.iters)
        n = w * h
        report(name, batch, med, mn, (9.5 * n + 8.0 * n / batch) * batch, note)

    if 'vga' in only:
        chain_case('configs[0] 640x480 chain', 640, 480, 256, 'batch 256, 236 MB of RGB per batch')
    if '4k' in only:
        chain_case('configs[2] 3840x2160 chain (one GPU of the 8)', 3840, 2160, 16, 'batch 16, 398 MB of RGB per batch')

    if 'streams' in only:
        # configs[3]: 64 camera streams 1280x720, ONE launch set per time step over the 64 current frames
        # (video_analysis_b200/streams.py): crop (per-stream rectangle position) + luma + apply-mask (per-stream
        # static mask) + threshold + label
        from video_analysis_b200.streams import MultiStreamSegmenter
        from video_analysis_b200.io.base import VideoBase
        w, h, n_streams = 1280, 720, 64
        cw, ch_ = 1024, 576
        rng = np.random.default_rng(4)
        rects = [(int(rng.integers(0, w - cw)), int(rng.integers(0, h - ch_)), cw, ch_) for _ in range(n_streams)]
        m = np.zeros((n_streams, ch_, cw), np.uint8)
        for s_ in range(n_streams):
            m[s_, 20 + s_ % 7:-30, 40:-10 - s_ % 5] = 1
        seg = MultiStreamSegmenter([VideoBase(size=(w, h), frame_count=1, is_color=True) for _ in range(n_streams)], rects, m,
                                   threshold=110)
        vids = [synth.generate(rt, 3, k * n_streams, n_streams, w, h, 6) for k in range(2)]
        med, mn = timed(t, lambda i: seg.step_device(vids[i % 2]), args.iters)
        n = cw * ch_
        report('configs[3] 64 x 1280x720 streams per launch: crop+mono, apply-mask, threshold, label', n_streams, med, mn,
               ((3 * n + n + n / 8) + (n / 8 + 4 * n)) * n_streams,
               'crop 1024x576 at a position per stream, mask per stream; bytes = fused front (3N RGB + N mask in, N/8 bits '
               'out) + label 4.125N per cropped frame')

        # end to end: frames of the 64 streams in host memory -> gather threads -> pinned block -> upload -> kernels ->
        # label images back on the host (chunk egress), three steps in flight
        import time

        class Camera(VideoBase):
            def __init__(self, data, steps):
                super(Camera, self).__init__(size=(w, h), frame_count=steps, is_color=True)
                self.data = data

            def get_frame(self, index):
                if index >= self.frame_count:
                    raise IndexError
                return self.data[index % len(self.data)]
        host = [synth_host(3, s_, 4, w, h) for s_ in range(n_streams)]
        for sparse in (True, False):
            steps = 40
            seg2 = MultiStreamSegmenter([Camera(host[s_], 5) for s_ in range(n_streams)], rects, m, threshold=110, sparse_egress=sparse)
            for _ in seg2:
                pass                                               # warm-up: buffers, streams, thread pool
            seg2.videos = [Camera(host[s_], steps) for s_ in range(n_streams)]
            seg2.egress_bytes = 0
            t.cuda.synchronize()
            t0 = time.perf_counter()
            sink = 0
            for lab, cnt in seg2:
                sink += int(cnt[0]) + int(lab[1, 5, 5])
            dt = time.perf_counter() - t0
            print(json.dumps({'config': 'configs[3] end to end, 64 x 1280x720 streams from host memory, %s egress' % ('chunk' if sparse else 'dense'),
                              'steps': steps, 'fps': round(steps * n_streams / dt, 1), 'ms_per_step': round(dt / steps * 1e3, 3),
                              'h2d_bytes_per_step': n_streams * w * h * 3,
                              'd2h_bytes_per_step': int(seg2.egress_bytes / steps) if sparse else n_streams * (cw * ch_ * 4 + 4)}), flush=True)

    if 'stencil' in only:
        # configs[4]: blur s=15 (91 taps) -> resize 0.5 -> threshold -> 7x7 close, open -> label
        w, h, batch = 1920, 1080, 32
        vids = [synth.generate(rt, 2, k * batch, batch, w, h, 12) for k in range(2)]

        def run(i):
            b = rt.luma_gauss(vids[i % 2], 15.0)
            hlf = rt.resize_half(b)
            bits = rt.threshold(hlf, 95)
            bits = rt.morph(bits, 'close', 'rect', 7)
            bits = rt.morph(bits, 'open', 'rect', 7)
            rt.label(bits, 4)
        med, mn = timed(t, run, args.iters)
        n = w * h
        q = n / 4
        report('configs[4] 1080p stencil-heavy: blur s=15, resize 1/2, threshold, 7x7 close+open, label', batch, med, mn,
               (4 * n + 1.25 * n + 1.125 * q + 2 * q / 4 + 4.125 * q) * batch,
               'bound by the 91-tap blur (ALU), not HBM')

        def blur_only(i):
            rt.luma_gauss(vids[i % 2], 15.0)
        med, mn = timed(t, blur_only, args.iters)
        report('configs[4] blur s=15 alone (fused luma)', batch, med, mn, 4 * n * batch, '91 taps: 2 x 91 MACs per pixel')


if __name__ == '__main__':
    main()
