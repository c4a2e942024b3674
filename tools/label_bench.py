"""
Labelling throughput on easy and hard masks, CUDA events, device-resident packed masks:
  blobs        what the chain's open produces on the synthetic video (about 95 % of the rows of words are empty)
  noise50      every pixel foreground with probability 1/2 (millions of tiny components per batch)
  checker      1-pixel checkerboard: N / 2 components, the maximum at 4-connectivity
  serpentine   one frame-spanning component that doubles back every 4 rows (long union-find chains)
  stress       serpentine in the upper half + a checkerboard tile + blobs (SURVEY's stress input)
Reports va_label_bits (forest + write), the two halves, and va_region_stats.
    python tools/label_bench.py [--w 1920 --h 1080 --batch 64]
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


def pack(mask):
    m = np.asarray(mask) != 0
    W = m.shape[-1]
    pad = (-W) % 32
    if pad:
        m = np.concatenate([m, np.zeros(m.shape[:-1] + (pad,), bool)], -1)
    return np.packbits(m, axis=-1, bitorder='little').view(np.uint32)


def timeit(fn, iters=10, warm=3):
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
    ap.add_argument('--peak', type=float, default=6538.9)
    a = ap.parse_args()
    W, H, B = a.w, a.h, a.batch
    N = W * H
    rt = get_runtime(0)
    rt.ensure(W, H, B)
    lib, h = rt.lib, rt._h
    masks = {}
    # blobs: the chain's own segment masks
    ch = SegmentChain((W, H), batch=B)
    mor = rt.empty_bits(B, H, W)
    for k in range(3):
        ch.run_device(synth.generate(rt, 0, k * B, B, W, H), morph=mor)
    masks['blobs'] = mor
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[:H, :W]
    one = {}
    one['noise50'] = rng.random((H, W)) < 0.5
    one['checker'] = ((yy + xx) & 1) == 1
    s = np.zeros((H, W), bool)
    s[::4] = True
    s[2::8, W - 1] = True
    s[1::8, W - 1] = True
    s[3::8, W - 1] = True
    s[5::8, 0] = True
    s[6::8, 0] = True
    s[7::8, 0] = True
    one['serpentine'] = s
    st = np.zeros((H, W), bool)
    st[:H // 2] = s[:H // 2]
    st[H // 2 + 8: H // 2 + 8 + H // 4, 16: 16 + W // 4] = one['checker'][:H // 4, :W // 4]
    for cx, cy, r in ((0.7, 0.75, 0.08), (0.5, 0.85, 0.05), (0.85, 0.6, 0.04)):
        st |= (xx - cx * W) ** 2 + (yy - cy * H) ** 2 <= (r * H) ** 2
    one['stress'] = st
    for name, m in one.items():
        words = pack(m)
        bits = rt.empty_bits(B, H, W)
        t = torch.from_numpy(words.view(np.int32)).to(rt.device)
        bits.t[:, :, :words.shape[1]] = t[None]
        if bits.t.shape[2] > words.shape[1]:
            bits.t[:, :, words.shape[1]:] = 0
        masks[name] = bits
    labels = rt.empty_i32(B, H, W)
    counts = torch.empty((B,), dtype=torch.int32, device=rt.device)
    MAXR = 4096
    stats = torch.empty((B, MAXR, 10), dtype=torch.int64, device=rt.device)
    largest = torch.empty((B,), dtype=torch.int32, device=rt.device)
    for name, m in masks.items():
        fg = float((rt.unpack_bits(m).t[:, :, :W] > 0).float().mean())

        def full():
            rt._check(lib.va_label_bits(h, rt.stream, *m.img(), *labels.img(), counts.data_ptr(), W, H, B, 4))

        def forest():
            rt._check(lib.va_label_forest(h, rt.stream, *m.img(), counts.data_ptr(), W, H, B, 4, 0))

        def write():
            rt._check(lib.va_label_write(h, rt.stream, *m.img(), *labels.img(), W, H, B, 0))

        def rstats():
            rt._check(lib.va_region_stats(h, rt.stream, *m.img(), stats.data_ptr(), MAXR, counts.data_ptr(), largest.data_ptr(), W, H, B, 4))
        t_full = timeit(full)
        t_forest = timeit(forest)
        t_write = timeit(write)
        t_stats = timeit(rstats)
        full()
        torch.cuda.synchronize()
        by = 4.125 * N * B
        print(json.dumps({'mask': name, 'size': '%dx%dx%d' % (W, H, B), 'foreground': round(fg, 4), 'regions_per_frame': int(counts[0]),
                          'label_ms': round(t_full, 4), 'forest_ms': round(t_forest, 4), 'write_ms': round(t_write, 4),
                          'region_stats_ms': round(t_stats, 4), 'alg_GBps': round(by / t_full / 1e6, 1),
                          'frac': round(by / t_full / 1e6 / a.peak, 3)}), flush=True)


if __name__ == '__main__':
    main()
