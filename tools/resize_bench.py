#!/usr/bin/env python
"""
Throughput of the FilterResize kernels (1080p grey frames, batch 32, CUDA events, median of 10): only the 1/2 INTER_AREA
kernel is tuned; the others are one thread per output byte and are reported so that nobody has to guess.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch as t
    from video_analysis_b200 import synth
    from video_analysis_b200.device import get_runtime
    rt = get_runtime(0)
    W, H, B = 1920, 1080, 32
    rgb = [synth.generate(rt, 0, k * B, B, W, H, 8) for k in range(2)]
    grey = [rt.luma(r) for r in rgb]
    cases = [('area 1/2 (resize_half)', lambda g: rt.resize_half(g), (960, 540)),
             ('area 1/3 (integer factors)', lambda g: rt.resize_area(g, 3, 3), (640, 360)),
             ('area -> 1280x720 (float tables)', lambda g: rt.resize_area_any(g, 1280, 720), (1280, 720)),
             ('linear -> 1280x720', lambda g: rt.resize_linear(g, 1280, 720), (1280, 720)),
             ('nearest -> 1280x720', lambda g: rt.resize_nearest(g, 1280, 720), (1280, 720)),
             ('cubic -> 2880x1620 (enlarging)', lambda g: rt.resize_cubic(g, 2880, 1620), (2880, 1620)),
             ('lanczos4 -> 1280x720', lambda g: rt.resize_lanczos4(g, 1280, 720), (1280, 720))]
    for name, fn, (dw, dh) in cases:
        for i in range(3):
            fn(grey[i & 1])
        t.cuda.synchronize()
        ms = []
        for i in range(10):
            a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            a.record()
            fn(grey[i & 1])
            b.record()
            b.synchronize()
            ms.append(a.elapsed_time(b))
        med = float(np.median(ms))
        byt = (W * H + dw * dh) * B
        print(json.dumps({'resize': name, 'ms_per_32_frames': round(med, 4), 'fps': round(B / med * 1e3, 1),
                          'in_plus_out_GBps': round(byt / med / 1e6, 1)}), flush=True)


if __name__ == '__main__':
    main()
