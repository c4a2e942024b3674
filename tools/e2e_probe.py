"""End-to-end probe of SegmentChain.process_blocks (1080p, pinned host frames in, dense int32 labels out on the host):
host threads of the chunk rebuild, ring depth, block size.   python tools/e2e_probe.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analysis_b200 import synth  # noqa: E402
from video_analysis_b200.chain import SegmentChain  # noqa: E402
from video_analysis_b200.device import get_runtime  # noqa: E402

W, H = 1920, 1080
rt = get_runtime(0)


def run(Be, depth, threads, steps=30, **kw):
    ring = 4
    host = torch.empty((ring * Be, H, W, 3), dtype=torch.uint8, pin_memory=True)
    for a in range(0, ring * Be, Be):
        host.view(ring * Be, H, W * 3)[a:a + Be].copy_(synth.generate(rt, 0, a, Be, W, H).t)
    torch.cuda.synchronize()
    hn = host.numpy()
    ch = SegmentChain((W, H), batch=Be, depth=depth, **kw)
    ch.host_threads = threads

    def blocks(n):
        for i in range(n):
            a = (i % ring) * Be
            yield hn[a:a + Be]
    sink = 0
    for lab, cnt in ch.process_blocks(blocks(5)):
        sink += int(cnt[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for lab, cnt in ch.process_blocks(blocks(steps)):
        sink += int(cnt[-1]) + int(lab[0, H // 2, W // 2])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({'frames_per_block': Be, 'depth': depth, 'host_threads': threads, 'kw': {k: str(v) for k, v in kw.items()},
                      'fps': round(steps * Be / dt, 1)}), flush=True)
    del ch, host


for Be, depth, threads in ((64, 3, 4), (64, 3, 8), (64, 4, 8), (32, 4, 8), (128, 3, 8), (64, 3, 2), (64, 3, 16)):
    run(Be, depth, threads)
run(64, 3, 4, sparse_egress=False)
