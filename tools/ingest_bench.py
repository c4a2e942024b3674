#!/usr/bin/env python
"""
Throughput of the raw-video ingest (video_analysis_b200/io/pipe.py): a file of packed rgb24 1080p frames on tmpfs,
read (a) directly and (b) through a pipe from a child process (`cat`, standing in for `ffmpeg -f image2pipe
-vcodec rawvideo -`), into the ring of (page-locked, when a GPU is there) frames; and (c) the same pipe feeding
`SegmentChain.process_blocks(max_regions=...)` on the GPU, region tables out.  One JSON line per case.

    python tools/ingest_bench.py [--frames 400] [--gpu]
"""

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

W, H = 1920, 1080


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=400)
    ap.add_argument('--gpu', action='store_true')
    args = ap.parse_args()
    from video_analysis_b200.io.pipe import VideoRawStream
    n = args.frames
    path = '/dev/shm/va_ingest_%d.rgb' % os.getpid()
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (8, H, W, 3), dtype=np.uint8)
    with open(path, 'wb') as f:
        for i in range(n):
            f.write(base[i % 8].tobytes())
    try:
        def drain(v):
            got, pos = 0, v.get_frame_pos()
            while True:
                b = v.frame_block(pos, pos + 64)
                if len(b) == 0:
                    return got
                got += len(b)
                pos += len(b)

        for name, factory, readers in (('file on tmpfs, one reader (file object)', lambda i: open(path, 'rb', buffering=0), 1),
                                       ('file on tmpfs by path, 2 readers (os.preadv)', lambda i: path, 2),
                                       ('file on tmpfs by path, 4 readers (os.preadv)', lambda i: path, 4),
                                       ('file on tmpfs by path, 8 readers (os.preadv)', lambda i: path, 8),
                                       ('pipe from a child process (cat)', lambda i: ['cat', path], 1)):
            v = VideoRawStream(factory, (W, H), n, ring_frames=384, pinned=args.gpu, readers=readers)
            drain(v)                                    # first pass touches the ring (page faults of a fresh allocation)
            t0 = time.perf_counter()
            v.set_frame_pos(0)                          # reopens the source
            got = drain(v)
            dt = time.perf_counter() - t0
            v.close()
            print(json.dumps({'case': 'VideoRawStream <- ' + name, 'frames': got, 'fps': round(got / dt, 1),
                              'GBps': round(got * W * H * 3 / dt / 1e9, 2), 'pinned_ring': bool(args.gpu)}), flush=True)
        if args.gpu:
            from video_analysis_b200.chain import SegmentChain
            ch = SegmentChain((W, H), batch=64)
            for name, src, readers in (('pipe', ['cat', path], 1), ('file by path, 8 readers', path, 8)):
                for warm in (True, False):
                    v = VideoRawStream(src, (W, H), n, ring_frames=384, readers=readers)
                    _, blocks = ch._blocks_of(v)
                    t0 = time.perf_counter()
                    regions = 0
                    for stats, counts, largest in ch.process_blocks(blocks, max_regions=256):
                        regions += int(counts.sum())
                    dt = time.perf_counter() - t0
                    v.close()
                    ch.reset()
                print(json.dumps({'case': name + ' -> ring -> SegmentChain.process_blocks(max_regions=256)', 'frames': n,
                                  'fps': round(n / dt, 1), 'regions': regions}), flush=True)
    finally:
        os.remove(path)


if __name__ == '__main__':
    main()
