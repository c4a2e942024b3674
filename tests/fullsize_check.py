#!/usr/bin/env python
"""
BASELINE.json configs[1] at its full size -- 1920x1080 RGB, 10 000 synthetic frames resident in HBM -- checked
through size-independent properties (the oracle cannot process 10 000 1080p frames in reasonable time):

  1. the fused, three-stream pipelined chain (what bench.py times) and the unfused single-stream chain
     (one kernel per stage, `va_chain_run` with fuse_luma_blur = 0) give identical label images, counts and
     background state over the whole video -- compared through 64-bit checksums per batch, computed on the device;
  2. labels are 0 exactly where the opened mask is 0 and 1..n elsewhere (max label == count) in every frame;
  3. the first 4 and the last 4 frames of the video are compared pixel by pixel with the oracle run on the host
     from the device's own background state (the recurrence makes the last frames depend on all 10 000).

    python tests/fullsize_check.py [--frames 10000] [--batch 128]
"""

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=10000)
    ap.add_argument('--batch', type=int, default=128)
    args = ap.parse_args()
    import torch as t
    from oracle import ops                                      # checker
    from video_analysis_b200 import synth
    from video_analysis_b200.chain import SegmentChain
    from video_analysis_b200.device import DeviceBatch, get_runtime
    W, H, B = 1920, 1080, args.batch
    n = args.frames // B * B
    rt = get_runtime(0)
    video = t.empty((n, H, W * 3), dtype=t.uint8, device=rt.device)
    for a in range(0, n, B):
        synth.generate(rt, 0, a, B, W, H, out=video[a:a + B])
    t.cuda.synchronize()
    batch_of = lambda k: DeviceBatch('u8', video[k * B:(k + 1) * B], B, H, W, 3)
    weights = t.arange(1, W + 1, device=rt.device, dtype=t.int64)

    def checksum(lab, cnt):
        x = lab.t[:, :, :W].to(t.int64)
        return int((x * weights).sum().item()), int(x.max().item()), int(cnt.to(t.int64).sum().item()), \
            bool((x.amax(dim=(1, 2)) == cnt.to(t.int64)).all().item())

    t0 = time.time()
    results = {}
    for name, fuse, piped in (('fused, three streams', True, True), ('unfused, one stream', False, False)):
        ch = SegmentChain((W, H), batch=B, fuse=fuse)
        labels = [rt.empty_i32(B, H, W) for _ in range(2)]
        counts = [t.empty((B,), dtype=t.int32, device=rt.device) for _ in range(2)]
        morph = rt.empty_bits(B, H, W)
        sums, ok_max, ok_zero = [], True, True
        prev = None
        for k in range(n // B):
            if piped:
                ch.run_device_pipelined(batch_of(k), labels[k & 1], counts[k & 1])
                if prev is not None:                            # check batch k - 1 while batch k is in flight
                    pass
                ch.pipeline_sync()
            else:
                ch.run_device(batch_of(k), labels[k & 1], counts[k & 1], morph=morph)
            t.cuda.synchronize()
            s = checksum(labels[k & 1], counts[k & 1])
            ok_max &= s[3]
            if not piped:                                       # labels are zero exactly where the opened mask is zero
                lab = labels[k & 1].t[:, :, :W]
                bits = morph.t[:, :, :(W + 31) // 32]
                sh = t.arange(32, device=rt.device, dtype=t.int32)
                fg = ((bits.unsqueeze(-1) >> sh) & 1).reshape(B, H, -1)[:, :, :W].bool()
                ok_zero &= bool(((lab != 0) == fg).all().item())
            sums.append(s[:3])
            prev = k
        results[name] = {'sums': sums, 'bg': ch.background.copy(), 'max_is_count': ok_max, 'zero_iff_background': ok_zero,
                         'chain': ch, 'labels': labels, 'counts': counts}
    a, b = results['fused, three streams'], results['unfused, one stream']
    same_sums = a['sums'] == b['sums']
    same_bg = np.array_equal(a['bg'].view(np.uint32), b['bg'].view(np.uint32))

    # ---- head and tail against the oracle ---------------------------------------------------------------
    def host_frames(lo, hi):
        return video[lo:hi].cpu().numpy().reshape(hi - lo, H, W, 3)
    head = ops.chain(host_frames(0, 4))
    ch = SegmentChain((W, H), batch=4)
    lab, cnt = ch.run_device(DeviceBatch('u8', video[0:4], 4, H, W, 3))
    t.cuda.synchronize()
    head_ok = np.array_equal(lab.t[:, :, :W].cpu().numpy(), head['labels']) and list(cnt.cpu().numpy()) == list(head['counts'])
    # tail: background state after frame n - 5 from the device (unfused run up to there), then the oracle on the last 4
    ch = SegmentChain((W, H), batch=B, fuse=False)
    for k in range(n // B - 1):
        ch.run_device(batch_of(k))
    rest = n - B                                                 # frames [rest, n - 4) one more partial batch
    ch.run_device(DeviceBatch('u8', video[rest:n - 4], B - 4, H, W, 3))
    bg_dev = ch.background.copy()
    tail = ops.chain(host_frames(n - 4, n), bg0=bg_dev)
    lab, cnt = ch.run_device(DeviceBatch('u8', video[n - 4:n], 4, H, W, 3))
    t.cuda.synchronize()
    tail_ok = np.array_equal(lab.t[:, :, :W].cpu().numpy(), tail['labels']) and list(cnt.cpu().numpy()) == list(tail['counts'])
    tail_bg_ok = np.array_equal(ch.background.view(np.uint32), tail['bg'].view(np.uint32))
    tail_same_as_full = np.array_equal(ch.background.view(np.uint32), b['bg'].view(np.uint32))

    out = {'frames': n, 'batch': B, 'regions_total': a['sums'] and sum(s[2] for s in a['sums']),
           'fused_pipelined_equals_unfused_sequential_checksums': same_sums, 'background_bit_identical': same_bg,
           'max_label_equals_count_every_frame': a['max_is_count'] and b['max_is_count'],
           'labels_zero_iff_opened_mask_zero': b['zero_iff_background'],
           'first_4_frames_equal_oracle': head_ok, 'last_4_frames_equal_oracle': tail_ok,
           'background_after_last_frame_equals_oracle': tail_bg_ok,
           'background_independent_of_batching': tail_same_as_full,
           'seconds': round(time.time() - t0, 1)}
    print(json.dumps(out))
    ok = all(v for k, v in out.items() if isinstance(v, bool))
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
