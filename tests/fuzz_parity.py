#!/usr/bin/env python
"""
Randomised differential test of the C ABI against the oracle (run on the GPU box; `-m gpu` runs a short version
through tests/test_filters_gpu.py).  Every case draws a frame size (ragged widths included), a batch, the chain
parameters (sigma, alpha, threshold, morphology op / element / size, connectivity, monochrome mode) and the input
statistics (blobs on noise, pure noise, sparse / dense salt, constant frames), runs `va_chain_run` fused and
unfused and compares every stage with the oracle bit for bit (background: float32 bit pattern).  Also draws
stand-alone cases for the resize modes, region statistics and the multi-stream front.

    python tests/fuzz_parity.py [--cases 300] [--seed 0] [--seconds 120]
"""

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_frames(rng, T, H, W, kind):
    if kind == 'blobs':
        from oracle import synth
        return synth.make_frames(int(rng.integers(0, 1 << 30)), 0, T, W, H, int(rng.integers(1, 9)))
    if kind == 'noise':
        return rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    if kind == 'salt':
        f = np.full((T, H, W, 3), int(rng.integers(0, 200)), np.uint8)
        m = rng.random((T, H, W)) < rng.choice([0.02, 0.3, 0.6])
        f[m] = 255
        return f
    f = np.empty((T, H, W, 3), np.uint8)
    f[:] = rng.integers(0, 256, (T, 1, 1, 1), dtype=np.uint8)      # constant frames, level changes over time
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', type=int, default=300)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--seconds', type=float, default=120.0)
    args = ap.parse_args()
    import cv2
    from oracle import ops
    from tests import harness as hz
    be = hz.CudaBackend()
    ctx = hz.Ctx(be, 1024, 512, 16)
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    done = {'chain': 0, 'resize': 0, 'regions': 0, 'streams': 0}
    fails = []
    for case in range(args.cases):
        if time.time() - t0 > args.seconds:
            break
        try:
            which = rng.choice(['chain', 'chain', 'chain', 'resize', 'regions', 'streams'])
            if which == 'chain':
                H = int(rng.integers(8, 200))
                W = int(rng.choice([rng.integers(8, 400), 16 * rng.integers(1, 25), 32 * rng.integers(1, 12)]))
                T = int(rng.integers(1, 9))
                kind = rng.choice(['blobs', 'noise', 'salt', 'const'])
                fr = make_frames(rng, T, H, W, kind)
                sigma = float(rng.choice([0.5, 0.8, 1.0, 1.5, 2.0, 2.5, 3.0, 4.2, 6.0]))
                alpha = float(rng.choice([0.05, 0.5, 0.013, 1.0]))
                thr = float(rng.choice([25, 0, 3.5, 100, 254.5]))
                op = [None, 'open', 'close', 'erode', 'dilate'][int(rng.integers(0, 5))]
                shape = rng.choice(['rect', 'cross', 'ellipse'])
                k = int(rng.choice([3, 5, 7, 2, 4, 9]))
                conn = int(rng.choice([4, 8]))
                ref = ops.chain(fr, sigma, alpha, thr, op, shape, k, conn)
                op = None if op is None else str(op)
                want = ('blur', 'mask', 'morph', 'labels') if op else ('blur', 'mask', 'labels')
                for fuse in (True, False):
                    got = hz.chain(ctx, fr, sigma=sigma, alpha=alpha, thr=thr, morph_op=op, shape=str(shape), k=k,
                                   connectivity=conn, fuse=fuse, want=want)
                    for key in want:
                        exp = ops.pack_bits(ref[key]) if key in ('mask', 'morph') else ref[key]
                        if not np.array_equal(got[key], exp):
                            raise AssertionError('stage %s differs (fuse=%s)' % (key, fuse))
                    if not np.array_equal(got['counts'], ref['counts']):
                        raise AssertionError('counts differ')
                    if not np.array_equal(got['bg'].view(np.uint32), ref['bg'].view(np.uint32)):
                        raise AssertionError('background differs')
                params = dict(H=H, W=W, T=T, kind=str(kind), sigma=sigma, alpha=alpha, thr=thr, op=op, shape=str(shape), k=k, conn=conn)
            elif which == 'resize':
                H, W = int(rng.integers(4, 160)), int(rng.integers(4, 200))
                dw, dh = int(rng.integers(1, 300)), int(rng.integers(1, 240))
                ch3 = bool(rng.integers(0, 2))
                fr = rng.integers(0, 256, (2, H, W, 3) if ch3 else (2, H, W), dtype=np.uint8)
                how = str(rng.choice(['area_any', 'linear', 'cubic', 'lanczos4']))
                interp = {'area_any': 'area', 'linear': 'linear', 'cubic': 'cubic', 'lanczos4': 'lanczos'}[how]
                if (dw, dh) == (W, H):
                    dw += 1
                cv2.ipp.setUseIPP(how != 'cubic')              # cubic: OpenCV's own arithmetic (the IPP routine is within 1 LSB)
                ref = np.stack([ops.resize(f, (dw, dh), interp) for f in fr]).reshape((2, dh, dw) + fr.shape[3:])
                cv2.ipp.setUseIPP(True)
                got = hz.resize_to(ctx, fr, dw, dh, how)
                if not np.array_equal(got, ref):
                    raise AssertionError('resize differs: max %d' % np.abs(got.astype(int) - ref).max())
                params = dict(H=H, W=W, dw=dw, dh=dh, ch3=ch3, how=how)
            elif which == 'regions':
                H, W = int(rng.integers(4, 150)), int(rng.integers(4, 300))
                m = ((rng.random((2, H, W)) < rng.choice([0.05, 0.4, 0.6, 0.9])) * 255).astype(np.uint8)
                conn = int(rng.choice([4, 8]))
                lab, cnt = hz.label(ctx, hz.pack_bits_np(m), W, conn)
                for t in range(2):
                    ref, n = ops.label(m[t], conn)
                    if cnt[t] != n or not np.array_equal(lab[t], ref):
                        raise AssertionError('labels differ')
                params = dict(H=H, W=W, conn=conn)
            else:
                H, W = int(rng.integers(20, 120)), 4 * int(rng.integers(10, 75))          # rows of whole words (3 W % 4 == 0)
                w, h = 32 * int(rng.integers(1, W // 32 + 1)), int(rng.integers(1, H + 1))
                S = int(rng.integers(1, 6))
                fr = rng.integers(0, 256, (S, H, W, 3), dtype=np.uint8)
                xy = np.stack([rng.integers(0, W - w + 1, S), rng.integers(0, H - h + 1, S)], axis=1)
                masks = ((rng.random((S, h, w)) < 0.7) * 255).astype(np.uint8)
                thr = int(rng.choice([110, 0, 200, -1, 255]))
                mode, name = [(-1, 'mean'), (0, 'blue'), (1, 'green'), (2, 'red')][int(rng.integers(0, 4))]
                mono = np.stack([ops.mono(ops.crop(fr[s], (int(xy[s, 0]), int(xy[s, 1]), w, h)), name) for s in range(S)])
                g = np.where(masks != 0, mono, 0)
                ref = hz.pack_bits_np(((g.astype(np.int32) > thr) * 255).astype(np.uint8))
                got = hz.streams_threshold(ctx, fr, xy, w, h, masks, thr, mode)
                if not np.array_equal(got, ref):
                    raise AssertionError('fused streams front differs')
                params = dict(H=H, W=W, w=w, h=h, S=S, thr=thr, mode=mode)
            done[which] += 1
        except Exception as err:                               # noqa: BLE001 -- every failure is reported with its parameters
            fails.append({'case': case, 'which': str(which), 'error': repr(err)[:300], 'params': locals().get('params')})
            if len(fails) >= 10:
                break
    ctx.close()
    out = {'seed': args.seed, 'cases_run': done, 'failures': fails, 'seconds': round(time.time() - t0, 1)}
    print(json.dumps(out))
    sys.exit(1 if fails else 0)


if __name__ == '__main__':
    main()
