"""
ORACLE (test infrastructure) -- seeded synthetic video generator, NumPy version.

The reference's own synthetic source (``VideoGaussianNoise``,
video/io/computed.py:15-41) draws unseeded ``np.random.randn`` noise and so
cannot be reproduced on a device.  SURVEY.md section 8d replaces it by an
integer hash of ``(seed, t, y, x, c)`` so that NumPy (here) and CUDA
(``va_synth_rgb`` in the product) produce identical bytes:

    frame[t,y,x,c] = clip(base(y,x,c) + noise(seed,t,y,x,c) + 90*inside_disc(t,y,x), 0, 255)

``base`` in [60,120) is a static texture, ``noise`` in [-8,8] changes per
frame, and ``n_blobs`` discs of radius 0.03..0.06*H move linearly with
wrap-around of their centres.  Everything is uint32 / int32 arithmetic.
"""

import numpy as np

GOLD = 0x9E3779B9
_M32 = 0xFFFFFFFF


def mix32(x):
    """lowbias32 integer finaliser on uint32 arrays (wraps mod 2**32)."""
    x = np.asarray(x, dtype=np.uint32).copy()
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _mix32_int(x):
    """same finaliser on a Python int"""
    x &= _M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & _M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & _M32
    x ^= x >> 16
    return x


def mulhi(h, n):
    """range reduction (h * n) >> 32 -> [0, n)"""
    return ((np.asarray(h, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int32)


def blob_table(seed, width, height, n_blobs):
    """int32 [n_blobs, 5] = (x0, y0, vx16, vy16, r); velocities in 1/16 px per frame"""
    tab = np.zeros((n_blobs, 5), dtype=np.int32)
    rmin = (height * 3) // 100
    for b in range(n_blobs):
        h = [_mix32_int(seed * 0x85EBCA6B + b * 0xC2B2AE35 + k + 1) for k in range(5)]
        tab[b, 0] = (h[0] * width) >> 32
        tab[b, 1] = (h[1] * height) >> 32
        tab[b, 2] = ((h[2] * 97) >> 32) - 48
        tab[b, 3] = ((h[3] * 97) >> 32) - 48
        tab[b, 4] = rmin + ((h[4] * (rmin + 1)) >> 32)
    return tab


def make_frames(seed, t0, n_frames, width, height, n_blobs=8):
    """uint8 [n_frames, height, width, 3] -- frames t0 .. t0+n_frames-1 of the video `seed`"""
    W, H = int(width), int(height)
    yy, xx, cc = np.meshgrid(np.arange(H, dtype=np.uint32), np.arange(W, dtype=np.uint32),
                             np.arange(3, dtype=np.uint32), indexing='ij')
    idx = (yy * np.uint32(W) + xx) * np.uint32(3) + cc
    kb = np.uint32(_mix32_int(seed * GOLD + 0x01234567))
    base = 60 + mulhi(mix32(idx ^ kb), 60)

    tab = blob_table(seed, W, H, n_blobs)
    ys = np.arange(H, dtype=np.int64)[:, None]
    xs = np.arange(W, dtype=np.int64)[None, :]

    out = np.empty((n_frames, H, W, 3), dtype=np.uint8)
    for i in range(n_frames):
        t = t0 + i
        kt = np.uint32(_mix32_int(seed ^ (((t + 1) * GOLD) & _M32)))
        noise = mulhi(mix32(idx + kt), 17) - 8
        inside = np.zeros((H, W), dtype=bool)
        for b in range(n_blobs):
            x0, y0, vx, vy, r = (int(v) for v in tab[b])
            cx = ((x0 * 16 + vx * t) % (16 * W)) >> 4
            cy = ((y0 * 16 + vy * t) % (16 * H)) >> 4
            inside |= ((xs - cx) ** 2 + (ys - cy) ** 2) <= r * r
        val = base + noise + 90 * inside[:, :, None].astype(np.int32)
        out[i] = np.clip(val, 0, 255).astype(np.uint8)
    return out
