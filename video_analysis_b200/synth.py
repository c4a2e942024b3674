"""
Seeded synthetic video generated on the GPU.  Stands in for the reference's
unseeded `VideoGaussianNoise` (video/io/computed.py:15-41): every byte is an integer hash of
(seed, t, y, x, c), so a frame can be regenerated anywhere, on any rank.

    frame[t,y,x,c] = clip(base(y,x,c) + noise(seed,t,y,x,c) + 90 * inside_disc(t,y,x), 0, 255)
"""

import numpy as np

from .device import DeviceBatch, get_runtime, torch
from .io.base import VideoBase

_M32 = 0xFFFFFFFF


def _mix32(x):
    x &= _M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & _M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & _M32
    x ^= x >> 16
    return x


def blob_table(seed, width, height, n_blobs):
    """ (x0, y0, vx16, vy16, r) per moving disc; velocities in 1/16 pixel per frame """
    rmin = (height * 3) // 100
    rows = []
    for b in range(n_blobs):
        h = [_mix32(seed * 0x85EBCA6B + b * 0xC2B2AE35 + k + 1) for k in range(5)]
        rows.append(((h[0] * width) >> 32, (h[1] * height) >> 32,
                     ((h[2] * 97) >> 32) - 48, ((h[3] * 97) >> 32) - 48,
                     rmin + ((h[4] * (rmin + 1)) >> 32)))
    return np.array(rows, dtype=np.int32).reshape(-1, 5)


def generate(rt, seed, t0, n, width, height, n_blobs=8, out=None):
    """ frames t0 .. t0+n-1 as a dense DeviceBatch (n, height, width, 3) """
    t = torch()
    if out is None:
        out = t.empty((n, height, width * 3), dtype=t.uint8, device=rt.device)
    batch = DeviceBatch('u8', out, n, height, width, 3)
    rt.synth_rgb(batch, t0, seed, blob_table(seed, width, height, n_blobs))
    return batch


class VideoSynthetic(VideoBase):
    """ seekable colour video whose frames are generated on the device on demand """

    seekable = True

    def __init__(self, size=(640, 480), frame_count=1000, seed=0, n_blobs=8, fps=25, device=None):
        super(VideoSynthetic, self).__init__(size=size, frame_count=frame_count, fps=fps, is_color=True)
        self.seed, self.n_blobs, self._device = seed, n_blobs, device

    def frame_block(self, start, stop):
        rt = get_runtime(self._device)
        t = torch()
        n = max(0, min(stop, self.frame_count) - start)
        if n == 0:
            return np.empty((0, self.size[1], self.size[0], 3), np.uint8)
        with t.cuda.device(rt.device):
            batch = generate(rt, self.seed, start, n, self.size[0], self.size[1], self.n_blobs)
            host = batch.t.cpu()
        return host.numpy().reshape(n, self.size[1], self.size[0], 3)

    def get_frame(self, index):
        if index < 0:
            index += self.frame_count
        if not 0 <= index < self.frame_count:
            raise IndexError('Cannot access frame %d.' % index)
        return self.frame_block(index, index + 1)[0]
