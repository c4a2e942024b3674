"""
Concurrent camera streams batched per kernel launch (BASELINE.json configs[3]): S videos of one frame size
advance in lock step; every time step is ONE set of launches over the S current frames (leading axis =
stream), so the latency of a result is one frame, not one batch of frames of the same stream:

    crop (per-stream rectangle position, common size) -> monochrome -> apply-mask (per-stream static mask)
         -> threshold -> connected-component labels

It is the same arithmetic as stacking FilterCrop / FilterMonochrome / FilterApplyMask / FilterThreshold /
FilterLabel on every stream separately (reference: video/filters.py:158-248, :348-374; threshold / apply-mask
as adopted in SURVEY.md 8c; labels video/analysis/regions.py:162) and is tested against exactly that.
"""

import numpy as np

from . import _lib
from .device import DeviceBatch, get_runtime, torch
from .filters import COLOR_CHANNELS, _check_coordinate


class MultiStreamSegmenter(object):
    """ videos: S colour videos of identical size; rects: one (left, top, width, height) per stream with the
    reference's coordinate rules -- all of the same width and height; masks: None, one (h, w) array for all
    streams or one per stream.  Iterating yields (labels (S, h, w) int32, counts (S,) int32) per time step
    until the first stream ends; the arrays are views of pinned buffers, valid until the next step. """

    def __init__(self, videos, rects, masks=None, threshold=110, mono_mode='mean', connectivity=4, device=None, fused=True):
        self.videos = list(videos)
        S = len(self.videos)
        if S == 0 or len(rects) != S:
            raise ValueError('one crop rectangle per stream is required')
        size = tuple(self.videos[0].size)
        if any(tuple(v.size) != size or not v.is_color for v in self.videos):
            raise ValueError('all streams must be colour videos of the same size')
        W, H = size
        xy, wh = [], set()
        for r in rects:
            left, top = _check_coordinate(r[0], W), _check_coordinate(r[1], H)
            width, height = _check_coordinate(r[2], W), _check_coordinate(r[3], H)
            if left + width > W or top + height > H:
                raise IndexError('crop rectangle %s leaves the %dx%d frame' % ((left, top, width, height), W, H))
            xy.append((left, top))
            wh.add((width, height))
        if len(wh) != 1:
            raise ValueError('streams batched in one launch need crop rectangles of one size, got %s' % sorted(wh))
        (self.w, self.h), = wh
        self.W, self.H, self.S = W, H, S
        mode = COLOR_CHANNELS.get(mono_mode.lower(), mono_mode.lower()) if isinstance(mono_mode, str) else mono_mode
        if mode != 'mean' and mode not in (0, 1, 2):
            raise ValueError('Unsupported conversion method to monochrome: %s' % mono_mode)
        self.mode = _lib.MONO_MEAN if mode == 'mean' else mode
        if connectivity not in (4, 8):
            raise ValueError('connectivity must be 4 or 8')
        self.threshold, self.connectivity = int(threshold), connectivity
        t = torch()
        self.rt = get_runtime(device)
        self.rt.ensure(max(W, self.w), max(H, self.h), S)
        dev = self.rt.device
        self._xy = t.tensor(xy, dtype=t.int32).to(dev)
        self._masks = None
        if masks is not None:
            m = np.asarray(masks)
            if m.shape not in ((self.h, self.w), (S, self.h, self.w)):
                raise ValueError('masks must have shape (%d, %d) or (%d, %d, %d)' % (self.h, self.w, S, self.h, self.w))
            self._masks = t.from_numpy(np.ascontiguousarray((m != 0).astype(np.uint8))).to(dev)
        # the fused front needs whole mask words per lane pair and aligned rows (dense frames are; masks are padded here)
        self._fused = fused and self.w % 32 == 0 and (W * 3) % 4 == 0
        self._host_in = t.empty((S, H, W * 3), dtype=t.uint8, pin_memory=True)
        self._dev_in = t.empty((S, H, W * 3), dtype=t.uint8, device=dev)
        self._labels = self.rt.empty_i32(S, self.h, self.w)
        self._counts = t.empty((S,), dtype=t.int32, device=dev)
        self._host_labels = t.empty(tuple(self._labels.t.shape), dtype=t.int32, pin_memory=True)
        self._host_counts = t.empty((S,), dtype=t.int32, pin_memory=True)

    def step_device(self, rgb):
        """ one time step on frames already on the device: DeviceBatch (S, H, W, 3) -> (labels DeviceBatch, counts) """
        rt = self.rt
        if self._fused:
            # crop + monochrome + mask + threshold in one pass (va_streams_threshold_bits)
            bits = rt.empty_bits(rgb.n, self.h, self.w)
            m = self._masks
            margs = (0, 0, 0) if m is None else (m.data_ptr(), m.stride(-2), m.stride(0) if m.dim() == 3 else 0)
            rt._check(rt.lib.va_streams_threshold_bits(rt._h, rt.stream, rgb.ptr, rgb.pitch, rgb.fstride, rgb.w, rgb.h,
                                                       *margs, *bits.img(), self.w, self.h, rgb.n, self.mode,
                                                       self.threshold, self._xy.data_ptr()))
        else:
            g = rt.luma_crop_multi(rgb, self._xy, self.w, self.h, self.mode)
            if self._masks is not None:
                g = rt.apply_mask(g, self._masks)
            bits = rt.threshold(g, self.threshold)
        rt._check(rt.lib.va_label_bits(rt._h, rt.stream, *bits.img(), *self._labels.img(), self._counts.data_ptr(),
                                       self.w, self.h, rgb.n, self.connectivity))
        return self._labels, self._counts

    def __iter__(self):
        t = torch()
        host = self._host_in.numpy().reshape(self.S, self.H, self.W, 3)
        its = [iter(v) for v in self.videos]
        with t.cuda.device(self.rt.device):
            while True:
                try:
                    for s, it in enumerate(its):
                        np.copyto(host[s], next(it))
                except StopIteration:
                    return
                self._dev_in.copy_(self._host_in, non_blocking=True)
                self.step_device(DeviceBatch('u8', self._dev_in, self.S, self.H, self.W, 3))
                self._host_labels.copy_(self._labels.t, non_blocking=True)
                self._host_counts.copy_(self._counts, non_blocking=True)
                t.cuda.current_stream(self.rt.device).synchronize()
                yield self._host_labels.numpy()[:, :, :self.w], self._host_counts.numpy()
