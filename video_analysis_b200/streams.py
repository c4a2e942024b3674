"""
Concurrent camera streams batched per kernel launch (BASELINE.json configs[3]): S videos of one frame size
advance in lock step; every time step is ONE set of launches over the S current frames (leading axis =
stream), so the latency of a result is one frame, not one batch of frames of the same stream:

    crop (per-stream rectangle position, common size) -> monochrome -> apply-mask (per-stream static mask)
         -> threshold -> connected-component labels

It is the same arithmetic as stacking FilterCrop / FilterMonochrome / FilterApplyMask / FilterThreshold /
FilterLabel on every stream separately (reference: video/filters.py:158-248, :348-374; threshold / apply-mask
as adopted in SURVEY.md 8c; labels video/analysis/regions.py:162) and is tested against exactly that.

Host loop: the S current frames are gathered into a page-locked block by a few threads, and three ring slots keep
the upload of step t + 1, the kernels of step t and the egress of step t - 1 (label images as their non-empty
chunks, chain.SparseLabelEgress) in flight at once; results are handed out in order, up to `depth` - 1 steps after
their frames were read.
"""

import collections
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib
from .device import DeviceBatch, get_runtime, torch
from .filters import COLOR_CHANNELS, _check_coordinate


class MultiStreamSegmenter(object):
    """ videos: S colour videos of identical size; rects: one (left, top, width, height) per stream with the
    reference's coordinate rules -- all of the same width and height; masks: None, one (h, w) array for all
    streams or one per stream.  Iterating yields (labels (S, h, w) int32, counts (S,) int32) per time step
    until the first stream ends; the arrays are views of pinned buffers, valid until the next step. """

    def __init__(self, videos, rects, masks=None, threshold=110, mono_mode='mean', connectivity=4, device=None, fused=True,
                 depth=3, gather_threads=8, sparse_egress=True):
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
        self.depth, self.gather_threads, self.sparse_egress = max(1, int(depth)), max(1, int(gather_threads)), bool(sparse_egress)
        self._labels = self.rt.empty_i32(S, self.h, self.w)
        self._counts = t.empty((S,), dtype=t.int32, device=dev)
        self._slots = None

    def step_device(self, rgb, labels=None, counts=None, bits=None):
        """ one time step on frames already on the device: DeviceBatch (S, H, W, 3) -> (labels DeviceBatch, counts);
        `bits` (optional) receives the thresholded packed mask """
        rt = self.rt
        labels = self._labels if labels is None else labels
        counts = self._counts if counts is None else counts
        if self._fused:
            # crop + monochrome + mask + threshold in one pass (va_streams_threshold_bits)
            if bits is None:
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
        rt._check(rt.lib.va_label_bits(rt._h, rt.stream, *bits.img(), *labels.img(), counts.data_ptr(),
                                       self.w, self.h, rgb.n, self.connectivity))
        self._last_bits = bits
        return labels, counts

    def _make_slots(self):
        from .chain import SparseLabelEgress
        t, rt = torch(), self.rt
        S, H, W = self.S, self.H, self.W
        self._slots = []
        policy = SparseLabelEgress.Policy()
        for _ in range(self.depth):
            s = {'host_in': t.empty((S, H, W * 3), dtype=t.uint8, pin_memory=True),
                 'dev_in': t.empty((S, H, W * 3), dtype=t.uint8, device=rt.device),
                 'labels': rt.empty_i32(S, self.h, self.w),
                 'counts': t.empty((S,), dtype=t.int32, device=rt.device),
                 'host_counts': t.empty((S,), dtype=t.int32, pin_memory=True),
                 'ev_in': t.cuda.Event(), 'ev_run': t.cuda.Event(), 'ev_out': t.cuda.Event()}
            if self.sparse_egress:
                s['egress'] = SparseLabelEgress(rt, S, self.h, self.w, s['labels'].t.shape[2], policy=policy)
            else:
                s['host_labels'] = t.empty(tuple(s['labels'].t.shape), dtype=t.int32, pin_memory=True)
            self._slots.append(s)
        self._s_in, self._s_run, self._s_out = (t.cuda.Stream(device=rt.device) for _ in range(3))
        self._pool = ThreadPoolExecutor(max_workers=min(self.gather_threads, S))
        self.egress_bytes = 0

    def _gather(self, its, host):
        """ the next frame of every stream into the page-locked block `host` (S, H, W, 3); False when a stream has ended.
        NumPy releases the GIL while it copies, so the threads overlap """
        n = min(self.gather_threads, self.S)

        def part(k):
            for s in range(k, self.S, n):
                try:
                    np.copyto(host[s], next(its[s]))
                except StopIteration:
                    return False
            return True
        return all(list(self._pool.map(part, range(n))))

    def __iter__(self):
        t, rt = torch(), self.rt
        if self._slots is None:
            self._make_slots()
        its = [iter(v) for v in self.videos]
        pending = collections.deque()
        k = 0
        with t.cuda.device(rt.device):
            while True:
                s = self._slots[k % self.depth]
                if len(pending) == self.depth:                     # the slot still belongs to a step that was not handed out
                    yield self._finish(pending.popleft())
                s['ev_in'].synchronize()                           # its previous upload has left the host block
                if not self._gather(its, s['host_in'].numpy().reshape(self.S, self.H, self.W, 3)):
                    break
                with t.cuda.stream(self._s_in):
                    self._s_in.wait_event(s['ev_run'])             # the kernels that read this input are done
                    s['dev_in'].copy_(s['host_in'], non_blocking=True)
                    s['ev_in'].record(self._s_in)
                with t.cuda.stream(self._s_run):
                    self._s_run.wait_event(s['ev_in'])
                    self._s_run.wait_event(s['ev_out'])            # the previous egress of this slot is done
                    self.step_device(DeviceBatch('u8', s['dev_in'], self.S, self.H, self.W, 3), s['labels'], s['counts'])
                    if self.sparse_egress:
                        s['egress'].enqueue(self._last_bits, s['labels'], self.S)
                    s['ev_run'].record(self._s_run)
                with t.cuda.stream(self._s_out):
                    self._s_out.wait_event(s['ev_run'])
                    if self.sparse_egress:
                        s['egress'].enqueue_copy(s['labels'], self.S)
                    else:
                        s['host_labels'].copy_(s['labels'].t, non_blocking=True)
                    s['host_counts'].copy_(s['counts'], non_blocking=True)
                    s['ev_out'].record(self._s_out)
                pending.append(s)
                k += 1
            while pending:
                yield self._finish(pending.popleft())

    def _finish(self, s):
        s['ev_out'].synchronize()
        if self.sparse_egress:
            labels = s['egress'].finish(self.S)
            self.egress_bytes += s['egress'].bytes + 4 * self.S
        else:
            labels = s['host_labels'].numpy()[:, :, :self.w]
        return labels, s['host_counts'].numpy()
