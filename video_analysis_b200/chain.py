"""
SegmentChain -- the whole filter -> segment chain of BASELINE.json configs 1-3 as one
batched operator:

    monochrome -> Gaussian blur -> running-average background / |diff| > thr
               -> binary morphology -> connected-component labels

One `va_chain_run` call per batch enqueues every kernel back to back on one stream, the
intermediates never leave the GPU.  It is the same arithmetic as stacking the filter
classes of `video_analysis_b200.filters` (and is tested against them and the oracle); what
it adds is (a) the fused RGB -> luma -> blur kernel and (b) a three-stream pipeline that
overlaps the host->device copy of batch k+1, the kernels of batch k and the device->host
copy of batch k-1 when frames come from / results go to host memory.
"""

import collections

import numpy as np

from . import _lib
from .device import DeviceBatch, get_runtime, torch
from .filters import COLOR_CHANNELS


class SparseLabelEgress(object):
    """ label images of one ring slot on their way to the host as non-empty 64-label chunks (csrc/va_export.cu): the
    device stores chunk data, chunk ids and per-frame chunk counts straight into page-locked host memory, `finish`
    rebuilds the dense int32 images in ordinary host memory -- clearing only the chunks the slot's previous batch left
    behind -- and returns them.  Chunks whose foreground is a single run (one label) travel as 16-byte records.

    Masks that are mostly noise (a plain threshold on an unblurred frame) have foreground in most chunks; chunks then
    cost more than a dense copy.  `policy` (shared by the slots of one pipeline) watches the chunk fraction of the
    finished batches and switches the following ones to dense copies above 35 %, back to chunks below 20 %; in dense
    mode the device still counts the chunks so that the way back is seen. """

    class Policy(object):
        def __init__(self):
            self.dense = False

        def update(self, fraction):
            if fraction > 0.35:
                self.dense = True
            elif fraction < 0.20:
                self.dense = False

    def __init__(self, rt, batch, h, w, pitch_e, host_threads=4, policy=None):
        t = torch()
        self.rt, self.batch, self.h, self.w, self.pitch_e = rt, batch, h, w, pitch_e
        self.cap = -(-w // 64) * h                       # worst case: every chunk non-empty
        self.ids = t.empty((batch, self.cap), dtype=t.int32, pin_memory=True)
        self.data = t.empty((batch, self.cap, 64), dtype=t.int32, pin_memory=True)
        self.n = t.zeros((batch,), dtype=t.int32, pin_memory=True)
        self.runs = t.empty((batch, self.cap, 4), dtype=t.int32, pin_memory=True)
        self.n_runs = t.zeros((batch,), dtype=t.int32, pin_memory=True)
        self.dense = t.zeros((batch, h, pitch_e), dtype=t.int32)
        self.dirty_ids = t.zeros((batch, self.cap), dtype=t.int32)
        self.n_dirty = t.zeros((batch,), dtype=t.int32)
        self.host_threads = host_threads
        self.policy = policy if policy is not None else SparseLabelEgress.Policy()
        self.bytes = 0
        self._mode_dense = False
        self._pinned_dense = None

    def enqueue(self, seg, labels, m):
        """ on the current stream: export the first m frames (`seg`: packed mask the labels were made from) """
        rt = self.rt
        self._mode_dense = self.policy.dense
        if self._mode_dense:
            if self._pinned_dense is None:
                self._pinned_dense = torch().empty((self.batch, self.h, self.pitch_e), dtype=torch().int32, pin_memory=True)
            rt._check(rt.lib.va_label_export_chunks(rt._h, rt.stream, *seg.img(), *labels.img(), self.w, self.h, m,
                                                    None, None, self.n.data_ptr(), None, None, self.n_runs.data_ptr(), self.cap))
        else:
            rt._check(rt.lib.va_label_export_chunks(rt._h, rt.stream, *seg.img(), *labels.img(), self.w, self.h, m,
                                                    self.ids.data_ptr(), self.data.data_ptr(), self.n.data_ptr(), None,
                                                    self.runs.data_ptr(), self.n_runs.data_ptr(), self.cap))

    def enqueue_copy(self, labels, m):
        """ on the egress stream, after the kernels: the dense device -> host copy of a batch that travels densely """
        if self._mode_dense:
            self._pinned_dense[:m].copy_(labels.t[:m], non_blocking=True)

    def finish(self, m):
        """ after the export has completed (event): dense (m, h, w) int32 view, valid until the slot is reused """
        n_chunks = int(self.n.numpy()[:m].sum())
        n_runs = int(self.n_runs.numpy()[:m].sum())
        # what the chunks cost on PCIe relative to the dense image: 260 bytes per raw chunk, 16 per run chunk
        self.policy.update((n_chunks + n_runs / 16.0) / float(max(1, m * self.cap)))
        if self._mode_dense:
            self.bytes = m * self.h * self.pitch_e * 4 + 4 * m
            return self._pinned_dense.numpy()[:m, :, :self.w]
        d = self.dense
        rc = self.rt.lib.va_host_densify_chunks(d.data_ptr(), d.stride(1), d.stride(0), self.w, self.h, m,
                                                self.ids.data_ptr(), self.data.data_ptr(), self.n.data_ptr(),
                                                self.runs.data_ptr(), self.n_runs.data_ptr(), self.cap,
                                                self.dirty_ids.data_ptr(), self.n_dirty.data_ptr(), self.host_threads)
        _lib.check(self.rt.lib, None, rc)
        # bytes the device stored over PCIe for this block: chunk data + chunk ids, run records, the two count vectors
        self.bytes = n_chunks * (64 * 4 + 4) + n_runs * 16 + 8 * m
        return d.numpy()[:m, :, :self.w]


class SegmentChain(object):
    def __init__(self, size, sigma=2.0, alpha=0.05, threshold=25.0, morph_op='open', morph_shape='rect',
                 morph_ksize=3, connectivity=4, mono_mode='mean', batch=64, device=None, fuse=True, depth=3,
                 label_dtype=np.int32, sparse_egress=True):
        self.w, self.h = int(size[0]), int(size[1])
        self.sigma, self.alpha, self.threshold = float(sigma), float(alpha), float(threshold)
        if morph_op is not None and morph_op not in _lib.MORPH_OPS:
            raise ValueError('unknown morphological operation %r' % (morph_op,))
        if morph_shape not in _lib.SE_SHAPES:
            raise ValueError('unknown structuring element shape %r' % (morph_shape,))
        if connectivity not in (0, 4, 8):
            raise ValueError('connectivity must be 4 or 8 (0 disables labelling)')
        mode = COLOR_CHANNELS.get(mono_mode.lower(), mono_mode.lower()) if isinstance(mono_mode, str) else mono_mode
        if mode != 'mean' and mode not in (0, 1, 2):
            raise ValueError('Unsupported conversion method to monochrome: %s' % mono_mode)
        self.mono_mode = _lib.MONO_MEAN if mode == 'mean' else mode
        self.morph_op, self.morph_shape = morph_op, morph_shape
        self.kx, self.ky = (morph_ksize, morph_ksize) if np.isscalar(morph_ksize) else morph_ksize
        self.connectivity = connectivity
        self.batch, self.fuse, self.depth = int(batch), bool(fuse), int(depth)
        # label_dtype: np.int32 (what ndimage.label returns by default) or np.int16 (ndimage.label(..., output=np.int16)):
        # half the bytes for the device to write and for PCIe to carry; frames with more than 32767 regions raise
        self.label_dtype = np.dtype(label_dtype)
        if self.label_dtype not in (np.dtype(np.int32), np.dtype(np.int16)):
            raise ValueError('label_dtype must be int32 or int16')
        # sparse_egress: `process_blocks` brings int32 label images to the host as their non-empty 64-label chunks
        # (written by the device straight into page-locked memory) and rebuilds the dense arrays there -- the same
        # arrays, a fraction of the PCIe traffic; False copies the dense images
        self.sparse_egress = bool(sparse_egress)
        self.host_threads = 4
        self.egress_bytes = 0
        self._egress_policy = SparseLabelEgress.Policy()
        self.rt = get_runtime(device)
        self.rt.ensure(self.w, self.h, self.batch)
        self._bg = self.rt.empty_f32(self.h, self.w)
        self._started = False
        self._slots = None

    # ---- state ---------------------------------------------------------------------------------
    def reset(self):
        """ forget the background model: the next frame initialises it """
        self._started = False

    @property
    def background(self):
        torch().cuda.synchronize(self.rt.device)
        return self._bg[:, :self.w].cpu().numpy()

    def set_background(self, bg):
        """ start from a given model (H, W) float32 -- e.g. the carry of the preceding frame range """
        t = torch()
        if isinstance(bg, np.ndarray):
            bg = t.from_numpy(np.ascontiguousarray(bg, dtype=np.float32))
        self._bg[:, :self.w].copy_(bg.to(self.rt.device)[:, :self.w])
        self._started = True

    # ---- one batch, everything on the device ---------------------------------------------------------
    def run_device(self, rgb, labels=None, counts=None, blur=None, mask=None, morph=None):
        """ rgb: DeviceBatch (n, h, w, 3).  Enqueues the chain on the current stream and returns
        (labels DeviceBatch, counts tensor).  Optional DeviceBatches `blur`, `mask`, `morph`
        receive the intermediates. """
        rt, t = self.rt, torch()
        n = rgb.n
        if (rgb.w, rgb.h, rgb.channels) != (self.w, self.h, 3):
            raise ValueError('chain built for %dx%d colour frames, got %dx%dx%d' % (self.w, self.h, rgb.w, rgb.h, rgb.channels))
        if labels is None and self.connectivity:
            labels = self._empty_labels(n)
        if counts is None and self.connectivity:
            counts = t.empty((n,), dtype=t.int32, device=rt.device)
        if labels is not None and labels.kind == 'i16':
            # int16 labels: the chain as separate calls (va_chain_run writes int32 labels), forest + int16 write
            b = self.blur_device(rgb, blur)
            m = mask if mask is not None else rt.empty_bits(n, self.h, self.w)
            rt._check(rt.lib.va_ema_diff_thresh(rt._h, rt.stream, *b.img(), self._bg.data_ptr(), self._bg.stride(0),
                                                *m.img(), self.w, self.h, n, self.alpha, self.threshold,
                                                0 if self._started else 1))
            self._started = True
            seg = m
            if self.morph_op:
                seg = morph if morph is not None else rt.empty_bits(n, self.h, self.w)
                rt._check(rt.lib.va_morph_bits(rt._h, rt.stream, *m.img(), *seg.img(), self.w, self.h, n,
                                               _lib.MORPH_OPS[self.morph_op], _lib.SE_SHAPES[self.morph_shape],
                                               int(self.kx), int(self.ky)))
            rt._check(rt.lib.va_label_forest(rt._h, rt.stream, *seg.img(), counts.data_ptr(), self.w, self.h, n,
                                             self.connectivity, 0))
            rt._check(rt.lib.va_label_write_i16(rt._h, rt.stream, *seg.img(), *labels.img(), self.w, self.h, n, 0))
            return labels, counts
        d = _lib.ChainDesc(w=self.w, h=self.h, batch=n, mono_mode=self.mono_mode, sigma=self.sigma,
                           alpha=self.alpha, thr=self.threshold, first_frame_inits=0 if self._started else 1,
                           morph_op=_lib.MORPH_OPS[self.morph_op] if self.morph_op else -1,
                           morph_shape=_lib.SE_SHAPES[self.morph_shape], morph_kx=int(self.kx), morph_ky=int(self.ky),
                           connectivity=self.connectivity, fuse_luma_blur=1 if self.fuse else 0)
        io = _lib.ChainIO()
        io.rgb, io.rgb_pitch, io.rgb_fstride = rgb.img()
        io.bg, io.bg_pitch_e = self._bg.data_ptr(), self._bg.stride(0)
        if blur is not None:
            io.blur, io.blur_pitch, io.blur_fstride = blur.img()
        if mask is not None:
            io.mask, io.mask_pitch_w, io.mask_fstride_w = mask.img()
        if morph is not None:
            io.morph, io.morph_pitch_w, io.morph_fstride_w = morph.img()
        if labels is not None:
            io.labels, io.labels_pitch_e, io.labels_fstride_e = labels.img()
        if counts is not None:
            io.counts = counts.data_ptr()
        rt.chain_run(d, io, self.w, self.h, n)
        self._started = True
        return labels, counts

    def _empty_labels(self, n):
        return self.rt.empty_i16(n, self.h, self.w) if self.label_dtype == np.int16 else self.rt.empty_i32(n, self.h, self.w)

    # ---- the two halves, used when the background state arrives between them (parallel.py) ------
    def blur_device(self, rgb, out=None):
        """ monochrome + blur of a device batch -> DeviceBatch 'u8' """
        rt = self.rt
        if out is None:
            out = rt.empty_u8(rgb.n, self.h, self.w)
        rt.ensure(self.w, self.h, rgb.n)
        if self.fuse and self.sigma >= 0.5 and 6 * self.sigma + 1 <= 127:
            rt._check(rt.lib.va_luma_gauss_u8(rt._h, rt.stream, *rgb.img(), *out.img(), self.w, self.h, rgb.n,
                                              self.mono_mode, self.sigma))
        else:
            mono = rt.luma(rgb, self.mono_mode)
            rt._check(rt.lib.va_gauss_u8(rt._h, rt.stream, *mono.img(), *out.img(), self.w, self.h, 1, rgb.n, self.sigma))
        return out

    def segment_device(self, blur, labels=None, counts=None):
        """ background / threshold / morphology / labelling of an already blurred batch """
        rt, t = self.rt, torch()
        mask = rt.ema_diff_thresh(blur, self._bg, self.alpha, self.threshold, not self._started)
        self._started = True
        if self.morph_op:
            mask = rt.morph(mask, self.morph_op, self.morph_shape, (self.kx, self.ky))
        if not self.connectivity:
            return mask, None
        if labels is None:
            labels = rt.empty_i32(blur.n, self.h, self.w)
        if counts is None:
            counts = t.empty((blur.n,), dtype=t.int32, device=rt.device)
        rt._check(rt.lib.va_label_bits(rt._h, rt.stream, *mask.img(), *labels.img(), counts.data_ptr(),
                                       self.w, self.h, blur.n, self.connectivity))
        return labels, counts

    def regions_device(self, rgb, stats=None, counts=None, largest=None, max_regions=256):
        """ the chain with per-region statistics as its result instead of a label image (SURVEY 8f
        rank 1): -> (stats int64 (n, max_regions, 10), counts int32 (n,), largest int32 (n,)) device
        tensors; rows as documented for va_region_stats in include/va_b200.h """
        rt, t = self.rt, torch()
        n = rgb.n
        if (rgb.w, rgb.h, rgb.channels) != (self.w, self.h, 3):
            raise ValueError('chain built for %dx%d colour frames, got %dx%dx%d' % (self.w, self.h, rgb.w, rgb.h, rgb.channels))
        if not self.connectivity:
            raise ValueError('region statistics need a connectivity of 4 or 8')
        p = getattr(self, '_reg', None)
        if p is None:
            p = self._reg = {'blur': rt.empty_u8(self.batch, self.h, self.w), 'mask': rt.empty_bits(self.batch, self.h, self.w),
                             'morph': rt.empty_bits(self.batch, self.h, self.w)}
        sub = lambda b: DeviceBatch(b.kind, b.t[:n], n, b.h, b.w, b.channels)
        blur, mask, morph = sub(p['blur']), sub(p['mask']), sub(p['morph'])
        if stats is None:
            stats = t.empty((n, int(max_regions), 10), dtype=t.int64, device=rt.device)
        if counts is None:
            counts = t.empty((n,), dtype=t.int32, device=rt.device)
        if largest is None:
            largest = t.empty((n,), dtype=t.int32, device=rt.device)
        lib, h = rt.lib, rt._h
        self.blur_device(rgb, blur)
        rt._check(lib.va_ema_diff_thresh(h, rt.stream, *blur.img(), self._bg.data_ptr(), self._bg.stride(0),
                                         *mask.img(), self.w, self.h, n, self.alpha, self.threshold,
                                         0 if self._started else 1))
        self._started = True
        seg = mask
        if self.morph_op:
            rt._check(lib.va_morph_bits(h, rt.stream, *mask.img(), *morph.img(), self.w, self.h, n,
                                        _lib.MORPH_OPS[self.morph_op], _lib.SE_SHAPES[self.morph_shape],
                                        int(self.kx), int(self.ky)))
            seg = morph
        rt._check(lib.va_region_stats(h, rt.stream, *seg.img(), stats.data_ptr(), int(stats.shape[1]), counts.data_ptr(),
                                      largest.data_ptr(), self.w, self.h, n, self.connectivity))
        return stats, counts, largest

    # ---- three-stream software pipeline over consecutive device batches -----------------------------------
    def run_device_pipelined(self, rgb, labels, counts, blur=None):
        """ Same result as `run_device`, but the chain is split over three internal streams:
        front = RGB -> luma -> blur -> background/threshold (the sequential state lives here),
        back  = morphology -> union-find forest of the labelling (counts are complete after it),
        write = the label image.  The issue-bound front kernels of batch k+2, the latency-bound forest
        kernels of batch k+1 and the store-bound label write of batch k then share the SMs (the forest
        scratch is double-buffered in the ctx: `va_label_forest` / `va_label_write` with slot k & 1).
        Call `pipeline_sync()` (or synchronise the device) before reading `labels` / `counts`.
        `blur`: the already blurred batch (DeviceBatch 'u8'), when the caller has it. """
        rt, t = self.rt, torch()
        n = rgb.n
        p = self.pipeline_streams()
        k = p['k']
        p['k'] = k + 1
        slot = k & 1
        caller = t.cuda.current_stream(rt.device)
        ev = t.cuda.Event()
        ev.record(caller)                                   # inputs / output buffers are ready on the caller's stream
        rt.ensure(self.w, self.h, n)
        lib, h = rt.lib, rt._h
        sub = lambda b: DeviceBatch(b.kind, b.t[:n], n, b.h, b.w, b.channels)
        have_blur = blur is not None
        blur, mask, morph = (blur if have_blur else sub(p['blur'])), sub(p['mask'][slot]), sub(p['morph'][slot])
        with t.cuda.stream(p['front']):
            p['front'].wait_event(ev)
            p['front'].wait_event(p['ev_back'][slot])        # the back half has finished reading this mask slot
            p['front'].wait_event(p['ev_write'][slot])       # ... and so has the label write (it reads the mask when
            if not have_blur:                                #     there is no morphology in between)
                self.blur_device(rgb, blur)
            rt._check(lib.va_ema_diff_thresh(h, rt.stream, *blur.img(), self._bg.data_ptr(), self._bg.stride(0),
                                             *mask.img(), self.w, self.h, n, self.alpha, self.threshold,
                                             0 if self._started else 1))
            self._started = True
            p['ev_front'][slot].record(p['front'])
        seg = mask
        with t.cuda.stream(p['back']):
            p['back'].wait_event(ev)
            p['back'].wait_event(p['ev_front'][slot])
            p['back'].wait_event(p['ev_write'][slot])        # the write of batch k - 2 used this morph buffer / scratch slot
            if self.morph_op:
                rt._check(lib.va_morph_bits(h, rt.stream, *mask.img(), *morph.img(), self.w, self.h, n,
                                            _lib.MORPH_OPS[self.morph_op], _lib.SE_SHAPES[self.morph_shape],
                                            int(self.kx), int(self.ky)))
                seg = morph
            rt._check(lib.va_label_forest(h, rt.stream, *seg.img(), counts.data_ptr(), self.w, self.h, n,
                                          self.connectivity, slot))
            p['ev_back'][slot].record(p['back'])
        with t.cuda.stream(p['write']):
            p['write'].wait_event(ev)
            p['write'].wait_event(p['ev_back'][slot])
            write = lib.va_label_write_i16 if labels.kind == 'i16' else lib.va_label_write
            rt._check(write(h, rt.stream, *seg.img(), *labels.img(), self.w, self.h, n, slot))
            p['ev_write'][slot].record(p['write'])
        return labels, counts

    def pipeline_streams(self):
        """ streams, events and intermediate buffers of `run_device_pipelined` (made on first use) """
        if getattr(self, '_pipe', None) is None:
            rt, t = self.rt, torch()
            self._pipe = {
                'front': t.cuda.Stream(device=rt.device), 'back': t.cuda.Stream(device=rt.device),
                'write': t.cuda.Stream(device=rt.device), 'k': 0,
                'blur': rt.empty_u8(self.batch, self.h, self.w),
                'mask': [rt.empty_bits(self.batch, self.h, self.w) for _ in range(2)],
                'morph': [rt.empty_bits(self.batch, self.h, self.w) for _ in range(2)],
                'ev_front': [t.cuda.Event(), t.cuda.Event()], 'ev_back': [t.cuda.Event(), t.cuda.Event()],
                'ev_write': [t.cuda.Event(), t.cuda.Event()],
            }
        return self._pipe

    def pipeline_sync(self):
        """ make the caller's current stream wait for everything `run_device_pipelined` enqueued """
        p = getattr(self, '_pipe', None)
        if p is not None:
            cur = torch().cuda.current_stream(self.rt.device)
            cur.wait_stream(p['front'])
            cur.wait_stream(p['back'])
            cur.wait_stream(p['write'])

    # ---- host frames in, host labels out: pipelined -------------------------------------------------------
    def _make_slots(self):
        t, rt = torch(), self.rt
        slots = []
        for _ in range(self.depth):
            s = {
                'in': t.empty((self.batch, self.h, self.w * 3), dtype=t.uint8, device=rt.device),
                'counts': t.empty((self.batch,), dtype=t.int32, device=rt.device),
                'ev_in': t.cuda.Event(), 'ev_run': t.cuda.Event(), 'ev_out': t.cuda.Event(),
            }
            s['counts_host'] = t.empty((self.batch,), dtype=t.int32, pin_memory=True)
            slots.append(s)
        self._slots = slots
        self._s_in = t.cuda.Stream(device=rt.device)
        self._s_run = t.cuda.Stream(device=rt.device)
        self._s_out = t.cuda.Stream(device=rt.device)

    def _slot_outputs(self, s, max_regions):
        """ result buffers of a slot, made on first use: the label image (max_regions is None) or the
        region table """
        t, rt = torch(), self.rt
        if max_regions is None:
            if 'labels' not in s:
                s['labels'] = self._empty_labels(self.batch)
                if self._sparse():
                    s['egress'] = SparseLabelEgress(rt, self.batch, self.h, self.w, s['labels'].t.shape[2], self.host_threads,
                                                    self._egress_policy)
                    s['seg_mask'] = rt.empty_bits(self.batch, self.h, self.w)
                    s['seg_morph'] = rt.empty_bits(self.batch, self.h, self.w)
                else:
                    s['labels_host'] = t.empty(tuple(s['labels'].t.shape), dtype=s['labels'].t.dtype, pin_memory=True)
        elif s.get('stats') is None or s['stats'].shape[1] != max_regions:
            s['stats'] = t.empty((self.batch, max_regions, 10), dtype=t.int64, device=rt.device)
            s['largest'] = t.empty((self.batch,), dtype=t.int32, device=rt.device)
            s['stats_host'] = t.empty((self.batch, max_regions, 10), dtype=t.int64, pin_memory=True)
            s['largest_host'] = t.empty((self.batch,), dtype=t.int32, pin_memory=True)

    def _sparse(self):
        return self.sparse_egress and self.label_dtype == np.int32 and bool(self.connectivity)

    def process_blocks(self, blocks, max_regions=None):
        """ blocks: iterable of host arrays (m, h, w, 3) uint8 with m <= batch (page-locked memory
        gives asynchronous copies).  Yields (labels (m, h, w) int32, counts (m,) int32) per block,
        in order -- or, with `max_regions`, (stats (m, max_regions, 10) int64, counts, largest): the
        per-region table of `regions_device` instead of the label image, a few kilobytes per frame
        over PCIe instead of 4 bytes per pixel.  The yielded arrays are views of a ring of pinned
        buffers: they are valid until the next block is requested (copy them if you keep them, as
        `process` does). """
        t, rt = torch(), self.rt
        if max_regions is not None:
            max_regions = int(max_regions)
        if self._slots is None:
            self._make_slots()
        pending = collections.deque()
        with t.cuda.device(rt.device):
            for k, block in enumerate(blocks):
                block = np.ascontiguousarray(block)
                m = len(block)
                if m == 0:
                    continue
                if m > self.batch or block.shape[1:] != (self.h, self.w, 3) or block.dtype != np.uint8:
                    raise ValueError('expected blocks of up to %d uint8 frames of shape (%d, %d, 3)' % (self.batch, self.h, self.w))
                s = self._slots[k % self.depth]
                if len(pending) == self.depth:             # the slot is still owned by an unyielded block
                    yield self._finish(pending.popleft())
                self._slot_outputs(s, max_regions)
                with t.cuda.stream(self._s_in):
                    self._s_in.wait_event(s['ev_run'])     # previous kernels that read this input are done
                    s['in'][:m].copy_(t.from_numpy(block.reshape(m, self.h, self.w * 3)), non_blocking=True)
                    s['ev_in'].record(self._s_in)
                with t.cuda.stream(self._s_run):
                    self._s_run.wait_event(s['ev_in'])
                    self._s_run.wait_event(s['ev_out'])    # previous download of this slot's labels is done
                    rgb = DeviceBatch('u8', s['in'][:m], m, self.h, self.w, 3)
                    if max_regions is None:
                        lab = DeviceBatch(s['labels'].kind, s['labels'].t[:m], m, self.h, self.w)
                        if self._sparse():
                            sub = lambda b: DeviceBatch(b.kind, b.t[:m], m, b.h, b.w, b.channels)
                            mask, morph = sub(s['seg_mask']), sub(s['seg_morph'])
                            self.run_device(rgb, lab, s['counts'][:m], mask=mask, morph=morph if self.morph_op else None)
                            seg = morph if self.morph_op else mask
                            s['egress'].enqueue(seg, lab, m)
                        else:
                            self.run_device(rgb, lab, s['counts'][:m])
                    else:
                        self.regions_device(rgb, s['stats'][:m], s['counts'][:m], s['largest'][:m])
                    s['ev_run'].record(self._s_run)
                with t.cuda.stream(self._s_out):
                    self._s_out.wait_event(s['ev_run'])
                    if max_regions is None:
                        if self._sparse():
                            s['egress'].enqueue_copy(s['labels'], m)
                        else:
                            s['labels_host'][:m].copy_(s['labels'].t[:m], non_blocking=True)
                    else:
                        s['stats_host'][:m].copy_(s['stats'][:m], non_blocking=True)
                        s['largest_host'][:m].copy_(s['largest'][:m], non_blocking=True)
                    s['counts_host'][:m].copy_(s['counts'][:m], non_blocking=True)
                    s['ev_out'].record(self._s_out)
                pending.append(self._finish_async((s, m, max_regions)))
            while pending:
                yield self._finish(pending.popleft())

    def _finish_async(self, item):
        """ hand a block whose work is enqueued to the finisher thread: it waits for the block's event and rebuilds the
        dense label images (host threads, GIL released) while the caller enqueues the next blocks """
        if not self._sparse() or item[2] is not None:
            return item
        import concurrent.futures
        if getattr(self, '_finisher', None) is None:
            self._finisher = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix='va-egress')
        return self._finisher.submit(self._finish_now, item)

    def _finish(self, item):
        if hasattr(item, 'result'):
            return item.result()
        return self._finish_now(item)

    def _finish_now(self, item):
        s, m, max_regions = item
        s['ev_out'].synchronize()
        if max_regions is None:
            if self.label_dtype == np.int16 and m and int(s['counts_host'].numpy()[:m].max()) > 32767:
                raise RuntimeError('insufficient bit-depth in requested output type')      # what ndimage.label raises
            if self._sparse():
                dense = s['egress'].finish(m)
                self.egress_bytes += s['egress'].bytes + 4 * m
                return dense, s['counts_host'].numpy()[:m]
            return s['labels_host'].numpy()[:m, :, :self.w], s['counts_host'].numpy()[:m]
        return s['stats_host'].numpy()[:m], s['counts_host'].numpy()[:m], s['largest_host'].numpy()[:m]

    def _blocks_of(self, frames):
        """ (frame count, generator of (m, h, w, 3) blocks of at most `batch` frames) for an ndarray or a video
        object with `frame_block` (VideoMemory, VideoRawStream: blocks may come shorter than asked for, and a
        stream's `frame_count` may be an estimate) """
        if hasattr(frames, 'frame_block'):
            video = frames
            n = video.frame_count

            def gen():
                a = 0
                while a < n:
                    block = video.frame_block(a, min(a + self.batch, n))
                    if len(block) == 0:
                        return
                    yield block
                    a += len(block)
            return n, gen()
        frames = np.asarray(frames)
        return len(frames), (frames[a:a + self.batch] for a in range(0, len(frames), self.batch))

    def process(self, frames):
        """ frames: ndarray (n, h, w, 3) uint8 or a video object -> (labels (n, h, w) int32,
        counts (n,) int32), continuing the background model from earlier calls """
        n, blocks = self._blocks_of(frames)
        t = torch()
        labels_t = t.empty((n, self.h, self.w), dtype=t.int16 if self.label_dtype == np.int16 else t.int32)   # filled by torch's multi-threaded host copy
        labels = labels_t.numpy()
        counts = np.empty((n,), np.int32)
        k = 0
        for lab, cnt in self.process_blocks(blocks):
            labels_t[k:k + len(lab)].copy_(t.from_numpy(lab))
            counts[k:k + len(cnt)] = cnt
            k += len(lab)
        return labels[:k], counts[:k]

    def annotate(self, frames, writer=None, channel='all', strength=128):
        """ annotated output (the reference's VideoComposer.highlight_mask, io/composer.py:131-154, fed from the
        device): runs monochrome -> blur -> background mask -> morphology on every block and highlights the
        resulting mask in the colour frames on the GPU; the mask never visits the host.  The annotated frames are
        written to `writer` (anything with `write_block`, e.g. io.pipe.RawStreamWriter on an encoder's stdin) or,
        without a writer, returned as one (n, h, w, 3) array. """
        rt, t = self.rt, torch()
        _, blocks = self._blocks_of(frames)
        out = []
        with t.cuda.device(rt.device):
            for block in blocks:
                rgb = rt.upload(block)
                blur = self.blur_device(rgb)
                mask = rt.ema_diff_thresh(blur, self._bg, self.alpha, self.threshold, not self._started)
                self._started = True
                if self.morph_op:
                    mask = rt.morph(mask, self.morph_op, self.morph_shape, (self.kx, self.ky))
                marked = rt.highlight_mask(rgb, mask, channel, strength)
                host = rt.download(marked)
                t.cuda.current_stream(rt.device).synchronize()
                view = rt.host_view(marked, host)
                if writer is not None:
                    writer.write_block(view)
                else:
                    out.append(np.array(view))
        if writer is not None:
            return writer
        return np.concatenate(out) if out else np.empty((0, self.h, self.w, 3), np.uint8)

    def process_regions(self, frames, max_regions=256):
        """ like `process`, but returns for every frame the list of its regions (`label`, `area`,
        `bbox`, `moments`: see analysis.regions.stats_to_regions) instead of a label image """
        from .analysis.regions import stats_to_regions
        _, blocks = self._blocks_of(frames)
        out = []
        for stats, counts, _ in self.process_blocks(blocks, max_regions=max_regions):
            for st, c in zip(stats, counts):
                if c > max_regions:
                    raise MemoryError('a frame has %d regions, max_regions is %d' % (c, max_regions))
                out.append(stats_to_regions(st, int(c)))
        return out
