"""
Device runtime: one `va_ctx` per (process, GPU) plus typed wrappers that enqueue the
kernels of libva_b200 on torch's current CUDA stream.

PyTorch is used for plumbing only -- device / pinned allocations, streams, events.
Every pixel is touched by the hand-written kernels behind include/va_b200.h.  There
is no CPU path: without the built library or without a GPU these calls raise.
"""

import ctypes
import threading

import numpy as np

from . import _lib

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def _round_up(v, m):
    return (v + m - 1) // m * m


REGION_FIELDS = 10          # VA_REGION_FIELDS of include/va_b200.h: m00 m10 m01 m20 m11 m02 xmin ymin xmax ymax


class DeviceBatch(object):
    """ a batch of images resident on the GPU.

    kind 'u8'  : uint8 tensor (n, h, pitch), pitch >= w * channels bytes
    kind 'bits': int32 tensor (n, h, pitch) of packed mask words, LSB = lowest x
    kind 'i32' : int32 tensor (n, h, pitch) label image
    """
    __slots__ = ('kind', 't', 'n', 'h', 'w', 'channels', 'extra')

    def __init__(self, kind, t, n, h, w, channels=1, extra=None):
        self.kind, self.t, self.n, self.h, self.w, self.channels = kind, t, n, h, w, channels
        self.extra = extra or {}

    @property
    def pitch(self):
        return self.t.stride(1)

    @property
    def fstride(self):
        return self.t.stride(0)

    @property
    def ptr(self):
        return self.t.data_ptr()

    def img(self):
        """ (pointer, pitch, frame stride) in the units the C ABI expects for this kind """
        return self.ptr, self.pitch, self.fstride


class DeviceRuntime(object):
    """ owns the va_ctx of one GPU; grows its scratch capacity on demand """

    def __init__(self, device=0):
        t = torch()
        if not t.cuda.is_available():
            raise _lib.VAError('no CUDA device available: video_analysis_b200 has no CPU fallback')
        self.lib = _lib.load()
        self.device = t.device('cuda', device if isinstance(device, int) else t.device(device).index or 0)
        self._h = ctypes.c_void_p()
        self._cap = (0, 0, 0)
        self._retired_launches = 0

    # ---- ctx management ------------------------------------------------------------------
    def ensure(self, w, h, n):
        """ make sure the ctx can take frames of w x h in batches of n.  The ctx of a GPU is shared by every chain of the
        process: growing it keeps the handle (va_reserve synchronises the device and reallocates the scratch), so other
        users are not disturbed """
        cw, ch, cn = self._cap
        if w <= cw and h <= ch and n <= cn and self._h:
            return
        cap = (max(w, cw), max(h, ch), max(n, cn))
        with torch().cuda.device(self.device):
            if self._h:
                rc = self.lib.va_reserve(self._h, cap[0], cap[1], cap[2])
                if rc != _lib.VA_OK:
                    _lib.check(self.lib, self._h, rc)
            else:
                rc = self.lib.va_create(ctypes.byref(self._h), self.device.index, cap[0], cap[1], cap[2])
                if rc != _lib.VA_OK:
                    self._cap = (0, 0, 0)
                    _lib.check(self.lib, None, rc)
        self._cap = cap

    @property
    def launches(self):
        """ kernels launched through this runtime so far """
        return self._retired_launches + (self.lib.va_launch_count(self._h) if self._h else 0)

    def close(self):
        if self._h:
            torch().cuda.synchronize(self.device)
            self._retired_launches += self.lib.va_launch_count(self._h)
            self.lib.va_destroy(self._h)
            self._h = ctypes.c_void_p()
            self._cap = (0, 0, 0)

    @property
    def stream(self):
        return torch().cuda.current_stream(self.device).cuda_stream

    def _check(self, rc):
        _lib.check(self.lib, self._h, rc)

    # ---- allocation ------------------------------------------------------------------------
    def empty_u8(self, n, h, w, channels=1):
        pitch = _round_up(w * channels, 16)
        t = torch().empty((n, h, pitch), dtype=torch().uint8, device=self.device)
        return DeviceBatch('u8', t, n, h, w, channels)

    def empty_bits(self, n, h, w):
        pitch = _round_up((w + 31) // 32, 4)
        t = torch().empty((n, h, pitch), dtype=torch().int32, device=self.device)
        return DeviceBatch('bits', t, n, h, w)

    def empty_i32(self, n, h, w):
        pitch = _round_up(w, 4)
        t = torch().empty((n, h, pitch), dtype=torch().int32, device=self.device)
        return DeviceBatch('i32', t, n, h, w)

    def empty_i16(self, n, h, w):
        pitch = _round_up(w, 8)
        t = torch().empty((n, h, pitch), dtype=torch().int16, device=self.device)
        return DeviceBatch('i16', t, n, h, w)

    def empty_f32(self, h, w):
        return torch().empty((h, _round_up(w, 4)), dtype=torch().float32, device=self.device)

    # ---- host <-> device -----------------------------------------------------------------------
    def upload(self, frames):
        """ frames: ndarray (n, h, w[, 3]) uint8, C-contiguous -> dense DeviceBatch.
        Page-locked arrays (VideoMemory.pin(), pinned torch buffers) go by async DMA. """
        t = torch()
        frames = np.ascontiguousarray(frames)
        if frames.dtype != np.uint8:
            raise ValueError('the device path handles uint8 frames, got %s' % frames.dtype)
        n, h, w = frames.shape[:3]
        ch = frames.shape[3] if frames.ndim == 4 else 1
        host = t.from_numpy(frames.reshape(n, h, w * ch))
        dev = t.empty((n, h, w * ch), dtype=t.uint8, device=self.device)
        dev.copy_(host, non_blocking=True)
        return DeviceBatch('u8', dev, n, h, w, ch, extra={'keepalive': host})

    def download(self, batch):
        """ DeviceBatch -> pinned host tensor of the same (n, h, pitch) layout (async; the
        caller synchronises) """
        t = torch()
        host = t.empty(tuple(batch.t.shape), dtype=batch.t.dtype, pin_memory=True)
        host.copy_(batch.t, non_blocking=True)
        return host

    @staticmethod
    def host_view(batch, host):
        """ numpy view (n, h, w[, 3]) of a downloaded batch """
        a = host.numpy()
        if batch.kind == 'u8':
            a = a[:, :, :batch.w * batch.channels]
            if batch.channels == 3:
                a = a.reshape(batch.n, batch.h, batch.w, 3) if a.flags['C_CONTIGUOUS'] else \
                    np.lib.stride_tricks.as_strided(a, (batch.n, batch.h, batch.w, 3),
                                                    (a.strides[0], a.strides[1], 3, 1))
            return a
        if batch.kind in ('i32', 'i16'):
            return a[:, :, :batch.w]
        raise ValueError('packed masks are unpacked on the device before download')

    # ---- kernels ------------------------------------------------------------------------------------
    def luma(self, src, mode=_lib.MONO_MEAN, rect=None):
        """ K1 (+ crop by pointer offset): (n,h,w,3) -> (n,h',w') """
        left, top, w, h = rect if rect is not None else (0, 0, src.w, src.h)
        self._check_rect(src, left, top, w, h)
        self.ensure(w, h, src.n)
        out = self.empty_u8(src.n, h, w)
        ptr = src.ptr + top * src.pitch + left * 3
        self._check(self.lib.va_luma_u8(self._h, self.stream, ptr, src.pitch, src.fstride,
                                        out.ptr, out.pitch, out.fstride, w, h, src.n, mode))
        return out

    @staticmethod
    def _check_rect(src, left, top, w, h):
        """ a rectangle is turned into a pointer offset: it must lie inside the frame """
        if left < 0 or top < 0 or w <= 0 or h <= 0 or left + w > src.w or top + h > src.h:
            raise ValueError('rectangle %s exceeds the %dx%d frame' % ((left, top, w, h), src.w, src.h))

    def crop(self, src, rect):
        left, top, w, h = rect
        self._check_rect(src, left, top, w, h)
        self.ensure(w, h, src.n)
        out = self.empty_u8(src.n, h, w, src.channels)
        ptr = src.ptr + top * src.pitch + left * src.channels
        self._check(self.lib.va_copy2d_u8(self._h, self.stream, ptr, src.pitch, src.fstride,
                                          out.ptr, out.pitch, out.fstride, w * src.channels, h, src.n))
        return out

    def gauss(self, src, sigma):
        self.ensure(src.w, src.h, src.n)
        out = self.empty_u8(src.n, src.h, src.w, src.channels)
        self._check(self.lib.va_gauss_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                         out.ptr, out.pitch, out.fstride, src.w, src.h, src.channels, src.n,
                                         float(sigma)))
        return out

    def luma_gauss(self, src, sigma, mode=_lib.MONO_MEAN):
        self.ensure(src.w, src.h, src.n)
        out = self.empty_u8(src.n, src.h, src.w)
        self._check(self.lib.va_luma_gauss_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                              out.ptr, out.pitch, out.fstride, src.w, src.h, src.n, mode,
                                              float(sigma)))
        return out

    def resize_half(self, src):
        self.ensure(src.w, src.h, src.n)
        out = self.empty_u8(src.n, src.h // 2, src.w // 2, src.channels)
        self._check(self.lib.va_resize_half_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                               out.ptr, out.pitch, out.fstride, src.w, src.h, src.channels, src.n))
        return out

    def resize_area(self, src, kx, ky):
        """ cv2.resize(INTER_AREA) by the integer factors 1/kx, 1/ky """
        if kx == 2 and ky == 2:
            return self.resize_half(src)
        self.ensure(src.w, src.h, src.n)
        out = self.empty_u8(src.n, src.h // ky, src.w // kx, src.channels)
        self._check(self.lib.va_resize_area_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                               out.ptr, out.pitch, out.fstride, src.w, src.h, src.channels, src.n,
                                               int(kx), int(ky)))
        return out

    def resize_nearest(self, src, dw, dh):
        """ cv2.resize(INTER_NEAREST) to (dw, dh) """
        self.ensure(max(src.w, dw), max(src.h, dh), src.n)
        out = self.empty_u8(src.n, dh, dw, src.channels)
        self._check(self.lib.va_resize_nearest_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                                  out.ptr, out.pitch, out.fstride, src.w, src.h, int(dw), int(dh),
                                                  src.channels, src.n))
        return out

    def resize_area_any(self, src, dw, dh):
        """ cv2.resize(INTER_AREA) to (dw, dh), any factors (where a direction enlarges, cv2's INTER_AREA is
        its linear interpolation with the area coefficient rule) """
        self.ensure(max(src.w, dw), max(src.h, dh), src.n)
        out = self.empty_u8(src.n, dh, dw, src.channels)
        self._check(self.lib.va_resize_area_any_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                                   out.ptr, out.pitch, out.fstride, src.w, src.h, int(dw), int(dh),
                                                   src.channels, src.n))
        return out

    def resize_linear(self, src, dw, dh):
        """ cv2.resize(INTER_LINEAR) to (dw, dh) """
        self.ensure(max(src.w, dw), max(src.h, dh), src.n)
        out = self.empty_u8(src.n, dh, dw, src.channels)
        self._check(self.lib.va_resize_linear_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                                 out.ptr, out.pitch, out.fstride, src.w, src.h, int(dw), int(dh),
                                                 src.channels, src.n))
        return out

    def resize_cubic(self, src, dw, dh):
        """ cv2.resize(INTER_CUBIC) to (dw, dh): OpenCV's own arithmetic (cv2 without IPP) """
        self.ensure(max(src.w, dw), max(src.h, dh), src.n)
        out = self.empty_u8(src.n, dh, dw, src.channels)
        self._check(self.lib.va_resize_cubic_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                                out.ptr, out.pitch, out.fstride, src.w, src.h, int(dw), int(dh),
                                                src.channels, src.n))
        return out

    def resize_lanczos4(self, src, dw, dh):
        """ cv2.resize(INTER_LANCZOS4) to (dw, dh) """
        self.ensure(max(src.w, dw), max(src.h, dh), src.n)
        out = self.empty_u8(src.n, dh, dw, src.channels)
        self._check(self.lib.va_resize_lanczos4_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                                   out.ptr, out.pitch, out.fstride, src.w, src.h, int(dw), int(dh),
                                                   src.channels, src.n))
        return out

    def apply_mask(self, src, mask_dev):
        """ mask_dev: uint8 device tensor (h, w) -- one static mask for the whole batch -- or (n, h, w): one per frame
        of the batch (the streams of a multi-stream batch) """
        self.ensure(src.w, src.h, src.n)
        out = self.empty_u8(src.n, src.h, src.w, src.channels)
        per_frame = mask_dev.dim() == 3
        if per_frame and mask_dev.shape[0] != src.n:
            raise ValueError('%d masks for a batch of %d frames' % (mask_dev.shape[0], src.n))
        self._check(self.lib.va_apply_mask_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                              mask_dev.data_ptr(), mask_dev.stride(-2), mask_dev.stride(0) if per_frame else 0,
                                              out.ptr, out.pitch, out.fstride, src.w, src.h, src.channels, src.n))
        return out

    def luma_crop_multi(self, src, xy_dev, w, h, mode=_lib.MONO_MEAN):
        """ per-frame crop position (device int32 tensor (n, 2) of left, top), common size (w, h), then monochrome """
        self.ensure(src.w, src.h, src.n)
        out = self.empty_u8(src.n, h, w)
        self._check(self.lib.va_luma_crop_multi_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride, src.w, src.h,
                                                   out.ptr, out.pitch, out.fstride, int(w), int(h), src.n, mode,
                                                   xy_dev.data_ptr()))
        return out

    def ema_diff_thresh(self, src, bg, alpha, thr, first_frame_inits):
        self.ensure(src.w, src.h, src.n)
        out = self.empty_bits(src.n, src.h, src.w)
        self._check(self.lib.va_ema_diff_thresh(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                                bg.data_ptr(), bg.stride(0), out.ptr, out.pitch, out.fstride,
                                                src.w, src.h, src.n, float(alpha), float(thr),
                                                1 if first_frame_inits else 0))
        return out

    def threshold(self, src, thr):
        self.ensure(src.w, src.h, src.n)
        out = self.empty_bits(src.n, src.h, src.w)
        self._check(self.lib.va_threshold_bits(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                               out.ptr, out.pitch, out.fstride, src.w, src.h, src.n, int(thr)))
        return out

    def pack_bits(self, src):
        self.ensure(src.w, src.h, src.n)
        out = self.empty_bits(src.n, src.h, src.w)
        self._check(self.lib.va_pack_bits_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                             out.ptr, out.pitch, out.fstride, src.w, src.h, src.n))
        return out

    def unpack_bits(self, src):
        self.ensure(src.w, src.h, src.n)
        out = self.empty_u8(src.n, src.h, src.w)
        self._check(self.lib.va_unpack_bits_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                               out.ptr, out.pitch, out.fstride, src.w, src.h, src.n))
        return out

    def morph(self, src, op, shape='rect', ksize=3):
        kx, ky = (ksize, ksize) if np.isscalar(ksize) else ksize
        try:
            op_id, shape_id = _lib.MORPH_OPS[op], _lib.SE_SHAPES[shape]
        except KeyError as e:
            raise ValueError('unknown morphological operation or shape: %s' % e)
        self.ensure(src.w, src.h, src.n)
        out = self.empty_bits(src.n, src.h, src.w)
        self._check(self.lib.va_morph_bits(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                           out.ptr, out.pitch, out.fstride, src.w, src.h, src.n,
                                           op_id, shape_id, int(kx), int(ky)))
        return out

    def label(self, src, connectivity=4, dtype=np.int32):
        """ -> (labels DeviceBatch 'i32' -- or 'i16' for dtype=np.int16 --, counts int32 device tensor (n,)) """
        self.ensure(src.w, src.h, src.n)
        counts = torch().empty((src.n,), dtype=torch().int32, device=self.device)
        if np.dtype(dtype) == np.int16:
            out = self.empty_i16(src.n, src.h, src.w)
            self._check(self.lib.va_label_forest(self._h, self.stream, src.ptr, src.pitch, src.fstride, counts.data_ptr(),
                                                 src.w, src.h, src.n, int(connectivity), 0))
            self._check(self.lib.va_label_write_i16(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                                    out.ptr, out.pitch, out.fstride, src.w, src.h, src.n, 0))
            return out, counts
        out = self.empty_i32(src.n, src.h, src.w)
        self._check(self.lib.va_label_bits(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                           out.ptr, out.pitch, out.fstride, counts.data_ptr(),
                                           src.w, src.h, src.n, int(connectivity)))
        return out, counts

    def region_areas(self, labels, max_labels):
        t = torch()
        areas = t.empty((labels.n, max_labels), dtype=t.int32, device=self.device)
        largest = t.empty((labels.n,), dtype=t.int32, device=self.device)
        self.ensure(labels.w, labels.h, labels.n)
        self._check(self.lib.va_region_areas(self._h, self.stream, labels.ptr, labels.pitch, labels.fstride,
                                             areas.data_ptr(), int(max_labels), largest.data_ptr(),
                                             labels.w, labels.h, labels.n))
        return areas, largest

    def region_stats(self, mask, connectivity=4, max_regions=1024):
        """ per-region raw moments and bounding boxes straight from a packed mask (no label image):
        -> (stats int64 device tensor (n, max_regions, 10), counts int32 (n,), largest int32 (n,)) """
        t = torch()
        self.ensure(mask.w, mask.h, mask.n)
        stats = t.empty((mask.n, int(max_regions), REGION_FIELDS), dtype=t.int64, device=self.device)
        counts = t.empty((mask.n,), dtype=t.int32, device=self.device)
        largest = t.empty((mask.n,), dtype=t.int32, device=self.device)
        self._check(self.lib.va_region_stats(self._h, self.stream, mask.ptr, mask.pitch, mask.fstride,
                                             stats.data_ptr(), int(max_regions), counts.data_ptr(), largest.data_ptr(),
                                             mask.w, mask.h, mask.n, int(connectivity)))
        return stats, counts, largest

    def lut(self, src, table):
        """ out = table[in] for uint8 frames (FilterNormalize) """
        tab = np.ascontiguousarray(table, dtype=np.uint8)
        self.ensure(src.w, src.h, src.n)
        out = self.empty_u8(src.n, src.h, src.w, src.channels)
        self._check(self.lib.va_lut_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                       out.ptr, out.pitch, out.fstride, src.w * src.channels, src.h, src.n,
                                       tab.ctypes.data))
        return out

    def highlight_mask(self, src, bits, channel='all', strength=128):
        """ VideoComposer.highlight_mask (io/composer.py:131-154) on the device: `src` u8 frames, `bits` the
        packed mask of the same batch; returns the annotated frames """
        if channel is None or channel == 'all':
            ch = -1
        elif src.channels == 3:
            try:
                ch = {0: 0, 'r': 0, 'red': 0, 1: 1, 'g': 1, 'green': 1, 2: 2, 'b': 2, 'blue': 2}[channel]
            except (KeyError, TypeError):
                raise ValueError('Unknown value `%s` for channel.' % (channel,))
        else:
            raise ValueError('Highlighting a specific channel is only supported for color videos.')
        if (bits.n, bits.h, bits.w) != (src.n, src.h, src.w):
            raise ValueError('mask and frames differ in shape')
        factor = (255 - strength) / 255                                  # the reference's expression on 0..255
        table = np.empty(256, np.uint8)
        table[:] = strength + factor * np.arange(256, dtype=np.uint8)
        self.ensure(src.w, src.h, src.n)
        out = self.empty_u8(src.n, src.h, src.w, src.channels)
        self._check(self.lib.va_highlight_mask_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                                  bits.ptr, bits.pitch, bits.fstride, out.ptr, out.pitch, out.fstride,
                                                  src.w, src.h, src.channels, src.n, ch, table.ctypes.data))
        return out

    def rot90(self, src, k):
        self.ensure(max(src.w, src.h), max(src.w, src.h), src.n)
        ow, oh = (src.h, src.w) if k & 1 else (src.w, src.h)
        out = self.empty_u8(src.n, oh, ow, src.channels)
        self._check(self.lib.va_rot90_u8(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                         out.ptr, out.pitch, out.fstride, src.w, src.h, src.channels, src.n, int(k)))
        return out

    def time_diff(self, src):
        """ src holds n + 1 frames -> int16 tensor (n, h, w * channels) of consecutive differences """
        t = torch()
        n = src.n - 1
        self.ensure(src.w, src.h, src.n)
        out = t.empty((n, src.h, src.w * src.channels), dtype=t.int16, device=self.device)
        self._check(self.lib.va_time_diff_i16(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                              out.data_ptr(), out.stride(1), out.stride(0),
                                              src.w * src.channels, src.h, n))
        return out

    def mean_update(self, src, mean, m2, n0):
        """ fold a batch of uint8 frames into float64 running statistics (in place) """
        self.ensure(src.w, src.h, src.n)
        self._check(self.lib.va_mean_update_f64(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                                mean.data_ptr(), None if m2 is None else m2.data_ptr(),
                                                mean.stride(0), src.w * src.channels, src.h, src.n, int(n0)))

    def ema_partial(self, src, S, alpha, accumulate):
        self.ensure(src.w, src.h, src.n)
        self._check(self.lib.va_ema_partial(self._h, self.stream, src.ptr, src.pitch, src.fstride,
                                            S.data_ptr(), S.stride(0), src.w, src.h, src.n, float(alpha),
                                            1 if accumulate else 0))

    def ema_fold(self, carry, S, scale, w, h):
        self.ensure(w, h, 1)
        self._check(self.lib.va_ema_fold(self._h, self.stream, carry.data_ptr(), S.data_ptr(), carry.stride(0),
                                         w, h, float(scale)))

    def synth_rgb(self, out, t0, seed, blobs):
        """ fill DeviceBatch `out` (n,h,w,3) with frames t0.. of the seeded synthetic video """
        tab = np.ascontiguousarray(blobs, dtype=np.int32).reshape(-1, 5)
        self.ensure(out.w, out.h, 1)
        self._check(self.lib.va_synth_rgb(self._h, self.stream, out.ptr, out.pitch, out.fstride, out.w, out.h,
                                          int(t0), out.n, int(seed) & 0xFFFFFFFF, tab.ctypes.data, len(tab)))

    def chain_run(self, desc, io, w, h, n):
        self.ensure(w, h, n)
        self._check(self.lib.va_chain_run(self._h, self.stream, ctypes.byref(desc), ctypes.byref(io)))


_runtimes = {}
_lock = threading.Lock()


def get_runtime(device=None):
    """ the process-wide runtime of `device` (default: torch's current CUDA device) """
    t = torch()
    if device is None:
        if not t.cuda.is_available():
            raise _lib.VAError('no CUDA device available: video_analysis_b200 has no CPU fallback')
        device = t.cuda.current_device()
    idx = device if isinstance(device, int) else (t.device(device).index or 0)
    with _lock:
        rt = _runtimes.get(idx)
        if rt is None:
            rt = _runtimes[idx] = DeviceRuntime(idx)
        return rt
