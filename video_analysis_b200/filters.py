"""
Filter classes of the reference (video/filters.py), backed by the B200 kernels.

Same class names, constructor keywords, metadata propagation and error behaviour as
the reference, so `FilterBlur(FilterMonochrome(FilterCrop(video, rect)))` is written
and iterated exactly as before:

    FilterCrop        video/filters.py:158-248   (pointer arithmetic; fused into the next kernel)
    FilterMonochrome  video/filters.py:348-374   K1  va_luma_u8
    FilterBlur        video/filters.py:378-392   K2  va_gauss_u8 / va_luma_gauss_u8
    FilterResize      video/filters.py:252-315   K2b va_resize_half_u8 (exact 1/2 INTER_AREA)
    FilterFunction    video/filters.py:56-72     host callback (breaks a device chain)

and the operators BASELINE.json's north star names that the reference does not have
(definitions: SURVEY.md 8c, oracle/ops.py):

    FilterApplyMask        K6  va_apply_mask_u8
    FilterBackgroundMask   K3  va_ema_diff_thresh   (running-average background, |diff| > thr)
    FilterThreshold            va_threshold_bits
    FilterMorphology       K4  va_morph_bits        (erode / dilate / open / close)
    FilterLabel            K5  va_label_bits        (scipy.ndimage.label semantics)

What is different from the reference, and why: a device filter does not process one
frame per pull.  The last device filter of a chain pulls `batch` frames ahead from the
first non-device source, runs every stage of the chain on the GPU for the whole batch and
then hands the frames out one by one (`get_frame_pos()` therefore reports frames handed
out, and the source's own cursor runs ahead).  Listeners of every stage still fire once
per frame, source first, when the frame is handed out.  `FilterBlur` never notifies its
listeners -- that is the reference's behaviour (filters.py:388-392) and is kept.
"""

import collections
import logging

import numpy as np

from . import _lib
from .device import DeviceBatch, get_runtime, torch
from .io.base import NotSeekableError, VideoFilterBase

logger = logging.getLogger('video')

# filters.py:32-34
COLOR_CHANNELS = {'blue': 0, 'b': 0, 0: 0,
                  'green': 1, 'g': 1, 1: 1,
                  'red': 2, 'r': 2, 2: 2}

DEFAULT_BATCH = 32


def get_color_range(dtype):
    """ filters.py:38-50 """
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        return info.min, info.max
    if np.issubdtype(dtype, np.floating):
        return 0, 1
    raise ValueError('Unsupported data type `%r`' % dtype)


def rect_to_slices(rect):
    """ video/analysis/regions.py:49-53 """
    return slice(rect[1], rect[3] + rect[1]), slice(rect[0], rect[2] + rect[0])


def _check_coordinate(value, max_value):
    """ filters.py:139-154: fractions in (-1, 1), negatives count from the far edge,
    anything outside [0, max) is an IndexError """
    if -1 < value < 1:
        value = int(value * max_value)
    if value < 0:
        value += max_value
    if not 0 <= value < max_value:
        raise IndexError('Coordinate %d is out of bounds [0, %d].' % (value, max_value))
    return value


class FilterFunction(VideoFilterBase):
    """ applies a host function to every frame (filters.py:56-72).  The callback needs the
    frame in host memory, so it ends a device chain. """

    def __init__(self, source, function):
        self._function = function
        super(FilterFunction, self).__init__(source)

    def _process_frame(self, frame):
        return super(FilterFunction, self)._process_frame(self._function(frame))


# =========================================================================================
# batched execution engine
# =========================================================================================
class _Inflight(object):
    __slots__ = ('event', 'n', 'result', 'result_host', 'taps', 'extras')


class DeviceFilterBase(VideoFilterBase):
    """ base of the GPU filters; subclasses implement `_device_process(rt, batch)` """

    temporal = False          # carries state from frame to frame (cannot seek)
    notifies = True           # whether produced frames are announced to listeners
    consumes = 'u8'           # kind of DeviceBatch the stage wants
    produces_dtype = np.uint8

    def __init__(self, source, batch=None, device=None, prefetch=True, **video_kwargs):
        super(DeviceFilterBase, self).__init__(source, **video_kwargs)
        self.batch = int(batch) if batch else getattr(source, 'batch', DEFAULT_BATCH)
        self._device = device if device is not None else getattr(source, '_device', None)
        self.prefetch = prefetch
        self._ready = collections.deque()
        self._inflight = None
        self._streams = None
        self._launch_no = 0
        self._prev_compute = None
        self._exhausted = False

    # ---- to be provided by subclasses -----------------------------------------------------
    def _device_process(self, rt, batch):
        raise NotImplementedError

    def _reset_state(self):
        pass

    # ---- chain discovery ---------------------------------------------------------------------
    def _stages(self):
        stages, src = [self], self._source
        while isinstance(src, DeviceFilterBase):
            stages.append(src)
            src = src._source
        stages.reverse()
        return src, stages

    @property
    def runtime(self):
        return get_runtime(self._device)

    def _flush(self):
        if self._inflight is not None:
            self._inflight.event.synchronize()
        self._inflight = None
        self._prev_compute = None
        self._ready.clear()
        self._exhausted = False

    # ---- cursor ----------------------------------------------------------------------------------
    def get_frame_pos(self):
        return self._frame_pos

    def set_frame_pos(self, index):
        if index < 0:
            index += self.frame_count
        root, stages = self._stages()
        if index != 0 and any(s.temporal for s in stages):
            raise NotSeekableError('Cannot seek to frame %d: the chain holds a running background model' % index)
        for s in stages:
            s._flush()
            s._reset_state()
            s._frame_pos = index
        root.set_frame_pos(index)

    # ---- batch execution -------------------------------------------------------------------------
    @staticmethod
    def _pull_block(root, n):
        """ up to n frames from the first non-device source as one (m, h, w[, 3]) array """
        if hasattr(root, 'frame_block') and hasattr(root, '_frame_pos'):
            p = root._frame_pos
            block = root.frame_block(p, min(p + n, root.frame_count))
            root._frame_pos = p + len(block)
            return block
        frames = []
        try:
            for _ in range(n):
                frames.append(np.asarray(root.get_next_frame()))
        except StopIteration:
            pass
        if not frames:
            return np.empty((0,))
        return np.stack(frames)

    def _run_stages(self, rt, stages, dev):
        """ enqueue every stage; returns (final DeviceBatch, {stage: DeviceBatch} for listener taps) """
        taps = {}
        i = 0
        while i < len(stages):
            st = stages[i]
            nxt = stages[i + 1] if i + 1 < len(stages) else None
            # peephole fusions: crop -> mono is pointer arithmetic; mono -> blur never writes the luma frame
            if isinstance(st, FilterCrop) and isinstance(nxt, FilterMonochrome) and not st._listeners \
                    and st.color_channel is None and dev.channels == 3:
                st._check_rect(dev)                        # same IndexError as the unfused crop
                dev = nxt._device_process(rt, dev, rect=st.rect)
                st, i = nxt, i + 1
            elif isinstance(st, FilterMonochrome) and isinstance(nxt, FilterBlur) and not st._listeners \
                    and dev.kind == 'u8' and dev.channels == 3 and nxt._fusable():
                dev = rt.luma_gauss(dev, nxt.sigma, st._mode_id())
                st, i = nxt, i + 1
            else:
                if st.consumes == 'bits' and dev.kind == 'u8':
                    dev = rt.pack_bits(dev)
                elif st.consumes == 'u8' and dev.kind == 'bits':
                    dev = rt.unpack_bits(dev)
                dev = st._device_process(rt, dev)
            if st._listeners and st.notifies and st is not stages[-1]:
                taps[st] = rt.unpack_bits(dev) if dev.kind == 'bits' else dev
            i += 1
        if dev.kind == 'bits':
            dev = rt.unpack_bits(dev)
        return dev, taps

    def _launch(self):
        """ pull one batch ahead and enqueue it; None when the source is exhausted """
        if self._exhausted:
            return None
        root, stages = self._stages()
        block = self._pull_block(root, self.batch)
        if len(block) == 0:
            # only an empty block ends the video: sources with `frame_block` may legitimately hand out
            # short blocks (VideoRawStream caps them at ring_frames // (hold + 2) and cuts them where its
            # ring wraps around), exactly as SegmentChain._blocks_of treats them
            self._exhausted = True
            return None
        t = torch()
        rt = self.runtime
        if self._streams is None:
            self._streams = [t.cuda.Stream(device=rt.device), t.cuda.Stream(device=rt.device)]
        stream = self._streams[self._launch_no & 1]
        self._launch_no += 1
        job = _Inflight()
        with t.cuda.device(rt.device), t.cuda.stream(stream):
            if self._prev_compute is not None:
                stream.wait_event(self._prev_compute)       # temporal state / scratch reuse
            dev = rt.upload(block)
            out, taps = self._run_stages(rt, stages, dev)
            self._prev_compute = t.cuda.Event()
            self._prev_compute.record(stream)
            job.result = out
            job.result_host = rt.download(out)
            job.taps = [(st, tb, rt.download(tb)) for st, tb in taps.items()]
            job.extras = self._collect_extras(rt, stages)
            job.event = t.cuda.Event()
            job.event.record(stream)
        job.n = len(block)
        return job

    def _collect_extras(self, rt, stages):
        return None

    def _advance(self):
        if self._inflight is None:
            self._inflight = self._launch()
        job, self._inflight = self._inflight, None
        if job is None:
            return
        if self.prefetch:
            self._inflight = self._launch()
        job.event.synchronize()
        rt_view = self.runtime.host_view(job.result, job.result_host)
        tap_views = [(st, self.runtime.host_view(tb, th)) for st, tb, th in job.taps]
        self._on_batch(job)
        for k in range(job.n):
            self._ready.append((rt_view[k], [(st, tv[k]) for st, tv in tap_views]))

    def _on_batch(self, job):
        pass

    def get_next_frame(self):
        if not self._ready:
            self._advance()
        if not self._ready:
            raise StopIteration
        frame, taps = self._ready.popleft()
        _, stages = self._stages()
        for st in stages:
            st._frame_pos += 1
        for st, tap_frame in taps:                 # listeners of the stages before this one, source first
            VideoFilterBase._process_frame(st, tap_frame)
        if self.notifies:
            return VideoFilterBase._process_frame(self, frame)
        return frame

    def get_frame(self, index):
        """ random access: runs the chain on that single frame """
        if index < 0:
            index += self.frame_count
        root, stages = self._stages()
        if any(s.temporal for s in stages):
            raise NotSeekableError('Random access is not possible through a running background model')
        t = torch()
        rt = self.runtime
        block = np.asarray(root.get_frame(index))[None]
        with t.cuda.device(rt.device):
            out, _ = self._run_stages(rt, stages, rt.upload(block))
            host = rt.download(out)
            t.cuda.current_stream(rt.device).synchronize()
        for st in stages:
            st._frame_pos = index
        frame = rt.host_view(out, host)[0]
        return VideoFilterBase._process_frame(self, frame) if self.notifies else frame

    def iter_batches(self):
        """ iterate over (m, h, w[, 3]) blocks of up to `batch` frames instead of single frames
        (same results, no per-frame Python overhead) """
        self.rewind()
        while True:
            if not self._ready:
                self._advance()
            if not self._ready:
                return
            frames = [f for f, _ in self._ready]
            self._ready.clear()
            _, stages = self._stages()
            for st in stages:
                st._frame_pos += len(frames)
            yield _as_block(frames)

    def copy(self, dtype=None, disp=False):
        """ materialise the filtered video (reference: io/base.py:248-269), batch-wise """
        from .io.memory import VideoMemory
        data = np.empty(self.shape, self.produces_dtype if dtype is None else dtype)
        k = 0
        for block in self.iter_batches():
            data[k:k + len(block)] = block
            k += len(block)
        return VideoMemory(data[:k], fps=self.fps, copy_data=False)

    def close(self, propagate=True):
        self._flush()
        super(DeviceFilterBase, self).close(propagate)


def _as_block(frames):
    """ frames handed out by one batch are consecutive views of one host buffer """
    first = frames[0]
    n = len(frames)
    if n > 1 and all(frames[i].__array_interface__['data'][0] - frames[i - 1].__array_interface__['data'][0]
                     == frames[1].__array_interface__['data'][0] - first.__array_interface__['data'][0]
                     for i in range(1, n)):
        step = frames[1].__array_interface__['data'][0] - first.__array_interface__['data'][0]
        return np.lib.stride_tricks.as_strided(first, (n,) + first.shape, (step,) + first.strides)
    return np.stack(frames)


# =========================================================================================
# the reference's filters
# =========================================================================================
class FilterCrop(DeviceFilterBase):
    """ crops the video to a rectangle (filters.py:158-248).  rect = (left, top, width,
    height), floats in (-1, 1) are fractions, negatives count from the far edge; or
    `region` built from 'left' / 'right' / 'upper' / 'lower'.  Nested crops collapse into
    one; `color_channel` additionally picks one channel (-> monochrome video). """

    def __init__(self, source, rect=None, region='', color_channel=None, size_alignment=1, **kwargs):
        source_width, source_height = source.size
        if rect is not None:
            left = _check_coordinate(rect[0], source_width)
            top = _check_coordinate(rect[1], source_height)
            width = _check_coordinate(rect[2], source_width)
            height = _check_coordinate(rect[3], source_height)
        else:
            region = region.lower()
            left, top, width, height = 0, 0, source_width, source_height
            if 'left' in region:
                width //= 2
            elif 'right' in region:
                width //= 2
                left = source_width - width
            if 'upper' in region:
                height //= 2
            elif 'lower' in region:
                height //= 2
                top = source_height - height

        while isinstance(source, FilterCrop):            # filters.py:209-215
            logger.debug('Combine this crop filter with the parent one.')
            left += source.rect[0]
            top += source.rect[1]
            if source.color_channel is not None:
                color_channel = source.color_channel
            source = source._source

        self.color_channel = COLOR_CHANNELS.get(color_channel, color_channel)
        is_color = None if color_channel is None else False
        if size_alignment != 1:
            width = int(round(width / size_alignment) * size_alignment)
            height = int(round(height / size_alignment) * size_alignment)
        self.rect = (left, top, width, height)
        # `_check_coordinate` validates every value on its own, so left + width (or a width enlarged by
        # `size_alignment`) can still leave the frame.  The reference's NumPy slicing would then silently clip and
        # hand out frames smaller than `size` says; the device path turns the rectangle into a pointer offset and
        # must not read past a row, so the rectangle is rejected here, with the error type of a bad coordinate
        # (collapsed nested crops are checked against the frame of the outermost source)
        if left + width > source.size[0] or top + height > source.size[1] or width <= 0 or height <= 0:
            raise IndexError('Crop rectangle %s exceeds the %dx%d frame' % (self.rect, source.size[0], source.size[1]))
        self.slices = rect_to_slices(self.rect)
        super(FilterCrop, self).__init__(source, size=self.rect[2:], is_color=is_color, **kwargs)
        logger.debug('Created filter for cropping to rectangle %s', self.rect)

    def _check_rect(self, batch):
        """ the constructor checked the rectangle against the source's `size`; this guards the pointer arithmetic
        against a source whose frames are smaller than its metadata says """
        left, top, w, h = self.rect
        if left < 0 or top < 0 or w <= 0 or h <= 0 or left + w > batch.w or top + h > batch.h:
            # not an IndexError: VideoIterator turns those into the end of the video (io/base.py:279-283)
            raise ValueError('Crop rectangle %s exceeds the %dx%d frames the source delivers' % (self.rect, batch.w, batch.h))

    def _device_process(self, rt, batch):
        self._check_rect(batch)
        if self.color_channel is None:
            return rt.crop(batch, self.rect)
        if batch.channels != 3:
            raise IndexError('too many indices: cannot pick a colour channel of a monochrome frame')
        if self.color_channel not in (0, 1, 2):
            raise IndexError('index %r is out of bounds for the colour axis' % (self.color_channel,))
        return rt.luma(batch, self.color_channel, rect=self.rect)


class FilterMonochrome(DeviceFilterBase):
    """ colour -> monochrome (filters.py:348-374): mode 'mean' or a channel name """

    def __init__(self, source, mode='mean', **kwargs):
        self.mode = COLOR_CHANNELS.get(mode.lower(), mode.lower())
        super(FilterMonochrome, self).__init__(source, is_color=False, **kwargs)
        logger.debug('Created filter for converting video to monochrome with method `%s`', mode)

    def _mode_id(self):
        if self.mode == 'mean':
            return _lib.MONO_MEAN
        if self.mode in (0, 1, 2):
            return self.mode
        raise ValueError('Unsupported conversion method to monochrome: %s' % self.mode)

    def _device_process(self, rt, batch, rect=None):
        mode = self._mode_id()
        if batch.channels != 3:
            raise ValueError('Unsupported conversion method to monochrome: %s' % self.mode)
        return rt.luma(batch, mode, rect=rect)


class FilterBlur(DeviceFilterBase):
    """ Gaussian blur (filters.py:378-392): cv2.GaussianBlur(frame.astype(uint8), (0, 0), sigma),
    bit-exact.  Like the reference it does not notify its listeners. """

    notifies = False

    def __init__(self, source, sigma=3, **kwargs):
        self.sigma = sigma
        super(FilterBlur, self).__init__(source, **kwargs)
        logger.debug('Created filter blurring the video with radius %g', sigma)

    def _fusable(self):
        taps = (6 * self.sigma + 1)
        return 3 <= taps <= 2 * 63 + 1 and self.sigma >= 0.5

    def _device_process(self, rt, batch):
        return rt.gauss(batch, self.sigma)


class FilterResize(DeviceFilterBase):
    """ resizes the video (filters.py:252-315).  The device path implements INTER_AREA ('auto' picks it when
    the pixel count shrinks; integer factors take OpenCV's exact integer / single-rounding form, other shrink
    factors its float32 area tables, and where a direction enlarges its linear interpolation with the area
    coefficient rule, as cv2 does), INTER_LINEAR (OpenCV's 11-bit fixed point)
    and INTER_NEAREST for any size, and INTER_CUBIC ('auto' when enlarging) with OpenCV's own arithmetic --
    bit-exact against cv2 with IPP switched off, within 1 LSB of the IPP routine the cv2 wheel uses by
    default -- and INTER_LANCZOS4 (OpenCV's 8-tap fixed point, bit-exact). """

    def __init__(self, source, size=None, interpolation='auto', even_dimensions=False, **kwargs):
        if hasattr(size, '__iter__'):
            width, height = size
        else:
            width = int(source.size[0] * size)
            height = int(source.size[1] * size)
        if even_dimensions:
            width += (width % 2)
            height += (height % 2)

        if (width, height) == tuple(source.size):
            self.interpolation = None
        elif interpolation == 'auto':
            self.interpolation = 'area' if width * height < source.size[0] * source.size[1] else 'cubic'
        elif interpolation in ('nearest', 'linear', 'area', 'cubic', 'lanczos'):
            self.interpolation = interpolation
        else:
            raise ValueError('Unknown interpolation method: %s' % interpolation)

        while isinstance(source, FilterResize):          # filters.py:299-301
            logger.debug('Combine this resize filter with the parent one.')
            source = source._source
        super(FilterResize, self).__init__(source, size=(width, height), **kwargs)
        logger.debug('Created filter for resizing to size %dx%d', width, height)

    def _device_process(self, rt, batch):
        if self.interpolation is None:
            return batch
        w, h = self.size
        if self.interpolation == 'area' and batch.w % w == 0 and batch.h % h == 0:
            return rt.resize_area(batch, batch.w // w, batch.h // h)
        if self.interpolation == 'area':
            return rt.resize_area_any(batch, w, h)
        if self.interpolation == 'linear':
            return rt.resize_linear(batch, w, h)
        if self.interpolation == 'nearest':
            return rt.resize_nearest(batch, w, h)
        if self.interpolation == 'cubic':
            return rt.resize_cubic(batch, w, h)
        return rt.resize_lanczos4(batch, w, h)


class FilterNormalize(DeviceFilterBase):
    """ clips to [vmin, vmax] and rescales to the colour range of `dtype` (filters.py:76-135).
    Bounds / dtype that are not given are taken from the first frame, as in the reference.
    The device path covers uint8 -> uint8 (one table look-up per byte, the table being the
    reference's own expression evaluated on 0..255); other dtypes raise NotImplementedError. """

    def __init__(self, source, vmin=None, vmax=None, dtype=None, **kwargs):
        self._fmin, self._fmax, self._dtype = vmin, vmax, dtype
        self._table = None
        super(FilterNormalize, self).__init__(source, **kwargs)

    def _build_table(self, first_frame):
        if self._dtype is None:
            self._dtype = first_frame.dtype
        if np.dtype(self._dtype) != np.uint8 or first_frame.dtype != np.uint8:
            raise NotImplementedError('FilterNormalize on the device handles uint8 frames and dtype=uint8')
        if self._fmin is None:
            self._fmin = first_frame.min()
        if self._fmax is None:
            self._fmax = first_frame.max()
        tmin, tmax = get_color_range(self._dtype)
        alpha = (tmax - tmin) / (self._fmax - self._fmin)
        ramp = np.arange(256, dtype=np.uint8)
        np.clip(ramp, self._fmin, self._fmax, out=ramp)
        self._table = ((ramp - self._fmin) * alpha + tmin).astype(self._dtype)     # filters.py:126-132

    def _device_process(self, rt, batch):
        if self._table is None:
            t = torch()
            first = batch.t[0, :, :batch.w * batch.channels].cpu().numpy()         # bounds come from the first frame
            self._build_table(first)
        return rt.lut(batch, self._table)


class FilterRotate(DeviceFilterBase):
    """ rotates the video counter-clockwise by 0 / 90 / 180 / 270 degrees (filters.py:319-344) """

    def __init__(self, source, angle=0, **kwargs):
        angle = angle % 360
        if angle in (0, 180):
            size = source.size
        elif angle in (90, 270):
            size = (source.size[1], source.size[0])
        else:
            raise ValueError('angle must be from [0, 90, 180, 270] but was %s' % angle)
        self.angle = angle
        super(FilterRotate, self).__init__(source, size=size, **kwargs)

    def _device_process(self, rt, batch):
        return rt.rot90(batch, self.angle // 90) if self.angle else batch


class FilterReplicate(VideoFilterBase):
    """ replicates the video `count` times (filters.py:396-430); pure index logic """

    def __init__(self, source, count=1):
        self.count = count
        super(FilterReplicate, self).__init__(source, frame_count=source.frame_count * count)

    def get_frame_pos(self):
        return self._frame_pos

    def set_frame_pos(self, index):
        if index < 0:
            index += self.frame_count
        if not 0 <= index < self.frame_count:
            raise IndexError('Cannot access frame %d.' % index)
        self._source.set_frame_pos(index % self._source.frame_count)
        self._frame_pos = index

    def get_frame(self, index):
        if index < 0:
            index += self.frame_count
        if not 0 <= index < self.frame_count:
            raise IndexError('Cannot access frame %d.' % index)
        return self._source.get_frame(index % self._source.frame_count)

    def get_next_frame(self):
        if self._frame_pos >= self.frame_count:
            raise StopIteration
        if self._frame_pos % self._source.frame_count == 0:
            self._source.set_frame_pos(0)
        frame = self._source.get_next_frame()
        self._frame_pos += 1
        return frame


class FilterDropFrames(VideoFilterBase):
    """ keeps every `compression`-th frame (filters.py:434-483); pure index logic """

    def __init__(self, source, compression=1):
        self._compression = compression
        frame_count = int((source.frame_count - 1) / compression) + 1
        super(FilterDropFrames, self).__init__(source, frame_count=frame_count, fps=source.fps / compression)

    def _source_index(self, index):
        return int(index * self._compression)

    def get_frame_pos(self):
        return self._frame_pos

    def set_frame_pos(self, index):
        if index < 0:
            index += self.frame_count
        if not 0 <= index < self.frame_count:
            raise IndexError('Cannot access frame %d.' % index)
        self._source.set_frame_pos(self._source_index(index))
        self._frame_pos = index

    def get_frame(self, index):
        if index < 0:
            index += self.frame_count
        frame = self._source[self._source_index(index)]
        self._frame_pos = index + 1
        return frame

    def get_next_frame(self):
        if self._frame_pos >= self.frame_count:
            raise StopIteration
        frame = self._source[self._source_index(self._frame_pos)]
        self._frame_pos += 1
        return frame


class FilterTimeDifference(VideoFilterBase):
    """ differences between consecutive frames, int16 by default (filters.py:492-568).
    Frame t of this video is source[t + 1] - source[t]; one frame shorter than the source.
    Iterating pulls `batch` + 1 source frames at a time and differences them on the GPU. """

    def __init__(self, source, dtype=np.int16, batch=DEFAULT_BATCH, device=None):
        if dtype is not None and np.dtype(dtype) != np.int16:
            raise NotImplementedError('FilterTimeDifference on the device produces int16 differences')
        self._dtype = dtype
        self.batch = batch
        self._device = device
        self._ready = collections.deque()
        self._last = None                      # last source frame of the previous block
        super(FilterTimeDifference, self).__init__(source, frame_count=source.frame_count - 1)

    def get_frame_pos(self):
        return self._frame_pos

    def set_frame_pos(self, index):
        if index < 0:
            index += self.frame_count
        self._source.set_frame_pos(index)
        self._ready.clear()
        self._last = None
        self._frame_pos = index

    def _diff(self, block):
        t = torch()
        rt = get_runtime(self._device)
        with t.cuda.device(rt.device):
            out = rt.time_diff(rt.upload(block))
            host = out.cpu().numpy()
        return host.reshape((len(block) - 1,) + block.shape[1:])

    def get_frame(self, index):
        if index < 0:
            index += self.frame_count
        pair = np.stack([np.asarray(self._source.get_frame(index)), np.asarray(self._source.get_frame(index + 1))])
        return self._process_frame(self._diff(pair)[0])

    def get_next_frame(self):
        if not self._ready:
            frames = [] if self._last is None else [self._last]
            try:
                while len(frames) < self.batch + 1:
                    frames.append(np.asarray(self._source.get_next_frame()))
            except StopIteration:
                pass
            if len(frames) < 2:
                raise StopIteration
            self._last = frames[-1]
            self._ready.extend(self._diff(np.stack(frames)))
        self._frame_pos += 1
        return self._process_frame(self._ready.popleft())


# =========================================================================================
# operators named by the north star that the reference does not ship
# =========================================================================================
class FilterApplyMask(DeviceFilterBase):
    """ out = frame where mask != 0 else 0; one static (H, W) mask for the whole video """

    def __init__(self, source, mask, **kwargs):
        mask = np.asarray(mask)
        if mask.shape != (source.size[1], source.size[0]):
            raise ValueError('mask shape %s does not match the frame shape %s'
                             % (mask.shape, (source.size[1], source.size[0])))
        self.mask = np.ascontiguousarray(mask != 0, dtype=np.uint8)
        self._mask_dev = None
        super(FilterApplyMask, self).__init__(source, **kwargs)

    def _device_process(self, rt, batch):
        if self._mask_dev is None:
            t = torch()
            pitch = (self.mask.shape[1] + 15) // 16 * 16
            self._mask_dev = t.zeros((self.mask.shape[0], pitch), dtype=t.uint8, device=rt.device)
            self._mask_dev[:, :self.mask.shape[1]].copy_(t.from_numpy(self.mask))
        return rt.apply_mask(batch, self._mask_dev)


class FilterBackgroundMask(DeviceFilterBase):
    """ running-average background model, difference and threshold in one kernel:

        bg_0 = frame_0, mask_0 = 0
        d = frame_t - bg_{t-1};  mask_t = |d| > threshold;  bg_t = bg_{t-1} + alpha * d   (float32)

    yields uint8 masks {0, 255}.  The model is sequential in time, so the filter can be
    iterated and rewound but not seeked.  `background` is the current model (H, W) float32. """

    temporal = True

    def __init__(self, source, alpha=0.05, threshold=25, **kwargs):
        if source.is_color:
            raise ValueError('the background model works on monochrome videos')
        self.alpha, self.threshold = float(alpha), float(threshold)
        self._bg = None
        self._started = False
        super(FilterBackgroundMask, self).__init__(source, **kwargs)

    def _reset_state(self):
        self._started = False

    def _device_process(self, rt, batch):
        if self._bg is None or self._bg.shape[0] != batch.h or self._bg.shape[1] < batch.w:
            self._bg = rt.empty_f32(batch.h, batch.w)
        out = rt.ema_diff_thresh(batch, self._bg, self.alpha, self.threshold, not self._started)
        self._started = True
        return out

    @property
    def background(self):
        if self._bg is None:
            return None
        torch().cuda.synchronize(self._bg.device)
        return self._bg[:, :self.size[0]].cpu().numpy()

    def set_background(self, bg):
        """ continue from a given model (e.g. the carry of the previous frame range) """
        rt = self.runtime
        bg = np.asarray(bg, dtype=np.float32)
        self._bg = rt.empty_f32(bg.shape[0], bg.shape[1])
        self._bg[:, :bg.shape[1]].copy_(torch().from_numpy(np.ascontiguousarray(bg)))
        self._started = True


class FilterThreshold(DeviceFilterBase):
    """ mask = frame > threshold (strict, as cv2.threshold THRESH_BINARY) -> uint8 {0, 255} """

    def __init__(self, source, threshold=127, **kwargs):
        if source.is_color:
            raise ValueError('thresholding works on monochrome videos')
        self.threshold = int(threshold)
        super(FilterThreshold, self).__init__(source, **kwargs)

    def _device_process(self, rt, batch):
        return rt.threshold(batch, self.threshold)


class FilterMorphology(DeviceFilterBase):
    """ binary erode / dilate / open / close with cv2.getStructuringElement(shape, ksize)
    semantics (reference call sites: video/analysis/image.py:248-256) """

    consumes = 'bits'

    def __init__(self, source, operation='open', shape='rect', ksize=3, **kwargs):
        if operation not in _lib.MORPH_OPS:
            raise ValueError('unknown morphological operation %r' % (operation,))
        if shape not in _lib.SE_SHAPES:
            raise ValueError('unknown structuring element shape %r' % (shape,))
        self.operation, self.shape_name, self.ksize = operation, shape, ksize
        super(FilterMorphology, self).__init__(source, **kwargs)

    def _device_process(self, rt, batch):
        return rt.morph(batch, self.operation, self.shape_name, self.ksize)


class FilterLabel(DeviceFilterBase):
    """ connected-component labelling with scipy.ndimage.label semantics
    (video/analysis/regions.py:162): int32 labels, 0 = background, 1..n in raster order of
    each component's first pixel.  `num_features` holds n for every frame handed out so far. """

    consumes = 'bits'
    produces_dtype = np.int32

    def __init__(self, source, connectivity=4, dtype=np.int32, **kwargs):
        if connectivity not in (4, 8):
            raise ValueError('connectivity must be 4 or 8')
        self.connectivity = connectivity
        # dtype=np.int16 is ndimage.label(mask, output=np.int16): the same numbering in half the bytes (frames with more
        # than 32767 regions raise RuntimeError, as scipy does)
        self.produces_dtype = np.dtype(dtype).type
        if self.produces_dtype not in (np.int32, np.int16):
            raise ValueError('labels are int32 or int16')
        self.num_features = []
        self._counts_dev = None
        super(FilterLabel, self).__init__(source, **kwargs)

    def _reset_state(self):
        self.num_features = []

    def _device_process(self, rt, batch):
        labels, counts = rt.label(batch, self.connectivity, self.produces_dtype)
        self._counts_dev = counts
        return labels

    def _collect_extras(self, rt, stages):
        if self._counts_dev is None:
            return None
        t = torch()
        host = t.empty(tuple(self._counts_dev.shape), dtype=t.int32, pin_memory=True)
        host.copy_(self._counts_dev, non_blocking=True)
        return host

    def _on_batch(self, job):
        if job.extras is not None:
            counts = [int(v) for v in job.extras.numpy()[:job.n]]
            if self.produces_dtype is np.int16 and counts and max(counts) > 32767:
                raise RuntimeError('insufficient bit-depth in requested output type')
            self.num_features.extend(counts)
