"""
Raw-video ingest and egress for the device path.

Reference anchors: `VideoFFmpeg` reads `ffmpeg -f image2pipe -vcodec rawvideo -pix_fmt rgb24 -`
from a pipe, `depth * w * h` bytes per frame, into a fresh `bytes` object per frame
(video/io/backend_ffmpeg.py:228-237, :273-323); `VideoPreprocessor` reads the next frame on a
worker thread while the caller works on the current one (video/io/parallel.py:386-489);
`VideoFileWriter` / `VideoWriterFFmpeg` write raw frames to the stdin of an encoder.

B200 version: the byte stream is read by a background thread with `readinto` **straight into a ring
of page-locked frames**, and device filters take whole batches from that ring through `frame_block`
-- no per-frame `bytes` object, no copy between the pipe and the DMA source.  Anything that produces
or consumes raw frames works: an `ffmpeg` command line (there is no ffmpeg binary in this image, so
the tests use a Python child process that writes the same byte layout), a file of raw frames, a
socket file object.

Short-read handling follows the reference (backend_ffmpeg.py:289-314): `frame_count` may be an
estimate; a short read near the announced end (fewer than 5 frames or 1 % left) ends the video,
a short read elsewhere repeats the last good frame with a warning, and a short read on the very
first frame raises `RawStreamError`.
"""

import collections
import logging
import os
import subprocess
import threading

import numpy as np

from .base import VideoBase, NotSeekableError

logger = logging.getLogger('video.io')


class RawStreamError(IOError):
    """ the stream ended before the first frame (FFmpegError in the reference) """


def _alloc_frames(n, shape, pinned):
    """ (n,) + shape uint8 array; page-locked when a GPU is there and `pinned` is set """
    if pinned:
        import torch
        if torch.cuda.is_available():
            t = torch.empty((n,) + tuple(shape), dtype=torch.uint8, pin_memory=True)
            return t.numpy(), t
    return np.empty((n,) + tuple(shape), np.uint8), None


class VideoRawStream(VideoBase):
    """ forward-only video over a byte stream of packed uint8 frames.

    `source`      a binary file object with `readinto`, a path of a raw file, or an argv list
                  (started with subprocess, frames read from its stdout, e.g. the reference's
                  ffmpeg command line); a callable `source(index)` returning one of those makes
                  the video reopenable and therefore seekable backwards (backend_ffmpeg.py:171-243)
    `size`        (width, height); `is_color` selects 3 bytes or 1 byte per pixel
    `frame_count` announced length (may be an estimate, see module docstring)
    `ring_frames` frames held in the page-locked ring: `hold` + 2 blocks of what the consumer pulls
    `hold`        a block handed out stays untouched until this many further blocks have been requested
                  (4 covers the three-deep pipeline of SegmentChain.process_blocks: it pulls block k + 3
                  before it waits for the results of block k, so the asynchronous upload of block k is
                  only known to be complete when block k + 4 is requested)
    `seek_max_frames`  forward seeks up to this distance skip frames instead of reopening
                  (backend_ffmpeg.py:255-268)
    `readers`     reader threads for a source that is the path of a regular file: a file can be read at any
                  offset, so the frames are fetched by `readers` threads with `os.preadv` (the GIL is released
                  during the copy) and handed out in order (default: half the cores, at most 8; 1080p rgb24 from
                  tmpfs on an 8-core host: 0.50 k frames/s with one reader, 1.9 k with 4, 3.5 k with 8).  Pipes,
                  sockets and file objects are single streams and keep the one reader thread
    """

    seekable = False

    def __init__(self, source, size, frame_count, fps=25, is_color=True, ring_frames=384, pinned=True,
                 seek_max_frames=100, hold=4, readers=None):
        super(VideoRawStream, self).__init__(size=size, frame_count=frame_count, fps=fps, is_color=is_color)
        self.depth = 3 if is_color else 1
        w, h = size
        self.frame_shape = (h, w, 3) if is_color else (h, w)
        self.frame_bytes = self.depth * w * h
        self.hold = int(hold)
        if self.frame_bytes <= 0 or self.hold < 1 or ring_frames < self.hold + 2:
            raise ValueError('VideoRawStream needs a non-empty frame size and a ring of at least hold + 2 frames')
        self.ring_frames = int(ring_frames)
        self.max_block = self.ring_frames // (self.hold + 2)
        self.seek_max_frames = seek_max_frames
        if readers is None:                                          # half the cores, at most 8
            readers = min(8, (os.cpu_count() or 2) // 2)
        self.readers = max(1, int(readers))
        self._factory = source if callable(source) else None
        self._ring, self._ring_owner = _alloc_frames(self.ring_frames, self.frame_shape, pinned)
        self._flat = self._ring.reshape(self.ring_frames, self.frame_bytes)
        self._cv = threading.Condition()
        self._thread = None
        self._threads = []
        self._fd = None
        self._proc = None
        self._stream = None
        self._owns_stream = False
        self.lastread = None
        self._open(source(0) if self._factory else source, 0)

    # ---- stream handling ---------------------------------------------------------------------
    def _open(self, source, index):
        self._shutdown_reader()
        if isinstance(source, (list, tuple)):
            self._proc = subprocess.Popen(list(source), stdout=subprocess.PIPE, bufsize=0)
            self._stream, self._owns_stream = self._proc.stdout, True
            try:                                        # 1 MiB pipe instead of 64 KiB: 16 times fewer read calls per frame
                import fcntl
                fcntl.fcntl(self._stream.fileno(), getattr(fcntl, 'F_SETPIPE_SZ', 1031), 1 << 20)
            except (ImportError, OSError):
                pass
        elif isinstance(source, (str, bytes)):
            if self.readers > 1 and os.path.isfile(source) and hasattr(os, 'preadv'):
                self._fd = os.open(source, os.O_RDONLY)          # parallel readers, see _file_reader
            else:
                self._stream, self._owns_stream = open(source, 'rb', buffering=0), True
        else:
            self._stream, self._owns_stream = source, False
        self._frame_pos = index
        self._base = index            # video index of the first frame this stream delivers
        self._produced = index        # frames [base, produced) have been read into the ring
        self._release = index         # ring slots of frames < release may be overwritten
        self._held = collections.deque()                            # starts of the blocks still promised intact
        self._eof = False
        self._error = None
        self._stop = False
        if self._fd is not None:
            self._claim = index       # next frame a reader thread will take
            self._done = {}           # frames read out of order: frame -> bytes obtained
            self._threads = [threading.Thread(target=self._file_reader, name='VideoRawStream-reader-%d' % i, daemon=True)
                             for i in range(self.readers)]
            for t in self._threads:
                t.start()
            return
        self._thread = threading.Thread(target=self._reader, name='VideoRawStream-reader', daemon=True)
        self._thread.start()

    def _shutdown_reader(self):
        if self._threads:
            with self._cv:
                self._stop = True
                self._cv.notify_all()
            for t in self._threads:
                t.join(timeout=10)
            self._threads = []
        if self._fd is not None:
            os.close(self._fd)
            self._fd = None
        if self._thread is not None:
            with self._cv:
                self._stop = True
                self._cv.notify_all()
            if self._proc is not None:
                self._proc.kill()
            if self._owns_stream and self._proc is None:
                try:
                    self._stream.close()
                except Exception:                       # the reader may sit in readinto on it
                    pass
            self._thread.join(timeout=10)
            self._thread = None
        if self._proc is not None:
            self._proc.stdout.close()
            self._proc.wait()
            self._proc = None
        elif self._owns_stream and self._stream is not None:
            self._stream.close()
        self._stream = None

    def _read_frame_into(self, row):
        """ fills one ring row from the stream; returns the number of bytes obtained """
        view = memoryview(row)
        got = 0
        while got < self.frame_bytes:
            n = self._stream.readinto(view[got:])
            if not n:
                break
            got += n
        return got

    def _short_read(self, slot, got):
        """ frame `_produced` came back short (lock held): the reference's rules (module docstring); ends the stream """
        remaining = self.frame_count - self._produced
        if remaining < 5 or remaining < 0.01 * self.frame_count:
            self._eof = True
        elif self._produced == self._base:
            self._error = RawStreamError('Failed to read the first frame of the raw stream '
                                         '(%d of %d bytes)' % (got, self.frame_bytes))
            self._eof = True
        else:
            logger.warning('%d bytes wanted but %d bytes read at frame %d/%d. Using the last '
                           'valid frame instead.', self.frame_bytes, got, self._produced,
                           self.frame_count)
            # the reference hands out `lastread` again without advancing; a stream that
            # stays short would never end, so the gap is filled once and the stream ends
            prev = (self._produced - 1) % self.ring_frames
            self._flat[slot] = self._flat[prev]
            self._produced += 1
            self._eof = True
        self._cv.notify_all()

    def _reader(self):
        """ background thread: stream -> ring, as far ahead as the ring allows """
        try:
            while True:
                with self._cv:
                    while not self._stop and self._produced - self._release >= self.ring_frames:
                        self._cv.wait()
                    if self._stop:
                        return
                    slot = self._produced % self.ring_frames
                got = self._read_frame_into(self._flat[slot])      # blocking I/O outside the lock
                with self._cv:
                    if self._stop:
                        return
                    if got != self.frame_bytes:
                        self._short_read(slot, got)
                        return
                    self._produced += 1
                    self._cv.notify_all()
        except Exception as err:                                   # surface I/O errors to the consumer
            with self._cv:
                if not self._stop:
                    self._error = err
                self._eof = True
                self._cv.notify_all()

    def _file_reader(self):
        """ one of `readers` threads over a regular file: takes the next frame index, reads that frame at its offset
        straight into its ring slot (os.preadv, GIL released), and advances `_produced` over the frames that are
        complete in order.  A frame that comes back short ends the stream exactly as in `_reader`. """
        try:
            while True:
                with self._cv:
                    while not self._stop and not self._eof and self._claim - self._release >= self.ring_frames:
                        self._cv.wait()
                    if self._stop or self._eof:
                        return
                    f = self._claim
                    self._claim += 1
                slot = f % self.ring_frames
                view = memoryview(self._flat[slot])
                off = (f - self._base) * self.frame_bytes
                got = 0
                while got < self.frame_bytes:
                    n = os.preadv(self._fd, [view[got:]], off + got)
                    if n <= 0:
                        break
                    got += n
                with self._cv:
                    if self._stop or self._eof:
                        return
                    self._done[f] = got
                    while self._produced in self._done:
                        g = self._done.pop(self._produced)
                        if g != self.frame_bytes:
                            self._short_read(self._produced % self.ring_frames, g)
                            return
                        self._produced += 1
                    self._cv.notify_all()
        except Exception as err:
            with self._cv:
                if not self._stop:
                    self._error = err
                self._eof = True
                self._cv.notify_all()

    def _wait_for(self, stop):
        """ blocks until frame stop-1 is in the ring or the stream has ended; returns the frames available """
        with self._cv:
            while self._produced < stop and not self._eof:
                self._cv.wait()
            if self._error is not None:
                raise self._error
            return min(stop, self._produced)

    def _advance_release(self, upto):
        with self._cv:
            if upto > self._release:
                self._release = upto
                self._cv.notify_all()

    # ---- VideoBase protocol ------------------------------------------------------------------
    def frame_block(self, start, stop):
        """ frames [start, stop) as one contiguous array in the page-locked ring (device filters upload
        it as it is) and moves the cursor behind them.  Forward-only: a `start` other than the cursor
        seeks first.  The block may be shorter than asked for -- at the end of the stream, where the
        ring wraps around, and it never exceeds ring_frames // (hold + 2).  It stays valid until `hold`
        further blocks have been requested. """
        if start != self._frame_pos:
            self.set_frame_pos(start)
        stop = min(stop, start + self.max_block)
        self._held.append(start)
        if len(self._held) > self.hold:                             # the block handed out `hold` calls ago expires
            self._held.popleft()
            self._advance_release(self._held[0])
        avail = self._wait_for(stop)
        if avail <= start:
            return self._ring[0:0]
        a = start % self.ring_frames
        n = min(avail - start, self.ring_frames - a)
        block = self._ring[a:a + n]
        self.lastread = block[-1]
        self._frame_pos = start + n                                 # the cursor follows the blocks handed out
        return block

    def get_next_frame(self):
        """ one frame, as a fresh array like the reference's reader (backend_ffmpeg.py:318-319): callers collect
        frames one by one before they stack them (filters._pull_block, FilterTimeDifference, analysis.video),
        and a view into the ring would be overwritten by the reader thread `hold` requests later.  Only
        `frame_block` hands out views (documented there). """
        block = self.frame_block(self._frame_pos, self._frame_pos + 1)
        if len(block) == 0:
            raise StopIteration
        return block[0].copy()

    def get_frame(self, index):
        if index < 0:
            index += self.frame_count
        if index == self._frame_pos - 1 and self.lastread is not None:
            return self.lastread.copy()                             # backend_ffmpeg.py:336-337
        self.set_frame_pos(index)
        return self.get_next_frame()

    def set_frame_pos(self, index):
        if index < 0:
            index += self.frame_count
        if index == self._frame_pos:
            return
        if index < self._frame_pos or index > self._frame_pos + self.seek_max_frames:
            if self._factory is None:
                if index < self._frame_pos:
                    raise NotSeekableError('Cannot seek to frame %d, because the stream is already at frame %d'
                                           % (index, self._frame_pos))
            else:
                self._open(self._factory(index), index)             # reopen at the position
                return
        while self._frame_pos < index:                              # skip frames (they are read and dropped)
            if len(self.frame_block(self._frame_pos, index)) == 0:
                raise IndexError('Seeking to frame %d was not possible.' % index)

    def close(self):
        self._shutdown_reader()

    def __del__(self):
        try:
            self._shutdown_reader()
        except Exception:
            pass


class VideoPreprocessor(object):
    """ reads a video on a worker thread and applies `functions` to every frame on further threads while
    the caller works on the previous result (video/io/parallel.py:386-489): iteration yields
    `{'raw': frame, name: functions[name](frame), ...}`.  The functions should release the GIL -- every
    call into libva_b200 does (ctypes). """

    def __init__(self, video, functions, preprocess=None, use_threads=True):
        if 'raw' in functions:
            raise KeyError('The key `raw` is reserved for the raw _frame and may not be used for functions.')
        from concurrent.futures import ThreadPoolExecutor
        self.length = len(video)
        self.video_iter = iter(video)
        self.functions = functions
        self.preprocess = preprocess
        self._pool = ThreadPoolExecutor(max_workers=len(functions) + 1) if use_threads else None
        self._frame = None
        self._results = None
        self._next = None
        self._init_next_processing(self._get_next_frame())

    def __len__(self):
        return self.length

    def _submit(self, fn, *args):
        if self._pool is not None:
            return self._pool.submit(fn, *args).result
        value = fn(*args)
        return lambda: value

    def _get_next_frame(self):
        try:
            frame = next(self.video_iter)
        except StopIteration:
            return None
        if self.preprocess:
            frame = self.preprocess(frame)
        return frame

    def _init_next_processing(self, frame_next):
        self._frame = frame_next
        if frame_next is None:
            return
        self._results = {name: self._submit(func, frame_next) for name, func in self.functions.items()}
        self._next = self._submit(self._get_next_frame)

    def __iter__(self):
        return self

    def __next__(self):
        if self._frame is None:
            if self._pool is not None:
                self._pool.shutdown(wait=False)
            raise StopIteration
        result = {name: get() for name, get in self._results.items()}
        result['raw'] = self._frame
        self._init_next_processing(self._next())
        return result

    next = __next__


class RawStreamWriter(object):
    """ writes frames as packed uint8 bytes to a binary file object, a path or the stdin of a command
    (an encoder: the reference pipes rgb24 frames into ffmpeg, video/io/backend_ffmpeg.py:385-489).
    `write_block` takes a whole (n, h, w[, 3]) result block of the device path in one call. """

    def __init__(self, sink, size, is_color=True):
        self.size = tuple(size)
        self.is_color = is_color
        self.frames_written = 0
        self._proc = None
        if isinstance(sink, (list, tuple)):
            self._proc = subprocess.Popen(list(sink), stdin=subprocess.PIPE, bufsize=0)
            self._stream, self._owns = self._proc.stdin, True
        elif isinstance(sink, (str, bytes)):
            self._stream, self._owns = open(sink, 'wb'), True
        else:
            self._stream, self._owns = sink, False

    @property
    def shape(self):
        w, h = self.size
        return (h, w, 3) if self.is_color else (h, w)

    def write_block(self, frames):
        frames = np.asarray(frames)
        if frames.dtype != np.uint8 or frames.shape[1:] != self.shape:
            raise ValueError('expected uint8 frames of shape %s, got %s %s'
                             % (self.shape, frames.dtype, frames.shape[1:]))
        self._stream.write(memoryview(np.ascontiguousarray(frames)).cast('B'))
        self.frames_written += len(frames)

    def write_frame(self, frame):
        frame = np.asarray(frame)
        if not self.is_color and frame.ndim == 3:
            raise ValueError('Cannot write a color frame to a monochrome stream')
        if self.is_color and frame.ndim == 2:                       # mono to RGB, as the reference's writer does
            frame = np.repeat(frame[:, :, None], 3, axis=2)
        self.write_block(frame[None])

    def close(self):
        if self._stream is None:
            return
        if self._owns:
            self._stream.close()
        else:
            self._stream.flush()
        if self._proc is not None:
            self._proc.wait()
        self._stream = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
