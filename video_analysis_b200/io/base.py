"""
Video protocol of the reference, restated for Python 3 (drop-in boundary).

Mirrors the public behaviour of video/io/base.py of david-zwicker/video-analysis:
`VideoBase` (:26-269), `VideoIterator` (:273-283), `VideoFilterBase` (:313-389)
and `VideoSlice` (:392-474).  A video is an object with `size = (width, height)`,
`frame_count`, `fps`, `is_color`, an internal cursor (`get_frame_pos` /
`set_frame_pos`), random access `get_frame(i)` where the source allows it, and
sequential access through `get_next_frame()` / iteration.  Frames are NumPy arrays
of shape (height, width[, 3]).

Python-3 deltas: iterators implement `__next__` (and keep `next`), `range` replaces
`xrange`, and `VideoSlice` resolves negative / open bounds against the source it is
given (the reference reads an undefined `self.source` there, base.py:401-409).
`VideoFork` (sync fan-out bookkeeping, no pixels) is out of scope.
"""

import logging

import numpy as np

logger = logging.getLogger('video.io')


class NotSeekableError(RuntimeError):
    pass


class SynchronizationError(RuntimeError):
    pass


class VideoBase(object):
    """ base of everything that yields frames; see module docstring """

    write_access = False
    seekable = False

    def __init__(self, size=(0, 0), frame_count=-1, fps=None, is_color=True):
        if len(size) != 2:
            raise ValueError('Videos must have two spatial dimensions.')
        self.size = size
        self.frame_count = frame_count
        self.fps = 25 if fps is None else fps
        self.is_color = is_color
        self._listeners = []
        self._frame_pos = 0

    # ---- description -----------------------------------------------------------------
    def get_property_list(self):
        return ('size=(%d, %d)' % tuple(self.size),
                'frame_count=%s' % self.frame_count,
                'fps=%s' % self.fps,
                'is_color=%s' % self.is_color)

    def _listener_suffix(self):
        n = len(self._listeners)
        if n == 1:
            return '[1 listener]'
        return '[%d listeners]' % n if n else ''

    def __str__(self):
        return '%s(%s)%s' % (self.__class__.__name__, ', '.join(self.get_property_list()),
                             self._listener_suffix())

    def info(self):
        return 'Video(%s)' % ', '.join(self.get_property_list())

    def __len__(self):
        return self.frame_count

    @property
    def width(self):
        return self.size[0]

    @property
    def height(self):
        return self.size[1]

    @property
    def bounds(self):
        return (0, 0, self.width, self.height)

    @property
    def shape(self):
        shape = (self.frame_count, self.size[1], self.size[0])
        return shape + (3,) if self.is_color else shape

    @property
    def video_format(self):
        return {'size': self.size, 'frame_count': self.frame_count,
                'fps': self.fps, 'is_color': self.is_color}

    # ---- listeners ----------------------------------------------------------------------
    def register_listener(self, listener_callback):
        self._listeners.append(listener_callback)

    def unregister_listener(self, listener_callback):
        self._listeners.remove(listener_callback)

    def _process_frame(self, frame):
        """ hook every produced frame passes through: notifies the listeners """
        for observer in self._listeners:
            observer(frame)
        return frame

    # ---- cursor ---------------------------------------------------------------------------
    def get_frame_pos(self):
        return self._frame_pos

    def set_frame_pos(self, index):
        if index < 0:
            index += self.frame_count
        if self.seekable:
            if not 0 <= index < self.frame_count:
                raise IndexError('Seeking to frame %d was not possible.' % index)
            self._frame_pos = index
        elif index >= self.get_frame_pos():
            # forward-only stream: consume frames until we are there
            for _ in range(self.get_frame_pos(), index):
                self.get_next_frame()
        else:
            raise NotSeekableError('Cannot seek to frame %d, because the video is already at frame %d'
                                   % (index, self.get_frame_pos()))

    def rewind(self):
        self.set_frame_pos(0)

    # ---- frames ---------------------------------------------------------------------------
    def get_frame(self, index):
        raise NotImplementedError

    def get_next_frame(self):
        try:
            frame = self.get_frame(self._frame_pos)
        except IndexError:
            raise StopIteration
        self._frame_pos += 1
        return frame

    def abort_iteration(self):
        pass

    def close(self):
        pass

    def __iter__(self):
        return VideoIterator(self)

    def __getitem__(self, key):
        if isinstance(key, slice):
            return VideoSlice(self, *key.indices(self.frame_count))
        if isinstance(key, (int, np.integer)):
            return self.get_frame(int(key))
        raise TypeError('Invalid key `%r` for indexing' % key)

    def __setitem__(self, key, value):
        raise ValueError('Writing to this video stream is prohibited.')

    def copy(self, dtype=np.uint8, disp=False):
        """ materialise the video into a VideoMemory (reference: base.py:248-269) """
        from .memory import VideoMemory
        logger.debug('Copy a video stream and store it in memory')
        data = np.empty(self.shape, dtype)
        for k, frame in enumerate(self):
            data[k, ...] = frame
        return VideoMemory(data, fps=self.fps, copy_data=False)


class VideoIterator(object):
    """ iterator protocol over a video; rewinds it first (reference: base.py:273-283) """

    def __init__(self, video):
        self._video = video
        self._video.rewind()

    def __iter__(self):
        return self

    def __next__(self):
        try:
            return self._video.get_next_frame()
        except IndexError:
            raise StopIteration

    next = __next__


class VideoFilterBase(VideoBase):
    """ a lazy per-frame view on another video; subclasses override `_process_frame`
    and finish with `super()._process_frame(frame)` (reference: base.py:313-389) """

    def __init__(self, source, size=None, frame_count=None, fps=None, is_color=None):
        self._source = source
        super(VideoFilterBase, self).__init__(
            size=source.size if size is None else size,
            frame_count=source.frame_count if frame_count is None else frame_count,
            fps=source.fps if fps is None else fps,
            is_color=source.is_color if is_color is None else is_color)

    def __str__(self):
        return '%s +%s%s' % (self._source, self.__class__.__name__, self._listener_suffix())

    @property
    def seekable(self):
        return self._source.seekable

    def abort_iteration(self):
        self._source.abort_iteration()

    def set_frame_pos(self, index):
        if index < 0:
            index += self.frame_count
        self._source.set_frame_pos(index)
        self._frame_pos = index

    def get_frame_pos(self):
        return self._source.get_frame_pos()

    def get_frame(self, index):
        if index < 0:
            index += self.frame_count
        frame = self._source.get_frame(index)
        self._frame_pos = index
        return self._process_frame(frame)

    def get_next_frame(self):
        frame = self._source.get_next_frame()
        self._frame_pos += 1
        return self._process_frame(frame)

    def close(self, propagate=True):
        if propagate and isinstance(self._source, VideoFilterBase):
            self._source.close(propagate=True)
        else:
            self._source.close()


class VideoSlice(VideoFilterBase):
    """ frames start:stop:step of a source; the frame-range partition primitive
    (reference: base.py:392-474) """

    def __init__(self, source, start=0, stop=None, step=1):
        n = source.frame_count
        self._start = start if start >= 0 else n + start
        if stop is None:
            self._stop = n
        else:
            self._stop = stop if stop >= 0 else n + stop
        if step == 0:
            raise ValueError('step argument must not be zero.')
        self._step = step
        frame_count = max(0, int(np.ceil((self._stop - self._start) / self._step)))
        if frame_count > 0:
            source.set_frame_pos(self._start)
        super(VideoSlice, self).__init__(source, frame_count=frame_count)
        logger.debug('Created video slice [%d:%d%s] of length %d.', self._start, self._stop,
                     '' if step == 1 else ':%d' % step, frame_count)
        if step < 0:
            logger.warning('Reversing a video can slow down the processing significantly.')

    def _source_index(self, index):
        if index < 0:
            index += self.frame_count
        if not 0 <= index < self.frame_count:
            raise IndexError('Cannot access frame %d in video of length %d' % (index, self.frame_count))
        return index, self._start + index * self._step

    def set_frame_pos(self, index):
        index, src = self._source_index(index)
        self._source.set_frame_pos(src)
        self._frame_pos = index

    def get_frame_pos(self):
        return self._frame_pos

    def get_frame(self, index):
        return self._source.get_frame(self._source_index(index)[1])

    def get_next_frame(self):
        if self._frame_pos >= self.frame_count:
            self.abort_iteration()
            raise StopIteration
        if self._step == 1:
            frame = self._source.get_next_frame()
        else:
            frame = self.get_frame(self._frame_pos)
        self._frame_pos += 1
        return frame
