"""
Host-side protocol of the drop-in boundary (video/io/base.py, video/io/memory.py,
constructor logic of video/filters.py): cursor, slicing, iteration, listeners, metadata
propagation, error behaviour.  CPU only -- nothing here launches a kernel.
"""

import numpy as np
import pytest

from video_analysis_b200 import filters
from video_analysis_b200.io.base import NotSeekableError, VideoBase, VideoFilterBase, VideoSlice
from video_analysis_b200.io.memory import VideoMemory


def video(t=10, h=12, w=16, color=True):
    shape = (t, h, w, 3) if color else (t, h, w)
    return VideoMemory(np.arange(np.prod(shape), dtype=np.uint32).reshape(shape).astype(np.uint8))


def test_metadata_and_shape():
    v = video()
    assert v.size == (16, 12) and v.width == 16 and v.height == 12
    assert v.shape == (10, 12, 16, 3) and len(v) == 10 and v.is_color and v.fps == 25
    assert v.bounds == (0, 0, 16, 12)
    assert v.video_format == {'size': (16, 12), 'frame_count': 10, 'fps': 25, 'is_color': True}
    assert str(v) == 'VideoMemory(size=(16, 12), frame_count=10, fps=25, is_color=True)'
    m = video(color=False)
    assert m.shape == (10, 12, 16) and not m.is_color
    assert VideoMemory(np.zeros((3, 4, 5, 1), np.uint8)).shape == (3, 4, 5)
    with pytest.raises(ValueError):
        VideoMemory(np.zeros((3, 4, 5, 2), np.uint8))
    with pytest.raises(ValueError):
        VideoBase(size=(1, 2, 3))


def test_frames_are_views_and_iteration_rewinds():
    v = video()
    assert v.get_frame(3).base is not None and np.array_equal(v.get_frame(-1), v.data[9])
    assert np.array_equal(v[2], v.data[2])
    frames = list(v)
    assert len(frames) == 10 and v.get_frame_pos() == 10
    assert len(list(v)) == 10                      # iterating again rewinds (VideoIterator)
    it = iter(v)
    assert np.array_equal(next(it), v.data[0]) and np.array_equal(it.next(), v.data[1])
    with pytest.raises(StopIteration):
        v.set_frame_pos(9); v.get_next_frame(); v.get_next_frame()


def test_seek_errors():
    v = video()
    v.set_frame_pos(-2)
    assert v.get_frame_pos() == 8
    with pytest.raises(IndexError):
        v.set_frame_pos(10)

    class Stream(VideoBase):
        def get_frame(self, index):
            if index >= 5:
                raise IndexError
            return np.full((2, 2), index, np.uint8)
    s = Stream(size=(2, 2), frame_count=5, is_color=False)
    s.set_frame_pos(3)                             # forward seek on a non-seekable video skips frames
    assert s.get_next_frame()[0, 0] == 3
    with pytest.raises(NotSeekableError):
        s.set_frame_pos(1)
    with pytest.raises(ValueError):
        s[0] = 1
    with pytest.raises(TypeError):
        s['a']


def test_slices():
    v = VideoBase.__getitem__(video(), slice(2, 8))
    assert isinstance(v, VideoSlice) and len(v) == 6
    base = video()
    s = VideoSlice(base, 2, 8, 2)
    assert len(s) == 3
    assert [f[0, 0, 0] for f in s] == [base.data[i][0, 0, 0] for i in (2, 4, 6)]
    assert np.array_equal(s.get_frame(-1), base.data[6])
    with pytest.raises(IndexError):
        s.get_frame(3)
    with pytest.raises(ValueError):
        VideoSlice(base, 0, 5, 0)
    t = VideoSlice(base, 4)                        # open end
    assert len(t) == 6 and np.array_equal(next(iter(t)), base.data[4])
    u = VideoSlice(base, -3, None)                 # negative start resolves against the source
    assert len(u) == 3
    assert str(s).endswith('+VideoSlice')


def test_listeners_and_filter_function():
    v = video()
    seen = []
    f = filters.FilterFunction(v, lambda fr: fr + 1)
    f.register_listener(lambda fr: seen.append(int(fr[0, 0, 0])))
    assert str(f).endswith('+FilterFunction[1 listener]')
    out = list(f)
    assert len(out) == 10 and seen[0] == int(v.data[0][0, 0, 0]) + 1 and len(seen) == 10
    assert np.array_equal(f.get_frame(-1), v.data[9] + 1)
    f.unregister_listener(f._listeners[0])
    assert str(f).endswith('+FilterFunction')
    assert f.seekable and f.get_frame_pos() == v.get_frame_pos()


def test_copy_materialises():
    v = video(t=4)
    c = filters.FilterFunction(v, lambda fr: 255 - fr).copy()
    assert isinstance(c, VideoMemory) and np.array_equal(c.data, 255 - v.data)


def test_crop_constructor_rules():
    v = video(h=48, w=64)
    c = filters.FilterCrop(v, rect=(10, 20, 30, 12))
    assert c.rect == (10, 20, 30, 12) and c.size == (30, 12) and c.is_color
    assert c.slices == (slice(20, 32), slice(10, 40))
    assert filters.FilterCrop(v, rect=(0.5, 0.25, 0.25, 0.5)).rect == (32, 12, 16, 24)
    assert filters.FilterCrop(v, rect=(-40, -30, 20, 10)).rect == (24, 18, 20, 10)
    assert filters.FilterCrop(v, region='lower right').rect == (32, 24, 32, 24)
    assert filters.FilterCrop(v, region='UPPER').rect == (0, 0, 64, 24)
    assert filters.FilterCrop(v, rect=(1, 1, 33, 21), size_alignment=4).rect == (1, 1, 32, 20)
    with pytest.raises(IndexError):
        filters.FilterCrop(v, rect=(0, 0, 64, 48))                      # full width is rejected like the reference
    with pytest.raises(IndexError):
        filters.FilterCrop(v, rect=(70, 0, 10, 10))
    inner = filters.FilterCrop(v, rect=(10, 20, 40, 20))
    outer = filters.FilterCrop(inner, rect=(5, 5, 10, 10), color_channel='red')
    assert outer.rect == (15, 25, 10, 10) and outer._source is v       # nested crops collapse
    assert outer.color_channel == 2 and not outer.is_color and outer.shape == (10, 10, 10)


def test_monochrome_blur_resize_constructors():
    v = video(h=48, w=64)
    m = filters.FilterMonochrome(v)
    assert not m.is_color and m.shape == (10, 48, 64) and m.mode == 'mean'
    assert filters.FilterMonochrome(v, 'Green').mode == 1 and filters.FilterMonochrome(v, 'r').mode == 2
    with pytest.raises(AttributeError):
        filters.FilterMonochrome(v, 1)                                  # reference quirk: int mode has no .lower()
    with pytest.raises(ValueError):
        filters.FilterMonochrome(v, 'hue')._mode_id()
    b = filters.FilterBlur(m)
    assert b.sigma == 3 and str(b).endswith('+FilterMonochrome +FilterBlur')
    r = filters.FilterResize(m, 0.5)
    assert r.size == (32, 24) and r.interpolation == 'area'
    assert filters.FilterResize(m, (64, 48)).interpolation is None
    assert filters.FilterResize(m, 2).interpolation == 'cubic'
    assert filters.FilterResize(m, 0.33, even_dimensions=True).size == (22, 16)
    assert filters.FilterResize(r, 0.25)._source is m                   # nested resizes collapse
    with pytest.raises(ValueError):
        filters.FilterResize(m, 0.5, interpolation='bogus')


def test_new_operator_constructors_validate():
    v = video(h=48, w=64, color=False)
    with pytest.raises(ValueError):
        filters.FilterApplyMask(v, np.ones((3, 3)))
    with pytest.raises(ValueError):
        filters.FilterBackgroundMask(video(), 0.05, 25)                 # needs monochrome input
    with pytest.raises(ValueError):
        filters.FilterMorphology(v, 'sharpen')
    with pytest.raises(ValueError):
        filters.FilterMorphology(v, 'open', 'star')
    with pytest.raises(ValueError):
        filters.FilterLabel(v, connectivity=6)
    chain = filters.FilterLabel(filters.FilterMorphology(filters.FilterBackgroundMask(v), 'open'))
    assert chain.shape == (10, 48, 64) and chain.batch == filters.DEFAULT_BATCH
    assert str(chain).endswith('+FilterBackgroundMask +FilterMorphology +FilterLabel')
    with pytest.raises(NotSeekableError):
        chain.set_frame_pos(3)
    assert filters.get_color_range(np.uint8) == (0, 255) and filters.get_color_range(np.float32) == (0, 1)


def test_replicate_and_drop_frames_index_logic():
    v = video(t=5)
    r = filters.FilterReplicate(v, 3)
    assert len(r) == 15 and [f[0, 0, 0] for f in r] == [v.data[i % 5][0, 0, 0] for i in range(15)]
    assert np.array_equal(r.get_frame(7), v.data[2]) and np.array_equal(r[-1], v.data[4])
    r.set_frame_pos(6)
    assert np.array_equal(r.get_next_frame(), v.data[1])
    with pytest.raises(IndexError):
        r.set_frame_pos(15)
    v = video(t=10)
    d = filters.FilterDropFrames(v, 3)
    assert len(d) == 4 and d.fps == 25 / 3
    assert [f[0, 0, 0] for f in d] == [v.data[i][0, 0, 0] for i in (0, 3, 6, 9)]
    assert np.array_equal(d.get_frame(-1), v.data[9])
    d2 = filters.FilterDropFrames(v, 2.5)
    assert len(d2) == 4 and [f[0, 0, 0] for f in d2] == [v.data[int(i * 2.5)][0, 0, 0] for i in range(4)]
    t = filters.FilterTimeDifference(video(t=6, color=False))
    assert len(t) == 5 and t.shape == (5, 12, 16)
    with pytest.raises(NotImplementedError):
        filters.FilterTimeDifference(v, dtype=np.float32)
    with pytest.raises(ValueError):
        filters.FilterRotate(v, 30)
    assert filters.FilterRotate(v, 270).size == (12, 16)


def test_device_filters_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from video_analysis_b200._lib import VAError
    with pytest.raises(VAError):
        next(iter(filters.FilterMonochrome(video())))


def test_central_moments_are_derived_like_cv2():
    # host half of regions.region_stats: integer raw moments -> cv2.moments' second-order entries
    import cv2
    from video_analysis_b200.analysis.regions import moments_from_raw
    rng = np.random.default_rng(3)
    for shape, p in (((40, 70), 0.5), ((300, 500), 0.3), ((7, 9), 0.9), ((1080, 1920), 0.6)):
        mask = (rng.random(shape) < p).astype(np.uint8)
        ys, xs = np.nonzero(mask)
        xs, ys = xs.astype(np.int64), ys.astype(np.int64)
        raw = [len(xs), xs.sum(), ys.sum(), (xs * xs).sum(), (xs * ys).sum(), (ys * ys).sum()]
        got, want = moments_from_raw(raw), cv2.moments(mask)
        # raw moments are exact; the central ones are differences of numbers as large as the raw
        # moments, so they agree to a few ulp of THOSE (cv2 itself rounds there)
        scale = {'m': 0.0, 'mu': 1e-15 * max(want['m20'], want['m11'], want['m02']),
                 'nu': 1e-15 * max(want['m20'], want['m11'], want['m02']) / want['m00'] ** 2}
        for key in got:
            assert got[key] == pytest.approx(want[key], rel=1e-13, abs=scale[key.rstrip('0123456789')]), key
