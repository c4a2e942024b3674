"""
Host logic of the raw-video ingest / egress (video_analysis_b200/io/pipe.py; reference:
video/io/backend_ffmpeg.py:273-323, video/io/parallel.py:386-489).  No GPU needed: the ring is
plain memory here and page-locked on the GPU box.
"""
import io
import os
import sys
import threading
import time

import numpy as np
import pytest

from video_analysis_b200.io.base import NotSeekableError
from video_analysis_b200.io.pipe import VideoRawStream, VideoPreprocessor, RawStreamWriter, RawStreamError

W, H = 20, 12


def frames_of(n, color=True, seed=3):
    shape = (n, H, W, 3) if color else (n, H, W)
    return np.random.default_rng(seed).integers(0, 256, shape, dtype=np.uint8)


def test_stream_from_file_object_frames_and_blocks():
    fr = frames_of(50)
    v = VideoRawStream(io.BytesIO(fr.tobytes()), (W, H), 50, ring_frames=12)
    assert v.shape == (50, H, W, 3) and len(v) == 50 and not v.seekable
    got = [f.copy() for f in v]
    assert len(got) == 50 and np.array_equal(np.stack(got), fr)
    v.close()
    # blocks: at most ring // (hold + 2) frames, contiguous, valid until `hold` further blocks were pulled
    v = VideoRawStream(io.BytesIO(fr.tobytes()), (W, H), 50, ring_frames=12, hold=2)
    pos, blocks = 0, []
    while True:
        b = v.frame_block(pos, pos + 7)
        if len(b) == 0:
            break
        assert len(b) <= 3 and b.flags['C_CONTIGUOUS'] and v.get_frame_pos() == pos + len(b)
        blocks.append((pos, b))
        time.sleep(0.01)                                            # let the reader run as far ahead as it may
        for p0, b0 in blocks[-2:]:                                  # hold = 2: this block and the one before it are intact
            assert np.array_equal(b0, fr[p0:p0 + len(b0)])
        assert v._produced - v._release <= 12
        pos += len(b)
    assert pos == 50
    v.close()


def test_stream_monochrome_and_child_process():
    fr = frames_of(30, color=False)
    path = os.path.join(os.path.dirname(__file__), '_raw_tmp.bin')
    try:
        fr.tofile(path)
        v = VideoRawStream(path, (W, H), 30, is_color=False, ring_frames=9)
        assert np.array_equal(np.stack([f.copy() for f in v]), fr)
        v.close()
        # a child process writing the byte layout of `ffmpeg -f image2pipe -vcodec rawvideo -`
        cmd = [sys.executable, '-c', 'import sys; sys.stdout.buffer.write(open(%r, "rb").read())' % path]
        v = VideoRawStream(cmd, (W, H), 30, is_color=False, ring_frames=9)
        assert np.array_equal(np.stack([f.copy() for f in v]), fr)
        v.close()
        # reopenable source -> backward seeks reopen, short forward seeks skip
        opened = []

        def factory(index):
            opened.append(index)
            f = open(path, 'rb', buffering=0)
            f.seek(index * W * H)
            return f
        v = VideoRawStream(factory, (W, H), 30, is_color=False, ring_frames=9, seek_max_frames=5)
        assert np.array_equal(v.get_frame(3), fr[3]) and opened == [0]
        assert np.array_equal(v.get_frame(3), fr[3])                # lastread
        assert np.array_equal(v.get_frame(20), fr[20]) and opened == [0, 20]
        assert np.array_equal(v.get_frame(-28), fr[2]) and opened == [0, 20, 2]
        assert np.array_equal(v.get_next_frame(), fr[3]) and v.get_frame_pos() == 4
        v.close()
    finally:
        if os.path.exists(path):
            os.remove(path)


def test_stream_short_reads_follow_the_reference():
    fr = frames_of(200)
    raw = fr.tobytes()
    # estimate too large, stream ends near the announced end -> plain end of video
    v = VideoRawStream(io.BytesIO(raw), (W, H), 202, ring_frames=12)
    assert sum(1 for _ in v) == 200
    # truncated in the middle: the last good frame is repeated once, then the video ends
    v = VideoRawStream(io.BytesIO(raw[:100 * W * H * 3 + 17]), (W, H), 200, ring_frames=12)
    got = [f.copy() for f in v]
    assert len(got) == 101 and np.array_equal(np.stack(got[:100]), fr[:100]) and np.array_equal(got[100], fr[99])
    # nothing at all
    v = VideoRawStream(io.BytesIO(b''), (W, H), 200, ring_frames=12)
    with pytest.raises(RawStreamError):
        v.get_next_frame()
    # forward-only stream cannot go back
    v = VideoRawStream(io.BytesIO(raw), (W, H), 200, ring_frames=12)
    v.set_frame_pos(40)
    assert np.array_equal(v.get_next_frame(), fr[40])
    with pytest.raises(NotSeekableError):
        v.set_frame_pos(3)
    v.close()


def test_reader_runs_ahead_of_the_consumer():
    fr = frames_of(40)

    class Slow(io.BytesIO):
        def readinto(self, b):
            time.sleep(0.002)
            return super().readinto(b)
    v = VideoRawStream(Slow(fr.tobytes()), (W, H), 40, ring_frames=60)
    time.sleep(0.3)                                                 # the reader fills the ring on its own
    assert v._produced >= 25
    t0 = time.perf_counter()
    b = v.frame_block(0, 10)
    assert len(b) == 10 and time.perf_counter() - t0 < 0.05 and np.array_equal(b, fr[:10])
    v.close()


def test_preprocessor_matches_reference_protocol():
    fr = frames_of(9, color=False)
    from video_analysis_b200.io.memory import VideoMemory
    for use_threads in (True, False):
        calls = []
        pp = VideoPreprocessor(VideoMemory(fr), {'neg': lambda f: 255 - f, 'sum': lambda f: int(f.sum())},
                               preprocess=lambda f: (calls.append(threading.get_ident()), f)[1], use_threads=use_threads)
        assert len(pp) == 9
        out = list(pp)
        assert len(out) == 9
        for i, d in enumerate(out):
            assert set(d) == {'raw', 'neg', 'sum'}
            assert np.array_equal(d['raw'], fr[i]) and np.array_equal(d['neg'], 255 - fr[i]) and d['sum'] == int(fr[i].sum())
        if use_threads:
            assert any(t != threading.get_ident() for t in calls)
    with pytest.raises(KeyError):
        VideoPreprocessor(VideoMemory(fr), {'raw': lambda f: f})


def test_writer_round_trip():
    fr = frames_of(7)
    sink = io.BytesIO()
    with RawStreamWriter(sink, (W, H)) as wr:
        wr.write_block(fr[:4])
        for f in fr[4:6]:
            wr.write_frame(f)
        wr.write_frame(fr[6, :, :, 0])                              # mono frame into a colour stream
        with pytest.raises(ValueError):
            wr.write_block(fr[:, :5])
        assert wr.frames_written == 7
    back = VideoRawStream(io.BytesIO(sink.getvalue()), (W, H), 7, ring_frames=6)
    got = np.stack([f.copy() for f in back])
    assert np.array_equal(got[:6], fr[:6]) and np.array_equal(got[6], np.repeat(fr[6, :, :, :1], 3, axis=2))


def test_single_frames_are_copies_and_survive_the_ring():
    """ get_next_frame / get_frame hand out fresh arrays like the reference's reader: collecting more
    frames than the ring holds before looking at them must not see overwritten data """
    fr = frames_of(200)
    v = VideoRawStream(io.BytesIO(fr.tobytes()), (W, H), 200, ring_frames=12)
    held = [v.get_next_frame() for _ in range(150)]                 # no copy by the caller
    time.sleep(0.05)
    assert np.array_equal(np.stack(held), fr[:150])
    assert all(f.base is None or not np.shares_memory(f, v._ring) for f in held)
    last = v.get_frame(149)                                         # lastread path
    assert np.array_equal(last, fr[149]) and not np.shares_memory(last, v._ring)
    v.close()
    # the collectors that stack single frames: FilterTimeDifference-style and analysis.video._blocks
    from video_analysis_b200.analysis.video import _blocks
    v = VideoRawStream(io.BytesIO(fr.tobytes()), (W, H), 200, ring_frames=12)
    got = np.concatenate([np.asarray(b) for b in _blocks(v, 32)])
    assert np.array_equal(got, fr)
    v.close()


@pytest.mark.parametrize('batch,ring', [(128, 384), (50, 384), (32, 100), (7, 12)])
def test_device_filter_pull_loop_sees_every_frame_of_a_stream(batch, ring):
    """ the loop DeviceFilterBase._launch runs (pull blocks until an EMPTY one comes back) over a raw
    stream whose blocks are capped at ring // (hold + 2) and cut where the ring wraps """
    from video_analysis_b200.filters import DeviceFilterBase
    fr = frames_of(1000, color=False, seed=5)
    v = VideoRawStream(io.BytesIO(fr.tobytes()), (W, H), 1000, is_color=False, ring_frames=ring, pinned=False)
    if batch == 32:
        v.set_frame_pos(5)                                          # a forward seek shifts the wrap position
    start = v.get_frame_pos()
    out, short = [], 0
    while True:
        block = DeviceFilterBase._pull_block(v, batch)
        if len(block) == 0:
            break
        short += len(block) < batch
        out.append(np.array(block))
    assert np.array_equal(np.concatenate(out), fr[start:])
    assert short > 1                                                # short blocks occurred and did not end the video
    v.close()


def test_file_source_with_parallel_readers():
    # the path of a regular file is read by several threads with os.preadv; frames come out in order, the ring is never
    # overrun, a file shorter / longer than announced behaves like the single-reader stream
    fr = frames_of(120)
    path = os.path.join(os.path.dirname(__file__), '_raw_par_tmp.bin')
    try:
        fr.tofile(path)
        for readers in (1, 2, 4, 7):
            v = VideoRawStream(path, (W, H), 120, ring_frames=12, readers=readers, pinned=False)
            assert (v._fd is not None) == (readers > 1)
            pos, out = 0, []
            while True:
                b = v.frame_block(pos, pos + 2)
                if len(b) == 0:
                    break
                out.append(b.copy())
                assert v._produced - v._release <= 12
                pos += len(b)
            assert np.array_equal(np.concatenate(out), fr), readers
            v.close()
        # length over-estimated by a frame: the stream ends where the file ends; iteration yields every frame
        v = VideoRawStream(path, (W, H), 121, ring_frames=12, pinned=False)
        assert np.array_equal(np.stack(list(v)), fr)
        v.close()
        # length under-estimated: `frame_count` is an estimate, the stream delivers what the file holds (like one reader)
        for readers in (1, 4):
            v = VideoRawStream(path, (W, H), 100, ring_frames=12, pinned=False, readers=readers)
            assert np.array_equal(np.stack(list(v)), fr)
            v.close()
        # a file that ends in the middle of a frame far from the announced end: the last good frame is repeated once
        with open(path, 'r+b') as f:
            f.truncate(60 * W * H * 3 + 17)
        v = VideoRawStream(path, (W, H), 120, ring_frames=12, pinned=False)
        got = np.stack(list(v))
        assert len(got) == 61 and np.array_equal(got[:60], fr[:60]) and np.array_equal(got[60], fr[59])
        v.close()
        # an empty file raises like the stream that ends before its first frame
        open(path, 'wb').close()
        v = VideoRawStream(path, (W, H), 120, ring_frames=12, pinned=False)
        with pytest.raises(RawStreamError):
            v.get_next_frame()
        v.close()
    finally:
        if os.path.exists(path):
            os.remove(path)
