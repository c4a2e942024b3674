"""
The reference-facing plug-in API on a real GPU: filter classes, SegmentChain and the
analysis helpers, compared with the oracle on the same seeded frames (bit-exact).
"""

import numpy as np
import pytest

from oracle import ops, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def frames():
    return synth.make_frames(0, 0, 40, 320, 240, 6)


@pytest.fixture(scope='module')
def ref(frames):
    return ops.chain(frames)


def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from video_analysis_b200 import filters
    from video_analysis_b200.io.memory import VideoMemory
    return filters, VideoMemory


def test_filter_chain_iteration_matches_reference_chain(frames, ref):
    F, VideoMemory = mods()
    v = VideoMemory(frames, copy_data=False)
    mono = F.FilterMonochrome(v, batch=16)
    assert np.array_equal(np.stack(list(mono)), ref['mono'])
    blur = F.FilterBlur(F.FilterMonochrome(v, batch=16), sigma=2)          # fused luma+blur path
    assert np.array_equal(np.stack(list(blur)), ref['blur'])
    mask = F.FilterBackgroundMask(F.FilterBlur(F.FilterMonochrome(v, batch=7), 2), 0.05, 25)
    assert np.array_equal(np.stack(list(mask)), ref['mask'])
    assert np.array_equal(mask.background.view(np.uint32), ref['bg'].view(np.uint32))
    assert np.array_equal(np.stack(list(mask)), ref['mask'])              # rewinding restarts the model
    full = F.FilterLabel(F.FilterMorphology(
        F.FilterBackgroundMask(F.FilterBlur(F.FilterMonochrome(v, batch=16), 2)), 'open', 'rect', 3))
    labels = np.stack(list(full))
    assert labels.dtype == np.int32 and np.array_equal(labels, ref['labels'])
    assert full.num_features == list(ref['counts'])
    assert full.get_frame_pos() == 40


def test_listeners_fire_per_frame_source_first_and_blur_stays_silent(frames, ref):
    F, VideoMemory = mods()
    v = VideoMemory(frames[:10], copy_data=False)
    log = []
    mono = F.FilterMonochrome(v, batch=4)
    mono.register_listener(lambda f: log.append(('mono', f.copy())))
    blur = F.FilterBlur(mono, 2)
    blur.register_listener(lambda f: log.append(('blur', None)))           # must never fire (filters.py:388-392)
    thr = F.FilterThreshold(blur, 100)
    thr.register_listener(lambda f: log.append(('thr', f.copy())))
    out = list(thr)
    assert [k for k, _ in log] == ['mono', 'thr'] * 10
    assert all(np.array_equal(log[2 * i][1], ref['mono'][i]) for i in range(10))
    assert all(np.array_equal(out[i], np.where(ref['blur'][i] > 100, 255, 0)) for i in range(10))


def test_crop_resize_applymask_and_random_access(frames):
    F, VideoMemory = mods()
    v = VideoMemory(frames[:9], copy_data=False)
    rect = (33, 10, 160, 120)
    crop = F.FilterCrop(v, rect, batch=4)
    assert np.array_equal(np.stack(list(crop)), np.stack([ops.crop(f, rect) for f in frames[:9]]))
    cm = F.FilterMonochrome(F.FilterCrop(v, rect, batch=4))
    exp = np.stack([ops.mono(ops.crop(f, rect)) for f in frames[:9]])
    assert np.array_equal(np.stack(list(cm)), exp)
    assert np.array_equal(cm.get_frame(5), exp[5]) and np.array_equal(cm[-1], exp[8])
    red = F.FilterCrop(v, rect, color_channel='red', batch=4)
    assert np.array_equal(np.stack(list(red)), np.stack([ops.crop(f, rect, 'red') for f in frames[:9]]))
    half = F.FilterResize(cm, 0.5)
    assert np.array_equal(np.stack(list(half)), np.stack([ops.resize(f, 0.5) for f in exp]))
    quarter = F.FilterResize(cm, 0.25)                                    # 160x120 -> 40x30: INTER_AREA by 4
    assert np.array_equal(np.stack(list(quarter)), np.stack([ops.resize(f, 0.25) for f in exp]))
    third = F.FilterResize(crop, (32, 40))                                # colour, factors 5 and 3
    assert np.array_equal(np.stack(list(third)), np.stack([ops.resize(ops.crop(f, rect), (32, 40)) for f in frames[:9]]))
    near = F.FilterResize(cm, (57, 201), interpolation='nearest')         # any size, also enlarging
    assert np.array_equal(np.stack(list(near)), np.stack([ops.resize(f, (57, 201), 'nearest') for f in exp]))
    frac = F.FilterResize(cm, 0.3)                                        # 48x36: INTER_AREA by 10/3
    assert np.array_equal(np.stack(list(frac)), np.stack([ops.resize(f, 0.3) for f in exp]))
    lin = F.FilterResize(crop, (201, 57), interpolation='linear')         # colour, enlarging x, shrinking y
    assert np.array_equal(np.stack(list(lin)), np.stack([ops.resize(ops.crop(f, rect), (201, 57), 'linear') for f in frames[:9]]))
    big = np.stack(list(F.FilterResize(cm, 2.0)))                          # 'auto' enlarges with INTER_CUBIC
    ref_big = np.stack([ops.resize(f, 2.0) for f in exp])
    assert np.abs(big.astype(np.int16) - ref_big).max() <= 1              # cv2's default is Intel IPP here: 1 LSB
    lz = F.FilterResize(cm, 1.5, interpolation='lanczos')
    assert np.array_equal(np.stack(list(lz)), np.stack([ops.resize(f, 1.5, 'lanczos') for f in exp]))
    m = np.zeros((120, 160), bool)
    m[20:90, 30:140] = True
    masked = F.FilterApplyMask(cm, m)
    assert np.array_equal(np.stack(list(masked)), np.stack([ops.apply_mask(f, m) for f in exp]))
    colour_blur = F.FilterBlur(crop, 1.5)
    assert np.array_equal(np.stack(list(colour_blur)), np.stack([ops.blur(ops.crop(f, rect), 1.5) for f in frames[:9]]))
    from video_analysis_b200.io.base import VideoSlice
    sl = F.FilterMonochrome(VideoSlice(v, 2, 7), batch=3)                  # VideoSlice as the root
    assert np.array_equal(np.stack(list(sl)), np.stack([ops.mono(f) for f in frames[2:7]]))


def test_copy_and_iter_batches(frames, ref):
    F, VideoMemory = mods()
    v = VideoMemory(frames, copy_data=False)
    blur = F.FilterBlur(F.FilterMonochrome(v, batch=16), 2)
    c = blur.copy()
    assert isinstance(c, VideoMemory) and np.array_equal(c.data, ref['blur'])
    blocks = list(blur.iter_batches())
    assert [len(b) for b in blocks] == [16, 16, 8] and np.array_equal(np.concatenate(blocks), ref['blur'])


def test_function_filter_breaks_the_device_chain(frames):
    F, VideoMemory = mods()
    v = VideoMemory(frames[:6], copy_data=False)
    chain = F.FilterBlur(F.FilterFunction(F.FilterMonochrome(v, batch=4), lambda f: 255 - f), 1)
    assert np.array_equal(np.stack(list(chain)), np.stack([ops.blur(255 - ops.mono(f), 1) for f in frames[:6]]))


def test_segment_chain_host_pipeline(frames, ref):
    mods()
    from video_analysis_b200.chain import SegmentChain
    for fuse in (True, False):
        ch = SegmentChain((320, 240), sigma=2, alpha=0.05, threshold=25, morph_op='open', morph_ksize=3,
                          batch=8, fuse=fuse)
        labels, counts = ch.process(frames)
        assert np.array_equal(labels, ref['labels']) and np.array_equal(counts, ref['counts'])
        assert np.array_equal(ch.background.view(np.uint32), ref['bg'].view(np.uint32))
        # continuing in two calls == one call
        ch.reset()
        l1, c1 = ch.process(frames[:13])
        l2, c2 = ch.process(frames[13:])
        assert np.array_equal(np.concatenate([l1, l2]), ref['labels'])


def test_segment_chain_region_tables(frames, ref):
    # same chain, per-region statistics as the result: equal to statistics of the oracle's label images
    mods()
    from video_analysis_b200.chain import SegmentChain
    ch = SegmentChain((320, 240), sigma=2, alpha=0.05, threshold=25, morph_op='open', morph_ksize=3, batch=8)
    regs = ch.process_regions(frames, max_regions=64)
    assert len(regs) == len(frames)
    assert np.array_equal(ch.background.view(np.uint32), ref['bg'].view(np.uint32))
    for t in range(len(frames)):
        lab, n = ref['labels'][t], ref['counts'][t]
        assert len(regs[t]) == n
        areas = ops.region_areas(lab, n)
        for reg in regs[t]:
            ys, xs = np.nonzero(lab == reg['label'])
            assert reg['area'] == areas[reg['label'] - 1]
            assert reg['bbox'] == (xs.min(), ys.min(), xs.max() - xs.min() + 1, ys.max() - ys.min() + 1)
            assert reg['moments']['m10'] == xs.sum() and reg['moments']['m02'] == (ys.astype(np.int64) ** 2).sum()
    ch.reset()
    with pytest.raises(MemoryError):
        ch.process_regions(frames[:30], max_regions=1)


def test_config5_like_chain(frames):
    """ stencil-heavy variant: sigma 5, 7x7 close then labelling with 8-connectivity """
    mods()
    from video_analysis_b200.chain import SegmentChain
    r = ops.chain(frames[:12], sigma=5, alpha=0.1, thr=15, morph_op='close', morph_shape='ellipse', morph_ksize=7, connectivity=8)
    ch = SegmentChain((320, 240), sigma=5, alpha=0.1, threshold=15, morph_op='close', morph_shape='ellipse',
                      morph_ksize=7, connectivity=8, batch=5)
    labels, counts = ch.process(frames[:12])
    assert np.array_equal(labels, r['labels']) and np.array_equal(counts, r['counts'])


def test_analysis_helpers(frames, ref):
    mods()
    from video_analysis_b200.analysis import image, regions
    m = ref['morph'][20]
    lab, n = regions.label(m)
    assert n == ref['counts'][20] and np.array_equal(lab, ref['labels'][20])
    lab8, n8 = regions.label(ref['mask'][20] > 0, connectivity=8)
    r8 = ops.label(ref['mask'][20], 8)
    assert n8 == r8[1] and np.array_equal(lab8, r8[0])
    big, area = regions.get_largest_region(m, ret_area=True)
    rbig, rarea = ops.get_largest_region(m, ret_area=True)
    assert area == rarea and np.array_equal(big, rbig)
    with pytest.raises(ValueError):
        regions.get_largest_region(np.zeros((20, 30), np.uint8))
    assert np.array_equal(image.opening(ref['mask'][20]), ops.morph(ref['mask'][20], 'open'))
    assert np.array_equal(image.erode(ref['mask'][20]), ops.morph(ref['mask'][20], 'erode', 'cross', 3))
    assert np.array_equal(image.closing(ref['mask'][20], 'ellipse', 5), ops.morph(ref['mask'][20], 'close', 'ellipse', 5))


def test_region_statistics_and_regionprops(frames, ref):
    mods()
    from video_analysis_b200.analysis import image, regions
    m = ref['morph'][25]
    lab, n = ops.label(m)
    regs = regions.region_stats(m)
    assert len(regs) == n and n >= 2
    for reg, mom in zip(regs, ops.region_moments(lab, n)):
        for key in ('m00', 'm10', 'm01', 'm20', 'm11', 'm02', 'mu20', 'mu11', 'mu02', 'nu20', 'nu11', 'nu02'):
            assert reg['moments'][key] == pytest.approx(mom[key], rel=1e-12, abs=1e-14 * mom['m20'] + 1e-12), key
        mask_l = lab == reg['label']
        assert reg['bbox'] == ops.find_bounding_box(mask_l) and reg['area'] == mask_l.sum()
        got, want = image.regionprops(moments=reg['moments']), ops.regionprops(mask=mask_l)
        assert got.centroid == pytest.approx(want.centroid, rel=1e-12)
        assert got.orientation == pytest.approx(want.orientation, rel=1e-9, abs=1e-12)
        assert got.eccentricity == pytest.approx(want.eccentricity, rel=1e-9, abs=1e-9)
        assert image.regionprops(mask=mask_l).moments['mu11'] == pytest.approx(want.moments['mu11'], rel=1e-12, abs=1e-14 * mom['m20'] + 1e-12)
    # the reference's bounding box of a mask with several regions: first block of rows / columns
    assert regions.find_bounding_box(m) == ops.find_bounding_box(m)
    two = np.zeros((40, 60), np.uint8)
    two[3:9, 30:41] = 1
    two[20:30, 5:12] = 1
    assert regions.find_bounding_box(two) == ops.find_bounding_box(two) == (5, 3, 7, 6)
    with pytest.raises(IndexError):
        regions.find_bounding_box(np.zeros((8, 8), np.uint8))
    # more regions than the table has rows: the helper runs again with a table of the exact size
    many = regions.region_stats((np.indices((16, 64)).sum(0) & 1), max_regions=16)
    assert len(many) == 512 and all(r['area'] == 1 for r in many)


def test_synthetic_video_source():
    mods()
    from video_analysis_b200.synth import VideoSynthetic
    v = VideoSynthetic((160, 120), 20, seed=3, n_blobs=5)
    exp = synth.make_frames(3, 0, 20, 160, 120, 5)
    assert np.array_equal(v.get_frame(7), exp[7]) and np.array_equal(v[-1], exp[19])
    assert np.array_equal(np.stack(list(v[4:9])), exp[4:9])


def test_full_size_properties_1080p():
    """ config-2 size: properties that need no CPU oracle at full size """
    mods()
    import torch
    from video_analysis_b200 import synth as dsynth
    from video_analysis_b200.chain import SegmentChain
    from video_analysis_b200.device import get_runtime
    rt = get_runtime(0)
    W, H, B = 1920, 1080, 16
    rgb = dsynth.generate(rt, 0, 0, B, W, H)
    ch = SegmentChain((W, H), batch=B)
    mask, morph, blur = rt.empty_bits(B, H, W), rt.empty_bits(B, H, W), rt.empty_u8(B, H, W)
    labels, counts = ch.run_device(rgb, blur=blur, mask=mask, morph=morph)
    # fused and unfused paths agree
    ch2 = SegmentChain((W, H), batch=B, fuse=False)
    labels2, counts2 = ch2.run_device(rgb)
    assert torch.equal(labels.t, labels2.t) and torch.equal(counts, counts2)
    lab = labels.t[:, :, :W]
    fg = rt.unpack_bits(morph).t[:, :, :W] > 0
    assert torch.equal(lab > 0, fg)                                        # labels cover exactly the mask
    assert torch.equal(lab.reshape(B, -1).max(1).values.int(), counts)     # labels are 1..n
    # opening is anti-extensive and idempotent
    m_u8, o_u8 = rt.unpack_bits(mask).t, rt.unpack_bits(morph).t
    assert bool((o_u8 <= m_u8).all())
    assert torch.equal(rt.morph(morph, 'open').t, morph.t)
    # raster-order numbering: first occurrences of 1..n are increasing (frame 5)
    flat = lab[5].reshape(-1)
    n = int(counts[5])
    first = torch.stack([(flat == k).nonzero()[0, 0] for k in range(1, n + 1)]) if n else torch.zeros(0)
    assert n > 0 and bool((first[1:] > first[:-1]).all())
    # one frame of the full-size batch against the oracle
    f5 = rgb.t[5].cpu().numpy().reshape(H, W, 3)
    assert np.array_equal(blur.t[5, :, :W].cpu().numpy(), ops.blur(ops.mono(f5), 2))
    rl, rn = ops.label(o_u8[5, :, :W].cpu().numpy())
    assert rn == n and np.array_equal(lab[5].cpu().numpy(), rl)


def test_config4_multi_stream_crop_mask_threshold_label():
    """ BASELINE.json configs[3] at reduced stream count: independent camera streams batched on the
    leading axis, per-stream crop rectangle and static mask, threshold, label (no background model) """
    F, VideoMemory = mods()
    W, H, T = 1280, 720, 6
    rng = np.random.default_rng(7)
    for stream in range(3):
        frames = synth.make_frames(10 + stream, 0, T, W, H, 5)
        rect = (int(rng.integers(0, 200)), int(rng.integers(0, 100)), 640 + 16 * stream, 360 + 8 * stream)
        m = np.zeros((rect[3], rect[2]), np.uint8)
        m[20:-30, 40:-10] = 1
        chain = F.FilterLabel(F.FilterThreshold(F.FilterApplyMask(
            F.FilterMonochrome(F.FilterCrop(VideoMemory(frames, copy_data=False), rect, batch=4)), m), 110))
        got = np.stack(list(chain))
        for t in range(T):
            g = ops.apply_mask(ops.mono(ops.crop(frames[t], rect)), m)
            lab, n = ops.label(g > 110)
            assert chain.num_features[t] == n and np.array_equal(got[t], lab), (stream, t)


def test_config5_stencil_heavy_chain_1080p():
    """ BASELINE.json configs[4]: blur sigma=15 (91 taps), threshold, 7x7 close then open, resize 0.5x,
    labelling of large regions -- two 1080p frames against the oracle """
    F, VideoMemory = mods()
    frames = synth.make_frames(2, 0, 2, 1920, 1080, 12)
    v = VideoMemory(frames, copy_data=False)
    blur = F.FilterBlur(F.FilterMonochrome(v, batch=2), sigma=15)
    half = F.FilterResize(blur, 0.5)
    chain = F.FilterLabel(F.FilterMorphology(F.FilterMorphology(F.FilterThreshold(half, 95), 'close', 'rect', 7), 'open', 'rect', 7))
    got = np.stack(list(chain))
    for t in range(2):
        b = ops.resize(ops.blur(ops.mono(frames[t]), 15), 0.5)
        mask = np.where(b > 95, 255, 0).astype(np.uint8)
        mo = ops.morph(ops.morph(mask, 'close', 'rect', 7), 'open', 'rect', 7)
        lab, n = ops.label(mo)
        assert n >= 1 and chain.num_features[t] == n
        assert np.array_equal(got[t], lab)
    assert np.array_equal(np.stack(list(blur)), np.stack([ops.blur(ops.mono(f), 15) for f in frames]))


def test_two_stream_pipelined_chain_equals_single_stream(frames, ref):
    mods()
    import torch
    from video_analysis_b200.chain import SegmentChain
    from video_analysis_b200.device import get_runtime
    rt = get_runtime(0)
    B = 8
    ch = SegmentChain((320, 240), batch=B)
    outs, cnts = [], []
    for a in range(0, 40, B):
        rgb = rt.upload(frames[a:a + B])
        lab, cnt = rt.empty_i32(B, 240, 320), torch.empty((B,), dtype=torch.int32, device=rt.device)
        ch.run_device_pipelined(rgb, lab, cnt)
        outs.append(lab); cnts.append(cnt)
    ch.pipeline_sync()
    torch.cuda.synchronize()
    got = np.concatenate([o.t[:, :, :320].cpu().numpy() for o in outs])
    assert np.array_equal(got, ref['labels'])
    assert np.array_equal(np.concatenate([c.cpu().numpy() for c in cnts]), ref['counts'])
    assert np.array_equal(ch.background.view(np.uint32), ref['bg'].view(np.uint32))


def test_remaining_filter_classes_and_temporal_statistics(frames):
    F, VideoMemory = mods()
    from video_analysis_b200.analysis import video as vstat
    v = VideoMemory(frames[:11], copy_data=False)
    mono = np.stack([ops.mono(f) for f in frames[:11]])
    # FilterNormalize: bounds from the first frame / explicit bounds
    n1 = F.FilterNormalize(F.FilterMonochrome(v, batch=4))
    assert np.array_equal(np.stack(list(n1)), ops.normalize(list(mono)))
    n2 = F.FilterNormalize(F.FilterMonochrome(v, batch=4), vmin=70, vmax=180)
    assert np.array_equal(np.stack(list(n2)), ops.normalize(list(mono), 70, 180))
    # FilterRotate
    for angle in (0, 90, 180, 270, 450):
        r = F.FilterRotate(F.FilterMonochrome(v, batch=4), angle)
        exp = np.stack([ops.rotate(f, angle) for f in mono])
        assert r.shape == exp.shape and np.array_equal(np.stack(list(r)), exp), angle
    rc = F.FilterRotate(F.FilterCrop(v, (8, 4, 100, 60), batch=4), 90)
    assert np.array_equal(np.stack(list(rc)), np.stack([ops.rotate(ops.crop(f, (8, 4, 100, 60)), 90) for f in frames[:11]]))
    with pytest.raises(ValueError):
        F.FilterRotate(v, 45)
    # FilterTimeDifference
    td = F.FilterTimeDifference(F.FilterMonochrome(v, batch=4), batch=4)
    exp = np.stack([ops.time_difference(mono[t + 1], mono[t]) for t in range(10)])
    got = np.stack(list(td))
    assert len(td) == 10 and got.dtype == np.int16 and np.array_equal(got, exp)
    assert np.array_equal(td.get_frame(3), exp[3]) and np.array_equal(td[-1], exp[9])
    # temporal statistics in float64, bit for bit
    mv = VideoMemory(mono, copy_data=False)
    assert np.array_equal(vstat.measure_mean(mv, batch=4).view(np.uint64), ops.measure_mean(list(mono)).view(np.uint64))
    m, s = vstat.measure_mean_std(mv, batch=4)
    mr, sr = ops.measure_mean_std(list(mono))
    assert np.array_equal(m.view(np.uint64), mr.view(np.uint64)) and np.array_equal(s.view(np.uint64), sr.view(np.uint64))
    acc = vstat.reduce_video(mv, lambda f, r: np.maximum(f, r))
    assert np.array_equal(acc, mono.max(0))


def test_raw_stream_ingest_and_annotated_egress(frames, ref):
    """ a raw-video byte stream read into the page-locked ring feeds the chain, and the
    annotated frames (mask highlighted on the device) leave through a raw-stream writer """
    import io
    F, VideoMemory = mods()
    from video_analysis_b200.chain import SegmentChain
    from video_analysis_b200.io.pipe import VideoRawStream, RawStreamWriter
    raw = frames.tobytes()
    # the filter classes over a stream source: same labels as from memory
    v = VideoRawStream(io.BytesIO(raw), (320, 240), len(frames), ring_frames=96)
    full = F.FilterLabel(F.FilterMorphology(
        F.FilterBackgroundMask(F.FilterBlur(F.FilterMonochrome(v, batch=8), 2)), 'open', 'rect', 3))
    assert np.array_equal(np.stack(list(full)), ref['labels'])
    v.close()
    # the pipelined chain straight from the ring (blocks shorter than the chain's batch)
    v = VideoRawStream(io.BytesIO(raw), (320, 240), len(frames) + 1, ring_frames=96)      # length over-estimated
    labels, counts = SegmentChain((320, 240), batch=16).process(v)
    assert np.array_equal(labels, ref['labels']) and list(counts) == list(ref['counts'])
    v.close()
    # annotated egress
    sink = io.BytesIO()
    v = VideoRawStream(io.BytesIO(raw), (320, 240), len(frames), ring_frames=96)
    with RawStreamWriter(sink, (320, 240)) as wr:
        SegmentChain((320, 240), batch=8).annotate(v, wr, channel='red', strength=100)
        assert wr.frames_written == len(frames)
    got = np.frombuffer(sink.getvalue(), np.uint8).reshape(frames.shape)
    exp = np.stack([ops.highlight_mask(f, m, 'red', 100) for f, m in zip(frames, ref['morph'])])
    assert np.array_equal(got, exp)
    marked = SegmentChain((320, 240), batch=16).annotate(frames[:10])
    assert np.array_equal(marked, np.stack([ops.highlight_mask(f, m) for f, m in zip(frames[:10], ref['morph'][:10])]))


def test_config2_full_size_properties_10k_frames():
    """ BASELINE.json configs[1] at its full size (1080p, 10 000 frames resident in HBM): tests/fullsize_check.py """
    mods()
    import json
    import os
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fullsize_check.py')
    p = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert out['frames'] == 9984 and all(v for v in out.values() if isinstance(v, bool)), out


def test_config4_streams_batched_per_launch():
    """ BASELINE.json configs[3]: the streams advance in lock step, one launch set per time step over all of them
    (per-stream crop position and mask) -- equal to the per-stream filter stacks of the oracle """
    F, VideoMemory = mods()
    from video_analysis_b200.streams import MultiStreamSegmenter
    W, H, T, S = 1280, 720, 4, 5
    rng = np.random.default_rng(11)
    vids = [synth.make_frames(20 + s, 0, T, W, H, 5) for s in range(S)]
    rects = [(int(rng.integers(0, 600)), int(rng.integers(0, 300)), 642, 363) for _ in range(S)]
    rects[1] = (-642, -363, 642, 363)                                     # negative: counted from the far edge
    masks = np.zeros((S, 363, 642), np.uint8)
    for s in range(S):
        masks[s, 10 + 5 * s:-20, 30:-10 - 7 * s] = 1
    seg = MultiStreamSegmenter([VideoMemory(v, copy_data=False) for v in vids], rects, masks, threshold=110)
    n_steps = 0
    for t, (labels, counts) in enumerate(seg):
        n_steps += 1
        for s in range(S):
            g = ops.apply_mask(ops.mono(ops.crop(vids[s][t], ops.crop_rect((W, H), rects[s]))), masks[s])
            lab, n = ops.label(g > 110)
            assert counts[s] == n and np.array_equal(labels[s], lab), (t, s)
    assert n_steps == T
    # dense copies instead of chunk egress, a shallower ring, one gather thread: same results
    ref_steps = [(l.copy(), c.copy()) for l, c in MultiStreamSegmenter([VideoMemory(v, copy_data=False) for v in vids], rects, masks,
                                                                       threshold=110)]
    alt = MultiStreamSegmenter([VideoMemory(v, copy_data=False) for v in vids], rects, masks, threshold=110, sparse_egress=False,
                               depth=2, gather_threads=1)
    for (la, ca), (lb, cb) in zip(ref_steps, [(l.copy(), c.copy()) for l, c in alt]):
        assert np.array_equal(la, lb) and np.array_equal(ca, cb)
    assert len(ref_steps) == T
    one = MultiStreamSegmenter([VideoMemory(v, copy_data=False) for v in vids], rects, masks[0], threshold=90, mono_mode='green',
                               connectivity=8)
    labels, counts = next(iter(one))
    for s in range(S):
        g = ops.apply_mask(ops.mono(ops.crop(vids[s][0], ops.crop_rect((W, H), rects[s])), 'green'), masks[0])
        lab, n = ops.label(g > 90, 8)
        assert counts[s] == n and np.array_equal(labels[s], lab)
    # crop width a multiple of 32: the front of the chain is one fused pass; same results as the separate kernels
    rects32 = [(r[0] % 600, r[1] % 300, 640, 352) for r in rects]
    outs = []
    for fused in (True, False):
        sg = MultiStreamSegmenter([VideoMemory(v, copy_data=False) for v in vids], rects32, masks[:, :352, :640], threshold=110,
                                  fused=fused)
        assert sg._fused == fused
        outs.append([(lab.copy(), cnt.copy()) for lab, cnt in sg])
    for (la, ca), (lb, cb) in zip(*outs):
        assert np.array_equal(la, lb) and np.array_equal(ca, cb)
    for s in range(S):
        g = ops.apply_mask(ops.mono(ops.crop(vids[s][2], ops.crop_rect((W, H), rects32[s]))), masks[s, :352, :640])
        lab, n = ops.label(g > 110)
        assert outs[0][2][1][s] == n and np.array_equal(outs[0][2][0][s], lab)
    with pytest.raises(ValueError):
        MultiStreamSegmenter([VideoMemory(v, copy_data=False) for v in vids], rects[:-1] + [(0, 0, 100, 100)])


def test_randomised_differential_parity():
    """ tests/fuzz_parity.py for 20 seconds: random sizes / parameters / input statistics, every stage of the chain, the
    resize modes, labelling and the multi-stream front against the oracle, bit for bit """
    mods()
    import json
    import os
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fuzz_parity.py')
    p = subprocess.run([sys.executable, script, '--cases', '100000', '--seconds', '20', '--seed', '7'],
                       capture_output=True, text=True, timeout=600)
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert p.returncode == 0 and not out['failures'] and sum(out['cases_run'].values()) > 200, out


def test_segment_chain_int16_labels(frames, ref):
    """ label_dtype=np.int16 (ndimage.label(..., output=np.int16)): same labels in half the bytes, through the host pipeline,
    the device call and the pipelined device call """
    mods()
    import torch
    from video_analysis_b200.chain import SegmentChain
    from video_analysis_b200.device import get_runtime
    seg = SegmentChain((320, 240), batch=16, label_dtype=np.int16)
    labels, counts = seg.process(frames)
    assert labels.dtype == np.int16 and np.array_equal(labels, ref['labels']) and list(counts) == list(ref['counts'])
    assert np.array_equal(seg.background.view(np.uint32), ref['bg'].view(np.uint32))
    rt = get_runtime()
    for piped in (False, True):
        ch = SegmentChain((320, 240), batch=8, label_dtype=np.int16)
        outs = []
        for a in range(0, len(frames), 8):
            rgb = rt.upload(frames[a:a + 8])
            lab = rt.empty_i16(rgb.n, 240, 320)
            cnt = torch.empty((rgb.n,), dtype=torch.int32, device=rt.device)
            if piped:
                ch.run_device_pipelined(rgb, lab, cnt)
                ch.pipeline_sync()
            else:
                ch.run_device(rgb, lab, cnt)
            torch.cuda.synchronize()
            outs.append(lab.t[:, :, :320].cpu().numpy())
        assert np.array_equal(np.concatenate(outs), ref['labels']), piped
    with pytest.raises(ValueError):
        SegmentChain((320, 240), label_dtype=np.int8)
    F, VideoMemory = mods()
    full = F.FilterLabel(F.FilterMorphology(F.FilterBackgroundMask(
        F.FilterBlur(F.FilterMonochrome(VideoMemory(frames, copy_data=False), batch=16), 2)), 'open', 'rect', 3), dtype=np.int16)
    got = np.stack(list(full))
    assert got.dtype == np.int16 and np.array_equal(got, ref['labels']) and full.num_features == list(ref['counts'])


@pytest.mark.parametrize('batch,ring', [(128, 384), (50, 384), (32, 100)])
def test_filter_chain_over_a_stream_with_short_blocks(batch, ring):
    """ a raw stream hands out blocks shorter than the filter's batch (capped at ring // (hold + 2), cut at
    the ring's wrap): every frame must still come out (round-1 advisor finding: the chain stopped at the
    first short block) """
    import io
    F, VideoMemory = mods()
    from video_analysis_b200.io.pipe import VideoRawStream
    fr = synth.make_frames(2, 0, 500, 64, 48, 3)
    v = VideoRawStream(io.BytesIO(fr.tobytes()), (64, 48), len(fr), ring_frames=ring)
    got = np.stack([f.copy() for f in F.FilterBlur(F.FilterMonochrome(v, batch=batch), 2)])
    assert got.shape[0] == 500
    assert np.array_equal(got, np.stack([ops.blur(ops.mono(f), 2) for f in fr]))
    v.close()
    # single-frame collectors over a stream: FilterTimeDifference
    v = VideoRawStream(io.BytesIO(fr[:120].tobytes()), (64, 48), 120, ring_frames=12)
    td = F.FilterTimeDifference(F.FilterFunction(v, lambda f: f[:, :, 1]), batch=40)
    mono = fr[:120, :, :, 1]
    assert np.array_equal(np.stack(list(td)), mono[1:].astype(np.int16) - mono[:-1])
    v.close()


def test_fused_crop_monochrome_checks_the_rectangle(frames):
    """ `_check_coordinate` validates each value alone: left + width may leave the frame, and size_alignment may
    enlarge it.  The constructor rejects such rectangles (IndexError, like a bad coordinate); the pointer-offset
    paths of the runtime -- fused crop -> monochrome included -- refuse them too (ValueError: VideoIterator would
    turn an IndexError into the end of the video) """
    F, VideoMemory = mods()
    v = VideoMemory(frames[:4], copy_data=False)
    for kw in (dict(rect=(300, 10, 100, 50)), dict(rect=(10, 200, 50, 100)),
               dict(rect=(250, 10, 69, 50), size_alignment=8)):          # 69 -> 72: 250 + 72 > 320
        with pytest.raises(IndexError):
            F.FilterCrop(v, batch=2, **kw)
    with pytest.raises(IndexError):
        F.FilterCrop(F.FilterCrop(v, rect=(200, 100, 100, 100)), rect=(70, 50, 60, 40))     # nested: 200 + 70 + 60 > 320
    ok = F.FilterCrop(v, rect=(250, 10, 67, 50), size_alignment=8, batch=2)                   # 67 -> 64: fits
    assert ok.rect == (250, 10, 64, 48) and np.stack(list(F.FilterMonochrome(ok))).shape == (4, 48, 64)
    # a source whose frames are smaller than its metadata: the runtime check fires on both paths
    liar = VideoMemory(frames[:4, :100, :100].copy(), copy_data=False)
    liar.size = (320, 240)
    for chain in (F.FilterMonochrome(F.FilterCrop(liar, rect=(90, 10, 50, 50), batch=2)),
                  F.FilterCrop(liar, rect=(90, 10, 50, 50), batch=2)):
        with pytest.raises(ValueError):
            list(chain)
    rt = F.get_runtime()
    dev = rt.upload(frames[:2])
    with pytest.raises(ValueError):
        rt.luma(dev, rect=(300, 0, 100, 10))
    with pytest.raises(ValueError):
        rt.crop(dev, (0, 230, 10, 20))


def test_interleaved_label_chains_share_the_runtime_safely(frames, ref):
    """ two label chains iterated in lock step (zip) run on different streams over ONE ctx: the library orders
    the uses of its union-find scratch with an event, so neither corrupts the other; the analysis helpers
    may be called in between """
    F, VideoMemory = mods()
    from video_analysis_b200.analysis import regions
    v1 = VideoMemory(frames, copy_data=False)
    v2 = VideoMemory(frames[::-1].copy(), copy_data=False)
    ref2 = ops.chain(frames[::-1])

    def chain(v, batch):
        return F.FilterLabel(F.FilterMorphology(F.FilterBackgroundMask(
            F.FilterBlur(F.FilterMonochrome(v, batch=batch), 2)), 'open', 'rect', 3))
    for rep in range(3):
        a, b = chain(v1, 5), chain(v2, 7)
        for t, (la, lb) in enumerate(zip(a, b)):
            assert np.array_equal(la, ref['labels'][t]), (rep, t)
            assert np.array_equal(lb, ref2['labels'][t]), (rep, t)
            if t % 9 == 0:
                lab, n = regions.label(ref['morph'][t])
                assert n == ref['counts'][t] and np.array_equal(lab, ref['labels'][t])


def test_region_helpers_beyond_the_table_size_and_value_weighting():
    """ more regions than the default table (4096 rows) and cv2.moments' weighting by the pixel value """
    mods()
    import cv2
    from video_analysis_b200.analysis import image, regions
    m = np.zeros((200, 300), np.uint8)
    m[::2, ::2] = 1                                                 # 15 000 single-pixel regions
    regs = regions.region_stats(m)
    assert len(regs) == 15000 and all(r['area'] == 1 for r in regs)
    assert regions.find_bounding_box(np.pad(np.ones((5, 9), np.uint8), 3)) == (3, 3, 9, 5)
    blob = np.zeros((60, 80), np.uint8)
    blob[10:30, 20:50] = 255
    blob[40:45, 5:70] = 255
    for mask in (blob, blob // 255, blob.astype(bool)):
        want = cv2.moments(mask.astype(np.uint8))
        p = image.regionprops(mask)
        assert p.area == want['m00']
        for k in ('m10', 'm01', 'm20', 'm11', 'm02'):
            assert p.moments[k] == want[k]
        for k in ('mu20', 'mu11', 'mu02'):
            assert abs(p.moments[k] - want[k]) <= 1e-9 * abs(want[k])


@pytest.mark.parametrize('world', [2, 3, 8])
def test_sharded_equals_sequential(world):
    """ the frame-sharded chain (what every N > 1 bench line runs) with R ranks emulated on ONE GPU --
    ShardedSegmentChain.pass1 for every rank, the partial states stacked in place of the NCCL all-gather, then
    pass2 for every rank -- against the sequential chain over the whole video: labels and counts identical,
    rank 0 bit-identical in the background too, later ranks within 1e-5 relative (re-association of the
    recurrence).  Shards are ragged (T not divisible by R or by the batch) and longer / shorter than the tail. """
    import torch
    mods()
    from video_analysis_b200 import synth as dsynth
    from video_analysis_b200.chain import SegmentChain
    from video_analysis_b200.device import get_runtime
    from video_analysis_b200.parallel import ShardedSegmentChain, shard_range
    rt = get_runtime()
    Wd, Hd, B = 320, 240, 16
    alpha = 0.2                                   # tail = 94 frames: shards of 8 ranks are shorter, of 2 ranks longer
    T = {2: 300, 3: 211, 8: 400}[world]

    def batches_of(a, b):
        out = []
        for t0 in range(a, b, B):
            n = min(B, b - t0)
            out.append(dsynth.generate(rt, 7, t0, n, Wd, Hd))
        return out

    # sequential run of the whole video
    seq = SegmentChain((Wd, Hd), batch=B, alpha=alpha)
    seq_labels, seq_counts, seq_bg_at = [], [], {}
    bounds = [shard_range(T, r, world) for r in range(world)]
    starts = {a for a, _ in bounds}
    for t0 in range(0, T, B):
        # batches are cut at shard boundaries so that the state before every shard can be recorded
        cuts = sorted({t0, min(t0 + B, T)} | {s for s in starts if t0 < s < min(t0 + B, T)})
        for a, b in zip(cuts[:-1], cuts[1:]):
            if a in starts and a > 0:
                torch.cuda.synchronize()
                seq_bg_at[a] = seq._bg.clone()
            lab, cnt = seq.run_device(dsynth.generate(rt, 7, a, b - a, Wd, Hd))
            seq_labels.append(lab.t[:, :, :Wd].clone())
            seq_counts.append(cnt.clone())
    seq_labels, seq_counts = torch.cat(seq_labels), torch.cat(seq_counts)
    final_bg = seq._bg.clone()

    # the ranks, emulated one after the other on this GPU
    ranks = []
    for r in range(world):
        a, b = bounds[r]
        ch = SegmentChain((Wd, Hd), batch=B, alpha=alpha)
        sh = ShardedSegmentChain(ch, rank=r, world=world)
        rgb = batches_of(a, b)
        S = sh.pass1(rgb)
        torch.cuda.synchronize()
        ranks.append((ch, sh, rgb, S.clone()))
    gathered = torch.stack([S for _, _, _, S in ranks])          # what all_gather_into_tensor delivers
    counts_per_rank = [b - a for a, b in bounds]
    for r, (ch, sh, rgb, _) in enumerate(ranks):
        a, b = bounds[r]
        outs = [rt.empty_i32(B, Hd, Wd) for _ in rgb]
        cnts = torch.zeros((len(rgb), B), dtype=torch.int32, device=rt.device)
        head = sh.preblur(rgb)
        sh.pass2(rgb, outs, cnts, gathered.clone(), counts_per_rank, None, head)
        torch.cuda.synchronize()
        got = torch.cat([o.t[:x.n, :, :Wd] for o, x in zip(outs, rgb)])
        gotc = torch.cat([cnts[i, :x.n] for i, x in enumerate(rgb)])
        assert torch.equal(gotc, seq_counts[a:b]), 'counts of rank %d' % r
        assert torch.equal(got, seq_labels[a:b]), 'labels of rank %d' % r
        if r + 1 < world:
            want = seq_bg_at[bounds[r + 1][0]]
        else:
            want = final_bg
        if r == 0:
            assert torch.equal(ch._bg[:, :Wd], want[:, :Wd])     # rank 0 is the sequential model itself
        else:
            rel = ((ch._bg[:, :Wd] - want[:, :Wd]).abs() / want[:, :Wd].abs().clamp(min=1)).max()
            assert float(rel) < 1e-5, float(rel)


def test_sparse_egress_gives_the_dense_label_images(frames, ref):
    """ SegmentChain.process_blocks brings the label images to the host as their non-empty chunks (default) or as a
    dense copy: same arrays, same counts, across buffer reuse (more blocks than ring slots), short last blocks and a
    frame full of foreground """
    mods()
    from video_analysis_b200.chain import SegmentChain
    for batch in (16, 7):
        a = SegmentChain((320, 240), batch=batch)
        b = SegmentChain((320, 240), batch=batch, sparse_egress=False)
        assert a._sparse() and not b._sparse()
        la, ca = a.process(frames)
        lb, cb = b.process(frames)
        assert np.array_equal(la, ref['labels']) and np.array_equal(lb, ref['labels'])
        assert list(ca) == list(ref['counts']) and list(cb) == list(ref['counts'])
        assert a.egress_bytes < la.nbytes // 4
    # noisy input: nearly every chunk is non-empty, then quiet again in the same buffers
    rng = np.random.default_rng(3)
    noisy = rng.integers(0, 256, (20, 240, 320, 3), dtype=np.uint8)
    calm = frames[:20]
    c = SegmentChain((320, 240), batch=8, sigma=0.5, threshold=10.0, morph_op=None)
    for video in (noisy, calm, noisy):
        want = ops.chain(video, sigma=0.5, alpha=0.05, thr=10.0, morph_op=None)
        c.reset()
        got, cnt = c.process(video)
        assert np.array_equal(got, want['labels']) and list(cnt) == list(want['counts'])


def test_ctx_grows_under_a_live_chain(frames, ref):
    """ the va_ctx of a GPU is shared by every chain of the process: a larger chain created while a smaller one is in
    the middle of its video grows the capacity in place (va_reserve) and neither loses its state """
    mods()
    from video_analysis_b200.chain import SegmentChain
    from video_analysis_b200.device import get_runtime
    small = SegmentChain((320, 240), batch=8)
    l1, c1 = small.process(frames[:16])
    rt = get_runtime()
    cap0 = rt._cap
    big_frames = synth.make_frames(4, 0, 6, cap0[0] + 64, cap0[1] + 32, 3)
    big = SegmentChain((cap0[0] + 64, cap0[1] + 32), batch=max(cap0[2], 8) + 3)         # grows every dimension of the ctx
    assert rt._cap[0] > cap0[0] and rt._cap[1] > cap0[1] and rt._cap[2] > cap0[2]
    lb, cb = big.process(big_frames)
    want = ops.chain(big_frames)
    assert np.array_equal(lb, want['labels']) and list(cb) == list(want['counts'])
    l2, c2 = small.process(frames[16:])                                               # continues from its own background
    assert np.array_equal(np.concatenate([l1, l2]), ref['labels'])
    assert list(np.concatenate([c1, c2])) == list(ref['counts'])
