"""
The CUDA path behind the reference's filter classes against tests/golden/golden_ref.npz --
vectors produced by THE REFERENCE'S OWN CODE (tests/golden/make_golden_ref.py ran the Python-3
conversion of /root/reference/video/filters.py, io/base.py, io/memory.py, analysis/*.py).
Same constructor calls as in the generator, bit-exact results required.
"""

import itertools
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'golden_ref.npz'))


def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from video_analysis_b200 import filters
    from video_analysis_b200.io.memory import VideoMemory
    return filters, VideoMemory


def stack(video):
    return np.stack([np.array(f) for f in video])


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)


def test_monochrome_and_crop_equal_the_reference():
    F, VM = mods()
    col, rag = GOLD['f_col'], GOLD['f_rag']
    assert same(stack(F.FilterMonochrome(VM(col), batch=4)), GOLD['f_mono_mean'])
    assert same(stack(F.FilterMonochrome(VM(col), 'green', batch=4)), GOLD['f_mono_green'])
    assert same(stack(F.FilterMonochrome(VM(rag), 'R', batch=2)), GOLD['f_mono_r'])
    assert same(stack(F.FilterCrop(VM(col), rect=(5, 7, 40, 30), batch=4)), GOLD['f_crop_rect'])
    assert same(stack(F.FilterCrop(VM(col), rect=(0.25, 0.5, 0.5, 0.25), batch=4)), GOLD['f_crop_frac'])
    assert same(stack(F.FilterCrop(VM(col), rect=(-30, -20, 17, 11), batch=4)), GOLD['f_crop_neg'])
    assert same(stack(F.FilterCrop(VM(rag), region='lower right', batch=4)), GOLD['f_crop_region'])
    assert same(stack(F.FilterCrop(VM(col), rect=(3, 2, 33, 21), color_channel='red', batch=4)), GOLD['f_crop_chan'])
    assert same(stack(F.FilterCrop(VM(col), rect=(1, 1, 37, 29), size_alignment=4, batch=4)), GOLD['f_crop_align'])
    nested = F.FilterCrop(F.FilterCrop(VM(col), rect=(4, 6, 50, 40)), rect=(3, 5, 20, 10))
    assert tuple(nested.rect) == tuple(GOLD['p_crop_nested_rect'])
    assert same(stack(nested), GOLD['f_crop_nested'])


def test_blur_equals_the_reference():
    F, VM = mods()
    col, rag = GOLD['f_col'], GOLD['f_rag']
    for s in (0.5, 1, 2, 3, 5):
        got = stack(F.FilterBlur(F.FilterMonochrome(VM(col), batch=4), sigma=s))
        assert same(got, GOLD['f_blur_%g' % s]), s
    assert same(stack(F.FilterBlur(F.FilterMonochrome(VM(rag), batch=3))), GOLD['f_blur_default_rag'])
    assert same(stack(F.FilterBlur(VM(col), sigma=1.5, batch=4)), GOLD['f_blur_color'])


def test_resize_equals_the_reference():
    F, VM = mods()
    col, rag = GOLD['f_col'], GOLD['f_rag']

    def mono(a):
        return F.FilterMonochrome(VM(a), batch=4)

    assert same(stack(F.FilterResize(mono(col), 0.5)), GOLD['f_resize_half'])
    assert same(stack(F.FilterResize(mono(col), 1 / 3., even_dimensions=True)), GOLD['f_resize_third_even'])
    assert same(stack(F.FilterResize(mono(col), (40, 27))), GOLD['f_resize_size'])
    up = stack(F.FilterResize(mono(rag), 1.5))                  # 'auto' enlarging: INTER_CUBIC
    # the cv2 wheel sends INTER_CUBIC to Intel IPP (unpublished rounding): 1 LSB, as in test_kernels.py
    assert up.shape == GOLD['f_resize_up_auto'].shape
    assert np.abs(up.astype(np.int16) - GOLD['f_resize_up_auto']).max() <= 1
    for interp in ('linear', 'area', 'lanczos'):
        got = stack(F.FilterResize(mono(col), (45, 31), interpolation=interp))
        assert same(got, GOLD['f_resize_%s' % interp]), interp
    cub = stack(F.FilterResize(mono(col), (45, 31), interpolation='cubic'))
    assert np.abs(cub.astype(np.int16) - GOLD['f_resize_cubic']).max() <= 1
    assert same(stack(F.FilterResize(VM(col), 0.5, batch=4)), GOLD['f_resize_color_half'])
    assert same(stack(F.FilterResize(mono(col), 1)), GOLD['f_resize_same'])


def test_remaining_filter_classes_equal_the_reference():
    F, VM = mods()
    col, rag = GOLD['f_col'], GOLD['f_rag']
    for a in (90, 180, 270, 450):
        assert same(stack(F.FilterRotate(F.FilterMonochrome(VM(rag), batch=2), a)), GOLD['f_rot_%d' % a]), a
    assert same(stack(F.FilterNormalize(F.FilterMonochrome(VM(col), batch=4), 70, 180)), GOLD['f_norm'])
    assert same(stack(F.FilterNormalize(F.FilterMonochrome(VM(col), batch=4))), GOLD['f_norm_auto'])
    rep = F.FilterReplicate(F.FilterMonochrome(VM(rag), batch=2), 3)
    assert same(np.stack([np.array(f) for f in itertools.islice(rep, rep.frame_count)]), GOLD['f_replicate'])
    assert same(stack(F.FilterDropFrames(F.FilterMonochrome(VM(col), batch=4), 2)), GOLD['f_drop'])
    td = F.FilterTimeDifference(F.FilterMonochrome(VM(col), batch=4))
    assert same(np.stack([td.get_frame(i) for i in range(td.frame_count)]), GOLD['f_timediff_get'])
    assert same(td.get_frame(-1), GOLD['f_timediff_get_neg'])
    chain = F.FilterResize(F.FilterBlur(F.FilterMonochrome(F.FilterCrop(VM(col), rect=(8, 4, 48, 40), batch=4)), 2), 0.5)
    assert str(chain) == str(GOLD['p_chain_str'])
    assert same(chain.copy().data, GOLD['f_chain_copy'])
    sl = F.FilterMonochrome(F.FilterCrop(VM(col), rect=(1, 2, 30, 20), batch=4))[1:5:2]
    assert same(stack(sl), GOLD['p_slice'])
    assert same(stack(F.FilterMonochrome(VM(col), batch=4)[4:0:-1]), GOLD['p_slice_rev'])
    assert same(F.FilterMonochrome(VM(col), batch=4)[-1], GOLD['p_neg_index'])


def test_region_helpers_equal_the_reference():
    mods()
    from video_analysis_b200.analysis import image, regions, video
    from video_analysis_b200 import filters as F
    from video_analysis_b200.io.memory import VideoMemory as VM
    masks = GOLD['a_masks']
    for i, m in enumerate(masks):
        big, area = regions.get_largest_region(m, ret_area=True)
        assert same(big, GOLD['a_largest'][i]) and int(area) == int(GOLD['a_largest_area'][i])
        assert tuple(regions.find_bounding_box(GOLD['a_largest'][i])) == tuple(GOLD['a_bbox'][i])
        p = image.regionprops(GOLD['a_largest'][i].astype(np.uint8))
        e1, e2 = p.inertia_tensor_eigvals
        got = np.array([p.area, p.centroid[0], p.centroid[1], p.orientation, e1, e2,
                        p.major_axis_length, p.minor_axis_length])
        assert np.allclose(got, GOLD['a_props'][i], rtol=1e-12, atol=1e-9)
    assert same(regions.get_largest_region(masks[0].astype(np.uint8) * 255), GOLD['a_largest_u8'])
    col = GOLD['f_col']
    assert same(video.measure_mean(F.FilterMonochrome(VM(col), batch=4)), GOLD['a_mean'])
    mean, std = video.measure_mean_std(F.FilterMonochrome(VM(col), batch=4))
    assert same(mean, GOLD['a_mean2']) and same(std, GOLD['a_std'])
