"""
The oracle against (a) the committed golden fixtures, (b) the integer restatements the CUDA
kernels implement, and (c) the library semantics SURVEY.md appendix B probed.  CPU only.
"""

import hashlib
import os

import cv2
import numpy as np
import pytest
from scipy import ndimage

from oracle import ops, synth

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'golden.npz'))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_synth_matches_golden_frames():
    assert np.array_equal(synth.make_frames(0, 0, 6, 64, 48, 4), GOLD['s_frames'])
    assert np.array_equal(synth.make_frames(3, 10, 5, 53, 37, 3), GOLD['r_frames'])


def test_synth_is_seekable_and_seeded():
    a = synth.make_frames(1, 0, 4, 40, 30, 2)
    assert np.array_equal(synth.make_frames(1, 2, 2, 40, 30, 2), a[2:])
    assert not np.array_equal(synth.make_frames(2, 0, 1, 40, 30, 2), a[:1])
    assert a.min() >= 52 and a.max() <= 218


@pytest.mark.parametrize('prefix,kw', [
    ('s_', dict(sigma=2, alpha=0.05, thr=25, morph_op='open', morph_ksize=3)),
    ('r_', dict(sigma=3, alpha=0.1, thr=12, morph_op='close', morph_shape='ellipse', morph_ksize=5, connectivity=8)),
])
def test_chain_matches_golden(prefix, kw):
    r = ops.chain(GOLD[prefix + 'frames'], **kw)
    for k in ('mono', 'blur', 'mask', 'morph', 'labels', 'counts'):
        assert np.array_equal(r[k], GOLD[prefix + k]), k
    assert np.array_equal(r['bg'].view(np.uint32), GOLD[prefix + 'bg'].view(np.uint32))


def test_vga_chain_hashes():
    fr = synth.make_frames(0, 0, 12, 640, 480, 8)
    r = ops.chain(fr)
    got = [sha(fr)] + [sha(r[k]) for k in ('mono', 'blur', 'mask', 'morph', 'labels', 'counts')]
    assert got == list(GOLD['vga_hashes'])
    assert r['counts'][1:].min() >= 1


def test_mono_is_integer_division():
    # np.mean(axis=2).astype(u8) == (c0+c1+c2)//3 for every possible sum (video/filters.py:366)
    s = np.arange(766)
    px = np.zeros((766, 1, 3), np.uint8)
    px[:, 0, 0] = np.minimum(s, 255)
    px[:, 0, 1] = np.clip(s - 255, 0, 255)
    px[:, 0, 2] = np.clip(s - 510, 0, 255)
    assert np.array_equal(ops.mono(px)[:, 0], s // 3)
    rng = np.random.default_rng(0)
    f = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    assert np.array_equal(ops.mono(f), f.astype(np.uint32).sum(2) // 3)
    for name, c in (('blue', 0), ('g', 1), ('red', 2)):
        assert np.array_equal(ops.mono(f, name), f[:, :, c])


def test_div3_by_half_precision_fma():
    # the kernels divide the channel sum by 3 with one fp16 fused multiply-add (csrc/va_device.cuh:
    # va_div3_h2): RN_fp16((1024 + s) * 0x3555 + 682.5) == 1024 + s // 3 for every sum s <= 765
    from fractions import Fraction
    c = Fraction(float(np.uint16(0x3555).view(np.float16)))
    b = Fraction(float(np.uint16(0x6155).view(np.float16)))
    for s in range(766):
        v = (1024 + s) * c + b
        f = np.float64(v.numerator) / np.float64(v.denominator)
        assert Fraction(float(f)) == v                      # exact in double, so one rounding to fp16
        assert int(np.float16(f).view(np.uint16)) == 0x6400 + s // 3


@pytest.mark.parametrize('sigma', [0.05, 0.3, 0.5, 1, 2, 3, 5, 15, 21])
@pytest.mark.parametrize('shape', [(37, 53), (120, 160)])
def test_integer_gaussian_equals_cv2(sigma, shape):
    rng = np.random.default_rng(int(sigma * 100))
    a = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(ops.blur(a, sigma), ops.blur_integer(a, sigma))


def test_gaussian_taps_golden():
    for s, row in zip(GOLD['tap_sigmas'], GOLD['taps']):
        k = ops.gauss_kernel_u8(float(s))
        assert np.array_equal(k, row[:len(k)]) and k.sum() == 256
    assert list(ops.gauss_kernel_u8(2)) == [1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1]


def test_gaussian_colour_is_per_channel():
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (30, 40, 3), dtype=np.uint8)
    ref = np.stack([ops.blur(np.ascontiguousarray(a[..., c]), 2) for c in range(3)], axis=2)
    assert np.array_equal(ops.blur(a, 2), ref)


def test_resize_half_is_rounded_mean():
    rng = np.random.default_rng(2)
    a = rng.integers(0, 256, (48, 64), dtype=np.uint8).astype(np.uint32)
    ref = (a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2
    assert np.array_equal(ops.resize(a.astype(np.uint8), 0.5), ref)


def test_label_is_raster_ordered():
    rng = np.random.default_rng(3)
    m = rng.random((60, 80)) < 0.45
    for conn in (4, 8):
        lab, n = ops.label(m, conn)
        first = [np.flatnonzero(lab.ravel() == k)[0] for k in range(1, n + 1)]
        assert first == sorted(first)
        assert lab.dtype == np.int32
    assert np.array_equal(ops.label(m)[0], ops.label(m.astype(np.uint8) * 255)[0])


def test_open_is_dilate_of_erode_with_cv_border():
    rng = np.random.default_rng(4)
    m = ((rng.random((40, 50)) < 0.7) * 255).astype(np.uint8)
    se = ops.structuring_element('rect', 3)
    assert np.array_equal(ops.morph(m, 'open'), cv2.dilate(cv2.erode(m, se), se))
    er = ndimage.binary_erosion(m > 0, np.ones((3, 3)), border_value=1)
    assert np.array_equal(ops.morph(m, 'erode') > 0, er)


def test_crop_rect_rules():
    assert ops.crop_rect((640, 480), rect=(10, 20, 100, 50)) == (10, 20, 100, 50)
    assert ops.crop_rect((640, 480), rect=(0.5, 0.25, 0.25, 0.5)) == (320, 120, 160, 240)
    assert ops.crop_rect((640, 480), rect=(-40, -30, 20, 10)) == (600, 450, 20, 10)
    assert ops.crop_rect((640, 480), region='lower right') == (320, 240, 320, 240)
    assert ops.crop_rect((640, 480), rect=(1, 1, 101, 51), size_alignment=4) == (1, 1, 100, 52)
    with pytest.raises(IndexError):
        ops.crop_rect((640, 480), rect=(0, 0, 640, 480))        # full width is rejected (filters.py:187)


def test_ema_float32_vs_float64():
    fr = [ops.blur(ops.mono(f), 2) for f in synth.make_frames(0, 0, 40, 64, 48, 4)]
    m32, bg32 = ops.background_ema(fr, 0.05, 25)
    m64, bg64 = ops.background_ema(fr, 0.05, 25, dtype=np.float64)
    assert np.allclose(bg32, bg64, rtol=1e-5)
    assert m32[0].max() == 0 and m32.shape == (40, 48, 64)


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(5)
    for w in (1, 31, 32, 33, 64, 100):
        m = ((rng.random((3, 7, w)) < 0.5) * 255).astype(np.uint8)
        p = ops.pack_bits(m)
        assert p.shape == (3, 7, (w + 31) // 32)
        assert np.array_equal(ops.unpack_bits(p, w), m)
        assert (p[..., 0] & 1 == (m[..., 0] != 0)).all()
