"""
Parity of every C-ABI entry point against the oracle (bit-exact unless stated).

Each test runs on the `cuda` backend (marked gpu: the parity gate, through
csrc/libva_b200.so on a B200) and on the `emu` backend (same kernel sources under the CPU
thread-emulation shim, small sizes only -- a logic check that runs without a GPU).
"""

import os

import numpy as np
from scipy import ndimage
import pytest

from oracle import ops, synth
from tests import harness as hz

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'golden.npz'))


@pytest.fixture
def ctx(be):
    big = be.name == 'cuda'
    c = hz.Ctx(be, 4096 if big else 1300, 2304 if big else 256, 16 if big else 8)
    yield c
    c.close()


def rng_frames(seed, shape):
    return np.random.default_rng(seed).integers(0, 256, shape, dtype=np.uint8)


def rmask(seed, shape, p):
    return ((np.random.default_rng(seed).random(shape) < p) * 255).astype(np.uint8)


def sizes(be, small, large):
    return small + (large if be.name == 'cuda' else [])


# ---- K1 -----------------------------------------------------------------------------------------
def test_luma_mean_and_channels(be, ctx):
    for (H, W) in sizes(be, [(13, 37), (24, 64), (9, 130)], [(480, 640), (1080, 1920), (271, 1003)]):
        fr = rng_frames(H * W, (2, H, W, 3))
        ref = np.stack([ops.mono(f) for f in fr])
        assert np.array_equal(hz.luma(ctx, fr), ref)
        assert np.array_equal(hz.luma(ctx, fr, in_pad=5, in_off=3, out_pad=3), ref)
        for c in (0, 1, 2):
            assert np.array_equal(hz.luma(ctx, fr, mode=c), fr[..., c])


def test_luma_every_sum(be, ctx):
    s = np.arange(766)
    fr = np.zeros((1, 1, 768, 3), np.uint8)
    fr[0, 0, :766, 0] = np.minimum(s, 255)
    fr[0, 0, :766, 1] = np.clip(s - 255, 0, 255)
    fr[0, 0, :766, 2] = np.clip(s - 510, 0, 255)
    assert np.array_equal(hz.luma(ctx, fr)[0, 0, :766], s // 3)


def test_crop_by_pointer_offset(be, ctx):
    for (H, W) in sizes(be, [(20, 50)], [(480, 640)]):
        fr = rng_frames(7, (2, H, W, 3))
        for rect in ((3, 2, W - 7, H - 5), (1, 0, 17, 9), (16, 4, 32, 8)):
            assert np.array_equal(hz.crop_luma(ctx, fr, rect), np.stack([ops.mono(ops.crop(f, rect)) for f in fr]))
            assert np.array_equal(hz.crop_luma(ctx, fr, rect, mode=2), np.stack([ops.crop(f, rect, 'red') for f in fr]))
            assert np.array_equal(hz.copy2d(ctx, fr, rect), np.stack([ops.crop(f, rect) for f in fr]))
        mono = fr[..., 0].copy()
        assert np.array_equal(hz.copy2d(ctx, mono, (16, 2, 32, 8)), mono[:, 2:10, 16:48])


def test_luma_every_channel_sum(be, ctx):
    # all 766 sums, at every position of a 4-pixel group and through the fused blur's converter
    s = np.arange(768 * 4) % 766
    px = np.zeros((1, 16, 192, 3), np.uint8)
    px[0, :, :, 0] = np.minimum(s, 255).reshape(16, 192)
    px[0, :, :, 1] = np.clip(s - 255, 0, 255).reshape(16, 192)
    px[0, :, :, 2] = np.clip(s - 510, 0, 255).reshape(16, 192)
    assert np.array_equal(hz.luma(ctx, px)[0], (s // 3).reshape(16, 192))
    assert np.array_equal(hz.luma_gauss(ctx, px, 0.5), np.stack([ops.blur(ops.mono(f), 0.5) for f in px]))


def test_luma_rejects_bad_arguments(be, ctx):
    fr = rng_frames(1, (1, 4, 8, 3))
    with pytest.raises(ValueError):
        hz.luma(ctx, fr, mode=7)


# ---- K2 -----------------------------------------------------------------------------------------
def test_gauss_bit_exact(be, ctx):
    for (H, W) in sizes(be, [(13, 37), (70, 140)], [(480, 640), (1080, 1920), (333, 1001)]):
        g = rng_frames(H + W, (2, H, W))
        for s in sizes(be, [0.3, 1, 2, 3], [0.5, 5]):
            assert np.array_equal(hz.gauss(ctx, g, s), np.stack([ops.blur(f, s) for f in g])), (H, W, s)
        assert np.array_equal(hz.gauss(ctx, g, 2, in_pad=3, out_pad=5), np.stack([ops.blur(f, 2) for f in g]))


def test_gauss_streaming_kernel(be, ctx, monkeypatch):
    # 16-byte aligned frames with W % 16 == 0 take the register-window kernel (radius <= 9): every radius,
    # strips cut by the right image edge, several row segments, images barely taller than the window,
    # fused and unfused, mean and single-channel luma
    monkeypatch.setenv('VA_GAUSS_MMA', '0')              # the dot-product kernels (fallback of the tensor-core kernel)
    cases = sizes(be, [(33, 160), (12, 16), (40, 640), (21, 336)], [(480, 640), (271, 1008), (67, 2064)])
    for (H, W) in cases:
        fr = rng_frames(H * W, (2, H, W, 3))
        g = fr[..., 1].copy()
        for s in (0.3, 0.5, 0.7, 0.9, 1.1, 1.3, 1.5, 2, 2.6, 3):
            want = np.stack([ops.blur(f, s) for f in g])
            for segs, nt in ((None, None), (1, 64), (3, 96), (2, 256)):
                for key, val in (('VA_GS_SEGS', segs), ('VA_GS_NT', nt)):
                    if val is None:
                        monkeypatch.delenv(key, raising=False)
                    else:
                        monkeypatch.setenv(key, str(val))
                assert np.array_equal(hz.gauss(ctx, g, s), want), (H, W, s, segs, nt)
                if segs is None or s == 2:
                    assert np.array_equal(hz.luma_gauss(ctx, fr, s, mode=1), want), (H, W, s, segs, nt)
                    assert np.array_equal(hz.luma_gauss(ctx, fr, s),
                                          np.stack([ops.blur(ops.mono(f), s) for f in fr])), (H, W, s, segs, nt)
    monkeypatch.setenv('VA_GAUSS_STREAM', '0')           # and the tile kernel on the same inputs
    assert np.array_equal(hz.gauss(ctx, g, 2), np.stack([ops.blur(f, 2) for f in g]))


def test_gauss_tensor_core_kernel(be, ctx, monkeypatch):
    # va_gauss_mma.cu (16-byte aligned frames, W % 16 == 0): every group count G = 2 .. 8 of the column pass (radius 1 .. 56),
    # both strip widths, 2 .. 4 TMA stages, one and several row segments, strips cut by the right image edge, images
    # barely larger than the radius, fused (mean and channel pick) and plain; identical to the dot-product kernels
    cases = sizes(be, [(40, 160), (70, 272), (21, 336), (130, 144)], [(480, 640), (1080, 1920), (271, 1008), (67, 2064)])
    # (the emulator takes one sigma per group count, and only three of them on the last two shapes)
    sig = [0.5, 1, 2, 2.6, 3, 3.5, 5, 6, 8, 9, 12, 15, 18] if be.name == 'cuda' else [0.5, 2, 3, 6, 9, 12, 15, 18]
    for n, (H, W) in enumerate(cases):
        fr = rng_frames(H * W, (2, H, W, 3))
        g = fr[..., 2].copy()
        for s in sig:
            if be.name != 'cuda' and n >= 2 and s not in (2, 15, 18):
                continue
            want = np.stack([ops.blur(f, s) for f in g])
            full = n == 0 and (be.name == 'cuda' or s in (2, 15))
            # the large radii (G > 2) run the CTA-per-strip kernel with two tiles per warp by default; VA_GM_CTA=0: the
            # warp-per-strip kernel for them as well; VA_GMC_TPW: one / four tiles per warp
            variants = [dict(), dict(VA_GM_SEGS=1), dict(VA_GM_SEGS=3, VA_GM_TILES=8), dict(VA_GM_STAGES=2),
                        dict(VA_GM_STAGES=3, VA_GM_MINB=3), dict(VA_GM_CTA=0), dict(VA_GM_CTA=0, VA_GM_SEGS=2, VA_GM_TILES=8),
                        dict(VA_GMC_TPW=1, VA_GM_MINB=2, VA_GM_SEGS=2, VA_GMC_STAGES=3), dict(VA_GMC_TPW=4),
                        dict(VA_GMC_STAGES=9, VA_GM_SEGS=3)] if full else \
                ([dict(), dict(VA_GM_CTA=0), dict(VA_GMC_TPW=1)] if be.name == 'cuda' else
                 [dict(), dict(VA_GM_CTA=0)] if s in (3, 15) else [dict()])
            for var in variants:
                for key in ('VA_GM_SEGS', 'VA_GM_TILES', 'VA_GM_STAGES', 'VA_GM_MINB', 'VA_GM_CTA', 'VA_GMC_STAGES', 'VA_GMC_TPW'):
                    monkeypatch.delenv(key, raising=False)
                for key, val in var.items():
                    monkeypatch.setenv(key, str(val))
                monkeypatch.setenv('VA_GAUSS_MMA', '2')       # 2: also for radius 9, where the streaming kernel is the default
                assert np.array_equal(hz.gauss(ctx, g, s), want), (H, W, s, var)
                assert np.array_equal(hz.luma_gauss(ctx, fr, s, mode=2), want), (H, W, s, var)
                if s in (1, 2, 15):
                    assert np.array_equal(hz.luma_gauss(ctx, fr, s), np.stack([ops.blur(ops.mono(f), s) for f in fr])), (H, W, s, var)
    for key in ('VA_GM_SEGS', 'VA_GM_TILES', 'VA_GM_STAGES', 'VA_GM_MINB', 'VA_GM_CTA', 'VA_GMC_STAGES', 'VA_GMC_TPW'):
        monkeypatch.delenv(key, raising=False)
    monkeypatch.setenv('VA_GAUSS_MMA', '0')
    assert np.array_equal(hz.gauss(ctx, g, 5), np.stack([ops.blur(f, 5) for f in g]))
    # unaligned rows and ragged widths fall back to the dot-product kernels (same bytes)
    monkeypatch.delenv('VA_GAUSS_MMA', raising=False)
    odd = rng_frames(9, (1, 40, 150))
    assert np.array_equal(hz.gauss(ctx, odd, 2), np.stack([ops.blur(f, 2) for f in odd]))
    al = rng_frames(10, (1, 40, 160))
    assert np.array_equal(hz.gauss(ctx, al, 2, in_pad=3, out_pad=5), np.stack([ops.blur(f, 2) for f in al]))


def test_gauss_large_sigma_identity_and_generic_paths(be, ctx):
    g = rng_frames(3, (1, 50, 70))
    for s in (0.05, 0.2, 10, 15, 21):          # ksize 1 (copy), tap 256 (generic), wide (TH=96)
        assert np.array_equal(hz.gauss(ctx, g, s), np.stack([ops.blur(f, s) for f in g])), s
    c = rng_frames(4, (1, 30, 45, 3))
    for s in (1, 2, 4):
        assert np.array_equal(hz.gauss(ctx, c, s), np.stack([ops.blur(f, s) for f in c])), s
    with pytest.raises(NotImplementedError):
        hz.gauss(ctx, g, 60)


def test_gauss_golden(be, ctx):
    assert np.array_equal(hz.gauss(ctx, GOLD['s_mono'], 2), GOLD['s_blur'])
    assert np.array_equal(hz.gauss(ctx, GOLD['r_mono'], 3), GOLD['r_blur'])


def test_luma_gauss_fused_equals_two_step(be, ctx):
    # aligned and ragged widths, tiles cut by the right / bottom image edge, images smaller than a tile
    for (H, W) in sizes(be, [(33, 150), (33, 160), (9, 16), (5, 144)], [(1080, 1920), (480, 640), (271, 1008)]):
        fr = rng_frames(5, (2, H, W, 3))
        for s in (1, 2, 3.3, 5):
            assert np.array_equal(hz.luma_gauss(ctx, fr, s), np.stack([ops.blur(ops.mono(f), s) for f in fr]))
    fr = rng_frames(6, (1, 20, 40, 3))
    assert np.array_equal(hz.luma_gauss(ctx, fr, 2, mode=1), np.stack([ops.blur(f[..., 1], 2) for f in fr]))


# ---- sparse egress of label images ----------------------------------------------------------------
def _chunk_kinds(frame, W):
    """ (non-empty 64-pixel chunks, those whose foreground is one horizontal run) of a boolean frame """
    pad = (-W) % 64
    f = np.concatenate([frame, np.zeros((frame.shape[0], pad), bool)], 1).reshape(frame.shape[0], -1, 64)
    any_ = f.any(-1)
    starts = (f & ~np.concatenate([np.zeros(f.shape[:2] + (1,), bool), f[..., :-1]], -1)).sum(-1)
    return int(any_.sum()), int((starts == 1).sum())


def test_label_export_chunks_and_host_densify(be, ctx):
    # the non-empty 64-label chunks written by the device (straight into page-locked host memory on the GPU) and the
    # host-side rebuild give the dense label image bit for bit; the result buffer is reused without being zeroed;
    # chunks whose foreground is a single run travel as 16-byte records when the caller asks for them
    for (H, W) in sizes(be, [(20, 64), (23, 200), (9, 130)], [(1080, 1920), (480, 640), (271, 1003)]):
        for use_runs in (True, False):
            rng = np.random.default_rng(H * W)
            state = None
            for rep, density in enumerate((0.02, 0.5, 0.0, 0.001, 1.0)):
                m = (rng.random((2, H, W)) < density)
                if density == 0.02:
                    m[0, H // 3: H // 2, W // 4: W // 2] = True
                words = ops.pack_bits(m)
                dense, n, state, direct, nr = hz.label_export_dense(ctx, words, W, reuse=state, lab_pad=(H + W) % 3,
                                                                    use_runs=use_runs)
                assert np.array_equal(dense, direct), (H, W, density)
                want = np.stack([ops.label(f)[0] for f in m])
                assert np.array_equal(dense, want), (H, W, density)
                kinds = [_chunk_kinds(f, W) for f in m]
                if use_runs:
                    assert list(nr) == [k[1] for k in kinds] and list(n) == [k[0] - k[1] for k in kinds]
                else:
                    assert list(n) == [k[0] for k in kinds] and not nr.any()
                assert max(k[0] for k in kinds) <= ((W + 63) // 64) * H
    # capacity smaller than the number of chunks: the true count is reported (the caller falls back to a dense copy)
    m = np.ones((1, 8, 128), bool)
    m[:, :, ::7] = False
    _, n, _, _, _ = hz.label_export_dense(ctx, ops.pack_bits(m), 128, cap=5)
    assert n[0] == 16
    m = np.ones((1, 8, 128), bool)
    _, n, _, _, nr = hz.label_export_dense(ctx, ops.pack_bits(m), 128, cap=5)
    assert n[0] == 0 and nr[0] == 16


# ---- K2b ----------------------------------------------------------------------------------------
def test_resize_half(be, ctx):
    for (H, W) in sizes(be, [(12, 40), (20, 64)], [(1080, 1920)]):
        g = rng_frames(H, (2, H, W))
        ref = np.stack([ops.resize(f, 0.5) for f in g])
        assert np.array_equal(hz.resize_half(ctx, g), ref)
        assert np.array_equal(hz.resize_half(ctx, g, in_pad=3), ref)
        c = rng_frames(W, (2, H, W, 3))
        assert np.array_equal(hz.resize_half(ctx, c), np.stack([ops.resize(f, 0.5) for f in c]))
    with pytest.raises(ValueError):
        hz.resize_half(ctx, rng_frames(0, (1, 7, 8)))


# ---- threshold / pack / unpack / apply-mask ------------------------------------------------------
def test_threshold_pack_unpack(be, ctx):
    for (H, W) in sizes(be, [(7, 37), (5, 64), (3, 1100)], [(1080, 1920)]):
        g = rng_frames(W, (2, H, W))
        assert np.array_equal(hz.threshold_bits(ctx, g, 100), ops.pack_bits(g > 100))
        assert np.array_equal(hz.threshold_bits(ctx, g, 100, pad=3), ops.pack_bits(g > 100))
        m = rmask(H, (2, H, W), 0.4) & g
        assert np.array_equal(hz.pack_bits(ctx, m), ops.pack_bits(m))
        p = ops.pack_bits(m)
        assert np.array_equal(hz.unpack_bits(ctx, p, W), ops.unpack_bits(p, W))
        assert np.array_equal(hz.unpack_bits(ctx, p, W, pad=5), ops.unpack_bits(p, W))


def test_apply_mask(be, ctx):
    for (H, W) in sizes(be, [(7, 37), (6, 64)], [(720, 1280)]):
        g, c = rng_frames(1, (2, H, W)), rng_frames(2, (2, H, W, 3))
        m = rmask(3, (H, W), 0.5) & rng_frames(4, (H, W))
        assert np.array_equal(hz.apply_mask(ctx, g, m), np.stack([ops.apply_mask(f, m) for f in g]))
        assert np.array_equal(hz.apply_mask(ctx, c, m), np.stack([ops.apply_mask(f, m) for f in c]))


# ---- synthetic source -------------------------------------------------------------------------------
def test_synth_equals_numpy_generator(be, ctx):
    for (H, W, nb) in sizes(be, [(20, 37, 3), (32, 64, 8)], [(480, 640, 8)]):
        tab = synth.blob_table(5, W, H, nb)
        assert np.array_equal(hz.synth(ctx, 5, 7, 3, W, H, tab), synth.make_frames(5, 7, 3, W, H, nb))
    assert np.array_equal(hz.synth(ctx, 0, 0, 6, 64, 48, synth.blob_table(0, 64, 48, 4)), GOLD['s_frames'])


def test_resize_area_integer_factors_and_nearest(be, ctx):
    rng = np.random.default_rng(8)
    for (H, W) in sizes(be, [(24, 60), (36, 48)], [(1080, 1920), (720, 1280)]):
        g = rng_frames(H, (2, H, W))
        c = rng_frames(W, (2, H, W, 3))
        for kx, ky in ((3, 3), (4, 4), (2, 3), (6, 2), (3, 1), (1, 2), (2, 2), (12, 12)):
            if W % kx or H % ky:
                continue
            for fr in (g, c):
                ref = np.stack([ops.resize(f, (W // kx, H // ky), 'area') for f in fr])
                assert np.array_equal(hz.resize_area(ctx, fr, kx, ky), ref), (H, W, kx, ky, fr.ndim)
        for dw, dh in ((W // 2, H // 2), (W // 3 + 1, H // 5 + 2), (W + 7, H + 3), (1, 1), (2 * W, 3)):
            for fr in (g, c):
                ref = np.stack([ops.resize(f, (dw, dh), 'nearest') for f in fr])
                assert np.array_equal(hz.resize_nearest(ctx, fr, dw, dh), ref), (H, W, dw, dh, fr.ndim)
    with pytest.raises(ValueError):
        hz.resize_area(ctx, rng_frames(1, (1, 10, 10)), 3, 3)


def test_resize_area_any_factor_and_linear(be, ctx):
    for (H, W) in sizes(be, [(24, 60), (37, 53)], [(1080, 1920), (271, 1003)]):
        g = rng_frames(H + 1, (2, H, W))
        c = rng_frames(W + 1, (2, H, W, 3))
        for dw, dh in ((W * 3 // 10, H * 3 // 10), (W - 1, H - 1), (W // 2 + 1, H // 3), (7, 5), (W, H // 2), (W // 3, H),
                       (W // 2, H // 2), (W * 2 // 3, H * 2 // 3), (1, 1)):
            for fr in (g, c):
                ref = np.stack([ops.resize(f, (dw, dh), 'area') for f in fr]).reshape((2, dh, dw) + fr.shape[3:])
                assert np.array_equal(hz.resize_to(ctx, fr, dw, dh, 'area_any'), ref), (H, W, dw, dh, fr.ndim)
        for dw, dh in ((W * 3 // 10, H * 3 // 10), (W - 1, H - 1), (W + 7, H + 3), (2 * W, 2 * H), (W // 2, H // 2), (7, 5),
                       (1, 1), (W + W // 2 + 1, H // 2), (W, H * 2)):
            for fr in (g, c):
                ref = np.stack([ops.resize(f, (dw, dh), 'linear') for f in fr]).reshape((2, dh, dw) + fr.shape[3:])
                assert np.array_equal(hz.resize_to(ctx, fr, dw, dh, 'linear'), ref), (H, W, dw, dh, fr.ndim)
        for dw, dh in ((W + 5, H // 2), (W // 2, H + 3), (2 * W, 2 * H), (W + 1, H + 1), (3 * W + 2, H)):   # INTER_AREA that enlarges
            for fr in (g, c):
                ref = np.stack([ops.resize(f, (dw, dh), 'area') for f in fr]).reshape((2, dh, dw) + fr.shape[3:])
                assert np.array_equal(hz.resize_to(ctx, fr, dw, dh, 'area_any'), ref), (H, W, dw, dh, fr.ndim)


def test_resize_cubic(be, ctx):
    import cv2
    ipp = cv2.ipp.useIPP()
    try:
        for (H, W) in sizes(be, [(24, 60), (37, 53)], [(1080, 1920), (271, 1003)]):
            g = rng_frames(H + 2, (2, H, W))
            c = rng_frames(W + 2, (2, H, W, 3))
            for dw, dh in ((W * 2, H * 2), (W + 7, H + 3), (W * 3 // 10, H * 3 // 10), (W + W // 2 + 1, H // 2), (7, 5), (1, 1)):
                for fr in (g, c):
                    out = hz.resize_to(ctx, fr, dw, dh, 'cubic')
                    for use_ipp, tol in ((False, 0), (True, 1)):     # OpenCV's own arithmetic: exact; Intel IPP's: 1 LSB
                        cv2.ipp.setUseIPP(use_ipp)
                        ref = np.stack([ops.resize(f, (dw, dh), 'cubic') for f in fr]).reshape((2, dh, dw) + fr.shape[3:])
                        assert np.abs(out.astype(np.int16) - ref).max() <= tol, (H, W, dw, dh, fr.ndim, use_ipp)
    finally:
        cv2.ipp.setUseIPP(ipp)


def test_resize_lanczos4(be, ctx):
    for (H, W) in sizes(be, [(24, 60), (37, 53)], [(1080, 1920), (271, 1003)]):
        g = rng_frames(H + 4, (2, H, W))
        c = rng_frames(W + 4, (2, H, W, 3))
        for dw, dh in ((W * 2, H * 2), (W + 7, H + 3), (W * 3 // 10, H * 3 // 10), (W + W // 2 + 1, H // 2), (7, 5), (1, 1), (W, H * 3)):
            for fr in (g, c):
                ref = np.stack([ops.resize(f, (dw, dh), 'lanczos') for f in fr]).reshape((2, dh, dw) + fr.shape[3:])
                assert np.array_equal(hz.resize_to(ctx, fr, dw, dh, 'lanczos4'), ref), (H, W, dw, dh, fr.ndim)


def test_luma_crop_per_stream_positions(be, ctx):
    for (H, W, w, h) in sizes(be, [(30, 41, 17, 11), (24, 64, 32, 8)], [(720, 1280, 642, 363), (1080, 1920, 1024, 576)]):
        S = 4
        fr = rng_frames(H * W + 1, (S, H, W, 3))
        rng = np.random.default_rng(w)
        xy = np.stack([rng.integers(0, W - w + 1, S), rng.integers(0, H - h + 1, S)], axis=1)
        xy[0] = (0, 0)
        xy[1] = (W - w, H - h)
        for mode, name in ((-1, 'mean'), (0, 'blue'), (2, 'red')):
            ref = np.stack([ops.mono(ops.crop(fr[s], (int(xy[s, 0]), int(xy[s, 1]), w, h)), name) for s in range(S)])
            for pad, out_pad in ((0, 3), (1, 3), (0, 0), (4, 16)):        # the last two take the 16-pixel path when w % 16 == 0
                got = hz.luma_crop_multi(ctx, fr, xy, w, h, mode, pad, out_pad)
                assert np.array_equal(got, ref), (H, W, w, h, mode, pad, out_pad)
    with pytest.raises(ValueError):
        hz.luma_crop_multi(ctx, rng_frames(1, (1, 8, 8, 3)), [(0, 0)], 9, 4)


def test_streams_fused_front(be, ctx):
    for (H, W, w, h) in sizes(be, [(30, 80, 32, 11), (24, 100, 64, 8)], [(720, 1280, 640, 352), (1080, 1920, 1024, 576)]):
        S = 3
        fr = rng_frames(H * W + 2, (S, H, W, 3))
        rng = np.random.default_rng(h)
        xy = np.stack([rng.integers(0, W - w + 1, S), rng.integers(0, H - h + 1, S)], axis=1)
        xy[0] = (W - w, H - h)
        masks = rmask(w * h, (S, h, w), 0.7)
        for mode, name in ((-1, 'mean'), (1, 'green')):
            mono = np.stack([ops.mono(ops.crop(fr[s], (int(xy[s, 0]), int(xy[s, 1]), w, h)), name) for s in range(S)])
            for mk in (None, masks[0], masks):
                for thr in (110, 0, 255, -1):
                    g = mono if mk is None else np.where((mk if mk.ndim == 3 else mk[None]) != 0, mono, 0)
                    ref = hz.pack_bits_np(((g.astype(np.int32) > thr) * 255).astype(np.uint8))
                    got = hz.streams_threshold(ctx, fr, xy, w, h, mk, thr, mode, in_pad=0 if thr else 4)
                    assert np.array_equal(got, ref), (H, W, w, h, mode, None if mk is None else mk.ndim, thr)
    with pytest.raises(NotImplementedError):
        hz.streams_threshold(ctx, rng_frames(1, (1, 20, 40, 3)), [(0, 0)], 17, 4, None, 10)


def test_highlight_mask(be, ctx):
    for (H, W) in sizes(be, [(13, 37), (24, 64)], [(1080, 1920), (271, 1003)]):
        g = rng_frames(H + 3, (2, H, W))
        c = rng_frames(W + 3, (2, H, W, 3))
        m = rmask(H * W, (2, H, W), 0.3)
        assert np.array_equal(hz.pack_bits_np(m), hz.pack_bits(ctx, m))
        for strength in (128, 0, 255, 77):
            table = np.empty(256, np.uint8)
            table[:] = strength + (255 - strength) / 255 * np.arange(256, dtype=np.uint8)
            for fr, channels in ((g, ('all',)), (c, ('all', 'red', 1, 'b'))):
                for ch in channels:
                    ref = np.stack([ops.highlight_mask(f, mm, ch, strength) for f, mm in zip(fr, m)])
                    ci = -1 if ch == 'all' else {'red': 0, 1: 1, 'b': 2}[ch]
                    for pad in (0, 3):
                        assert np.array_equal(hz.highlight_mask(ctx, fr, m, ci, table, pad), ref), (H, W, strength, ch, pad)
    with pytest.raises(ValueError):
        hz.highlight_mask(ctx, rng_frames(1, (1, 8, 8)), rmask(1, (1, 8, 8), .5), 2, np.zeros(256, np.uint8))


# ---- K3 -----------------------------------------------------------------------------------------
def noisy_video(seed, shape):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 256, shape) // 8 + 100 + rng.integers(0, 2, shape) * 60).astype(np.uint8)


def test_ema_diff_threshold_bit_exact(be, ctx):
    for (B, H, W) in sizes(be, [(6, 9, 37), (9, 5, 600), (5, 64, 1100)], [(8, 1080, 1920), (7, 480, 640)]):
        g = noisy_video(B, (B, H, W))
        m_ref, bg_ref = ops.background_ema(list(g), 0.05, 25)
        for pad in (0, 4):
            m, bg = hz.ema_diff_thresh(ctx, g, 0.05, 25, pad=pad)
            assert np.array_equal(m, ops.pack_bits(m_ref))
            assert np.array_equal(bg.view(np.uint32), bg_ref.view(np.uint32))       # float32 state, bit for bit
        m2_ref, bg2_ref = ops.background_ema(list(g), 0.05, 25, bg0=bg_ref)         # continuation batch
        m2, bg2 = hz.ema_diff_thresh(ctx, g, 0.05, 25, bg0=bg_ref)
        assert np.array_equal(m2, ops.pack_bits(m2_ref))
        assert np.array_equal(bg2.view(np.uint32), bg2_ref.view(np.uint32))


def test_ema_flat_group_mapping(be, ctx, monkeypatch):
    # widths that are multiples of 32 but not of a whole warp step take the raster-order group mapping
    # (warps straddle rows, the last warp has idle lanes); both pixels-per-thread variants, and the
    # row-chunk mapping on the same input as the cross-check
    for px in ('4', '8', '16'):
        monkeypatch.setenv('VA_EMA_PX', px)
        for (B, H, W) in sizes(be, [(5, 9, 96), (4, 7, 160), (3, 5, 1056), (6, 1, 32)], [(5, 1080, 1920), (4, 720, 1280)]):
            g = noisy_video(W, (B, H, W))
            m_ref, bg_ref = ops.background_ema(list(g), 0.05, 25)
            for noflat in (False, True):
                if noflat:
                    monkeypatch.setenv('VA_EMA_NOFLAT', '1')
                else:
                    monkeypatch.delenv('VA_EMA_NOFLAT', raising=False)
                m, bg = hz.ema_diff_thresh(ctx, g, 0.05, 25)
                assert np.array_equal(m, ops.pack_bits(m_ref)), (px, B, H, W, noflat)
                assert np.array_equal(bg.view(np.uint32), bg_ref.view(np.uint32))
    monkeypatch.delenv('VA_EMA_NOFLAT', raising=False)


def test_ema_unrolled_ring_rounds(be, ctx, monkeypatch):
    # batches longer than the cp.async ring: whole turns of the unrolled frame loop plus a tail, on shapes where every
    # lane is on the aligned path (the unrolled variant) and on one where it is not (the general loop), for every
    # pixels-per-thread variant; continuation batches start from the state of the first
    for px in ('16', '8', '4'):
        monkeypatch.setenv('VA_EMA_PX', px)
        for (B, H, W) in sizes(be, [(19, 4, 512), (9, 2, 1024), (17, 3, 96), (33, 2, 256)], [(19, 480, 640), (35, 64, 1024)]):
            g = noisy_video(B + W, (B, H, W))
            m_ref, bg_ref = ops.background_ema(list(g), 0.05, 25)
            m, bg = hz.ema_diff_thresh(ctx, g, 0.05, 25)
            assert np.array_equal(m, ops.pack_bits(m_ref)), (px, B, H, W)
            assert np.array_equal(bg.view(np.uint32), bg_ref.view(np.uint32))
            m2_ref, bg2_ref = ops.background_ema(list(g[::-1]), 0.05, 25, bg0=bg_ref)
            m2, bg2 = hz.ema_diff_thresh(ctx, g[::-1].copy(), 0.05, 25, bg0=bg_ref)
            assert np.array_equal(m2, ops.pack_bits(m2_ref)), (px, B, H, W)
            assert np.array_equal(bg2.view(np.uint32), bg2_ref.view(np.uint32))
    monkeypatch.delenv('VA_EMA_PX', raising=False)


def test_ema_threshold_edge_cases(be, ctx):
    # differences that hit the threshold exactly (alpha = 0.5 keeps the state on a dyadic grid), zero,
    # negative and signed-zero thresholds; a non-finite threshold is rejected
    rng = np.random.default_rng(11)
    g = rng.integers(0, 8, (9, 6, 64), dtype=np.uint8) * 4
    for alpha, thr in ((0.5, 2.0), (0.5, 1.0), (0.25, 3.0), (0.05, 0.0), (0.05, -0.0), (0.5, -1.0), (1.0, 4.0)):
        m_ref, bg_ref = ops.background_ema(list(g), alpha, thr)
        m, bg = hz.ema_diff_thresh(ctx, g, alpha, thr)
        assert np.array_equal(m, ops.pack_bits(m_ref)), (alpha, thr)
        assert np.array_equal(bg.view(np.uint32), bg_ref.view(np.uint32))
    with pytest.raises(ValueError):
        hz.ema_diff_thresh(ctx, g, 0.05, float('inf'))


def test_ema_golden(be, ctx):
    m, bg = hz.ema_diff_thresh(ctx, GOLD['s_blur'], 0.05, 25)
    assert np.array_equal(m, ops.pack_bits(GOLD['s_mask']))
    assert np.array_equal(bg.view(np.uint32), GOLD['s_bg'].view(np.uint32))


def test_ema_partial_and_fold(be, ctx):
    # frame-sharded recurrence: tolerance 1e-5 relative (re-association), SURVEY.md 8e
    g = noisy_video(1, (12, 10, 50))
    alpha = 0.05
    _, bg_seq = ops.background_ema(list(g), alpha, 25)
    # rank 0 owns frames 0..6 (true state), rank 1 folds frames 7..11 from a zero state
    _, bg_r0 = ops.background_ema(list(g[:7]), alpha, 25)
    S1 = hz.ema_partial(ctx, g[7:], alpha)
    state = hz.ema_fold(ctx, bg_r0, S1, (1 - alpha) ** 5)
    assert np.allclose(state, bg_seq, rtol=1e-5, atol=1e-4)
    # accumulate form: two halves equal one pass
    Sa = hz.ema_partial(ctx, g[7:9], alpha)
    Sb = hz.ema_partial(ctx, g[9:], alpha, S0=Sa)
    assert np.allclose(Sb, S1, rtol=1e-6)


# ---- K4 -----------------------------------------------------------------------------------------
SES = [('rect', 3), ('cross', 3), ('rect', 5), ('ellipse', 5), ('ellipse', 7), ('rect', (2, 2)), ('rect', (4, 3)),
       ('cross', (5, 7)), ('ellipse', (9, 5)), ('rect', 7)]


def test_morphology_bit_exact(be, ctx):
    for (H, W) in sizes(be, [(9, 37), (40, 64), (70, 100)], [(1080, 1920)]):
        for p in (0.3, 0.7, 0.95):
            m = rmask(int(p * 100) + W, (2, H, W), p)
            packed = ops.pack_bits(m)
            for op in ('erode', 'dilate', 'open', 'close'):
                for shape, k in (SES if H < 1000 else SES[:3]):
                    ref = np.stack([ops.morph(f, op, shape, k) for f in m])
                    assert np.array_equal(hz.morph(ctx, packed, W, op, shape, k), ops.pack_bits(ref)), (H, W, p, op, shape, k)


def test_morphology_square_elements_strips_and_segments(be, ctx):
    # the register-streaming kernel: several word strips per row (ragged last word), several row
    # segments per frame, images shorter than the element
    for (H, W) in ((100, 1003), (67, 1920), (2, 40), (1, 33), (5, 961), (37, 2200), (9, 4001)):
        m = rmask(H + W, (2, H, W), 0.8)
        for op in ('erode', 'dilate', 'open', 'close'):
            for k in (3, 5, 7):
                ref = np.stack([ops.morph(f, op, 'rect', k) for f in m])
                assert np.array_equal(hz.morph(ctx, ops.pack_bits(m), W, op, 'rect', k), ops.pack_bits(ref)), (H, W, op, k)


def test_morphology_large_elements(be, ctx):
    m = rmask(9, (1, 50, 90), 0.9)
    for k in (9, 15, 31):
        for shape in ('ellipse', 'rect', 'cross'):
            for op in ('erode', 'open', 'close'):
                ref = np.stack([ops.morph(f, op, shape, k) for f in m])
                assert np.array_equal(hz.morph(ctx, ops.pack_bits(m), 90, op, shape, k), ops.pack_bits(ref)), (k, shape, op)
    with pytest.raises(NotImplementedError):
        hz.morph(ctx, ops.pack_bits(m), 90, 'open', 'rect', 65)


def test_morphology_golden_and_edges(be, ctx):
    assert np.array_equal(hz.morph(ctx, ops.pack_bits(GOLD['s_mask']), 64, 'open', 'rect', 3), ops.pack_bits(GOLD['s_morph']))
    assert np.array_equal(hz.morph(ctx, ops.pack_bits(GOLD['r_mask']), 53, 'close', 'ellipse', 5), ops.pack_bits(GOLD['r_morph']))
    for fill in (0, 255):
        m = np.full((1, 11, 45), fill, np.uint8)
        for op in ('erode', 'dilate', 'open', 'close'):
            assert np.array_equal(hz.morph(ctx, ops.pack_bits(m), 45, op), ops.pack_bits(m))


# ---- K5 -----------------------------------------------------------------------------------------
def check_labels(ctx, m, conns=(4, 8)):
    B, H, W = m.shape
    for conn in conns:
        lab, cnt = hz.label(ctx, ops.pack_bits(m), W, conn)
        refs = [ops.label(f, conn) for f in m]
        assert np.array_equal(cnt, np.array([r[1] for r in refs], np.int32)), conn
        assert np.array_equal(lab, np.stack([r[0] for r in refs])), conn


def test_label_random_masks(be, ctx):
    for (H, W) in sizes(be, [(9, 37), (40, 64), (12, 1100)], [(1080, 1920), (2160, 3840)]):
        for p in (0.1, 0.5, 0.6, 0.9):
            check_labels(ctx, rmask(int(p * 10) + H, (2 if H < 2000 else 1, H, W), p))


def adversarial_masks(H, W):
    out = {}
    out['zeros'] = np.zeros((H, W), np.uint8)
    out['ones'] = np.full((H, W), 255, np.uint8)
    out['checker'] = ((np.indices((H, W)).sum(0) % 2) * 255).astype(np.uint8)      # N/2 components at 4-conn
    u = np.zeros((H, W), np.uint8)                                                 # U and n shapes: late merges
    u[5:H - 8, 10] = u[5:H - 8, 30] = 255
    u[H - 9, 10:31] = 255
    u[5:H - 8, 50] = u[5:H - 8, 70] = 255
    u[5, 50:71] = 255
    out['u_shapes'] = u
    s = np.zeros((H, W), np.uint8)                                                 # frame-spanning serpentine
    for i in range(0, H, 4):
        s[i, :] = 255
    for i in range(0, H - 4, 8):
        s[i:i + 5, W - 1] = 255
        s[i + 4:min(i + 9, H), 0] = 255
    out['serpentine'] = s
    c = np.zeros((H, W), np.uint8)                                                 # comb: long run over many short ones
    c[1::2, ::2] = 255
    c[0::4, :] = 255
    out['comb'] = c
    d = np.zeros((H, W), np.uint8)                                                 # diagonal: 8-conn only
    idx = np.arange(min(H, W))
    d[idx, idx] = 255
    d[idx, W - 1 - idx] = 255
    out['diagonals'] = d
    return out


def test_label_adversarial(be, ctx):
    for (H, W) in sizes(be, [(48, 96), (33, 75)], [(1080, 1920)]):
        for name, m in adversarial_masks(H, W).items():
            check_labels(ctx, m[None])


def test_label_noise_next_to_blobs(be, ctx):
    # rows with hundreds of runs next to rows with a few, components that cross from one kind into the other,
    # widths with several 32-word chunks
    rng = np.random.default_rng(21)
    for (H, W) in sizes(be, [(70, 1300)], [(400, 1920), (90, 4000)]):
        m = np.zeros((2, H, W), np.uint8)
        m[0, :H // 3] = rng.random((H // 3, W)) < 0.5                      # noise on top
        m[0, H // 3 - 2:, W // 4:W // 4 + 70] = 1                          # a bar growing out of the noise
        m[0, H // 2:H // 2 + 11, 5:W - 5] = 1                              # a wide slab
        m[0, H // 2 + 10:, W - 40:W - 8] = 1
        m[1, 3::8] = rng.random((len(range(3, H, 8)), W)) < 0.5            # one noisy row per band
        m[1, :, ::97] = 1                                                  # vertical lines through all bands
        m[1, H - 9:, :] |= (rng.random((9, W)) < 0.55).astype(np.uint8)
        check_labels(ctx, m)


def test_label_golden_and_pitch(be, ctx):
    lab, cnt = hz.label(ctx, ops.pack_bits(GOLD['s_morph']), 64, 4, lab_pad=4)
    assert np.array_equal(lab, GOLD['s_labels']) and np.array_equal(cnt, GOLD['s_counts'])
    lab, cnt = hz.label(ctx, ops.pack_bits(GOLD['r_morph']), 53, 8)
    assert np.array_equal(lab, GOLD['r_labels']) and np.array_equal(cnt, GOLD['r_counts'])


def test_label_forest_and_write_halves_with_two_scratch_slots(be, ctx):
    for (H, W) in sizes(be, [(40, 70)], [(1080, 1920)]):
        ma, mb = rmask(5, (2, H, W), 0.55), rmask(6, (2, H, W), 0.35)
        for conn in (4, 8):
            (la, ca), (lb, cb) = hz.label_two_batches_split(ctx, hz.pack_bits_np(ma), hz.pack_bits_np(mb), W, conn)
            for lab, cnt, m in ((la, ca, ma), (lb, cb, mb)):
                for t in range(2):
                    ref, n = ops.label(m[t], conn)
                    assert cnt[t] == n and np.array_equal(lab[t], ref), (H, W, conn, t)
    with pytest.raises(ValueError):
        hz.label_two_batches_split(ctx, hz.pack_bits_np(rmask(1, (1, 8, 8), .5)), hz.pack_bits_np(rmask(1, (1, 8, 8), .5)), 8, 5)


def test_label_int16_output(be, ctx):
    """ ndimage.label(mask, output=np.int16): same numbering, half the bytes """
    for (H, W) in sizes(be, [(40, 70), (9, 131)], [(1080, 1920), (271, 1003)]):
        for p in (0.03, 0.55):
            m = rmask(H + int(100 * p), (2, H, W), p)
            if H * W > 100000:            # keep the component count of the large frames below 32767: 8 x 8 blocks of a coarse mask
                coarse = rmask(H + int(100 * p), (2, (H + 7) // 8, (W + 7) // 8), p)
                m = np.kron(coarse, np.ones((1, 8, 8), np.uint8))[:, :H, :W]
            for pad in (0, 3, 4):
                lab, cnt = hz.label_i16(ctx, hz.pack_bits_np(m), W, 4, pad)
                assert lab.dtype == np.int16
                for t in range(2):
                    ref = np.zeros((H, W), np.int16)
                    n = ndimage.label(m[t], output=ref)
                    assert cnt[t] == n and np.array_equal(lab[t], ref), (H, W, p, pad, t)


def test_label_capacity_error(be, ctx):
    m = np.zeros((17, 4, 40), np.uint8)
    with pytest.raises(MemoryError):
        hz.label(ctx, ops.pack_bits(m), 40)


def synth_masks(n, H, W):
    fr = synth.make_frames(0, 0, n, W, H, 6)
    return ops.chain(fr)['morph'].astype(bool)


def test_region_areas_and_largest(be, ctx):
    fr = synth.make_frames(0, 0, 6, 200, 120, 6)
    r = ops.chain(fr)
    lab, cnt = hz.label(ctx, ops.pack_bits(r['morph']), 200, 4)
    areas, largest = hz.region_areas(ctx, lab, 64)
    for b in range(6):
        ra = ops.region_areas(lab[b], cnt[b])
        assert np.array_equal(areas[b, :cnt[b]], ra)
        assert largest[b] == (np.argmax(ra) + 1 if cnt[b] else 0)


def test_region_stats_match_cv2_moments_and_bounding_boxes(be, ctx):
    # raw moments are exact integers; cv2.moments returns the same numbers as doubles
    rng = np.random.default_rng(5)
    cases = [(synth_masks(6, 120, 200), 200), ((rng.random((3, 40, 70)) < 0.4), 70),
             (np.ones((1, 33, 97), bool), 97), (np.zeros((2, 9, 40), bool), 40)]
    for masks, W in cases:
        for conn in (4, 8):
            B = masks.shape[0]
            cap = 2048
            st, cnt, largest = hz.region_stats(ctx, ops.pack_bits(masks), W, cap, conn)
            for b in range(B):
                lab, n = ops.label(masks[b], conn)
                assert cnt[b] == n
                moms = ops.region_moments(lab, n)
                for l in range(n):
                    got = st[b, l]
                    for k, key in enumerate(('m00', 'm10', 'm01', 'm20', 'm11', 'm02')):
                        assert float(got[k]) == moms[l][key], (b, l, key)
                    ys, xs = np.nonzero(lab == l + 1)
                    assert tuple(got[6:10]) == (xs.min(), ys.min(), xs.max(), ys.max())
                    assert ops.find_bounding_box(lab == l + 1) == (got[6], got[7], got[8] - got[6] + 1, got[9] - got[7] + 1)
                areas = ops.region_areas(lab, n)
                assert largest[b] == (np.argmax(areas) + 1 if n else 0)
                assert (st[b, n:].view(np.uint8) == 0x5A).all()        # rows of absent regions are not touched
    # more regions than the caller's capacity: counts still report all of them
    checker = (np.indices((1, 16, 64)).sum(0) & 1).astype(bool)
    st, cnt, _ = hz.region_stats(ctx, ops.pack_bits(checker), 64, 8, 4)
    assert cnt[0] == 512 and (st[0, :, 0] == 1).all()


# ---- remaining filter bodies and temporal statistics -----------------------------------
def test_normalize_lut_rotate_time_difference(be, ctx):
    for (H, W) in sizes(be, [(9, 37), (12, 64)], [(480, 640)]):
        g = rng_frames(H, (3, H, W))
        c = rng_frames(W, (3, H, W, 3))
        # FilterNormalize on uint8 is a table: build it with the reference expression, apply on the device
        for vmin, vmax in ((None, None), (40, 200)):
            ref = ops.normalize(list(g), vmin, vmax)
            lo = g[0].min() if vmin is None else vmin
            hi = g[0].max() if vmax is None else vmax
            table = ops.normalize([np.arange(256, dtype=np.uint8)], lo, hi)[0]
            assert np.array_equal(hz.lut(ctx, g, table), ref)
        for k in range(4):
            assert np.array_equal(hz.rot90(ctx, g, k), np.stack([ops.rotate(f, 90 * k) for f in g])), k
            assert np.array_equal(hz.rot90(ctx, c, k), np.stack([ops.rotate(f, 90 * k) for f in c])), k
        d = hz.time_diff(ctx, g)
        assert d.dtype == np.int16
        assert np.array_equal(d, np.stack([ops.time_difference(g[t + 1], g[t]) for t in range(2)]))
        assert np.array_equal(hz.time_diff(ctx, c), np.stack([ops.time_difference(c[t + 1], c[t]) for t in range(2)]))


def test_temporal_mean_and_std_float64(be, ctx):
    g = rng_frames(11, (9, 10, 33))
    mean_ref = ops.measure_mean(list(g))
    mu, _ = hz.mean_update(ctx, g[:4], np.zeros((10, 33)))
    mu, _ = hz.mean_update(ctx, g[4:], mu, n0=4)                     # continuation
    assert np.array_equal(mu.view(np.uint64), mean_ref.view(np.uint64))   # float64, bit for bit
    m_ref, s_ref = ops.measure_mean_std(list(g))
    mu2, m2 = hz.mean_update(ctx, g, np.zeros((10, 33)), np.zeros((10, 33)))
    assert np.array_equal(mu2.view(np.uint64), m_ref.view(np.uint64))
    assert np.array_equal(np.sqrt(m2 / 8).view(np.uint64), s_ref.view(np.uint64))


# ---- the whole chain ---------------------------------------------------------------------------------
@pytest.mark.parametrize('fuse', [False, True])
def test_chain_run_matches_oracle_stagewise_and_end_to_end(be, ctx, fuse):
    W, H, T = (200, 120, 8) if be.name == "emu" else (640, 480, 24)
    fr = synth.make_frames(0, 0, T, W, H, 6)
    ref = ops.chain(fr)
    want = ('blur', 'mask', 'morph', 'labels') if fuse else ('mono', 'blur', 'mask', 'morph', 'labels')
    cut = T // 2 + 1
    a = hz.chain(ctx, fr[:cut], fuse=fuse, want=want)
    b = hz.chain(ctx, fr[cut:], fuse=fuse, want=want, bg0=a['bg'])          # second batch continues the model
    for k in want:
        got = np.concatenate([a[k], b[k]])
        exp = ops.pack_bits(ref[k]) if k in ('mask', 'morph') else ref[k]
        assert np.array_equal(got, exp), k
    assert np.array_equal(np.concatenate([a['counts'], b['counts']]), ref['counts'])
    assert np.array_equal(b['bg'].view(np.uint32), ref['bg'].view(np.uint32))
    assert ref['counts'][1:].min() >= 1


def test_chain_golden_ragged(be, ctx):
    out = hz.chain(ctx, GOLD['r_frames'], sigma=3, alpha=0.1, thr=12, morph_op='close', shape='ellipse', k=5, connectivity=8)
    assert np.array_equal(out['labels'], GOLD['r_labels'])
    assert np.array_equal(out['counts'], GOLD['r_counts'])
    assert np.array_equal(out['blur'], GOLD['r_blur'])
