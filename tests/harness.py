"""
Test harness: calls the C ABI of include/va_b200.h on NumPy inputs and returns
NumPy outputs, on one of two backends

  cuda  csrc/libva_b200.so on cuda:0 (tests marked `gpu`; the parity gate)
  emu   tests/emu/libva_b200_emu.so, the same kernel sources compiled for the CPU
        with the thread-emulation shim (development-time logic check only)

Nothing here is product code.
"""

import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from video_analysis_b200 import _lib as valib  # noqa: E402


class EmuBackend(object):
    name = 'emu'

    def __init__(self):
        from tests.emu import build_emu
        self.lib = valib.bind(ctypes.CDLL(build_emu.build()))
        self.stream = None

    def empty(self, shape, dtype, fill=None):
        a = np.empty(shape, dtype)
        a[...] = 0xCD if fill is None else fill
        return a

    def to_dev(self, a):
        return np.array(a, copy=True, order='C')

    def to_host(self, d):
        return np.array(d, copy=True)

    def ptr(self, d, offset_bytes=0):
        return d.ctypes.data + offset_bytes

    def sync(self):
        pass


class CudaBackend(object):
    name = 'cuda'

    def __init__(self):
        import torch
        self.torch = torch
        self.lib = valib.load()
        self.dev = torch.device('cuda:0')
        self.stream = None

    def empty(self, shape, dtype, fill=None):
        a = np.empty(shape, dtype)
        a[...] = 0xCD if fill is None else fill
        return self.to_dev(a)

    def to_dev(self, a):
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint32:
            return self.torch.from_numpy(a.view(np.int32)).to(self.dev)
        return self.torch.from_numpy(a).to(self.dev)

    def to_host(self, d):
        self.torch.cuda.synchronize()
        return d.cpu().numpy()

    def ptr(self, d, offset_bytes=0):
        return d.data_ptr() + offset_bytes

    def sync(self):
        self.torch.cuda.synchronize()


class Ctx(object):
    def __init__(self, be, max_w, max_h, max_batch):
        self.be = be
        self.lib = be.lib
        h = ctypes.c_void_p()
        rc = self.lib.va_create(ctypes.byref(h), 0, max_w, max_h, max_batch)
        assert rc == 0, rc
        self.h = h

    def check(self, rc):
        valib.check(self.lib, self.h, rc)

    def close(self):
        if self.h:
            self.lib.va_destroy(self.h)
            self.h = None


def _round_up(v, m):
    return (v + m - 1) // m * m


class Img(object):
    """ a batch of u8 / u32 / i32 / f32 images on the backend with explicit pitch """

    def __init__(self, be, batch, h, row_elems, dtype, pitch_elems=None, data=None, offset_elems=0, fill=None):
        self.be = be
        self.dtype = np.dtype(dtype)
        self.batch, self.h, self.row = batch, h, row_elems
        self.pitch = pitch_elems or row_elems
        self.offset = offset_elems
        host = np.empty((batch, h, self.pitch), self.dtype)
        host.view(np.uint8)[...] = 0xCD if fill is None else fill
        if data is not None:
            host[:, :, offset_elems:offset_elems + row_elems] = np.asarray(data).reshape(batch, h, row_elems)
        self.dev = be.to_dev(host)

    @property
    def ptr(self):
        return self.be.ptr(self.dev, self.offset * self.dtype.itemsize)

    @property
    def fstride(self):
        return self.h * self.pitch

    def get(self):
        host = self.be.to_host(self.dev)
        if self.dtype == np.uint32:
            host = host.view(np.uint32)
        return host.reshape(self.batch, self.h, self.pitch)[:, :, self.offset:self.offset + self.row].copy()

    def raw(self):
        host = self.be.to_host(self.dev)
        if self.dtype == np.uint32:
            host = host.view(np.uint32)
        return host.reshape(self.batch, self.h, self.pitch)


# ---------------------------------------------------------------------------------------
# one wrapper per entry point: NumPy in, NumPy out
# ---------------------------------------------------------------------------------------
def luma(ctx, frames, mode=-1, in_pad=0, in_off=0, out_pad=0):
    be = ctx.be
    B, H, W, _ = frames.shape
    src = Img(be, B, H, 3 * W, np.uint8, 3 * W + in_pad + in_off, frames.reshape(B, H, 3 * W), in_off)
    dst = Img(be, B, H, W, np.uint8, W + out_pad)
    ctx.check(ctx.lib.va_luma_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch, dst.fstride,
                                 W, H, B, mode))
    raw = dst.raw()
    assert (raw[:, :, W:] == 0xCD).all(), 'luma wrote outside its rows'
    return dst.get()


def crop_luma(ctx, frames, rect, mode=-1):
    """ crop by pointer offset into a larger frame, then luma """
    be = ctx.be
    B, H, W, _ = frames.shape
    left, top, w, h = rect
    src = Img(be, B, H, 3 * W, np.uint8, data=frames.reshape(B, H, 3 * W))
    dst = Img(be, B, h, w, np.uint8)
    ptr = be.ptr(src.dev, top * 3 * W + 3 * left)
    ctx.check(ctx.lib.va_luma_u8(ctx.h, be.stream, ptr, 3 * W, H * 3 * W, dst.ptr, dst.pitch, dst.fstride, w, h, B, mode))
    return dst.get()


def copy2d(ctx, frames, rect):
    be = ctx.be
    B, H = frames.shape[:2]
    rowb = int(np.prod(frames.shape[2:]))
    ch = frames.shape[3] if frames.ndim == 4 else 1
    left, top, w, h = rect
    src = Img(be, B, H, rowb, np.uint8, data=frames.reshape(B, H, rowb))
    dst = Img(be, B, h, w * ch, np.uint8)
    ptr = be.ptr(src.dev, top * rowb + left * ch)
    ctx.check(ctx.lib.va_copy2d_u8(ctx.h, be.stream, ptr, rowb, H * rowb, dst.ptr, dst.pitch, dst.fstride, w * ch, h, B))
    out = dst.get()
    return out.reshape(B, h, w, ch) if frames.ndim == 4 else out


def gauss(ctx, frames, sigma, in_pad=0, out_pad=0):
    be = ctx.be
    if frames.ndim == 3:
        B, H, W = frames.shape
        ch = 1
    else:
        B, H, W, ch = frames.shape
    src = Img(be, B, H, W * ch, np.uint8, W * ch + in_pad, frames.reshape(B, H, W * ch))
    dst = Img(be, B, H, W * ch, np.uint8, W * ch + out_pad)
    ctx.check(ctx.lib.va_gauss_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch, dst.fstride,
                                  W, H, ch, B, float(sigma)))
    raw = dst.raw()
    assert (raw[:, :, W * ch:] == 0xCD).all(), 'gauss wrote outside its rows'
    return dst.get().reshape(frames.shape)


def luma_gauss(ctx, frames, sigma, mode=-1):
    be = ctx.be
    B, H, W, _ = frames.shape
    src = Img(be, B, H, 3 * W, np.uint8, data=frames.reshape(B, H, 3 * W))
    dst = Img(be, B, H, W, np.uint8)
    ctx.check(ctx.lib.va_luma_gauss_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch,
                                       dst.fstride, W, H, B, mode, float(sigma)))
    return dst.get()


def gauss_taps(lib, sigma):
    buf = (ctypes.c_int * 1024)()
    n = lib.va_gauss_taps(float(sigma), buf, 1024)
    assert n > 0, n
    return np.array(buf[:n], dtype=np.int64)


def resize_half(ctx, frames, in_pad=0):
    be = ctx.be
    if frames.ndim == 3:
        B, H, W = frames.shape
        ch = 1
    else:
        B, H, W, ch = frames.shape
    src = Img(be, B, H, W * ch, np.uint8, W * ch + in_pad, frames.reshape(B, H, W * ch))
    dst = Img(be, B, H // 2, (W // 2) * ch, np.uint8)
    ctx.check(ctx.lib.va_resize_half_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch,
                                        dst.fstride, W, H, ch, B))
    out = dst.get()
    return out if frames.ndim == 3 else out.reshape(B, H // 2, W // 2, ch)


def resize_area(ctx, frames, kx, ky):
    be = ctx.be
    B, H, W = frames.shape[:3]
    ch = 1 if frames.ndim == 3 else frames.shape[3]
    src = Img(be, B, H, W * ch, np.uint8, W * ch + 5, frames.reshape(B, H, W * ch))
    dst = Img(be, B, H // ky, (W // kx) * ch, np.uint8, (W // kx) * ch + 3)
    ctx.check(ctx.lib.va_resize_area_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch, dst.fstride,
                                        W, H, ch, B, kx, ky))
    assert (dst.raw()[:, :, (W // kx) * ch:] == 0xCD).all()
    out = dst.get()
    return out if frames.ndim == 3 else out.reshape(B, H // ky, W // kx, ch)


def resize_nearest(ctx, frames, dw, dh):
    be = ctx.be
    B, H, W = frames.shape[:3]
    ch = 1 if frames.ndim == 3 else frames.shape[3]
    src = Img(be, B, H, W * ch, np.uint8, data=frames.reshape(B, H, W * ch))
    dst = Img(be, B, dh, dw * ch, np.uint8)
    ctx.check(ctx.lib.va_resize_nearest_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch, dst.fstride,
                                           W, H, dw, dh, ch, B))
    out = dst.get()
    return out if frames.ndim == 3 else out.reshape(B, dh, dw, ch)


def resize_to(ctx, frames, dw, dh, how):
    """ how: 'area_any' | 'linear' """
    be = ctx.be
    B, H, W = frames.shape[:3]
    ch = 1 if frames.ndim == 3 else frames.shape[3]
    src = Img(be, B, H, W * ch, np.uint8, W * ch + 5, frames.reshape(B, H, W * ch))
    dst = Img(be, B, dh, dw * ch, np.uint8, dw * ch + 3)
    fn = getattr(ctx.lib, 'va_resize_%s_u8' % how)
    ctx.check(fn(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch, dst.fstride, W, H, dw, dh, ch, B))
    assert (dst.raw()[:, :, dw * ch:] == 0xCD).all()
    out = dst.get()
    return out if frames.ndim == 3 else out.reshape(B, dh, dw, ch)


def pack_bits_np(masks):
    """ (B, H, W) nonzero = set -> (B, H, ceil(W / 32)) uint32 words, bit i of word j = pixel 32 j + i """
    B, H, W = masks.shape
    Wp = (W + 31) // 32
    m = np.zeros((B, H, Wp * 32), np.uint8)
    m[:, :, :W] = masks != 0
    by = np.packbits(m, axis=2, bitorder='little')
    return np.ascontiguousarray(by).view('<u4').reshape(B, H, Wp)


def highlight_mask(ctx, frames, masks, channel, table, in_pad=0):
    """ frames (B, H, W[, 3]) u8, masks (B, H, W) nonzero = set """
    be = ctx.be
    B, H, W = frames.shape[:3]
    ch = 1 if frames.ndim == 3 else frames.shape[3]
    src = Img(be, B, H, W * ch, np.uint8, W * ch + in_pad, frames.reshape(B, H, W * ch))
    bits = Img(be, B, H, mask_words(W), np.uint32, mask_words(W) + 1, pack_bits_np(masks))
    dst = Img(be, B, H, W * ch, np.uint8, W * ch + in_pad)
    tab = np.ascontiguousarray(table, np.uint8)
    ctx.check(ctx.lib.va_highlight_mask_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, bits.ptr, bits.pitch, bits.fstride,
                                           dst.ptr, dst.pitch, dst.fstride, W, H, ch, B, channel, tab.ctypes.data))
    out = dst.get()
    return out if frames.ndim == 3 else out.reshape(B, H, W, ch)


def luma_crop_multi(ctx, frames, xy, w, h, mode=-1, in_pad=0, out_pad=3):
    """ frames (S, H, W, 3); xy (S, 2) left, top """
    be = ctx.be
    S, H, W, _ = frames.shape
    src = Img(be, S, H, 3 * W, np.uint8, 3 * W + in_pad, frames.reshape(S, H, 3 * W))
    tab = Img(be, 1, 1, 2 * S, np.int32, data=np.asarray(xy, np.int32).reshape(1, 1, 2 * S))
    dst = Img(be, S, h, w, np.uint8, w + out_pad)
    ctx.check(ctx.lib.va_luma_crop_multi_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, W, H,
                                            dst.ptr, dst.pitch, dst.fstride, w, h, S, mode, tab.ptr))
    assert (dst.raw()[:, :, w:] == 0xCD).all()
    return dst.get()


def streams_threshold(ctx, frames, xy, w, h, masks, thr, mode=-1, in_pad=0):
    """ frames (S, H, W, 3); xy (S, 2); masks None, (h, w) or (S, h, w) u8 -> packed words (S, h, ceil(w / 32)) """
    be = ctx.be
    S, H, W, _ = frames.shape
    src = Img(be, S, H, 3 * W, np.uint8, 3 * W + in_pad, frames.reshape(S, H, 3 * W))
    tab = Img(be, 1, 1, 2 * S, np.int32, data=np.asarray(xy, np.int32).reshape(1, 1, 2 * S))
    margs = (None, 0, 0)
    if masks is not None:
        m3 = masks if masks.ndim == 3 else masks[None]
        mk = Img(be, m3.shape[0], h, w, np.uint8, w + 16, m3)
        margs = (mk.ptr, mk.pitch, mk.fstride if masks.ndim == 3 else 0)
    dst = Img(be, S, h, mask_words(w), np.uint32, mask_words(w) + 1)
    ctx.check(ctx.lib.va_streams_threshold_bits(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, W, H, *margs,
                                                dst.ptr, dst.pitch, dst.fstride, w, h, S, mode, thr, tab.ptr))
    return dst.get()


def mask_words(W):
    return (W + 31) // 32


def ema_diff_thresh(ctx, frames, alpha, thr, bg0=None, pad=0):
    """ returns (packed masks [B,H,Wp] uint32, bg float32 [H,W]) """
    be = ctx.be
    B, H, W = frames.shape
    src = Img(be, B, H, W, np.uint8, W + pad, frames)
    first = bg0 is None
    bg = Img(be, 1, H, W, np.float32, W + (pad // 4) * 4, None if first else np.asarray(bg0, np.float32)[None])
    Wp = mask_words(W)
    msk = Img(be, B, H, Wp, np.uint32, Wp + (1 if pad else 0))
    ctx.check(ctx.lib.va_ema_diff_thresh(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, bg.ptr, bg.pitch,
                                         msk.ptr, msk.pitch, msk.fstride, W, H, B, float(alpha), float(thr), int(first)))
    return msk.get(), bg.get()[0]


def threshold_bits(ctx, frames, thr, pad=0):
    be = ctx.be
    B, H, W = frames.shape
    src = Img(be, B, H, W, np.uint8, W + pad, frames)
    Wp = mask_words(W)
    msk = Img(be, B, H, Wp, np.uint32)
    ctx.check(ctx.lib.va_threshold_bits(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, msk.ptr, msk.pitch,
                                        msk.fstride, W, H, B, int(thr)))
    return msk.get()


def pack_bits(ctx, masks, pad=0):
    be = ctx.be
    B, H, W = masks.shape
    src = Img(be, B, H, W, np.uint8, W + pad, masks)
    Wp = mask_words(W)
    msk = Img(be, B, H, Wp, np.uint32)
    ctx.check(ctx.lib.va_pack_bits_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, msk.ptr, msk.pitch,
                                      msk.fstride, W, H, B))
    return msk.get()


def unpack_bits(ctx, words, W, pad=0):
    be = ctx.be
    B, H, Wp = words.shape
    msk = Img(be, B, H, Wp, np.uint32, data=words)
    dst = Img(be, B, H, W, np.uint8, W + pad)
    ctx.check(ctx.lib.va_unpack_bits_u8(ctx.h, be.stream, msk.ptr, msk.pitch, msk.fstride, dst.ptr, dst.pitch,
                                        dst.fstride, W, H, B))
    raw = dst.raw()
    assert (raw[:, :, W:] == 0xCD).all()
    return dst.get()


def morph(ctx, words, W, op, shape='rect', ksize=3):
    be = ctx.be
    B, H, Wp = words.shape
    kx, ky = (ksize, ksize) if np.isscalar(ksize) else ksize
    src = Img(be, B, H, Wp, np.uint32, data=words)
    dst = Img(be, B, H, Wp, np.uint32)
    ctx.check(ctx.lib.va_morph_bits(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch, dst.fstride,
                                    W, H, B, valib.MORPH_OPS[op], valib.SE_SHAPES[shape], int(kx), int(ky)))
    return dst.get()


def label(ctx, words, W, connectivity=4, lab_pad=0):
    """ returns (labels int32 [B,H,W], counts int32 [B]) """
    be = ctx.be
    B, H, Wp = words.shape
    src = Img(be, B, H, Wp, np.uint32, data=words)
    lab = Img(be, B, H, W, np.int32, W + lab_pad)
    cnt = Img(be, 1, 1, B, np.int32)
    ctx.check(ctx.lib.va_label_bits(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, lab.ptr, lab.pitch, lab.fstride,
                                    cnt.ptr, W, H, B, connectivity))
    return lab.get(), cnt.get()[0, 0]


def label_export_dense(ctx, words, W, connectivity=4, lab_pad=0, cap=None, reuse=None, threads=3, use_runs=True):
    """ va_label_bits, va_label_export_chunks into page-locked host memory (plain memory on the emulator) and
    va_host_densify_chunks: returns (dense labels [B,H,W], n_chunks [B], state, dense copy, n_runs [B]) -- `state` =
    (dense buffer, dirty ids, dirty counts) to be passed back as `reuse` so that the next call rebuilds into the same
    buffer.  use_runs: single-run chunks travel as 16-byte records (n_chunks then counts the other chunks only) """
    be = ctx.be
    B, H, Wp = words.shape
    src = Img(be, B, H, Wp, np.uint32, data=words)
    pitch = W + lab_pad + ((W + lab_pad) & 1)
    lab = Img(be, B, H, W, np.int32, pitch)
    cnt = Img(be, 1, 1, B, np.int32)
    ctx.check(ctx.lib.va_label_bits(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, lab.ptr, lab.pitch, lab.fstride,
                                    cnt.ptr, W, H, B, connectivity))
    cpr = (W + 63) // 64
    cap = cap or cpr * H
    if be.name == 'cuda':
        t = be.torch
        ids_t = t.empty((B, cap), dtype=t.int32, pin_memory=True)
        data_t = t.empty((B, cap, 64), dtype=t.int32, pin_memory=True)
        n_t = t.empty((B,), dtype=t.int32, pin_memory=True)
        runs_t = t.empty((B, cap, 4), dtype=t.int32, pin_memory=True)
        nr_t = t.zeros((B,), dtype=t.int32, pin_memory=True)
        ids, data, n, runs, nr = ids_t.numpy(), data_t.numpy(), n_t.numpy(), runs_t.numpy(), nr_t.numpy()
        keep = (ids_t, data_t, n_t, runs_t, nr_t)
    else:
        ids, data, n = np.empty((B, cap), np.int32), np.empty((B, cap, 64), np.int32), np.empty((B,), np.int32)
        runs, nr = np.empty((B, cap, 4), np.int32), np.zeros((B,), np.int32)
        keep = None
    rp, nrp = (runs.ctypes.data, nr.ctypes.data) if use_runs else (None, None)
    ndev = Img(be, 1, 1, B, np.int32)
    ctx.check(ctx.lib.va_label_export_chunks(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, lab.ptr, lab.pitch, lab.fstride,
                                             W, H, B, ids.ctypes.data, data.ctypes.data, n.ctypes.data, ndev.ptr, rp, nrp, cap))
    be.sync()
    assert np.array_equal(ndev.get()[0, 0], n)
    if (n > cap).any() or (nr > cap).any():   # more chunks than the export buffers hold: the caller copies densely instead
        rc = ctx.lib.va_host_densify_chunks(lab.get().ctypes.data, pitch, H * pitch, W, H, B, ids.ctypes.data, data.ctypes.data,
                                            n.ctypes.data, rp, nrp, cap, ids.ctypes.data, n.ctypes.data, 1)
        assert rc == valib.VA_ERR_CAPACITY
        return None, n.copy(), None, lab.get(), nr.copy()
    if reuse is None:
        dense = np.zeros((B, H, pitch), np.int32)
        dirty, n_dirty = np.zeros((B, cap), np.int32), np.zeros((B,), np.int32)
    else:
        dense, dirty, n_dirty = reuse
    rc = ctx.lib.va_host_densify_chunks(dense.ctypes.data, pitch, H * pitch, W, H, B, ids.ctypes.data, data.ctypes.data,
                                        n.ctypes.data, rp, nrp, cap, dirty.ctypes.data, n_dirty.ctypes.data, threads)
    assert rc == 0, rc
    del keep
    return dense[:, :, :W].copy(), n.copy(), (dense, dirty, n_dirty), lab.get(), nr.copy()


def label_two_batches_split(ctx, words_a, words_b, W, connectivity=4):
    """ forest(a, slot 0), forest(b, slot 1), write(b, slot 1), write(a, slot 0): the two scratch sets are independent """
    be = ctx.be
    out = []
    imgs = []
    for slot, words in ((0, words_a), (1, words_b)):
        B, H, Wp = words.shape
        src = Img(be, B, H, Wp, np.uint32, data=words)
        lab = Img(be, B, H, W, np.int32, W + 4)
        cnt = Img(be, 1, 1, B, np.int32)
        ctx.check(ctx.lib.va_label_forest(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, cnt.ptr, W, H, B, connectivity, slot))
        imgs.append((slot, src, lab, cnt, B, H))
    for slot, src, lab, cnt, B, H in reversed(imgs):
        ctx.check(ctx.lib.va_label_write(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, lab.ptr, lab.pitch, lab.fstride,
                                         W, H, B, slot))
    for slot, src, lab, cnt, B, H in imgs:
        out.append((lab.get(), cnt.get()[0, 0]))
    return out


def label_i16(ctx, words, W, connectivity=4, lab_pad=0):
    """ forest + int16 label write -> (labels int16 [B, H, W], counts int32 [B]) """
    be = ctx.be
    B, H, Wp = words.shape
    src = Img(be, B, H, Wp, np.uint32, data=words)
    lab = Img(be, B, H, W, np.int16, W + lab_pad)
    cnt = Img(be, 1, 1, B, np.int32)
    ctx.check(ctx.lib.va_label_forest(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, cnt.ptr, W, H, B, connectivity, 0))
    ctx.check(ctx.lib.va_label_write_i16(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, lab.ptr, lab.pitch, lab.fstride,
                                         W, H, B, 0))
    return lab.get(), cnt.get()[0, 0]


def region_areas(ctx, labels, max_labels):
    be = ctx.be
    B, H, W = labels.shape
    lab = Img(be, B, H, W, np.int32, data=labels)
    areas = Img(be, 1, B, max_labels, np.int32)
    largest = Img(be, 1, 1, B, np.int32)
    ctx.check(ctx.lib.va_region_areas(ctx.h, be.stream, lab.ptr, lab.pitch, lab.fstride, areas.ptr, max_labels,
                                      largest.ptr, W, H, B))
    return areas.get()[0], largest.get()[0, 0]


def region_stats(ctx, words, W, max_regions, connectivity=4):
    """ returns (stats int64 [B, max_regions, 10], counts int32 [B], largest int32 [B]) """
    be = ctx.be
    B, H, Wp = words.shape
    src = Img(be, B, H, Wp, np.uint32, data=words)
    st = Img(be, 1, B, max_regions * 10, np.int64, fill=0x5A)
    cnt = Img(be, 1, 1, B, np.int32)
    largest = Img(be, 1, 1, B, np.int32)
    ctx.check(ctx.lib.va_region_stats(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, st.ptr, max_regions, cnt.ptr,
                                      largest.ptr, W, H, B, connectivity))
    return st.get()[0].reshape(B, max_regions, 10), cnt.get()[0, 0], largest.get()[0, 0]


def apply_mask(ctx, frames, mask):
    be = ctx.be
    if frames.ndim == 3:
        B, H, W = frames.shape
        ch = 1
    else:
        B, H, W, ch = frames.shape
    src = Img(be, B, H, W * ch, np.uint8, data=frames.reshape(B, H, W * ch))
    dst = Img(be, B, H, W * ch, np.uint8)
    static = mask.ndim == 2
    m = Img(be, 1 if static else B, H, W, np.uint8, data=mask[None] if static else mask)
    ctx.check(ctx.lib.va_apply_mask_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, m.ptr, m.pitch,
                                       0 if static else m.fstride, dst.ptr, dst.pitch, dst.fstride, W, H, ch, B))
    return dst.get().reshape(frames.shape)


def ema_partial(ctx, frames, alpha, S0=None):
    be = ctx.be
    B, H, W = frames.shape
    src = Img(be, B, H, W, np.uint8, data=frames)
    S = Img(be, 1, H, W, np.float32, data=None if S0 is None else np.asarray(S0, np.float32)[None])
    ctx.check(ctx.lib.va_ema_partial(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, S.ptr, S.pitch, W, H, B,
                                     float(alpha), 0 if S0 is None else 1))
    return S.get()[0]


def ema_fold(ctx, carry, S, scale):
    be = ctx.be
    H, W = carry.shape
    c = Img(be, 1, H, W, np.float32, data=carry[None])
    s = Img(be, 1, H, W, np.float32, data=S[None])
    ctx.check(ctx.lib.va_ema_fold(ctx.h, be.stream, c.ptr, s.ptr, c.pitch, W, H, float(scale)))
    return c.get()[0]


def synth(ctx, seed, t0, n, W, H, blobs):
    be = ctx.be
    dst = Img(be, n, H, 3 * W, np.uint8)
    tab = np.ascontiguousarray(blobs, dtype=np.int32)
    ctx.check(ctx.lib.va_synth_rgb(ctx.h, be.stream, dst.ptr, dst.pitch, dst.fstride, W, H, t0, n, seed,
                                   tab.ctypes.data, len(tab)))
    return dst.get().reshape(n, H, W, 3)


def lut(ctx, frames, table):
    be = ctx.be
    B, H = frames.shape[:2]
    rowb = int(np.prod(frames.shape[2:]))
    src = Img(be, B, H, rowb, np.uint8, data=frames.reshape(B, H, rowb))
    dst = Img(be, B, H, rowb, np.uint8)
    tab = np.ascontiguousarray(table, dtype=np.uint8)
    ctx.check(ctx.lib.va_lut_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch, dst.fstride,
                                rowb, H, B, tab.ctypes.data))
    return dst.get().reshape(frames.shape)


def time_diff(ctx, frames):
    """ frames (T, H, ...) -> int16 (T-1, H, ...) """
    be = ctx.be
    T, H = frames.shape[:2]
    rowe = int(np.prod(frames.shape[2:]))
    src = Img(be, T, H, rowe, np.uint8, data=frames.reshape(T, H, rowe))
    dst = Img(be, T - 1, H, rowe, np.int16)
    ctx.check(ctx.lib.va_time_diff_i16(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch, dst.fstride,
                                       rowe, H, T - 1))
    return dst.get().reshape((T - 1,) + frames.shape[1:])


def rot90(ctx, frames, k):
    be = ctx.be
    if frames.ndim == 3:
        B, H, W = frames.shape
        ch = 1
    else:
        B, H, W, ch = frames.shape
    oh, ow = (W, H) if k & 1 else (H, W)
    src = Img(be, B, H, W * ch, np.uint8, data=frames.reshape(B, H, W * ch))
    dst = Img(be, B, oh, ow * ch, np.uint8)
    ctx.check(ctx.lib.va_rot90_u8(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, dst.ptr, dst.pitch, dst.fstride,
                                  W, H, ch, B, k))
    out = dst.get()
    return out.reshape(B, oh, ow) if frames.ndim == 3 else out.reshape(B, oh, ow, ch)


def mean_update(ctx, frames, mean, m2=None, n0=0):
    be = ctx.be
    B, H = frames.shape[:2]
    rowe = int(np.prod(frames.shape[2:]))
    src = Img(be, B, H, rowe, np.uint8, data=frames.reshape(B, H, rowe))
    mu = Img(be, 1, H, rowe, np.float64, data=np.asarray(mean, np.float64).reshape(1, H, rowe))
    q = None if m2 is None else Img(be, 1, H, rowe, np.float64, data=np.asarray(m2, np.float64).reshape(1, H, rowe))
    ctx.check(ctx.lib.va_mean_update_f64(ctx.h, be.stream, src.ptr, src.pitch, src.fstride, mu.ptr,
                                         None if q is None else q.ptr, mu.pitch, rowe, H, B, n0))
    shape = frames.shape[1:]
    return mu.get()[0].reshape(shape), (None if q is None else q.get()[0].reshape(shape))


def chain(ctx, frames, sigma=2.0, alpha=0.05, thr=25.0, morph_op='open', shape='rect', k=3, connectivity=4,
          bg0=None, fuse=False, want=('mono', 'blur', 'mask', 'morph', 'labels')):
    be = ctx.be
    B, H, W, _ = frames.shape
    Wp = mask_words(W)
    src = Img(be, B, H, 3 * W, np.uint8, data=frames.reshape(B, H, 3 * W))
    first = bg0 is None
    bg = Img(be, 1, H, W, np.float32, data=None if first else np.asarray(bg0, np.float32)[None])
    bufs = {
        'mono': Img(be, B, H, W, np.uint8), 'blur': Img(be, B, H, W, np.uint8),
        'mask': Img(be, B, H, Wp, np.uint32), 'morph': Img(be, B, H, Wp, np.uint32),
        'labels': Img(be, B, H, W, np.int32),
    }
    cnt = Img(be, 1, 1, B, np.int32)
    d = valib.ChainDesc(w=W, h=H, batch=B, mono_mode=-1, sigma=float(sigma), alpha=alpha, thr=thr,
                        first_frame_inits=int(first), morph_op=valib.MORPH_OPS[morph_op] if morph_op else -1,
                        morph_shape=valib.SE_SHAPES[shape], morph_kx=k, morph_ky=k, connectivity=connectivity,
                        fuse_luma_blur=int(fuse))
    io = valib.ChainIO()
    io.rgb, io.rgb_pitch, io.rgb_fstride = src.ptr, src.pitch, src.fstride
    io.bg, io.bg_pitch_e = bg.ptr, bg.pitch
    for name in want:
        b = bufs[name]
        setattr(io, name, b.ptr)
        if name in ('mono', 'blur'):
            setattr(io, name + '_pitch', b.pitch)
            setattr(io, name + '_fstride', b.fstride)
        elif name in ('mask', 'morph'):
            setattr(io, name + '_pitch_w', b.pitch)
            setattr(io, name + '_fstride_w', b.fstride)
        else:
            io.labels_pitch_e, io.labels_fstride_e = b.pitch, b.fstride
    io.counts = cnt.ptr
    ctx.check(ctx.lib.va_chain_run(ctx.h, be.stream, ctypes.byref(d), ctypes.byref(io)))
    out = {name: bufs[name].get() for name in want}
    out['counts'] = cnt.get()[0, 0]
    out['bg'] = bg.get()[0]
    return out
