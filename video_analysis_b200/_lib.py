"""
ctypes binding of libva_b200.so (include/va_b200.h).

There is deliberately no fallback: if the CUDA library has not been built, or a
call fails, the product raises.  `bind()` only declares signatures on an already
opened library object; `load()` is the one place the product opens a library and
it only ever opens csrc/libva_b200.so.
"""

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_uint32, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'csrc', 'libva_b200.so')

VA_OK = 0
VA_ERR_INVALID = -1
VA_ERR_CUDA = -2
VA_ERR_NOMEM = -3
VA_ERR_CAPACITY = -4
VA_ERR_UNSUPPORTED = -5

MORPH_OPS = {'erode': 0, 'dilate': 1, 'open': 2, 'close': 3}
SE_SHAPES = {'rect': 0, 'cross': 1, 'ellipse': 2}
MONO_MEAN = -1


class ChainDesc(ctypes.Structure):
    _fields_ = [('w', c_int), ('h', c_int), ('batch', c_int), ('mono_mode', c_int),
                ('sigma', c_double), ('alpha', c_float), ('thr', c_float),
                ('first_frame_inits', c_int), ('morph_op', c_int), ('morph_shape', c_int),
                ('morph_kx', c_int), ('morph_ky', c_int), ('connectivity', c_int),
                ('fuse_luma_blur', c_int)]


class ChainIO(ctypes.Structure):
    _fields_ = [('rgb', c_void_p), ('rgb_pitch', c_size_t), ('rgb_fstride', c_size_t),
                ('bg', c_void_p), ('bg_pitch_e', c_size_t),
                ('mono', c_void_p), ('mono_pitch', c_size_t), ('mono_fstride', c_size_t),
                ('blur', c_void_p), ('blur_pitch', c_size_t), ('blur_fstride', c_size_t),
                ('mask', c_void_p), ('mask_pitch_w', c_size_t), ('mask_fstride_w', c_size_t),
                ('morph', c_void_p), ('morph_pitch_w', c_size_t), ('morph_fstride_w', c_size_t),
                ('labels', c_void_p), ('labels_pitch_e', c_size_t), ('labels_fstride_e', c_size_t),
                ('counts', c_void_p)]


_IMG = [c_void_p, c_size_t, c_size_t]      # pointer, pitch, frame stride

# name -> (restype, argtypes); every symbol include/va_b200.h declares
SIGNATURES = {
    'va_version': (c_int, []),
    'va_status_string': (c_char_p, [c_int]),
    'va_create': (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int]),
    'va_destroy': (c_int, [c_void_p]),
    'va_reserve': (c_int, [c_void_p, c_int, c_int, c_int]),
    'va_last_error': (c_char_p, [c_void_p]),
    'va_launch_count': (c_longlong, [c_void_p]),
    'va_luma_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int]),
    'va_copy2d_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int]),
    'va_gauss_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_double]),
    'va_gauss_taps': (c_int, [c_double, ctypes.POINTER(c_int), c_int]),
    'va_luma_gauss_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_double]),
    'va_resize_half_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int]),
    'va_ema_diff_thresh': (c_int, [c_void_p, c_void_p] + _IMG + [c_void_p, c_size_t] + _IMG +
                           [c_int, c_int, c_int, c_float, c_float, c_int]),
    'va_threshold_bits': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int]),
    'va_unpack_bits_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int]),
    'va_pack_bits_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int]),
    'va_morph_bits': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    'va_label_bits': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_void_p, c_int, c_int, c_int, c_int]),
    'va_region_areas': (c_int, [c_void_p, c_void_p] + _IMG + [c_void_p, c_int, c_void_p, c_int, c_int, c_int]),
    'va_region_stats': (c_int, [c_void_p, c_void_p] + _IMG + [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
    'va_apply_mask_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + _IMG + [c_int, c_int, c_int, c_int]),
    'va_ema_partial': (c_int, [c_void_p, c_void_p] + _IMG + [c_void_p, c_size_t, c_int, c_int, c_int, c_float, c_int]),
    'va_ema_fold': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_float]),
    'va_synth_rgb': (c_int, [c_void_p, c_void_p] + _IMG + [c_int, c_int, c_int, c_int, c_uint32, c_void_p, c_int]),
    'va_resize_area_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_int, c_int]),
    'va_resize_nearest_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_int, c_int]),
    'va_resize_area_any_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_int, c_int]),
    'va_resize_linear_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_int, c_int]),
    'va_resize_cubic_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_int, c_int]),
    'va_highlight_mask_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'va_resize_lanczos4_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_int, c_int]),
    'va_label_forest': (c_int, [c_void_p, c_void_p] + _IMG + [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    'va_label_write': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int]),
    'va_luma_crop_multi_u8': (c_int, [c_void_p, c_void_p] + _IMG + [c_int, c_int] + _IMG + [c_int, c_int, c_int, c_int, c_void_p]),
    'va_streams_threshold_bits': (c_int, [c_void_p, c_void_p] + _IMG + [c_int, c_int] + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'va_label_write_i16': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int]),
    'va_lut_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_void_p]),
    'va_time_diff_i16': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int]),
    'va_rot90_u8': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_int, c_int]),
    'va_mean_update_f64': (c_int, [c_void_p, c_void_p] + _IMG + [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_longlong]),
    'va_label_export_chunks': (c_int, [c_void_p, c_void_p] + _IMG + _IMG + [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int]),
    'va_host_densify_chunks': (c_int, [c_void_p, c_size_t, c_size_t, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int]),
    'va_chain_run': (c_int, [c_void_p, c_void_p, ctypes.POINTER(ChainDesc), ctypes.POINTER(ChainIO)]),
}


class VAError(RuntimeError):
    """ a libva_b200 call failed for a reason that is not the caller's arguments """


def bind(cdll):
    """ declare restype / argtypes for every exported symbol; raises AttributeError
    if the library does not export one of them """
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(cdll, name)
        fn.restype = restype
        fn.argtypes = argtypes
    return cdll


_lib = None


def load():
    """ open csrc/libva_b200.so (the CUDA library).  No fallback. """
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VAError('%s is missing: build it with `python -m video_analysis_b200.build` '
                          '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
        _lib = bind(ctypes.CDLL(LIB_PATH))
    return _lib


def check(lib, ctx, status):
    """ map a va_status to the exception the reference raises for that situation """
    if status == VA_OK:
        return
    msg = lib.va_last_error(ctx).decode() if ctx else ''
    msg = msg or lib.va_status_string(status).decode()
    if status == VA_ERR_INVALID:
        raise ValueError(msg)
    if status == VA_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if status in (VA_ERR_NOMEM, VA_ERR_CAPACITY):
        raise MemoryError(msg)
    raise VAError(msg)
