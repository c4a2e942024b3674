"""
ORACLE (test infrastructure) -- CPU restatement of the filter -> segment path.

Every function issues the library call the reference issues at the cited line
(paths relative to the reference root).  Frames are NumPy arrays ``(H, W, 3)``
or ``(H, W)``, ``size`` tuples are ``(width, height)`` as in
video/io/base.py:119-125.

Ops that do not exist in the reference snapshot (EMA background, threshold,
open/close, apply-mask) are marked ADOPTED: their semantics are the ones
SURVEY.md section 8c fixes; the reference pins nothing for them.
"""

import cv2
import numpy as np
from scipy import ndimage

# video/filters.py:32-34
COLOR_CHANNELS = {'blue': 0, 'b': 0, 0: 0,
                  'green': 1, 'g': 1, 1: 1,
                  'red': 2, 'r': 2, 2: 2}


# --------------------------------------------------------------------------
# monochrome -- video/filters.py:348-374
# --------------------------------------------------------------------------
def mono(frame, mode='mean'):
    mode = COLOR_CHANNELS.get(mode.lower(), mode.lower())      # filters.py:352
    if mode == 'mean':
        return np.mean(frame, axis=2).astype(frame.dtype)      # filters.py:366
    return frame[:, :, mode]                                   # filters.py:368


# --------------------------------------------------------------------------
# crop -- video/filters.py:139-248, video/analysis/regions.py:49-53
# --------------------------------------------------------------------------
def check_coordinate(value, max_value):
    """ filters.py:139-154 """
    if -1 < value < 1:
        value = int(value * max_value)
    if value < 0:
        value += max_value
    if not 0 <= value < max_value:
        raise IndexError('Coordinate %d is out of bounds [0, %d].' % (value, max_value))
    return value


def crop_rect(source_size, rect=None, region='', size_alignment=1):
    """ resolves (left, top, width, height) the way FilterCrop.__init__ does
    (filters.py:181-226) """
    source_width, source_height = source_size
    if rect is not None:
        left = check_coordinate(rect[0], source_width)
        top = check_coordinate(rect[1], source_height)
        width = check_coordinate(rect[2], source_width)
        height = check_coordinate(rect[3], source_height)
    else:
        region = region.lower()
        left, top = 0, 0
        width, height = source_width, source_height
        if 'left' in region:
            width //= 2
        elif 'right' in region:
            width //= 2
            left = source_width - width
        if 'upper' in region:
            height //= 2
        elif 'lower' in region:
            height //= 2
            top = source_height - height
    if size_alignment != 1:
        width = int(round(width / size_alignment) * size_alignment)
        height = int(round(height / size_alignment) * size_alignment)
    return (left, top, width, height)


def rect_to_slices(rect):
    """ regions.py:49-53 """
    slice_x = slice(rect[0], rect[2] + rect[0])
    slice_y = slice(rect[1], rect[3] + rect[1])
    return slice_y, slice_x


def crop(frame, rect, color_channel=None):
    """ filters.py:238-248 """
    slices = rect_to_slices(rect)
    if color_channel is None:
        return frame[slices]
    channel = COLOR_CHANNELS.get(color_channel, color_channel)
    return frame[slices[0], slices[1], channel]


# --------------------------------------------------------------------------
# Gaussian blur -- video/filters.py:378-392
# --------------------------------------------------------------------------
def blur(frame, sigma=3):
    return cv2.GaussianBlur(frame.astype(np.uint8), (0, 0), sigma)   # filters.py:392


def gauss_ksize(sigma):
    """ kernel size OpenCV picks for ksize=(0,0) on 8-bit images """
    return int(round(sigma * 6 + 1)) | 1


def gauss_kernel_u8(sigma):
    """ the 8-bit fixed-point kernel (sum == 256) OpenCV's uint8 Gaussian uses:
    float64 taps, 8 fractional bits, error diffused from the ends inwards
    (probed bit-exact against cv2 4.13.0) """
    ksize = gauss_ksize(sigma)
    k64 = cv2.getGaussianKernel(ksize, sigma, cv2.CV_64F)[:, 0]
    K = np.zeros(ksize, dtype=np.int64)
    err = 0.0
    for i in range(ksize // 2):
        adj = k64[i] * 256 + err
        K[i] = K[ksize - 1 - i] = int(np.rint(adj))
        err = adj - K[i]
    K[ksize // 2] = 256 - 2 * K[:ksize // 2].sum()
    return K


def blur_integer(frame, sigma=3):
    """ integer restatement of cv2.GaussianBlur on uint8:
    out = (sum_ky sum_kx K[ky] K[kx] src + 32768) >> 16, border REFLECT_101.
    Must equal `blur` bit for bit; the CUDA kernel implements this arithmetic. """
    K = gauss_kernel_u8(sigma)
    r = len(K) // 2
    src = np.asarray(frame, dtype=np.uint8)
    if src.ndim == 3:
        return np.stack([blur_integer(src[..., c], sigma) for c in range(src.shape[2])], axis=2)
    pad = np.pad(src.astype(np.int64), r, mode='reflect')
    H, W = src.shape
    row = np.zeros((H + 2 * r, W), dtype=np.int64)
    for k in range(len(K)):
        row += K[k] * pad[:, k:k + W]
    acc = np.zeros((H, W), dtype=np.int64)
    for k in range(len(K)):
        acc += K[k] * row[k:k + H, :]
    return ((acc + 32768) >> 16).astype(np.uint8)


# --------------------------------------------------------------------------
# resize -- video/filters.py:252-315
# --------------------------------------------------------------------------
_INTERPOLATIONS = {'nearest': cv2.INTER_NEAREST, 'linear': cv2.INTER_LINEAR,
                   'area': cv2.INTER_AREA, 'cubic': cv2.INTER_CUBIC,
                   'lanczos': cv2.INTER_LANCZOS4}


def resize_target(source_size, size, even_dimensions=False):
    """ filters.py:265-273 """
    if hasattr(size, '__iter__'):
        width, height = size
    else:
        width = int(source_size[0] * size)
        height = int(source_size[1] * size)
    if even_dimensions:
        width += (width % 2)
        height += (height % 2)
    return width, height


def resize(frame, size, interpolation='auto', even_dimensions=False):
    """ filters.py:275-311; `size` is (width, height) or a scalar factor """
    src_size = (frame.shape[1], frame.shape[0])
    width, height = resize_target(src_size, size, even_dimensions)
    if (width, height) == src_size:
        return frame                                            # filters.py:276-277
    if interpolation == 'auto':
        if width * height < src_size[0] * src_size[1]:
            interp = cv2.INTER_AREA                             # filters.py:279-281
        else:
            interp = cv2.INTER_CUBIC
    else:
        try:
            interp = _INTERPOLATIONS[interpolation]
        except KeyError:
            raise ValueError('Unknown interpolation method: %s' % interpolation)
    return cv2.resize(frame, (width, height), interpolation=interp)   # filters.py:311


# --------------------------------------------------------------------------
# frame difference -- video/filters.py:564-568
# --------------------------------------------------------------------------
def time_difference(this_frame, prev_frame, dtype=np.int16):
    if dtype is not None:
        this_frame = this_frame.astype(dtype)
    return this_frame - prev_frame


# --------------------------------------------------------------------------
# normalize / rotate -- video/filters.py:76-135, :319-344
# --------------------------------------------------------------------------
def normalize(frames, vmin=None, vmax=None, dtype=None):
    """ FilterNormalize over a sequence: bounds and dtype are fixed by the first frame when not
    given (filters.py:104-124); clip, scale to the dtype's colour range, cast (filters.py:126-132) """
    fmin, fmax, tmin, alpha = vmin, vmax, None, None
    out = []
    for frame in frames:
        frame = np.array(frame, copy=True)            # the reference clips its input in place
        if dtype is None:
            dtype = frame.dtype
        if fmin is None:
            fmin = frame.min()
        if fmax is None:
            fmax = frame.max()
        if tmin is None:
            if np.issubdtype(dtype, np.integer):
                tmin, tmax = np.iinfo(dtype).min, np.iinfo(dtype).max
            else:
                tmin, tmax = 0, 1
            alpha = (tmax - tmin) / (fmax - fmin)
        np.clip(frame, fmin, fmax, out=frame)
        out.append(((frame - fmin) * alpha + tmin).astype(dtype))
    return np.stack(out)


def rotate(frame, angle):
    """ filters.py:338-344 """
    return np.rot90(frame, (angle % 360) // 90)


# --------------------------------------------------------------------------
# temporal folds -- video/analysis/video.py:14-35
# --------------------------------------------------------------------------
def measure_mean_std(frames):
    """ analysis/video.py:39-55 (incremental mean / M2, float64) """
    mean = np.zeros(np.shape(frames[0]))
    M2 = np.zeros(np.shape(frames[0]))
    n = -1
    for n, frame in enumerate(frames):
        delta = frame - mean
        mean = mean + delta / (n + 1)
        M2 = M2 + delta * (frame - mean)
    if n < 2:
        return frames[-1], 0
    return mean, np.sqrt(M2 / n)



def measure_mean(frames):
    """ analysis/video.py:26-35 (cumulative mean, float64) """
    mean = np.zeros(np.shape(frames[0]))
    for n, frame in enumerate(frames):
        mean = mean * n / (n + 1) + frame / (n + 1)
    return mean


def background_ema(frames, alpha=0.05, thr=25, dtype=np.float32, bg0=None):
    """ ADOPTED (no EMA / threshold in the reference; fold shape and first-frame
    initialisation follow analysis/video.py:14-22, signed difference follows
    filters.py:564-568):

        bg_0 = float(blur_0), mask_0 = 0
        d_t = float(blur_t) - bg_{t-1};  mask_t = |d_t| > thr;  bg_t = bg_{t-1} + alpha * d_t

    evaluated in `dtype` with separately rounded multiply and add.  If `bg0` is
    given it is the state *before* frames[0] (continuation of a longer video).
    Returns (masks uint8 {0,255} [T,H,W], bg after the last frame). """
    alpha = dtype(alpha)
    thr = dtype(thr)
    masks = np.zeros((len(frames),) + np.shape(frames[0]), dtype=np.uint8)
    bg = None if bg0 is None else np.array(bg0, dtype=dtype)
    for t, frame in enumerate(frames):
        x = np.asarray(frame).astype(dtype)
        if bg is None:
            bg = x.copy()                     # first frame initialises the state
            continue
        d = x - bg
        masks[t] = np.where(np.abs(d) > thr, 255, 0)
        bg = bg + alpha * d
    return masks, bg


# --------------------------------------------------------------------------
# apply-mask -- ADOPTED (idioms: io/composer.py:154,186,208)
# --------------------------------------------------------------------------
def apply_mask(frame, mask):
    m = np.asarray(mask) != 0
    if frame.ndim == 3:
        m = m[:, :, None]
    return np.where(m, frame, 0).astype(frame.dtype)


# --------------------------------------------------------------------------
# binary morphology -- anchors analysis/image.py:248-256 (cv2.erode/dilate with
# 3x3 MORPH_CROSS), analysis/image.py:164-165 (MORPH_ELLIPSE)
# --------------------------------------------------------------------------
_SHAPES = {'rect': cv2.MORPH_RECT, 'cross': cv2.MORPH_CROSS, 'ellipse': cv2.MORPH_ELLIPSE}


def structuring_element(shape='rect', ksize=3):
    kx, ky = (ksize, ksize) if np.isscalar(ksize) else ksize
    return cv2.getStructuringElement(_SHAPES[shape], (int(kx), int(ky)))


def morph(mask, op, shape='rect', ksize=3):
    """ erode / dilate as called at image.py:250-251; open / close ADOPTED as
    cv2.morphologyEx (== dilate(erode) / erode(dilate) with OpenCV's default
    border, i.e. pixels outside the image never win the min / max) """
    se = structuring_element(shape, ksize)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    if op == 'erode':
        return cv2.erode(mask, se)
    if op == 'dilate':
        return cv2.dilate(mask, se)
    if op == 'open':
        return cv2.morphologyEx(mask, cv2.MORPH_OPEN, se)
    if op == 'close':
        return cv2.morphologyEx(mask, cv2.MORPH_CLOSE, se)
    raise ValueError('unknown morphological operation %r' % op)


# --------------------------------------------------------------------------
# labelling -- video/analysis/regions.py:159-174
# --------------------------------------------------------------------------
def label(mask, connectivity=4):
    """ regions.py:162: ndimage.measurements.label(mask) -> (int32 labels, n);
    default structure is the 4-connected cross; connectivity=8 uses ones(3,3) """
    structure = None if connectivity == 4 else np.ones((3, 3), dtype=int)
    labels, n = ndimage.label(mask, structure=structure)
    return labels.astype(np.int32, copy=False), int(n)


def region_areas(labels, n):
    """ regions.py:165-166 (same numbers, without the O(n*N) loop) """
    return np.bincount(labels.ravel(), minlength=n + 1)[1:]


def get_largest_region(mask, ret_area=False):
    """ regions.py:159-174 """
    labels, num_features = label(mask)
    areas = region_areas(labels, num_features)
    label_max = np.argmax(areas) + 1
    if ret_area:
        return labels == label_max, areas[label_max - 1]
    return labels == label_max


def find_bounding_box(mask):
    """ regions.py:113-149: (left, top, width, height) of the first block of non-empty rows and the
    first block of non-empty columns (the bounding box when the mask holds one connected region) """
    rows = np.any(mask, axis=1)
    cols = np.any(mask, axis=0)
    top = 0
    while not rows[top]:
        top += 1
    bottom = top + 1
    while bottom < mask.shape[0] and rows[bottom]:
        bottom += 1
    left = 0
    while not cols[left]:
        left += 1
    right = left + 1
    while right < mask.shape[1] and cols[right]:
        right += 1
    return (left, top, right - left, bottom - top)


def region_moments(labels, n):
    """ image.py:350: cv2.moments(mask.astype(np.uint8)) for every region of a label image """
    return [cv2.moments((labels == l).astype(np.uint8)) for l in range(1, n + 1)]


class regionprops(object):
    """ image.py:310-405, the properties derived from the moments (same formulae, same order) """

    def __init__(self, mask=None, moments=None):
        self.moments = moments if moments is not None else cv2.moments(mask.astype(np.uint8))

    @property
    def area(self):
        return self.moments['m00']

    @property
    def centroid(self):
        m = self.moments
        return (m['m10'] / m['m00'], m['m01'] / m['m00'])

    @property
    def orientation(self):
        m = self.moments
        a, b, c = m['mu20'], m['mu11'], m['mu02']
        if a - c == 0:
            return -np.pi / 4 if b > 0 else np.pi / 4
        return -np.arctan2(2 * b, (a - c)) / 2

    @property
    def inertia_tensor_eigvals(self):
        m = self.moments
        a, b, c = m['mu20'] / m['m00'], -m['mu11'] / m['m00'], m['mu02'] / m['m00']
        e1 = (a + c) + np.sqrt(4 * b ** 2 + (a - c) ** 2)
        e2 = (a + c) - np.sqrt(4 * b ** 2 + (a - c) ** 2)
        return e1, e2

    @property
    def eccentricity(self):
        e1, e2 = self.inertia_tensor_eigvals
        return 0 if e1 == 0 else np.sqrt(1 - e2 / e1)


# --------------------------------------------------------------------------
# packed bit masks (device layout): 32-bit words, LSB = lowest x
# --------------------------------------------------------------------------
def pack_bits(mask):
    """ (..., H, W) nonzero -> (..., H, ceil(W/32)) uint32 """
    m = (np.asarray(mask) != 0)
    W = m.shape[-1]
    Wp = (W + 31) // 32
    pad = Wp * 32 - W
    if pad:
        m = np.concatenate([m, np.zeros(m.shape[:-1] + (pad,), dtype=bool)], axis=-1)
    by = np.packbits(m, axis=-1, bitorder='little')
    return np.ascontiguousarray(by).view(np.uint32)


def unpack_bits(words, width):
    """ inverse of pack_bits -> uint8 {0,255} """
    by = np.ascontiguousarray(words, dtype=np.uint32).view(np.uint8)
    bits = np.unpackbits(by, axis=-1, bitorder='little')[..., :width]
    return (bits * np.uint8(255)).astype(np.uint8)


# --------------------------------------------------------------------------
# the whole chain of BASELINE.json configs 1-3
# --------------------------------------------------------------------------
def chain(frames, sigma=2, alpha=0.05, thr=25, morph_op='open', morph_shape='rect',
          morph_ksize=3, connectivity=4, bg0=None, keep=('mono', 'blur', 'mask', 'morph', 'labels')):
    """ mono -> blur -> EMA background / |diff| > thr -> open -> label, one frame
    per iteration like the reference's lazy chain (io/base.py:377-380).
    Returns a dict of per-stage stacks plus 'counts' and the final 'bg'. """
    out = {k: [] for k in keep}
    counts = []
    alpha32, thr32 = np.float32(alpha), np.float32(thr)
    bg = None if bg0 is None else np.array(bg0, dtype=np.float32)
    for frame in frames:
        m = mono(frame) if frame.ndim == 3 else frame
        b = blur(m, sigma)
        x = b.astype(np.float32)
        if bg is None:
            bg = x.copy()
            mask = np.zeros(b.shape, dtype=np.uint8)
        else:
            d = x - bg
            mask = np.where(np.abs(d) > thr32, 255, 0).astype(np.uint8)
            bg = bg + alpha32 * d
        mo = morph(mask, morph_op, morph_shape, morph_ksize) if morph_op else mask
        lab, n = label(mo, connectivity)
        counts.append(n)
        for k, v in (('mono', m), ('blur', b), ('mask', mask), ('morph', mo), ('labels', lab)):
            if k in out:
                out[k].append(v)
    res = {k: np.stack(v) for k, v in out.items()}
    res['counts'] = np.array(counts, dtype=np.int32)
    res['bg'] = bg
    return res


# --------------------------------------------------------------------------
# annotated output -- video/io/composer.py:131-154 (VideoComposer.highlight_mask, zoom factor 1)
# --------------------------------------------------------------------------
_CHANNEL_NAMES = {0: 0, 'r': 0, 'red': 0, 1: 1, 'g': 1, 'green': 1, 2: 2, 'b': 2, 'blue': 2}   # composer.py:43-45


def highlight_mask(frame, mask, channel='all', strength=128):
    """ returns a copy of `frame` with the non-zero entries of `mask` highlighted """
    frame = frame.copy()
    is_color = frame.ndim == 3
    if channel is None or channel == 'all':
        channel = slice(0, 3) if is_color else 0
    elif is_color:
        try:
            channel = _CHANNEL_NAMES[channel]
        except KeyError:
            raise ValueError('Unknown value `%s` for channel.' % channel)
    else:
        raise ValueError('Highlighting a specific channel is only supported for color videos.')
    mask = np.asarray(mask).astype(bool)
    factor = (255 - strength) / 255                                   # composer.py:153 (true division)
    if is_color:
        frame[mask, channel] = strength + factor * frame[mask, channel]
    else:
        frame[mask] = strength + factor * frame[mask]                 # a monochrome _frame has no channel axis
    return frame
