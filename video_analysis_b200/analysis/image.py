"""
Binary morphology helpers (reference call sites: video/analysis/image.py:248-256 use
cv2.erode / cv2.dilate with a 3x3 MORPH_CROSS element; image.py:164-165 builds MORPH_ELLIPSE).
Results are bit-identical to cv2.erode / cv2.dilate / cv2.morphologyEx with
cv2.getStructuringElement(shape, ksize) and OpenCV's default border.
"""

import numpy as np

from ..device import get_runtime, torch
from .regions import _mask_to_device


def morphology(mask, operation, shape='rect', ksize=3, device=None):
    """ mask (H, W) nonzero = foreground -> uint8 {0, 255} """
    rt = get_runtime(device)
    t = torch()
    with t.cuda.device(rt.device):
        out = rt.unpack_bits(rt.morph(_mask_to_device(rt, mask), operation, shape, ksize))
        host = rt.download(out)
        t.cuda.current_stream(rt.device).synchronize()
    return np.array(rt.host_view(out, host)[0])


def erode(mask, shape='cross', ksize=3, device=None):
    return morphology(mask, 'erode', shape, ksize, device)


def dilate(mask, shape='cross', ksize=3, device=None):
    return morphology(mask, 'dilate', shape, ksize, device)


def opening(mask, shape='rect', ksize=3, device=None):
    return morphology(mask, 'open', shape, ksize, device)


def closing(mask, shape='rect', ksize=3, device=None):
    return morphology(mask, 'close', shape, ksize, device)


class regionprops(object):
    """ properties of a region in a binary image from its moments (reference: image.py:310-405,
    which credits scikit-image for the formulae).  `mask` is reduced to moments on the device
    (`regions.region_stats`: all foreground pixels of the mask count as one region, as in
    cv2.moments(mask)); `moments` takes a cv2.moments-style dict, e.g. one entry of
    `regions.region_stats(...)`. """

    def __init__(self, mask=None, contour=None, moments=None, device=None):
        if moments is not None:
            self.moments = moments
        elif mask is not None:
            from .regions import RAW_KEYS, moments_from_raw, region_stats
            parts = region_stats(mask, connectivity=8, device=device)
            # cv2.moments(mask.astype(np.uint8)) (image.py:350) weights every pixel by its value: a mask of
            # 0 / 255 has 255 times the moments of the same mask of 0 / 1
            values = np.unique(np.asarray(mask).astype(np.uint8))
            values = values[values != 0]
            if len(values) > 1:
                raise ValueError('regionprops(mask=...) on the device takes a two-valued mask; '
                                 'grey-value weighted moments are outside the filter -> segment path')
            weight = int(values[0]) if len(values) else 1
            raw = [weight * sum(int(p['moments'][k]) for p in parts) for k in RAW_KEYS]
            self.moments = moments_from_raw(raw)
        elif contour is not None:
            raise NotImplementedError('contour moments are outside the filter -> segment path')
        else:
            raise ValueError('Either the mask or the moments must be given')

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
        root = np.sqrt(4 * b ** 2 + (a - c) ** 2)
        return (a + c) + root, (a + c) - root

    @property
    def eccentricity(self):
        e1, e2 = self.inertia_tensor_eigvals
        return 0 if e1 == 0 else np.sqrt(1 - e2 / e1)

    @property
    def major_axis_length(self):
        return 4 * np.sqrt(self.inertia_tensor_eigvals[0])

    @property
    def minor_axis_length(self):
        return 4 * np.sqrt(self.inertia_tensor_eigvals[1])
