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
