"""
Region helpers on the segment step (reference: video/analysis/regions.py).

    rect_to_slices        regions.py:49-53
    label                 regions.py:162    ndimage.measurements.label(mask)
    get_largest_region    regions.py:159-174

Single-frame convenience wrappers: the mask goes to the GPU, is packed, labelled by the
union-find kernels (va_label_bits) and the result comes back.  For throughput use the
batched `FilterLabel` / `SegmentChain` instead.
"""

import numpy as np

from ..device import get_runtime, torch


def rect_to_slices(rect):
    """ creates slices for an array from a rectangle (left, top, width, height) """
    slice_x = slice(rect[0], rect[2] + rect[0])
    slice_y = slice(rect[1], rect[3] + rect[1])
    return slice_y, slice_x


def _mask_to_device(rt, mask):
    mask = np.asarray(mask)
    if mask.ndim != 2:
        raise ValueError('mask must be two-dimensional')
    if mask.dtype != np.uint8:
        mask = (mask != 0).astype(np.uint8)
    return rt.pack_bits(rt.upload(np.ascontiguousarray(mask)[None]))


def label(mask, connectivity=4, device=None):
    """ -> (labels int32 (H, W), num_features), scipy.ndimage.label semantics """
    rt = get_runtime(device)
    t = torch()
    with t.cuda.device(rt.device):
        labels, counts = rt.label(_mask_to_device(rt, mask), connectivity)
        host = rt.download(labels)
        n = int(counts.cpu()[0])
        t.cuda.current_stream(rt.device).synchronize()
    return np.array(rt.host_view(labels, host)[0]), n


def region_areas(mask, connectivity=4, device=None):
    """ -> (labels, num_features, areas[num_features]) with the areas reduced on the device """
    rt = get_runtime(device)
    t = torch()
    with t.cuda.device(rt.device):
        labels, counts = rt.label(_mask_to_device(rt, mask), connectivity)
        n = int(counts.cpu()[0])
        areas, largest = rt.region_areas(labels, max(n, 1))
        host = rt.download(labels)
        areas_h = areas.cpu().numpy()[0, :n]
        t.cuda.current_stream(rt.device).synchronize()
    return np.array(rt.host_view(labels, host)[0]), n, areas_h


def get_largest_region(mask, ret_area=False, device=None):
    """ returns a mask only containing the largest region (regions.py:159-174) """
    labels, n, areas = region_areas(mask, device=device)
    if n == 0:
        raise ValueError('attempt to get argmax of an empty sequence')   # what np.argmax([]) raises
    label_max = int(np.argmax(areas)) + 1
    if ret_area:
        return labels == label_max, int(areas[label_max - 1])
    return labels == label_max
