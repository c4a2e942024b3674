"""
Region helpers on the segment step (reference: video/analysis/regions.py).

    rect_to_slices        regions.py:49-53
    label                 regions.py:162    ndimage.measurements.label(mask)
    get_largest_region    regions.py:159-174
    find_bounding_box     regions.py:113-149
    region_stats          per-region moments / bounding boxes reduced on the device

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


# ---------------------------------------------------------------------------------------------
# per-region statistics (no label image leaves the device)
# ---------------------------------------------------------------------------------------------
RAW_KEYS = ('m00', 'm10', 'm01', 'm20', 'm11', 'm02')


def moments_from_raw(raw):
    """ the second-order part of a cv2.moments dict from exact integer raw moments: central moments
    as OpenCV derives them (cx = m10 / m00 via the reciprocal, mu20 = m20 - m10 cx, ...) and the
    normalised ones (nu = mu / m00^2) """
    m = {k: float(v) for k, v in zip(RAW_KEYS, raw)}
    inv = 1.0 / m['m00'] if abs(m['m00']) > np.finfo(np.float64).eps else 0.0
    cx, cy = m['m10'] * inv, m['m01'] * inv
    m['mu20'] = m['m20'] - m['m10'] * cx
    m['mu11'] = m['m11'] - m['m10'] * cy
    m['mu02'] = m['m02'] - m['m01'] * cy
    s2 = inv * inv
    m['nu20'], m['nu11'], m['nu02'] = m['mu20'] * s2, m['mu11'] * s2, m['mu02'] * s2
    return m


def stats_to_regions(stats, n):
    """ rows of va_region_stats -> list of dicts {label, area, bbox=(left, top, width, height), moments} """
    regions = []
    for l in range(n):
        row = [int(v) for v in stats[l]]
        regions.append({'label': l + 1, 'area': row[0],
                        'bbox': (row[6], row[7], row[8] - row[6] + 1, row[9] - row[7] + 1),
                        'moments': moments_from_raw(row[:6])})
    return regions


def region_stats(mask, connectivity=4, max_regions=4096, device=None):
    """ regions of a 2-D mask, numbered like `label` numbers them, with area, bounding box and the
    moments that `analysis.image.regionprops` consumes; only a few bytes per region are copied back """
    rt = get_runtime(device)
    t = torch()
    with t.cuda.device(rt.device):
        dev = _mask_to_device(rt, mask)
        stats, counts, _ = rt.region_stats(dev, connectivity, max_regions)
        n = int(counts.cpu()[0])
        if n > max_regions:
            # a noisy mask can hold more regions than the table has rows (the count is exact either way):
            # run once more with a table of the right size, as the reference's per-label loop has no limit
            stats, counts, _ = rt.region_stats(dev, connectivity, n)
        host = stats[0, :n].cpu().numpy()
    return stats_to_regions(host, n)


def find_bounding_box(mask, device=None):
    """ finds the rectangle, which bounds a white region in a mask, as [left, top, width, height]
    (regions.py:113-149).  Like the reference it reports the first block of non-empty rows and the
    first block of non-empty columns; every connected region covers a gap-free range of rows and of
    columns, so both blocks follow from the per-region boxes. """
    regions = region_stats(mask, connectivity=8, device=device)
    if not regions:
        raise IndexError('index out of bounds: the mask is empty')      # the reference runs off the array

    def first_block(intervals):
        intervals = sorted(intervals)
        lo, hi = intervals[0]
        for a, b in intervals[1:]:
            if a > hi + 1:
                break
            hi = max(hi, b)
        return lo, hi - lo + 1

    left, width = first_block([(r['bbox'][0], r['bbox'][0] + r['bbox'][2] - 1) for r in regions])
    top, height = first_block([(r['bbox'][1], r['bbox'][1] + r['bbox'][3] - 1) for r in regions])
    return (left, top, width, height)
