"""
Temporal statistics of a video (reference: video/analysis/video.py:14-55), with the per-pixel
float64 recurrences evaluated on the GPU in the reference's operation order (bit-identical):

    measure_mean       mean = mean * n / (n + 1) + frame / (n + 1)                    (:26-35)
    measure_mean_std   delta = frame - mean; mean += delta / (n + 1); M2 += delta * (frame - mean)   (:39-55)

Frames are pulled in blocks (no per-frame kernel launches); uint8 videos only.
"""

import numpy as np

from ..device import get_runtime, torch


def reduce_video(video, function, initial_value=None):
    """ applies function to consecutive frames (video.py:14-22); host callback, as in the reference """
    result = initial_value
    for frame in video:
        result = frame if result is None else function(frame, result)
    return result


def _blocks(video, batch):
    video.rewind()
    while True:
        frames = []
        try:
            for _ in range(batch):
                frames.append(np.asarray(video.get_next_frame()))
        except StopIteration:
            pass
        if frames:
            yield np.stack(frames)
        if len(frames) < batch:
            return


def _fold(video, with_m2, batch, device):
    rt = get_runtime(device)
    t = torch()
    shape = tuple(video.shape[1:])
    rowe = int(np.prod(shape[1:]))
    n = 0
    last = None
    with t.cuda.device(rt.device):
        mean = t.zeros((shape[0], rowe), dtype=t.float64, device=rt.device)
        m2 = t.zeros((shape[0], rowe), dtype=t.float64, device=rt.device) if with_m2 else None
        for block in _blocks(video, batch):
            if block.dtype != np.uint8:
                raise NotImplementedError('temporal statistics on the device handle uint8 videos')
            rt.mean_update(rt.upload(block), mean, m2, n)
            n += len(block)
            last = block[-1]
        t.cuda.synchronize(rt.device)
        mean_h = mean.cpu().numpy().reshape(shape)
        m2_h = m2.cpu().numpy().reshape(shape) if with_m2 else None
    return mean_h, m2_h, n, last


def measure_mean(video, batch=32, device=None):
    """ measures the mean of each movie pixel over time """
    return _fold(video, False, batch, device)[0]


def measure_mean_std(video, batch=32, device=None):
    """ mean and standard deviation of each pixel over time; like the reference, the variance is
    divided by the index of the last frame and videos shorter than three frames return
    (last frame, 0) """
    mean, m2, count, last = _fold(video, True, batch, device)
    n = count - 1
    if n < 2:
        return last, 0
    return mean, np.sqrt(m2 / n)
