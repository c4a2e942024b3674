"""Times the label kernels on masks of different density (ncu --metrics gpu__time_duration.sum)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analysis_b200 import synth
from video_analysis_b200.device import get_runtime
W, H, B = 1920, 1080, 64
rt = get_runtime(0); rt.ensure(W, H, B)
rgb = synth.generate(rt, 0, 0, B, W, H)
mono = rt.luma(rgb)
blur = rt.gauss(mono, 2.0)
for thr in (255, 150, 118):
    mask = rt.threshold(blur, thr)
    frac = float((rt.unpack_bits(mask).t[:, :, :W] > 0).float().mean())
    for _ in range(2):
        lab, cnt = rt.label(mask, 4)
    torch.cuda.synchronize()
    print('thr', thr, 'foreground fraction %.4f' % frac, 'regions/frame', cnt[:4].tolist(), flush=True)
