"""Two launches of va_gauss_u8 at sigma 15 (and 5) under the current environment (for `ncu -k regex:gauss_mma`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analysis_b200 import synth  # noqa: E402
from video_analysis_b200.device import get_runtime  # noqa: E402

W, H, B = 1920, 1080, int(os.environ.get('PROF_BATCH', '32'))
rt = get_runtime(0)
rt.ensure(W, H, B)
rgb = synth.generate(rt, 0, 0, B, W, H)
mono = rt.luma(rgb)
for s in (15.0, 5.0, 15.0, 5.0):
    a = rt.gauss(mono, s)
torch.cuda.synchronize()
print('ok', int(a.t.sum()))
