"""Two launches of va_luma_gauss_u8 and va_gauss_u8 under the current environment (for `ncu -k regex:gauss`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analysis_b200 import synth  # noqa: E402
from video_analysis_b200.device import get_runtime  # noqa: E402

W, H, B = 1920, 1080, int(os.environ.get('PROF_BATCH', '64'))
rt = get_runtime(0)
rt.ensure(W, H, B)
rgbs = [synth.generate(rt, 0, i * B, B, W, H) for i in range(2)]
monos = [rt.luma(r) for r in rgbs]
for i in range(2):
    a = rt.luma_gauss(rgbs[i], 2.0)
    b = rt.gauss(monos[i], 2.0)
torch.cuda.synchronize()
print('ok', int(a.t.sum()), int(b.t.sum()))
