"""
Launches every kernel of the chain exactly once (after a warm-up of 19 launches) so that
`ncu -s 19 -c 13` captures one profile per kernel:
  luma, gauss, luma_gauss, ema_diff_thresh, morph, label x5 (init merge flatten scan write),
  region statistics x3 (init stats largest; their forest kernels come first and are skipped by -k)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analysis_b200 import synth  # noqa: E402
from video_analysis_b200.chain import SegmentChain  # noqa: E402
from video_analysis_b200.device import get_runtime  # noqa: E402

W, H, B = 1920, 1080, int(os.environ.get('PROF_BATCH', '32'))
rt = get_runtime(0)
rt.ensure(W, H, B)
rgbs = [synth.generate(rt, 0, i * B, B, W, H) for i in range(2)]          # 2 launches
for fuse in (True, False):                                                  # 8 + 9 launches
    ch = SegmentChain((W, H), batch=B, fuse=fuse)
    ch.run_device(rgbs[0])
torch.cuda.synchronize()
# ---- profiled region: 10 launches
mono = rt.luma(rgbs[1])
blur = rt.gauss(mono, 2.0)
blur2 = rt.luma_gauss(rgbs[1], 2.0)
bg = rt.empty_f32(H, W)
bg.copy_(ch._bg)
mask = rt.ema_diff_thresh(blur, bg, 0.05, 25.0, False)
mo = rt.morph(mask, 'open', 'rect', 3)
lab, cnt = rt.label(mo, 4)
stats, cnt2, big = rt.region_stats(mo, 4, 256)
# round 2: the tensor-core blur at a large radius (sigma 15: seven 16-row groups, 64-column strips) and the chunk export
if os.environ.get('PROF_EXTRA', '1') != '0':
    n15 = min(B, 32)
    from video_analysis_b200.device import DeviceBatch
    m15 = DeviceBatch('u8', mono.t[:n15], n15, H, W, 1)
    blur15 = rt.gauss(m15, 15.0)
    import torch as _t
    ids = _t.empty((B, 30 * H), dtype=_t.int32, pin_memory=True)
    data = _t.empty((B, 30 * H, 64), dtype=_t.int32, pin_memory=True)
    nn = _t.empty((B,), dtype=_t.int32, pin_memory=True)
    runs = _t.empty((B, 30 * H, 4), dtype=_t.int32, pin_memory=True)
    nr = _t.empty((B,), dtype=_t.int32, pin_memory=True)
    # PROF_EXPORT_RAW=1: every chunk travels raw (260 bytes), as before the run chunks
    raw = os.environ.get('PROF_EXPORT_RAW', '0') == '1'
    rt._check(rt.lib.va_label_export_chunks(rt._h, rt.stream, *mo.img(), *lab.img(), W, H, B, ids.data_ptr(), data.data_ptr(),
                                            nn.data_ptr(), None, None if raw else runs.data_ptr(), None if raw else nr.data_ptr(),
                                            30 * H))
torch.cuda.synchronize()
print('ok', cnt[:4].tolist(), rt.launches)
