import time, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from video_analysis_b200 import filters
from video_analysis_b200.io.memory import VideoMemory
from video_analysis_b200.chain import SegmentChain
from oracle import synth
W, H, n = 1920, 1080, 256
fr = synth.make_frames(0, 0, 8, W, H, 6)
frames = np.concatenate([fr] * (n // 8))
v = VideoMemory(frames)
v.pin() if hasattr(v, 'pin') else None
def chain(v):
    c = filters.FilterMonochrome(v)
    c = filters.FilterBlur(c, 2)
    c = filters.FilterBackgroundMask(c, alpha=0.05, threshold=25)
    c = filters.FilterMorphology(c, 'open', 'rect', 3)
    return filters.FilterLabel(c)
for rep in range(2):
    c = chain(v)
    t0 = time.perf_counter(); k = 0
    for f in c:
        k += 1
    dt = time.perf_counter() - t0
    print('filter classes, per-frame iteration: %d frames %.3f s = %.0f fps' % (k, dt, k / dt))
    ch = SegmentChain((W, H), batch=64)
    t0 = time.perf_counter()
    lab, cnt = ch.process(v)
    dt = time.perf_counter() - t0
    print('SegmentChain.process: %.0f fps' % (n / dt))
    t0 = time.perf_counter()
    regs = ch.process_regions(v)
    dt = time.perf_counter() - t0
    print('SegmentChain.process_regions: %.0f fps' % (n / dt))
