import time, sys, numpy as np, torch, cProfile, pstats
sys.path.insert(0, '/root/repo')
from video_analysis_b200 import filters
from video_analysis_b200.io.memory import VideoMemory
from oracle import synth
W, H, n = 1920, 1080, 256
fr = synth.make_frames(0, 0, 8, W, H, 6)
frames = np.concatenate([fr] * (n // 8))
v = VideoMemory(frames); v.pin()
def chain(v, batch):
    c = filters.FilterMonochrome(v, batch=batch)
    c = filters.FilterBlur(c, 2)
    c = filters.FilterBackgroundMask(c, alpha=0.05, threshold=25)
    c = filters.FilterMorphology(c, 'open', 'rect', 3)
    return filters.FilterLabel(c)
for batch in (32, 64):
    for rep in range(3):
        c = chain(v, batch)
        t0 = time.perf_counter(); k = 0
        for f in c: k += 1
        dt = time.perf_counter() - t0
        print('batch %d rep %d: %.0f fps' % (batch, rep, k / dt))
c = chain(v, 32)
pr = cProfile.Profile(); pr.enable()
for f in c: pass
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
