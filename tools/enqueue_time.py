import sys, time, torch
sys.path.insert(0, '/root/repo')
from video_analysis_b200 import synth
from video_analysis_b200.chain import SegmentChain
from video_analysis_b200.device import get_runtime
W, H, B = 1920, 1080, 64
rt = get_runtime(0); rt.ensure(W, H, B)
rgbs = [synth.generate(rt, 0, i * B, B, W, H) for i in range(4)]
ch = SegmentChain((W, H), batch=B)
labels = [rt.empty_i32(B, H, W) for _ in range(2)]
counts = torch.empty((B,), dtype=torch.int32, device=rt.device)
for i in range(10): ch.run_device_pipelined(rgbs[i % 4], labels[i & 1], counts)
ch.pipeline_sync(); torch.cuda.synchronize()
for K in (50, 200):
    t0 = time.perf_counter()
    for i in range(K): ch.run_device_pipelined(rgbs[i % 4], labels[i & 1], counts)
    t1 = time.perf_counter()
    ch.pipeline_sync(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    print('K=%d enqueue %.1f ms (%.1f us/step), total %.1f ms (%.1f us/step)' % (K, (t1 - t0) * 1e3, (t1 - t0) / K * 1e6, (t2 - t0) * 1e3, (t2 - t0) / K * 1e6))
t0 = time.perf_counter()
for i in range(200): ch.run_device(rgbs[i % 4], labels[i & 1], counts)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('run_device: enqueue %.1f us/step total %.1f us/step' % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
