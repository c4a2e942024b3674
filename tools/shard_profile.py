"""Times the phases of ShardedSegmentChain.run_device_range on one GPU (world size 1)."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analysis_b200 import synth
from video_analysis_b200.chain import SegmentChain
from video_analysis_b200.device import get_runtime
from video_analysis_b200 import parallel

os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29533')
dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', 0))
rt = get_runtime(0)
W, H, B, K = 1920, 1080, 64, 20
batches = [synth.generate(rt, 0, i * B, B, W, H) for i in range(K)]
ch = SegmentChain((W, H), batch=B)
sh = parallel.ShardedSegmentChain(ch)
labels = [rt.empty_i32(B, H, W) for _ in range(2)]
counts = torch.empty((B,), dtype=torch.int32, device=rt.device)

def t(fn, name):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print('%-28s %.2f ms' % (name, (time.perf_counter() - t0) * 1e3), flush=True); return r

for rep in range(2):
    t(lambda: sh.run_device_range(batches, labels, counts), 'run_device_range (%d steps)' % K)
m = sh.tail_batches(K)
blurs = t(lambda: [ch.blur_device(b, sh._blurs[i]) for i, b in enumerate(batches[K - m:])], 'pass1 blur (%d tail batches)' % m)
t(lambda: sh._partial_state(blurs, m == K), 'partial state')
S = sh._partial_state(blurs, m == K)
t(lambda: parallel.exchange_carry(S, K * B, ch.alpha, lambda c, s, sc: rt.ema_fold(c, s, sc, W, H)), 'exchange')
ch.reset()
t(lambda: [ch.run_device(b, labels[i & 1], counts) for i, b in enumerate(batches)], 'fused chain (reference)')
dist.destroy_process_group()
