"""
Phase timings of the frame-sharded chain (ShardedSegmentChain.run_device_range), CUDA events, max over ranks.
Run under torchrun on N GPUs (or plainly on one):

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/shard_profile.py [--steps 20] [--batch 128]

Prints, per range of `steps` batches: the whole range, its phases enqueued one by one with a device
synchronisation in between (pass 1 = tail blurs + partial state, exchange = all-gather, pass 2 = carry fold +
pipelined chain), and the unsharded pipelined chain over the same batches on the same GPU for comparison.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analysis_b200 import synth  # noqa: E402
from video_analysis_b200.chain import SegmentChain  # noqa: E402
from video_analysis_b200.device import get_runtime  # noqa: E402
from video_analysis_b200 import parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--width', type=int, default=1920)
    ap.add_argument('--height', type=int, default=1080)
    ap.add_argument('--reps', type=int, default=5)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29533')
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local))
    rt = get_runtime(local)
    W, H, B, K = args.width, args.height, args.batch, args.steps
    batches = [synth.generate(rt, 0, (rank * K + i) * B, B, W, H) for i in range(K)]
    ch = SegmentChain((W, H), batch=B)
    sh = parallel.ShardedSegmentChain(ch)
    sh.reserve(K)
    labels = [rt.empty_i32(B, H, W) for _ in range(2)]
    counts = torch.empty((B,), dtype=torch.int32, device=rt.device)

    def timed(fn):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        ch.pipeline_sync()                     # the phases run on the chain's own streams
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=rt.device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    def report(name, ms):
        if rank == 0:
            print('%-44s %8.3f ms  (%.4f ms / step)' % (name, ms, ms / K), flush=True)

    shard_counts = [K * B] * world
    for rep in range(args.reps):
        ms, _ = timed(lambda: sh.run_device_range(batches, labels, counts))
        report('sharded range, %d GPUs, rep %d' % (world, rep), ms)
    ms1, S = timed(lambda: sh.pass1(batches))
    report('  pass 1 (%d tail blurs + partial state)' % sh.tail_batches(K), ms1)
    def exch():
        G, work = sh.exchange(S)
        work.wait()                            # stream-side: the current stream continues after the gather
        return G
    msx, G = timed(exch)
    report('  exchange (all-gather of one frame per rank)', msx)
    msh, head = timed(lambda: sh.preblur(batches))
    report('  pre-blur of %d head batches' % len(head), msh)
    ms2, _ = timed(lambda: sh.pass2(batches, labels, counts, G, shard_counts, None, head))
    report('  pass 2 (carry fold + pipelined chain)', ms2)
    blur0 = ch.blur_device(batches[0])
    S2 = rt.empty_f32(H, W)
    for rep in range(3):
        msf, _ = timed(lambda: rt.ema_partial(blur0, S2, ch.alpha, True))
    report('  (one va_ema_partial over %d frames, alone)' % B, msf)
    single = SegmentChain((W, H), batch=B)

    def plain():
        for i, b in enumerate(batches):
            single.run_device_pipelined(b, labels[i & 1], counts)
        single.pipeline_sync()
    for rep in range(3):
        msp, _ = timed(plain)
    report('unsharded pipelined chain, same batches', msp)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
