"""
Multi-GPU correctness check (run under torchrun on N GPUs): a synthetic video is sharded by
contiguous frame ranges; every rank runs ShardedSegmentChain on its range and the labels /
counts are compared with a sequential single-GPU run of the whole video on rank 0's device
data (each rank recomputes the sequential result for its own range, the generator is seeded).
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/mgpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_analysis_b200 import synth  # noqa: E402
from video_analysis_b200.chain import SegmentChain  # noqa: E402
from video_analysis_b200.device import DeviceBatch, get_runtime  # noqa: E402
from video_analysis_b200.parallel import ShardedSegmentChain, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rt = get_runtime(local)
    W, H, B, T = 640, 480, 16, 16 * 4 * world
    a, b = shard_range(T, rank, world)

    # sequential reference on this GPU: whole video up to the end of my range
    seq = SegmentChain((W, H), batch=B)
    ref_labels, ref_counts, ref_masks = [], [], []
    for t0 in range(0, b, B):
        rgb = synth.generate(rt, 0, t0, B, W, H)
        mask = rt.empty_bits(B, H, W)
        lab, cnt = seq.run_device(rgb, mask=mask)
        if t0 >= a:
            ref_labels.append(lab.t.clone()); ref_counts.append(cnt.clone()); ref_masks.append(mask.t.clone())
    ref_bg = seq._bg.clone()

    # sharded run: only my frames
    ch = SegmentChain((W, H), batch=B)
    sh = ShardedSegmentChain(ch)
    batches = [synth.generate(rt, 0, t0, B, W, H) for t0 in range(a, b, B)]
    outs = [rt.empty_i32(B, H, W) for _ in batches]
    counts = torch.empty((len(batches), B), dtype=torch.int32, device=rt.device)

    sh.run_device_range(batches, outs, counts)
    torch.cuda.synchronize()

    got = torch.stack([o.t for o in outs])
    exp = torch.stack(ref_labels)
    same_labels = float((got == exp).float().mean())
    same_counts = float((counts == torch.stack(ref_counts)).float().mean())
    bg_err = float(((ch._bg - ref_bg).abs() / ref_bg.abs().clamp(min=1)).max())
    exact = torch.equal(got, exp)
    print('rank %d frames [%d,%d): labels identical=%s (pixel agreement %.6f), counts agreement %.4f, bg max rel err %.2e'
          % (rank, a, b, exact, same_labels, same_counts, bg_err), flush=True)
    ok = bg_err < 1e-5 and same_labels > 0.9999 and same_counts > 0.99 and (rank > 0 or (exact and same_counts == 1.0))
    # temporal statistics of a frame-sharded video (merge of per-rank mean / M2) against the sequential recurrence
    import numpy as np
    from video_analysis_b200.analysis.video import measure_mean_std
    from video_analysis_b200.io.memory import VideoMemory
    from video_analysis_b200.parallel import measure_mean_std_sharded
    vid = VideoMemory(np.random.default_rng(5).integers(0, 256, (37, 48, 64, 3), dtype=np.uint8), copy_data=False)
    m_par, s_par = measure_mean_std_sharded(vid, batch=8, device=local)
    m_seq, s_seq = measure_mean_std(vid, batch=8, device=local)
    stats_err = max(float(np.abs(m_par - m_seq).max()), float(np.abs(s_par - s_seq).max()))
    print('rank %d temporal mean / std of a sharded video: max abs deviation from the sequential recurrence %.2e' % (rank, stats_err),
          flush=True)
    ok = ok and stats_err < 1e-9
    flag = torch.tensor([1 if ok else 0], device=rt.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print('MGPU_CHECK', 'PASS' if int(flag.item()) else 'FAIL')
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == '__main__':
    main()
