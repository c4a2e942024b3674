"""
Frame-range sharding over the GPUs of one box.

Frames are independent except for the running-average background, an affine recurrence

    bg_t = a * bg_{t-1} + alpha * x_t,      a = 1 - alpha.

Rank r owns the contiguous frames [r*T/R, (r+1)*T/R) (the `VideoSlice` partition of the
reference, video/io/base.py:392-474).  Pass 1: every rank blurs its frames and folds them
from a ZERO state into S_r (rank 0 from its first frame, as the sequential model does).
One `all_gather` of S_r (one float32 frame per rank) over NCCL / NVLink is the only data
that crosses GPUs; rank r then combines

    carry_r = sum_{j<r} a^(n_{j+1} + ... + n_{r-1}) * S_j

locally (at most R-1 AXPYs, va_ema_fold) and pass 2 runs the ordinary chain from that state.  Rank 0 is bit-identical to the single-GPU run; later ranks differ by the
re-association of the recurrence (<= 1e-5 relative on the background), which
decays as a^k into their chunk.

Terms older than `tail` frames are below float32 resolution (a^tail < 2^-30) and are not
folded: S_r only reads the last `tail` blurred frames of a shard, so pass 1 only blurs those
(and keeps them for pass 2).
"""

import math


def shard_range(total, rank, world):
    """ frames [a, b) of rank `rank` """
    return rank * total // world, (rank + 1) * total // world


def ema_tail(alpha, bits=30):
    """ number of trailing frames whose weight a^k is still above 2^-bits """
    if not 0 < alpha < 1:
        return 1
    return int(math.ceil(-bits * math.log(2.0) / math.log(1.0 - alpha)))


def exchange_carry(S_local, n_local, alpha, fold, group=None):
    """ all-gather the per-rank partial states and fold the ones of the ranks before us.

    S_local : torch tensor (H, pitch) float32 -- this rank's partial (rank 0: its true state)
    n_local : number of frames this rank owns
    fold    : callable(carry, S, scale) doing carry = scale * carry + S in place
    returns the background state before this rank's first frame, or None on rank 0 """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = [torch.empty_like(S_local) for _ in range(world)]
    dist.all_gather(gathered, S_local.contiguous(), group=group)
    n_t = torch.tensor([n_local], dtype=torch.int64, device=S_local.device)
    counts = [torch.empty_like(n_t) for _ in range(world)]
    dist.all_gather(counts, n_t, group=group)
    counts = [int(c.item()) for c in counts]
    if rank == 0:
        return None
    a = 1.0 - alpha
    carry = gathered[0].clone()
    for j in range(1, rank):
        fold(carry, gathered[j], a ** counts[j])
    return carry


def merge_mean_m2(mean, m2, n_local, group=None):
    """ temporal statistics of a frame-sharded video (video/analysis/video.py:26-55): every
    rank holds the running mean and the sum of squared deviations M2 of ITS frame range (float64 torch tensors,
    `m2` may be None for the mean alone) and the number of frames; one all-gather of those per-rank frames
    and the pairwise combination of Chan et al. in rank order,

        n = nA + nB,  d = meanB - meanA,  mean = meanA + d nB / n,  M2 = M2A + M2B + d^2 nA nB / n,

    give every rank the statistics of the whole video (the same exchange pattern as the background carry:
    one frame-sized message per rank, nothing else crosses GPUs).  Agrees with the sequential recurrence to
    float64 rounding, not bit for bit.  Returns (mean, m2, n). """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    means = [torch.empty_like(mean) for _ in range(world)]
    dist.all_gather(means, mean.contiguous(), group=group)
    m2s = None
    if m2 is not None:
        m2s = [torch.empty_like(m2) for _ in range(world)]
        dist.all_gather(m2s, m2.contiguous(), group=group)
    n_t = torch.tensor([n_local], dtype=torch.int64, device=mean.device)
    counts = [torch.empty_like(n_t) for _ in range(world)]
    dist.all_gather(counts, n_t, group=group)
    counts = [int(c.item()) for c in counts]
    tot_mean, tot_m2, tot_n = None, None, 0
    for j in range(world):
        nb = counts[j]
        if nb == 0:
            continue
        if tot_n == 0:
            tot_mean, tot_m2, tot_n = means[j].clone(), (m2s[j].clone() if m2s else None), nb
            continue
        n = tot_n + nb
        d = means[j] - tot_mean
        if tot_m2 is not None:
            tot_m2 += m2s[j] + d * d * (tot_n * nb / n)
        tot_mean += d * (nb / n)
        tot_n = n
    return tot_mean, tot_m2, tot_n


def measure_mean_std_sharded(video, batch=32, device=None, group=None):
    """ `analysis.video.measure_mean_std` over a video whose frames are sharded by rank (`shard_range`): each
    rank folds its own frames on its GPU, `merge_mean_m2` combines.  Returns what the reference returns:
    (mean, sqrt(M2 / index of the last frame)). """
    import numpy as np
    import torch
    import torch.distributed as dist
    from .analysis.video import _fold
    from .io.base import VideoSlice
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = shard_range(video.frame_count, rank, world)
    part = VideoSlice(video, a, b)
    mean, m2, n, last = _fold(part, True, batch, device)
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    tm, tm2, tn = merge_mean_m2(torch.from_numpy(mean).to(dev), torch.from_numpy(m2).to(dev), n, group)
    if tn - 1 < 2:
        return last, 0
    return tm.cpu().numpy(), np.sqrt(tm2.cpu().numpy() / (tn - 1))


class ShardedSegmentChain(object):
    """ runs a SegmentChain over this rank's frame range of a video sharded over the
    ranks of a torch.distributed group (NCCL) """

    def __init__(self, chain, group=None, tail=None):
        self.chain = chain
        self.group = group
        self.tail = tail if tail is not None else ema_tail(chain.alpha)
        self._blurs = []
        self._S = None
        self._scratch_mask = None

    def tail_batches(self, n_batches):
        """ how many of the last batches of a shard carry weight into the next shard """
        return min(n_batches, -(-self.tail // self.chain.batch) + 1)

    def reserve(self, n_batches):
        """ allocate the blurred-frame storage of pass 1 up front (keeps cudaMalloc out of the hot loop) """
        ch, rt = self.chain, self.chain.rt
        while len(self._blurs) < self.tail_batches(n_batches):
            self._blurs.append(rt.empty_u8(ch.batch, ch.h, ch.w))
        if self._S is None:
            self._S = rt.empty_f32(ch.h, ch.w)

    def _partial_state(self, blurs, covers_shard):
        """ S of this rank from (the last of) its blurred batches (device) """
        import torch.distributed as dist
        ch, rt = self.chain, self.chain.rt
        if self._S is None:
            self._S = rt.empty_f32(ch.h, ch.w)
        self._S.zero_()
        n = sum(b.n for b in blurs)
        rank = dist.get_rank(self.group)
        # only the last `tail` frames matter
        skip = max(0, n - self.tail)
        first = True
        seen = 0
        for b in blurs:
            lo = max(0, skip - seen)
            seen += b.n
            if lo >= b.n:
                continue
            part = b if lo == 0 else _slice_batch(b, lo, b.n)
            if first and rank == 0 and skip == 0 and covers_shard:
                # the sequential model starts from its first frame: S = float(x_0)
                if self._scratch_mask is None:
                    self._scratch_mask = rt.empty_bits(1, ch.h, ch.w)
                rt.ema_diff_thresh(_slice_batch(part, 0, 1), self._S, ch.alpha, ch.threshold, True)
                if part.n > 1:
                    rt.ema_partial(_slice_batch(part, 1, part.n), self._S, ch.alpha, True)
            else:
                rt.ema_partial(part, self._S, ch.alpha, not first)
            first = False
        return self._S

    def run_device_range(self, rgb_batches, labels_ring, counts):
        """ rgb_batches: this rank's frames as a list of DeviceBatch (n, h, w, 3), in order.
        labels_ring: list of label DeviceBatches reused round-robin; counts: int32 tensor (batch,)
        shared by all batches, or (len(rgb_batches), batch) with one row per batch.
        Results are complete once the caller's stream has passed `chain.pipeline_sync()` (done here). """
        ch, rt = self.chain, self.chain.rt
        nb = len(rgb_batches)
        # pass 1: blur the last batches of the shard -- the only frames whose weight a^k reaches the
        # next rank -- and fold them into the partial state; the blurred frames are kept for pass 2
        m = self.tail_batches(nb)
        self.reserve(nb)
        blurs = []
        for i in range(nb - m, nb):
            rgb = rgb_batches[i]
            buf = self._blurs[i - (nb - m)]
            blurs.append(ch.blur_device(rgb, buf if rgb.n == ch.batch else _slice_batch(buf, 0, rgb.n)))
        S = self._partial_state(blurs, covers_shard=(m == nb))
        n_total = sum(b.n for b in rgb_batches)
        carry = exchange_carry(S, n_total, ch.alpha, lambda c, s, sc: rt.ema_fold(c, s, sc, ch.w, ch.h), self.group)
        # pass 2: the ordinary pipelined chain (three streams) from the exact incoming state
        if carry is None:
            ch.reset()
        else:
            ch.set_background(carry)
        for i, rgb in enumerate(rgb_batches):
            ch.run_device_pipelined(rgb, labels_ring[i % len(labels_ring)], counts[i] if counts.dim() == 2 else counts,
                                    blur=blurs[i - (nb - m)] if i >= nb - m else None)
        ch.pipeline_sync()


def _slice_batch(b, lo, hi):
    from .device import DeviceBatch
    return DeviceBatch(b.kind, b.t[lo:hi], hi - lo, b.h, b.w, b.channels)
