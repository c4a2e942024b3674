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


def ema_tail(alpha, bits=26):
    """ number of trailing frames whose weight a^k is still above 2^-bits.  float32 carries 24 bits: a term of
    weight 2^-26 times a value <= 255 changes a background value of the usual size by less than half a unit in
    the last place """
    if not 0 < alpha < 1:
        return 1
    return int(math.ceil(-bits * math.log(2.0) / math.log(1.0 - alpha)))


def fold_carry(gathered, counts, rank, alpha, fold):
    """ the background state before rank `rank`'s first frame from the partial states of the ranks before it:

        carry = S_0;  carry = a^(n_j) * carry + S_j   for j = 1 .. rank-1          (a = 1 - alpha)

    gathered : sequence / tensor indexed by rank of the (H, pitch) float32 partials
    counts   : frames owned by every rank (host integers -- every rank knows every `shard_range`)
    fold     : callable(carry, S, scale) doing carry = scale * carry + S in place
    The fold runs in place in gathered[0] (the gather buffer belongs to this rank).  None on rank 0. """
    if rank == 0:
        return None
    a = 1.0 - alpha
    carry = gathered[0]
    for j in range(1, rank):
        fold(carry, gathered[j], a ** int(counts[j]))
    return carry


def exchange_carry(S_local, counts, alpha, fold, group=None, out=None):
    """ all-gather the per-rank partial states and fold the ones of the ranks before us.

    S_local : torch tensor (H, pitch) float32 -- this rank's partial (rank 0: its true state)
    counts  : frames owned by every rank, host integers (an int means "the same on every rank");
              nothing about the counts crosses the network and nothing is read back from the device
    fold    : callable(carry, S, scale) doing carry = scale * carry + S in place
    out     : optional preallocated (world, H, pitch) gather buffer
    One `all_gather_into_tensor` of one frame per rank is the only communication.
    returns the background state before this rank's first frame, or None on rank 0 """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if isinstance(counts, int):
        counts = [counts] * world
    if len(counts) != world:
        raise ValueError('%d frame counts for %d ranks' % (len(counts), world))
    S_local = S_local.contiguous()
    if out is None:
        out = torch.empty((world,) + tuple(S_local.shape), dtype=S_local.dtype, device=S_local.device)
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(out, S_local, group=group)
    else:                                               # gloo (CPU tests): gather into the rows of the same buffer
        dist.all_gather(list(out.unbind(0)), S_local, group=group)
    return fold_carry(out, counts, rank, alpha, fold)


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
    """ runs a SegmentChain over this rank's frame range of a video sharded over the ranks of a
    torch.distributed group (NCCL).

    The work of one range is three phases, all enqueued without a host synchronisation:

      pass1     on the chain's front stream: blur the last `tail` frames of the range and fold them into the
                partial state S (the blurred batches are kept for pass 2)
      exchange  one asynchronous `all_gather_into_tensor` of S (a frame per rank); while it is in flight the
                front stream already blurs the first batches of the range (only the EMA needs the carry)
      pass2     fold the carry from the gathered partials (<= world-2 AXPYs), then the ordinary three-stream
                pipelined chain from that state

    `run_device_range` strings them together.  They are separate methods so that a test can emulate R ranks
    on ONE GPU: pass1 for every rank, stack the partials in place of the all-gather, pass2 for every rank. """

    PREBLUR = 1               # head batches blurred while the all-gather is in flight (one 128-frame blur outlasts the gather)

    def __init__(self, chain, group=None, tail=None, rank=None, world=None):
        self.chain = chain
        self.group = group
        self.tail = tail if tail is not None else ema_tail(chain.alpha)
        if rank is None or world is None:
            import torch.distributed as dist
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.rank, self.world = int(rank), int(world)
        self._blurs = []
        self._head = []
        self._S = None
        self._G = None

    def tail_batches(self, n_batches, last_n=None):
        """ how many of the last batches of a shard carry weight into the next shard: the ones that hold its
        last `tail` frames (`last_n`: frames in the last batch when it is a ragged one) """
        B = self.chain.batch
        last_n = B if last_n is None else last_n
        need = max(0, self.tail - last_n)
        return min(n_batches, 1 + -(-need // B))

    def reserve(self, n_batches):
        """ allocate the storage of pass 1 / the exchange up front (keeps cudaMalloc out of the hot loop) """
        ch, rt = self.chain, self.chain.rt
        m = min(n_batches, self.tail_batches(n_batches) + 1)      # + 1: a ragged last batch needs one more
        while len(self._blurs) < m:
            self._blurs.append(rt.empty_u8(ch.batch, ch.h, ch.w))
        while len(self._head) < min(self.PREBLUR, n_batches - m):
            self._head.append(rt.empty_u8(ch.batch, ch.h, ch.w))
        if self._S is None:
            self._S = rt.empty_f32(ch.h, ch.w)
        if self._G is None:
            t = torch_mod()
            self._G = t.empty((self.world,) + tuple(self._S.shape), dtype=t.float32, device=rt.device)
        ch.pipeline_streams()

    def _fold_batch(self, b, lo, first, init_from_first_frame):
        """ fold frames lo.. of the blurred batch `b` into S (current stream) """
        ch, rt = self.chain, self.chain.rt
        part = b if lo == 0 else _slice_batch(b, lo, b.n)
        if init_from_first_frame:
            # the sequential model starts from its first frame: S = float(x_0)
            rt.ema_diff_thresh(_slice_batch(part, 0, 1), self._S, ch.alpha, ch.threshold, True)
            if part.n > 1:
                rt.ema_partial(_slice_batch(part, 1, part.n), self._S, ch.alpha, True)
        else:
            rt.ema_partial(part, self._S, ch.alpha, not first)

    # ---- the three phases ---------------------------------------------------------------------------
    def pass1(self, rgb_batches):
        """ blur the last batches of the shard -- the only frames whose weight a^k reaches the next rank -- and
        fold them into the partial state S (returned; valid on the chain's front stream).  The blurs run on the
        front stream; the fold of batch i (memory-bound) runs on the back stream beside the blur of batch i + 1
        (issue-bound). """
        ch, rt, t = self.chain, self.chain.rt, torch_mod()
        nb = len(rgb_batches)
        m = self.tail_batches(nb, rgb_batches[-1].n)
        self.reserve(nb)
        streams = ch.pipeline_streams()
        front, back = streams['front'], streams['back']
        ready = t.cuda.Event()
        ready.record(t.cuda.current_stream(rt.device))
        n = sum(rgb_batches[i].n for i in range(nb - m, nb))
        skip = max(0, n - self.tail)                      # only the last `tail` frames matter
        covers_shard = (m == nb)
        blurs, seen, first = [], 0, True
        front.wait_event(ready)
        back.wait_event(ready)
        with t.cuda.stream(back):
            self._S.zero_()
        for i in range(nb - m, nb):
            rgb = rgb_batches[i]
            buf = self._blurs[i - (nb - m)]
            with t.cuda.stream(front):
                b = ch.blur_device(rgb, buf if rgb.n == ch.batch else _slice_batch(buf, 0, rgb.n))
                done = t.cuda.Event()
                done.record(front)
            blurs.append(b)
            lo = max(0, skip - seen)
            seen += b.n
            if lo >= b.n:
                continue
            with t.cuda.stream(back):
                back.wait_event(done)
                self._fold_batch(b, lo, first, first and self.rank == 0 and skip == 0 and covers_shard)
            first = False
        folded = t.cuda.Event()
        folded.record(back)
        front.wait_event(folded)                          # S is complete for whatever `front` does next
        self._p1 = (nb, m, blurs)
        return self._S

    def exchange(self, S):
        """ start the all-gather of the partial states (asynchronous); -> (gather buffer, work handle) """
        import torch.distributed as dist
        t = torch_mod()
        front = self.chain.pipeline_streams()['front']
        with t.cuda.stream(front):                       # NCCL's stream waits for what `front` holds now: S is complete
            work = dist.all_gather_into_tensor(self._G, S, group=self.group, async_op=True)
        return self._G, work

    def preblur(self, rgb_batches):
        """ blur the first batches of the range on the front stream (they do not need the carry) """
        ch, t = self.chain, torch_mod()
        nb, m, _ = self._p1
        front = ch.pipeline_streams()['front']
        head = []
        with t.cuda.stream(front):
            for i in range(min(len(self._head), nb - m)):
                rgb = rgb_batches[i]
                buf = self._head[i]
                head.append(ch.blur_device(rgb, buf if rgb.n == ch.batch else _slice_batch(buf, 0, rgb.n)))
        return head

    def pass2(self, rgb_batches, labels_ring, counts, gathered, shard_counts, work=None, head=()):
        """ fold the carry out of `gathered` ((world, H, pitch) partial states) and run the pipelined chain over
        the range.  `shard_counts`: frames owned by every rank (host integers).  Ends with `pipeline_sync()`. """
        ch, rt, t = self.chain, self.chain.rt, torch_mod()
        nb, m, blurs = self._p1
        front = ch.pipeline_streams()['front']
        with t.cuda.stream(front):
            if work is not None:
                work.wait()                               # stream-side wait: `front` continues after the gather
            carry = fold_carry(gathered, shard_counts, self.rank, ch.alpha,
                               lambda c, s, sc: rt.ema_fold(c, s, sc, ch.w, ch.h))
            if carry is None:
                ch.reset()
            else:
                ch.set_background(carry)
        for i, rgb in enumerate(rgb_batches):
            blur = blurs[i - (nb - m)] if i >= nb - m else (head[i] if i < len(head) else None)
            ch.run_device_pipelined(rgb, labels_ring[i % len(labels_ring)], counts[i] if counts.dim() == 2 else counts,
                                    blur=blur)
        ch.pipeline_sync()

    def run_device_range(self, rgb_batches, labels_ring, counts, shard_counts=None):
        """ rgb_batches: this rank's frames as a list of DeviceBatch (n, h, w, 3), in order.
        labels_ring: list of label DeviceBatches reused round-robin; counts: int32 tensor (batch,)
        shared by all batches, or (len(rgb_batches), batch) with one row per batch.
        shard_counts: frames owned by every rank (default: every rank owns as many as this one).
        Results are complete once the caller's stream has passed `chain.pipeline_sync()` (done here). """
        n_total = sum(b.n for b in rgb_batches)
        if shard_counts is None:
            shard_counts = [n_total] * self.world
        S = self.pass1(rgb_batches)
        gathered, work = self.exchange(S)
        head = self.preblur(rgb_batches)
        self.pass2(rgb_batches, labels_ring, counts, gathered, shard_counts, work, head)


def torch_mod():
    from .device import torch
    return torch()


def _slice_batch(b, lo, hi):
    from .device import DeviceBatch
    return DeviceBatch(b.kind, b.t[lo:hi], hi - lo, b.h, b.w, b.channels)
