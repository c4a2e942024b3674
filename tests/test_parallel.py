"""
Frame-range sharding: partition arithmetic and the EMA carry exchange.  The collective
plumbing runs here on CPU with the gloo backend (world_size 2 and 3); the fold callable
is NumPy so that no kernel is needed.  The GPU version of the same path is exercised by
`bench.py --gpus N`, tools/mgpu_check.py (real ranks) and, with R ranks emulated on one GPU,
tests/test_filters_gpu.py::test_sharded_equals_sequential.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ops
from video_analysis_b200.parallel import ema_tail, exchange_carry, shard_range


def test_shard_ranges_partition_the_video():
    for total in (10, 64, 1000, 20000):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_tail_length():
    assert ema_tail(0.05) == 352
    assert (1 - 0.05) ** ema_tail(0.05) < 2 ** -26 <= (1 - 0.05) ** (ema_tail(0.05) - 1)
    assert ema_tail(0.05, bits=30) == 406
    assert ema_tail(0.5) == 26


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partial(frames, alpha, init_first):
    """ NumPy statement of va_ema_partial (float32, a*S + alpha*x) """
    a, al = np.float32(1 - alpha), np.float32(alpha)
    S = frames[0].astype(np.float32) if init_first else np.zeros(frames[0].shape, np.float32)
    for f in (frames[1:] if init_first else frames):
        S = a * S + al * f.astype(np.float32)
    return S


def _worker(rank, world, port, total, alpha, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    video = rng.integers(60, 200, (total, 6, 10)).astype(np.uint8)
    a, b = shard_range(total, rank, world)
    S = torch.from_numpy(_partial(video[a:b], alpha, rank == 0))

    def fold(carry, s, scale):
        carry.mul_(np.float32(scale)).add_(s)
    counts = [shard_range(total, j, world)[1] - shard_range(total, j, world)[0] for j in range(world)]
    carry = exchange_carry(S, counts, alpha, fold)           # counts are host knowledge: nothing but S is exchanged
    # the state before this rank's first frame, from the sequential float32 model
    if rank == 0:
        assert carry is None
    else:
        _, bg = ops.background_ema(list(video[:a]), alpha, 25)
        assert np.allclose(carry.numpy(), bg, rtol=1e-5, atol=1e-4), np.abs(carry.numpy() - bg).max()
        # masks from that state agree with the sequential run except within tolerance of the threshold
        m_seq, _ = ops.background_ema(list(video[a:b]), alpha, 25, bg0=bg)
        m_par, _ = ops.background_ema(list(video[a:b]), alpha, 25, bg0=carry.numpy())
        assert (m_seq != m_par).mean() < 1e-3
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


@pytest.mark.parametrize('world', [2, 3])
def test_carry_exchange_gloo(world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 50, 0.05, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5) for _ in range(world)) == list(range(world))


def _stats_worker(rank, world, port, total, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from video_analysis_b200.parallel import merge_mean_m2
    rng = np.random.default_rng(3)
    video = rng.integers(0, 256, (total, 5, 7)).astype(np.uint8)
    a, b = shard_range(total, rank, world)
    mean, m2 = np.zeros(video.shape[1:]), np.zeros(video.shape[1:])
    for n, frame in enumerate(video[a:b]):                     # the reference's recurrence on this rank's frames
        delta = frame - mean
        mean = mean + delta / (n + 1)
        m2 = m2 + delta * (frame - mean)
    tm, tm2, tn = merge_mean_m2(torch.from_numpy(mean), torch.from_numpy(m2), b - a)
    ref_mean, ref_std = ops.measure_mean_std(video)
    assert tn == total
    assert np.allclose(tm.numpy(), ref_mean, rtol=1e-12, atol=1e-10)
    assert np.allclose(np.sqrt(tm2.numpy() / (tn - 1)), ref_std, rtol=1e-10, atol=1e-9)
    only_mean, none, _ = merge_mean_m2(torch.from_numpy(mean), None, b - a)
    assert none is None and np.allclose(only_mean.numpy(), ref_mean, rtol=1e-12, atol=1e-10)
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


@pytest.mark.parametrize('world,total', [(2, 41), (4, 3)])
def test_temporal_statistics_merge_gloo(world, total):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stats_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5) for _ in range(world)) == list(range(world))
