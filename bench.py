#!/usr/bin/env python
"""
bench.py -- frames/sec of the full filter + segment chain (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- 1920x1080 RGB uint8 synthetic video,
chain mono -> blur sigma=2 -> running background (alpha .05) -> |diff| > 25 -> 3x3 open ->
label (4-conn).  A step is one batch of `--batch` frames through the whole chain.

  value      device-resident: the synthetic video lives in HBM (generated there by
             va_synth_rgb), every step reads a different batch of it (working set >> L2)
  e2e        the same chain through the plug-in API (SegmentChain.process_blocks) with the
             frames in pinned HOST memory and the int32 labels + counts copied back to host,
             copies inside the timed region
  roofline   the kernel with the largest share of the step, timed with CUDA events around
             its launches on the launch stream in a second, instrumented pass over the same
             steps; achieved = algorithmic bytes / duration (DESIGN.md states the bytes)
  cpu_baseline  the oracle (same cv2 / NumPy / SciPy calls as the reference) on this box's
             host cores, one process, one frame per iteration, on a bounded sample

`--impl reference` times that CPU chain alone, frame-sharded over all host cores.
N > 1 (torchrun): frames are sharded by contiguous ranges, one process per GPU, the only
exchange is the background-EMA carry (video_analysis_b200/parallel.py).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
CHAIN = dict(sigma=2.0, alpha=0.05, threshold=25.0, morph_op='open', morph_shape='rect', morph_ksize=3, connectivity=4)
WORKLOAD = '1920x1080 RGB uint8 synthetic video, mono->blur s=2->EMA bg a=.05->|diff|>25->3x3 open->label(4)'


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def pcie_roofline(h2d_bytes, d2h_bytes, fps_per_gpu, frames_per_step):
    """ the end-to-end legs are bound by the host link: bytes per step in the busier direction against the pinned-copy
    bandwidth measured on this pool with both directions running (tools/pcie_bench.py -> profiles/pcie_r2.txt) """
    peak = None
    try:
        for line in open(os.path.join(ROOT, 'profiles', 'pcie_r2.txt')):
            d = json.loads(line)
            if d.get('n_gpus') == 1:
                peak = (d['h2d_alone_GBps_per_gpu'], d['d2h_alone_GBps_per_gpu'], d['both_GBps_per_gpu'])
    except Exception:
        pass
    if peak is None:
        return None
    both = d2h_bytes > 0.25 * h2d_bytes                 # a thin return stream leaves the upload at its stand-alone rate
    link = peak[2] if both else peak[0]
    busy = max(h2d_bytes, d2h_bytes)
    achieved = busy * fps_per_gpu / frames_per_step / 1e9
    return {'bound': 'pcie', 'achieved': round(achieved, 2), 'peak': link, 'unit': 'GB/s', 'frac': round(achieved / link, 3),
            'peak_source': 'profiles/pcie_r2.txt (N=1: %s GB/s in the busier direction)' % ('both directions busy' if both else 'upload alone')}


# ------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    """ SM clock and throttle reasons while the timed region runs.  NVML is polled from a thread every millisecond (the
    timed region of a 20-step run is 12 ms: `nvidia-smi -lms` would not deliver a single sample inside it); samples
    are stamped, and the ones between `mark_begin` and `mark_end` are reported.  Falls back to `nvidia-smi -lms 20`
    when NVML cannot be loaded. """
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.samples, self.t0, self.t1 = [], None, None
        self._stop = False
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber the devices: address the GPU by its PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.index), 'pci_bus_id') else None
            handle = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                        handle = h
                        break
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml, self.handle = pynvml, handle
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        get_reasons = getattr(n, 'nvmlDeviceGetCurrentClocksEventReasons', None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                reasons = int(get_reasons(self.handle))
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.nvml is not None:
            time.sleep(0.01)
            self._stop = True
            self.thread.join(timeout=1)
            inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= (self.t1 or s[0])]
            where = 'inside the timed region'
            if not inside and self.samples:                        # a region shorter than one poll: the samples around it
                mid = 0.5 * ((self.t0 or 0) + (self.t1 or 0))
                inside = sorted(self.samples, key=lambda s: abs(s[0] - mid))[:3]
                where = 'nearest to the timed region'
            n = self.nvml
            bits = {'hw_slowdown': getattr(n, 'nvmlClocksEventReasonHwSlowdown', 0x8),
                    'hw_thermal_slowdown': getattr(n, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                    'sw_thermal_slowdown': getattr(n, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
                    'sw_power_cap': getattr(n, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
            reasons = [k for k, bit in bits.items() if any(s[2] & bit for s in inside)]
            return {'sm_mhz': float(np.median([s[1] for s in inside])) if inside else None, 'sm_max_mhz': self.max_mhz,
                    'reasons': reasons, 'samples': len(inside), 'source': 'NVML polled every ms, samples ' + where}
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm), 'source': 'nvidia-smi -lms 20'}


# ------------------------------------------------------------------------------------------
# CPU reference chain (oracle): the only places bench.py touches oracle/
# ------------------------------------------------------------------------------------------
_W = {}


def cpu_frames(n_unique, w=None, h=None):
    from oracle import synth
    return synth.make_frames(0, 0, n_unique, w or W, h or H, 8)


def _reference_front():
    """ the reference's OWN FilterMonochrome / FilterBlur classes (oracle/_ref, converted by oracle/build_ref.py) when that
    tree travelled with the repository; None otherwise (the oracle restates the same two calls) """
    try:
        from oracle import build_ref
        return build_ref.import_ref()
    except Exception:
        return None


def cpu_chain_run(frames, n_frames, bg=None, stages=None):
    """ reference-style loop: one frame per iteration; returns (seconds, bg).  `stages` (dict) accumulates seconds per
    stage.  Monochrome and blur run through the reference's own filter classes when oracle/_ref is there; the steps
    the reference does not have (EMA / threshold / open) and the label call are the oracle's. """
    from oracle import ops
    alpha32, thr32 = np.float32(CHAIN['alpha']), np.float32(CHAIN['threshold'])
    ref = _W.get('ref', False)
    if ref is False:
        ref = _W['ref'] = _reference_front()
    tick = time.perf_counter
    acc = stages if stages is not None else {}
    for k in ('mono', 'blur', 'ema_diff_thresh', 'morph_open', 'label'):
        acc.setdefault(k, 0.0)
    if ref is not None:
        class _One(object):                    # a one-frame source for the reference's filter objects
            size, frame_count, fps, is_color = (frames.shape[2], frames.shape[1]), 1, 25, True
        mono_f = ref.filters.FilterMonochrome(_One())
        blur_f = ref.filters.FilterBlur(mono_f, CHAIN['sigma'])
    t0 = tick()
    for i in range(n_frames):
        f = frames[i % len(frames)]
        t1 = tick()
        m = mono_f._process_frame(f) if ref is not None else ops.mono(f)
        t2 = tick()
        b = blur_f._process_frame(m) if ref is not None else ops.blur(m, CHAIN['sigma'])
        t3 = tick()
        x = b.astype(np.float32)
        if bg is None:
            bg = x.copy()
            mask = np.zeros(b.shape, np.uint8)
        else:
            d = x - bg
            mask = np.where(np.abs(d) > thr32, 255, 0).astype(np.uint8)
            bg = bg + alpha32 * d
        t4 = tick()
        mo = ops.morph(mask, CHAIN['morph_op'], CHAIN['morph_shape'], CHAIN['morph_ksize'])
        t5 = tick()
        ops.label(mo, CHAIN['connectivity'])
        t6 = tick()
        acc['mono'] += t2 - t1
        acc['blur'] += t3 - t2
        acc['ema_diff_thresh'] += t4 - t3
        acc['morph_open'] += t5 - t4
        acc['label'] += t6 - t5
    return tick() - t0, bg


def cpu_baseline_single(budget_s=12.0):
    import cv2
    frames = cpu_frames(8)
    dt, bg = cpu_chain_run(frames, 8)                 # warm-up + rate estimate
    n = int(max(16, min(400, budget_s / (dt / 8))))
    stages = {}
    dt, _ = cpu_chain_run(frames, n, bg, stages)
    kind = 'reference' if _W.get('ref') is not None else 'port'
    # configs[0] (the reference's own CPU-runnable case): 640x480, bounded sample
    vga = cpu_frames(8, 640, 480)
    dv, bgv = cpu_chain_run(vga, 8)
    nv = int(max(32, min(1000, 4.0 / (dv / 8))))
    vstages = {}
    dv, _ = cpu_chain_run(vga, nv, bgv, vstages)
    return {'value': round(n / dt, 2), 'unit': 'frames/s', 'cores': int(cv2.getNumThreads()), 'kind': kind,
            'sample': '%d frames 1080p (8 unique synthetic frames cycled), one process, one frame per iteration, '
                      'cv2 threads=%d; monochrome + blur through %s, EMA / threshold / open / label through the oracle '
                      '(absent from the reference, SURVEY 8c)'
                      % (n, cv2.getNumThreads(), "the reference's own filter classes (oracle/_ref)" if kind == 'reference' else 'the oracle'),
            'per_stage_ms': {k: round(v / n * 1e3, 3) for k, v in stages.items()},
            'vga_640x480': {'value': round(nv / dv, 2), 'unit': 'frames/s', 'sample': '%d frames' % nv,
                            'per_stage_ms': {k: round(v / nv * 1e3, 3) for k, v in vstages.items()}}}


def _worker_init():
    import cv2
    cv2.setNumThreads(1)
    _W['frames'] = cpu_frames(4)
    _W['bg'] = None


def _worker_step(n):
    dt, _W['bg'] = cpu_chain_run(_W['frames'], n, _W['bg'])
    return dt


def reference_arm(args):
    """ the reference's CPU chain on all host cores: frames sharded over processes """
    import multiprocessing as mp
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    procs = max(1, min(os.cpu_count() or 1, 64))
    per = 4                                           # frames per process per step
    ctx = mp.get_context('fork')
    with ctx.Pool(procs, initializer=_worker_init) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_worker_step, [per] * procs)
        steps = min(args.steps, 20)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_worker_step, [per] * procs)
        dt = time.perf_counter() - t0
    fps = steps * procs * per / dt
    line = {
        'impl': 'reference', 'metric': 'frames/sec full filter+segment chain', 'value': round(fps, 2), 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': max(args.warmup, 1), 'ms_per_step': round(dt / steps * 1e3, 3),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'frames_per_step': procs * per},
        'cpu_baseline': {'value': round(fps, 2), 'unit': 'frames/s', 'cores': procs,
                         'kind': 'reference' if _reference_front() is not None else 'port',
                         'sample': '%d steps x %d processes x %d frames 1080p, frame-sharded, cv2 threads=1 per process; '
                                   'monochrome + blur through the reference\'s own filter classes where oracle/_ref is present '
                                   '(Python-3 conversion of /root/reference, oracle/build_ref.py), the steps the reference lacks '
                                   'through the oracle' % (steps, procs, per)},
        'e2e': {'value': round(fps, 2), 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from video_analysis_b200 import synth
    from video_analysis_b200.chain import SegmentChain
    from video_analysis_b200.device import DeviceBatch, get_runtime

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner (NCCL_DEBUG=VERSION in this image) with a
        # plain printf when the communicator is created, so file descriptor 1 points at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    rt = get_runtime(local)
    B, K, Wm = args.batch, args.steps, args.warmup
    N = W * H
    hbm_peak, peak_src = peaks()

    # ---- the synthetic video of this rank, resident in HBM ------------------------------------
    n_frames = min(args.frames, (K + Wm) * B)
    n_frames = max(B, n_frames // B * B)
    free, _ = torch.cuda.mem_get_info()
    while n_frames * N * 3 > free * 0.45 and n_frames > 2 * B:
        n_frames = n_frames // 2 // B * B
    t0_rank = rank * args.frames                       # rank r owns frames [r*T, (r+1)*T) of the global video
    video = torch.empty((n_frames, H, W * 3), dtype=torch.uint8, device=rt.device)
    for a in range(0, n_frames, B):
        synth.generate(rt, 0, t0_rank + a, B, W, H, out=video[a:a + B])
    torch.cuda.synchronize()

    def batch_of(step):
        a = (step * B) % n_frames
        return DeviceBatch('u8', video[a:a + B], B, H, W, 3)

    chain = SegmentChain((W, H), batch=B, fuse=not args.no_fuse, **CHAIN)
    labels = [rt.empty_i32(B, H, W) for _ in range(2)]
    counts = torch.empty((B,), dtype=torch.int32, device=rt.device)

    if world > 1:
        from video_analysis_b200.parallel import ShardedSegmentChain
        sharded = ShardedSegmentChain(chain)
        sharded.reserve(K)

    def run_steps(first, n):
        if world > 1:
            sharded.run_device_range([batch_of(first + i) for i in range(n)], labels, counts)
        else:
            for i in range(n):
                if args.no_overlap:
                    chain.run_device(batch_of(first + i), labels[i & 1], counts)
                else:       # blur + EMA of batch k+2, forest of k+1 and label write of k overlap (three streams)
                    chain.run_device_pipelined(batch_of(first + i), labels[i & 1], counts)
            if not args.no_overlap:
                chain.pipeline_sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident ------------------------------------------------------------------
    run_steps(0, Wm)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = rt.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.mark_begin()
    ev0.record()
    run_steps(Wm, K)
    ev1.record()
    barrier()
    if rank == 0:
        sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    launches = rt.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([ms], device=rt.device)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    fps = world * K * B / (ms * 1e-3)

    # ---- the same device-resident steps with int16 label images (ndimage.label(..., output=np.int16)): extra key only
    fps_i16 = None
    if world == 1 and not args.no_overlap:
        ch16 = SegmentChain((W, H), batch=B, fuse=not args.no_fuse, label_dtype=np.int16, **CHAIN)
        labels16 = [rt.empty_i16(B, H, W) for _ in range(2)]
        for i in range(Wm):
            ch16.run_device_pipelined(batch_of(i), labels16[i & 1], counts)
        ch16.pipeline_sync()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            ch16.run_device_pipelined(batch_of(Wm + i), labels16[i & 1], counts)
        ch16.pipeline_sync()
        e1.record()
        torch.cuda.synchronize()
        fps_i16 = K * B / (e0.elapsed_time(e1) * 1e-3)
        del ch16, labels16

    # ---- roofline of the dominant kernel: instrumented pass, events around each launch group ------
    roof = None
    if rank == 0:
        roof = instrumented_pass(rt, chain, batch_of, labels, counts, Wm, K, B, N, hbm_peak, peak_src, args)

    # ---- e2e: host frames -> labels on host ---------------------------------------------------------
    Be = min(B, args.e2e_batch)             # the host pipeline is PCIe-bound at any batch size; smaller blocks pin less memory
    ring = max(2, min(4, n_frames // Be))
    host = torch.empty((ring * Be, H, W, 3), dtype=torch.uint8, pin_memory=True)
    host.view(ring * Be, H, W * 3).copy_(video[:ring * Be])
    torch.cuda.synchronize()
    host_np = host.numpy()
    e2e_chain = SegmentChain((W, H), batch=Be, fuse=not args.no_fuse, **CHAIN)
    Ke = max(3, min(K, args.e2e_steps))

    def blocks(n):
        for i in range(n):
            a = (i % ring) * Be
            yield host_np[a:a + Be]

    sink = [0]
    E2E_REPEATS = 2

    def e2e_leg(ch, touch, **kw):
        """ warm-up blocks, then Ke timed blocks through process_blocks, E2E_REPEATS times: (best frames/s, every run).
        The host side of these boxes is shared (PCIe copies were seen at half speed for a whole run), so the line
        carries the best run and lists all of them """
        for res in ch.process_blocks(blocks(max(3, min(Wm, 5))), **kw):
            sink[0] += int(res[1][0])
        runs = []
        for _ in range(E2E_REPEATS):
            barrier()
            ch.egress_bytes = 0
            t0 = time.perf_counter()
            for res in ch.process_blocks(blocks(Ke), **kw):
                sink[0] += touch(res)                                  # touch the results on the host
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                tmax = torch.tensor([dt], device=rt.device)
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dt = float(tmax.item())
            runs.append(round(world * Ke * Be / dt, 1))
        return max(runs), runs

    def touch_labels(res):
        return int(res[1][-1]) + int(res[0][0, H // 2, W // 2])

    e2e_fps, e2e_runs = e2e_leg(e2e_chain, touch_labels)
    e2e_d2h = e2e_chain.egress_bytes // Ke                       # what the device stored into host memory per step

    # ---- the same with a dense device -> host copy of the label images (round 1's e2e), for comparison
    dense_chain = SegmentChain((W, H), batch=Be, fuse=not args.no_fuse, sparse_egress=False, **CHAIN)
    dense_fps, dense_runs = e2e_leg(dense_chain, touch_labels)
    del dense_chain

    # ---- e2e with the region table as the result: same chain, same host input, but
    # per-region moments / boxes come back instead of the 4-bytes-per-pixel label image
    MAXR = 256
    reg_fps, reg_runs = e2e_leg(e2e_chain, lambda res: int(res[1][-1]) + int(res[0][0, 0, 0]) + int(res[2][0]), max_regions=MAXR)

    # ---- e2e with int16 labels (ndimage.label(..., output=np.int16)): the same label image in half the bytes
    i16_chain = SegmentChain((W, H), batch=Be, fuse=not args.no_fuse, label_dtype=np.int16, **CHAIN)
    i16_fps, i16_runs = e2e_leg(i16_chain, touch_labels)

    if rank == 0:
        line = {
            'metric': 'frames/sec full filter+segment chain', 'value': round(fps, 1), 'unit': 'frames/s',
            'n_gpus': world, 'steps': K, 'warmup': Wm, 'ms_per_step': round(ms / K, 4),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'frames_per_step': B, 'resident_frames_per_gpu': n_frames,
                       'parallelism': 'frame-range shards x%d, EMA carry via NCCL all-gather' % world if world > 1 else 'single GPU',
                       'l2': 'every step reads a different %d MB batch of the resident video (>> 126 MB L2)' % (B * N * 3 >> 20),
                       'fused_luma_blur': not args.no_fuse,
                       'two_stream_overlap': (not args.no_overlap) or world > 1,
                       'pipeline_streams': 1 if (args.no_overlap and world == 1) else 3},
            'clocks': clocks,
            'e2e': {'value': round(e2e_fps, 1), 'unit': 'frames/s', 'h2d_bytes_per_step': Be * N * 3,
                    'd2h_bytes_per_step': int(e2e_d2h), 'steps': Ke, 'frames_per_step': Be, 'runs': e2e_runs,
                    'policy': 'best of %d runs of `steps` blocks (all listed in `runs`)' % E2E_REPEATS,
                    'api': 'SegmentChain.process_blocks (pinned host frames in; dense int32 label images + counts out, '
                           'brought over PCIe as their non-empty 64-label chunks and rebuilt on the host)',
                    'roofline': pcie_roofline(Be * N * 3, int(e2e_d2h), e2e_fps / world, Be)},
            'e2e_dense_copy': {'value': round(dense_fps, 1), 'unit': 'frames/s', 'h2d_bytes_per_step': Be * N * 3,
                               'd2h_bytes_per_step': Be * N * 4 + Be * 4, 'steps': Ke, 'frames_per_step': Be, 'runs': dense_runs,
                               'api': 'SegmentChain(sparse_egress=False).process_blocks: the label images copied densely',
                               'roofline': pcie_roofline(Be * N * 3, Be * N * 4 + Be * 4, dense_fps / world, Be)},
            'e2e_region_table': {'value': round(reg_fps, 1), 'unit': 'frames/s', 'h2d_bytes_per_step': Be * N * 3,
                                 'd2h_bytes_per_step': Be * (MAXR * 80 + 8), 'steps': Ke, 'frames_per_step': Be, 'runs': reg_runs,
                                 'api': 'SegmentChain.process_blocks(max_regions=%d): pinned host frames in, per-region '
                                        'moments + bounding boxes + counts out (no label image crosses PCIe)' % MAXR},
            'e2e_labels_int16': {'value': round(i16_fps, 1), 'unit': 'frames/s', 'h2d_bytes_per_step': Be * N * 3,
                                 'd2h_bytes_per_step': Be * N * 2 + Be * 4, 'steps': Ke, 'frames_per_step': Be, 'runs': i16_runs,
                                 'api': 'SegmentChain(label_dtype=np.int16).process_blocks: pinned host frames in, int16 labels '
                                        '(ndimage.label(..., output=np.int16)) + counts out'},
            'value_labels_int16': None if fps_i16 is None else round(fps_i16, 1),
            'gpu_launches': int(launches),
            'roofline': roof,
        }
        if world == 1 and not args.no_cpu:
            line['cpu_baseline'] = cpu_baseline_single()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def instrumented_pass(rt, chain, batch_of, labels, counts, Wm, K, B, N, hbm_peak, peak_src, args):
    """ same steps, one C-ABI call per kernel with CUDA events in between """
    import torch
    from video_analysis_b200 import _lib
    fuse = not args.no_fuse
    blur, mono = rt.empty_u8(B, H, W), rt.empty_u8(B, H, W)
    mask, morph = rt.empty_bits(B, H, W), rt.empty_bits(B, H, W)
    bg = rt.empty_f32(H, W)
    lib, h = rt.lib, rt._h
    # one entry per launch group; the labelling is timed in its two halves -- the union-find forest (four small kernels,
    # latency- / atomics-bound, N/8 bytes in) and the write of the label image (one kernel, 4N bytes out) -- and also
    # reported as the unit SURVEY 8d defines (`label`: N/8 + 4N)
    stages = (['luma_gauss'] if fuse else ['luma', 'gauss']) + ['ema_diff_thresh', 'morph_open', 'label_forest', 'label_write']
    alg_bytes = {'luma_gauss': 4 * N, 'luma': 4 * N, 'gauss': 2 * N, 'ema_diff_thresh': N + N / 8 + 8 * N / B,
                 'morph_open': N / 4, 'label_forest': N / 8, 'label_write': N / 8 + 4 * N, 'label': N / 8 + 4 * N}
    kernel_names = {'luma_gauss': 'gauss_mma_kernel<fused luma>', 'luma': 'luma_fast_kernel', 'gauss': 'gauss_mma_kernel',
                    'ema_diff_thresh': 'ema_diff_thresh_kernel', 'morph_open': 'morph_stream_kernel<3,2>',
                    'label_forest': 'label_init + label_merge + label_flatten + label_scan', 'label_write': 'label_write_kernel'}
    tot = {s: 0.0 for s in stages}
    n_timed = 0
    all_evs = []            # the steps are enqueued back to back (no host sync in between): an event then sits directly
                            # between two kernels on the stream and no launch latency of an idle GPU leaks into a duration
    for step in range(Wm + K):
        rgb = batch_of(step)
        lab = labels[step & 1]
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
        evs[0].record()
        i = 1
        if fuse:
            rt._check(lib.va_luma_gauss_u8(h, rt.stream, *rgb.img(), *blur.img(), W, H, B, -1, CHAIN['sigma'])); evs[i].record(); i += 1
        else:
            rt._check(lib.va_luma_u8(h, rt.stream, *rgb.img(), *mono.img(), W, H, B, -1)); evs[i].record(); i += 1
            rt._check(lib.va_gauss_u8(h, rt.stream, *mono.img(), *blur.img(), W, H, 1, B, CHAIN['sigma'])); evs[i].record(); i += 1
        rt._check(lib.va_ema_diff_thresh(h, rt.stream, *blur.img(), bg.data_ptr(), bg.stride(0), *mask.img(), W, H, B,
                                         CHAIN['alpha'], CHAIN['threshold'], 1 if step == 0 else 0)); evs[i].record(); i += 1
        rt._check(lib.va_morph_bits(h, rt.stream, *mask.img(), *morph.img(), W, H, B, _lib.MORPH_OPS['open'],
                                    _lib.SE_SHAPES['rect'], 3, 3)); evs[i].record(); i += 1
        rt._check(lib.va_label_forest(h, rt.stream, *morph.img(), counts.data_ptr(), W, H, B, 4, 0)); evs[i].record(); i += 1
        rt._check(lib.va_label_write(h, rt.stream, *morph.img(), *lab.img(), W, H, B, 0)); evs[i].record()
        if step >= Wm:
            all_evs.append(evs)
    torch.cuda.synchronize()
    for evs in all_evs:
        n_timed += 1
        for j, s in enumerate(stages):
            tot[s] += evs[j].elapsed_time(evs[j + 1])
    avg = {s: tot[s] / n_timed for s in stages}
    step_ms = sum(avg.values())
    dom = max(avg, key=avg.get)                     # the launch (group) with the largest share of the step
    achieved = alg_bytes[dom] * B / (avg[dom] * 1e-3) / 1e9
    per_kernel = {s: {'ms': round(avg[s], 4), 'share': round(avg[s] / step_ms, 3),
                      'alg_GBps': round(alg_bytes[s] * B / (avg[s] * 1e-3) / 1e9, 1),
                      'frac': round(alg_bytes[s] * B / (avg[s] * 1e-3) / 1e9 / hbm_peak, 3)} for s in stages}
    lab_ms = avg['label_forest'] + avg['label_write']
    per_kernel['label'] = {'ms': round(lab_ms, 4), 'share': round(lab_ms / step_ms, 3),
                           'alg_GBps': round(alg_bytes['label'] * B / (lab_ms * 1e-3) / 1e9, 1),
                           'frac': round(alg_bytes['label'] * B / (lab_ms * 1e-3) / 1e9 / hbm_peak, 3),
                           'note': 'label_forest + label_write: the unit SURVEY 8d accounts (N/8 + 4N bytes)'}
    traffic = None
    try:        # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed ncu --set full capture
        for tr in json.load(open(os.path.join(ROOT, 'profiles', 'traffic_r2.json'))).get(dom, []):
            if tr.get('batch') == B:
                traffic = int(tr['dram_bytes_read'] + tr['dram_bytes_write'])
    except Exception:
        pass
    return {'bound': 'hbm', 'kernel': dom, 'kernel_name': kernel_names.get(dom, dom), 'achieved': round(achieved, 1), 'peak': hbm_peak, 'unit': 'GB/s',
            'frac': round(achieved / hbm_peak, 3), 'traffic': traffic, 'peak_source': peak_src,
            'launch_ms': round(avg[dom], 4), 'alg_bytes_per_launch': int(alg_bytes[dom] * B),
            'per_kernel': per_kernel}


def main():
    global W, H, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--batch', type=int, default=128, help='frames per step of the device-resident chain')
    ap.add_argument('--e2e-batch', type=int, default=64, help='frames per block of the host pipeline (e2e)')
    ap.add_argument('--frames', type=int, default=10000, help='frames of the synthetic video per GPU')
    ap.add_argument('--e2e-steps', type=int, default=30)
    ap.add_argument('--no-fuse', action='store_true')
    ap.add_argument('--no-overlap', action='store_true', help='one stream, one va_chain_run call per step')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--width', type=int, default=W, help='frame width (default: configs[1], 1920; configs[2] is 3840)')
    ap.add_argument('--height', type=int, default=H, help='frame height (default 1080; configs[2] is 2160)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if (args.width, args.height) != (W, H):
        WORKLOAD = WORKLOAD.replace('%dx%d' % (W, H), '%dx%d' % (args.width, args.height))
        W, H = args.width, args.height
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
