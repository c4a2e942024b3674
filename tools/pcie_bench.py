"""
Host <-> device copy ceilings of this box: pinned H2D alone, D2H alone, and both at once, per GPU and summed over
all ranks running at the same time.  The end-to-end legs of bench.py move 6.2 MB per 1080p frame up and the result
down, so these numbers are the roofline of `e2e`.

    python tools/pcie_bench.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_bench.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29541')
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local))
    MB = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n = MB << 20
    h_up = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_dn = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_up = torch.empty(n, dtype=torch.uint8, device='cuda')
    d_dn = torch.empty(n, dtype=torch.uint8, device='cuda')
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    reps = 8

    def run(up, dn):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_up.wait_event(e0)
        s_dn.wait_event(e0)
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s_up):
                    d_up.copy_(h_up, non_blocking=True)
            if dn:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_up)
        torch.cuda.current_stream().wait_stream(s_dn)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # all ranks copy at the same time; the slowest one counts
        return reps * n / (float(ms.item()) * 1e-3) / 1e9

    out = {'n_gpus': world, 'buffer_MB': MB, 'reps': reps}
    for name, up, dn in (('h2d_alone', True, False), ('d2h_alone', False, True), ('both', True, True)):
        run(up, dn)
        g = run(up, dn)
        out[name + '_GBps_per_gpu'] = round(g, 2)
        out[name + '_GBps_all'] = round(g * world, 2)
    if rank == 0:
        out['note'] = '`both`: GB/s in EACH direction while the two run concurrently'
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
