"""Summarise an .ncu-rep: one line per kernel launch with the counters that matter here.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls]"""
import csv
import subprocess
import sys

M = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
     'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
     'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
     'launch__registers_per_thread', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
     'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_bytes.sum',
     'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
     'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
     'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.sum',
     'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
STALL = 'smsp__average_warps_issue_stalled_'


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0][:34]
        print('== %s' % name)
        for m in M:
            if m in idx:
                print('   %-70s %s %s' % (m, r[idx[m]], units[idx[m]]))
        if '--stalls' in sys.argv:
            st = [(float(r[i]), h[len(STALL):]) for h, i in idx.items() if h.startswith(STALL) and h.endswith('_per_issue_active.ratio') and r[i]]
            for v, h in sorted(st, reverse=True)[:6]:
                print('   stall %-40s %.2f' % (h.replace('_per_issue_active.ratio', ''), v))


if __name__ == '__main__':
    main()
