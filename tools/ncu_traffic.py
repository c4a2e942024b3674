"""
DRAM traffic per chain stage from an `ncu --set full` capture of tools/prof_once.py: the LAST profiled launch of every
kernel, grouped into the stages bench.py times (luma_gauss, ema_diff_thresh, morph_open, label = its five kernels).
    python tools/ncu_traffic.py gpurun_out/prof_all_r2.ncu-rep 128 > profiles/traffic_r2.json
"""
import csv
import json
import subprocess
import sys

STAGES = {'luma_gauss': ['gauss_mma_kernel<1, 2', 'gauss_stream_kernel<6, 1'], 'gauss': ['gauss_mma_kernel<0, 2', 'gauss_stream_kernel<6, 0'],
          'gauss_sigma15': ['gauss_mma_cta_kernel<7', 'gauss_mma_kernel<0, 7'], 'luma': ['luma_fast_kernel'],
          'ema_diff_thresh': ['ema_diff_thresh_kernel'], 'morph_open': ['morph_stream_kernel'],
          'label': ['label_init_kernel', 'label_merge_kernel', 'label_flatten_kernel', 'label_scan_kernel', 'label_write_kernel'],
          'label_forest': ['label_init_kernel', 'label_merge_kernel', 'label_flatten_kernel', 'label_scan_kernel'],
          'label_write': ['label_write_kernel'],
          'export_chunks': ['export_count_kernel', 'export_scan_kernel', 'export_write_kernel']}


def main():
    rep, batch = sys.argv[1], int(sys.argv[2])
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    last = {}
    for r in rows[2:]:
        last[r[idx['Kernel Name']]] = r

    def num(r, key):
        return float(r[idx[key]].replace(',', ''))
    unit_r = rows[1][idx['dram__bytes_read.sum']]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    res = {}
    for stage, names in STAGES.items():
        rd = wr = us = 0.0
        found = []
        for n in names:
            for k, r in last.items():
                if k.startswith('void ' + n) or k.startswith(n):
                    rd += num(r, 'dram__bytes_read.sum') * scale.get(rows[1][idx['dram__bytes_read.sum']], 1)
                    wr += num(r, 'dram__bytes_write.sum') * scale.get(rows[1][idx['dram__bytes_write.sum']], 1)
                    us += num(r, 'gpu__time_duration.sum') * {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6}.get(rows[1][idx['gpu__time_duration.sum']], 1.0)
                    found.append(k.split('(')[0])
                    break
        if found:
            res[stage] = [{'batch': batch, 'dram_bytes_read': int(rd), 'dram_bytes_write': int(wr), 'kernels': found,
                           'gpu_time_us_under_ncu': round(us, 1), 'source': 'ncu --set full on tools/prof_once.py (PROF_BATCH=%d)' % batch}]
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
