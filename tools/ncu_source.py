"""Per-source-line hot spots of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_source.py rep.ncu-rep <kernel-regex> [launch-skip] [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else '0'
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name',
                      'regex:' + kern, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = []
cur_file = ''
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
        ii = hdr.index('Instructions Executed')
        si = hdr.index('# Samples')
    elif hdr and r[0].isdigit():
        try:
            lines.append((int(r[ii] or 0), int(r[si] or 0), cur_file, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot_i = sum(l[0] for l in lines) or 1
tot_s = sum(l[1] for l in lines) or 1
print('total warp-instr %d  samples %d' % (tot_i, tot_s))
print('--- by instructions')
for l in sorted(lines, reverse=True)[:top]:
    print('%5.1f%% instr %5.1f%% smp  %s:%d  %s' % (100.0 * l[0] / tot_i, 100.0 * l[1] / tot_s, l[2], l[3], l[4][:90]))
print('--- by stall samples')
for l in sorted(lines, key=lambda l: -l[1])[:top]:
    print('%5.1f%% smp %5.1f%% instr  %s:%d  %s' % (100.0 * l[1] / tot_s, 100.0 * l[0] / tot_i, l[2], l[3], l[4][:90]))
