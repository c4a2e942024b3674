"""Per-source-line hot spots of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_source.py rep.ncu-rep <kernel-regex> [launch-skip]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else '0'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + kern,
                      '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] in ('Address', '#', 'Line')]
print('sections', [(i, rows[i][:3]) for i in hi][:4])
