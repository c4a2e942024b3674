"""SASS excerpts of the kernels the bench line names (cuobjdump -sass of csrc/libva_b200.so): opcode histogram of the
whole kernel, the special instructions present, and a window of the listing around the first occurrence of each marker.
    python tools/sass_excerpt.py > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'video_analysis_b200', 'csrc', 'libva_b200.so')

KERNELS = [
    # (substring of the mangled name, title, [(marker regex, lines before, lines after, caption)])
    ('gauss_mma_kernelILb1ELi2ELi8ELi4E', 'gauss_mma_kernel<fused luma, G=2, 8 tiles, min 4 CTAs/SM>  (K1+K2, the dominant launch)',
     [(r'UTMALDG', 6, 6, 'TMA box load issued by lane 0 (cp.async.bulk.tensor.3d) and its mbarrier'),
      (r'SYNCS\.PHASECHK', 4, 10, 'mbarrier wait of the warp, then the row pass: LDS.64 B fragments, IMMA.16832.U8.U8'),
      (r'IDP\.4A', 6, 24, 'RGB -> luma of the staged rows: IDP.4A channel sums, HFMA2 division by 3, PRMT packing'),
      ]),
    ('ema_diff_thresh_kernelILi16ELi32ELi1ELb1E', 'ema_diff_thresh_kernel<16 px, one warp per block, packed float32, unrolled ring>  (K3)',
     [(r'FFMA2', 14, 22, 'one frame of the background update on packed float32 pairs: PRMT byte -> 2^23 + b, FADD2 (- 2^23, - bg), '
                         'FADD thr - |d| + SHF mask bit, FFMA2 alpha * d (+ -0), FADD2 bg + ...'),
      (r'LDGSTS', 3, 6, 'next frame into the per-thread ring: predicated LDGSTS (cp.async) at an immediate slot offset'),
      ]),
]


def listing(sym):
    out = subprocess.run(['cuobjdump', '-sass', '-fun', sym, LIB], capture_output=True, text=True).stdout
    return [l.rstrip() for l in out.splitlines() if re.match(r'^\s+/\*[0-9a-f]{4}\*/', l)]


def main():
    syms = subprocess.run(['cuobjdump', '-elf', LIB], capture_output=True, text=True).stdout
    names = sorted(set(re.findall(r'\.text\.(_Z\w+)', syms)))
    for key, title, marks in KERNELS:
        sym = [n for n in names if key in n]
        if not sym:
            print('# %s: not found in %s' % (key, LIB))
            continue
        ins = listing(sym[0])
        ops = collections.Counter(re.sub(r'^(@!?U?P\w+\s+)?', '', l.split('*/', 1)[1].strip()).split()[0].split('.')[0].rstrip(';') for l in ins)
        special = sorted(set(m for l in ins for m in re.findall(r'(IMMA\.\w+(?:\.\w+)*|UTMALDG\.\w+|SYNCS\.[\w.]+|FADD2|FFMA2|LDGSTS(?:\.\w+)*|IDP\.\w+(?:\.\w+)*)', l)))
        print('# ' + '=' * 110)
        print('# %s' % title)
        print('# %s, %d instructions' % (sym[0], len(ins)))
        print('# opcode histogram: ' + ', '.join('%s %d' % kv for kv in ops.most_common(22)))
        print('# special instructions present: ' + ', '.join(special))
        for rx, before, after, cap in marks:
            idx = next((i for i, l in enumerate(ins) if re.search(rx, l)), None)
            if idx is None:
                print('\n# --- %s: marker %s not found' % (cap, rx))
                continue
            print('\n# --- %s' % cap)
            for l in ins[max(0, idx - before):idx + after + 1]:
                print(re.sub(r'\s+/\* 0x[0-9a-f]+ \*/$', '', l))
        print()


if __name__ == '__main__':
    sys.exit(main())
