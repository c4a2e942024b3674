"""
Generates tests/golden/golden.npz: outputs of the oracle (the reference's library calls,
oracle/ops.py) on seeded synthetic frames, run in the build container with
numpy 2.3.5 / OpenCV 4.13.0 / SciPy 1.18.1.  The reference itself ships no golden vectors
and cannot be imported (Python 2), so these pin the *library behaviour* the oracle relies
on; tests/test_oracle.py recomputes them and tests/test_kernels.py compares the CUDA path
against them.      python tests/golden/make_golden.py
"""

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ops, synth  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = {}
    # small case stored in full: 6 frames 64x48, the config-1 chain
    fr = synth.make_frames(0, 0, 6, 64, 48, 4)
    r = ops.chain(fr, sigma=2, alpha=0.05, thr=25, morph_op='open', morph_ksize=3)
    out['s_frames'] = fr
    for k in ('mono', 'blur', 'mask', 'morph', 'labels', 'counts', 'bg'):
        out['s_' + k] = r[k]
    # ragged size 53x37, sigma 3, 8-connectivity, close with a 5x5 ellipse
    fr = synth.make_frames(3, 10, 5, 53, 37, 3)
    r = ops.chain(fr, sigma=3, alpha=0.1, thr=12, morph_op='close', morph_shape='ellipse', morph_ksize=5, connectivity=8)
    out['r_frames'] = fr
    for k in ('mono', 'blur', 'mask', 'morph', 'labels', 'counts', 'bg'):
        out['r_' + k] = r[k]
    # config 1 proper (640x480), 12 frames: hashes only
    fr = synth.make_frames(0, 0, 12, 640, 480, 8)
    r = ops.chain(fr)
    out['vga_hashes'] = np.array([sha(fr)] + [sha(r[k]) for k in ('mono', 'blur', 'mask', 'morph', 'labels', 'counts')])
    out['vga_counts'] = r['counts']
    # gaussian taps
    sig = np.array([0.3, 0.5, 1, 1.5, 2, 2.5, 3, 5, 7.5, 10, 15, 20])
    out['tap_sigmas'] = sig
    out['taps'] = np.array([np.pad(ops.gauss_kernel_u8(s), (0, 128 - len(ops.gauss_kernel_u8(s)))) for s in sig])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
