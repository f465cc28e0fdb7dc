#!/usr/bin/env python3
"""Regenerates tests/golden/golden_v4_frac.npz (fractional-pel refinement grid: two-pass 8-tap interpolation + SAD / SATD) from the UNMODIFIED
reference (oracle/_ref).  Run in the build container only:  python tests/golden/make_golden_frac.py   (AVX2 == scalar asserted)"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases as C
from _libs import refshim, P, PO


def main():
    R = refshim()
    out = {}
    for ci, (seed, bd) in enumerate(C.FRAC_CASES):
        case = C.frac_case(seed, bit_depth=bd)
        S = case['stride']; base = case['margin'] * S + case['margin']
        for li, (fam, w, h, b) in enumerate(case['lists']):
            rt, alt = C.frac_filter_of(li)
            res = []
            for opt, simd in ((0, b'SCALAR'), (1, b'AVX2')):
                R.refshim_set_simd(simd)
                t = np.zeros((len(b), 7, 7), dtype=np.uint32)
                R.refshim_frac_cost_grid(opt, PO(case['org'], base), S, PO(case['ref'], base), S, P(np.ascontiguousarray(b)), len(b), fam, bd, rt, alt, P(t))
                res.append(t)
            assert np.array_equal(res[0], res[1]), ('scalar != AVX2', seed, fam, w, h)
            out['c%d_l%d' % (ci, li)] = res[1]
    path = os.path.join(HERE, 'golden_v4_frac.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(out), 'tables', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
