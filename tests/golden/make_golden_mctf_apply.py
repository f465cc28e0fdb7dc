#!/usr/bin/env python3
"""Regenerates tests/golden/golden_v3_mctf_apply.npz (MCTF apply stage: applyFrac + applyPlanarCorrection + applyBlock per block, calcVar) from the
UNMODIFIED reference (oracle/_ref).  Run in the build container only:  python tests/golden/make_golden_mctf_apply.py
Expected pictures come from the AVX2 kernels and are cross-checked against the scalar ones (float results included)."""
import os, sys, ctypes
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases as C
import impls
from _libs import refshim, P


def main():
    R = refshim()
    R.refshim_mctf_calc_var.restype = ctypes.c_double
    out = {}
    for k, (seed, W, H, refs, bs, bd, tap4, planar) in enumerate(C.MCTF_APPLY_CASES):
        case = C.mctf_apply_case(seed, W, H, 24, refs, bs, bd)
        res = []
        for opt, simd in ((0, b'SCALAR'), (1, b'AVX2')):
            R.refshim_set_simd(simd)
            res.append(impls.mctf_apply_expected(R, 'refshim', case, tap4, planar, opt))
        assert np.array_equal(res[0], res[1]), ('scalar != AVX2', seed)
        out['apply_%d' % k] = res[1]
    rs = np.random.RandomState(99)
    shapes = [(8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (8, 16), (32, 8)]
    plane = rs.randint(0, 1024, size=(96, 160)).astype(np.int16)
    blocks = np.array([[int(rs.randint(0, 160 - w + 1)) // 8 * 8, int(rs.randint(0, 96 - h + 1)), w, h] for (w, h) in shapes * 4], dtype=np.int32)
    var = np.zeros(len(blocks), dtype=np.float64)
    for i, (x, y, w, h) in enumerate(blocks):
        blk = C.aligned((h, w), np.int16); blk[:] = plane[y:y + h, x:x + w]
        vals = []
        for opt, simd in ((0, b'SCALAR'), (1, b'AVX2')):
            R.refshim_set_simd(simd)
            vals.append(R.refshim_mctf_calc_var(opt, P(blk), int(w), int(w), int(h)))
        assert vals[0] == vals[1], ('scalar != AVX2', x, y, w, h)
        var[i] = vals[1]
    out['var_plane'] = plane; out['var_blocks'] = blocks; out['var_expect'] = var
    path = os.path.join(HERE, 'golden_v3_mctf_apply.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items()}, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
