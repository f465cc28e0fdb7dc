#!/usr/bin/env python3
"""Regenerates tests/golden/golden_v1.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference
by oracle/Makefile.ref).  Run in the build container only:  python tests/golden/make_golden.py
Every expected value is produced by the reference's AVX2 path and cross-checked against its scalar path."""
import os, sys, ctypes
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases as C
from _libs import refshim, P, PO

I32 = ctypes.c_int32


def main():
    R = refshim()
    out = {}
    # --- distortion kernels
    rows = C.dist_cases(); exp = np.zeros(len(rows), dtype=np.uint64)
    for i, row in enumerate(rows):
        fam, w, h, so, sc, ss, ko, kc, bd, seed = [int(v) for v in row]
        o, c = C.dist_inputs(row)
        a = R.refshim_dist(1, fam, P(o), so, P(c), sc, w, h, bd, ss)
        b = R.refshim_dist(0, fam, P(o), so, P(c), sc, w, h, bd, ss)
        assert a == b, ('scalar != AVX2', row)
        exp[i] = a
    out['dist_rows'] = rows; out['dist_expect'] = exp
    # --- transform + quant
    rows = C.tq_cases(); coefs = []; qs = []; meta = np.zeros((len(rows), 3), dtype=np.int32)
    for i, row in enumerate(rows):
        th, tv, w, h, st, amp, qp, irap, bd, seed = [int(v) for v in row]
        resi = C.tq_inputs(row)
        res = []
        for simd in (b'SCALAR', b'AVX2'):
            R.refshim_set_simd(simd)
            coef = np.zeros((h, w), dtype=np.int32); q = np.zeros((h, w), dtype=np.int16); s = I32(); lp = I32()
            assert R.refshim_transform_quant(th, tv, P(resi), st, w, h, bd, qp, irap, P(coef), P(q), ctypes.byref(s), ctypes.byref(lp)) == 0
            nr = R.refshim_need_rdoq(P(coef), w, h, bd, qp, seed & 1)
            res.append((coef, q, s.value, lp.value, nr))
        assert all(np.array_equal(res[0][k], res[1][k]) for k in range(5)), ('scalar != AVX2', row)
        coefs.append(res[1][0].reshape(-1)); qs.append(res[1][1].reshape(-1)); meta[i] = res[1][2:]
    out['tq_rows'] = rows; out['tq_coef'] = np.concatenate(coefs); out['tq_q'] = np.concatenate(qs); out['tq_meta'] = meta
    # --- MCTF
    rows = C.mctf_cases(); exp = np.zeros(len(rows), dtype=np.int32)
    for i, row in enumerate(rows):
        w, h, mvx, mvy, tap4, bd, seed = [int(v) for v in row]
        org, buf = C.mctf_inputs(row)
        m = C.MCTF_MARGIN; S = buf.shape[1]
        desc = np.array([[m, m, mvx, mvy, w, h]], dtype=np.int32)
        vals = []
        for opt in (0, 1):
            o = np.zeros(1, dtype=np.int32)
            # org is addressed at (m,m) too: shift its base pointer back so plane coords line up
            R.refshim_mctf_err_list(opt, tap4, PO(org, -(m * org.shape[1] + m)), org.shape[1], P(buf), S, P(desc), 1, bd, P(o), 1)
            vals.append(int(o[0]))
        assert vals[0] == vals[1], ('scalar != AVX2', row, vals)
        exp[i] = vals[1]
    out['mctf_rows'] = rows; out['mctf_expect'] = exp
    # --- affine
    rows = C.affine_cases(); sob = []; eqs = np.zeros((len(rows), 49), dtype=np.int64)
    for i, row in enumerate(rows):
        w, h, ps, ds, six, seed = [int(v) for v in row]
        pred, resi, gx, gy = C.affine_inputs(row)
        d = [[np.zeros((h, ds), dtype=np.int16) for _ in range(2)] for _ in range(2)]
        e = [np.zeros(49, dtype=np.int64) for _ in range(2)]
        for opt in (0, 1):
            for vert in (0, 1):
                R.refshim_sobel(opt, vert, P(pred), ps, P(d[opt][vert]), ds, w, h)
            R.refshim_equal_coeff(opt, six, P(resi), ps, P(gx), P(gy), ds, w, h, P(e[opt]))
        assert np.array_equal(d[0][0], d[1][0]) and np.array_equal(d[0][1], d[1][1]) and np.array_equal(e[0], e[1]), row
        sob.append(d[1][0][:, :w].reshape(-1)); sob.append(d[1][1][:, :w].reshape(-1)); eqs[i] = e[1]
    out['aff_rows'] = rows; out['aff_sobel'] = np.concatenate(sob); out['aff_eq'] = eqs
    # --- full search replay
    sc = C.search_case()
    n = len(sc['blk']); S = sc['stride']; m = sc['margin']; base = m * S + m
    for ss in (0, 1):
        res = []
        for opt in (0, 1):
            o = np.zeros((n, 4), dtype=np.int32)
            R.refshim_full_search(opt, PO(sc['org'], base), S, PO(sc['ref'], base), S, P(sc['blk']), n, 10, ss, sc['lam'], sc['cost_scale'],
                                  sc['imv_shift'], P(o), None, 0, 1, 0)
            res.append(o)
        assert np.array_equal(res[0], res[1])
        out['search_best_ss%d' % ss] = res[1]
    # --- MV rate
    rs = np.random.RandomState(4242)
    mv = np.zeros((512, 6), dtype=np.int32); mvb = np.zeros(512, dtype=np.uint32); mvc = np.zeros(512, dtype=np.uint64)
    for i in range(512):
        mv[i] = [rs.randint(-300, 300), rs.randint(-300, 300), rs.randint(-1200, 1200), rs.randint(-1200, 1200), rs.randint(0, 3), rs.randint(0, 3)]
        a = [int(v) for v in mv[i]]
        mvb[i] = R.refshim_mv_bits(*a); mvc[i] = R.refshim_mv_cost(57.25 + i, *a)
    out['mv_rows'] = mv; out['mv_bits'] = mvb; out['mv_cost'] = mvc
    np.savez_compressed(os.path.join(HERE, 'golden_v1.npz'), **out)
    print('wrote golden_v1.npz:', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
